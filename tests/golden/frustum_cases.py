"""Seeded inputs of the frustum-selection / keyframe-overlap goldens (shared by make_frustum_golden.py and the tests)."""
import numpy as np


def _pose(rng, centre, yaw, pitch):
    cy_, sy_, cp, sp = np.cos(yaw), np.sin(yaw), np.cos(pitch), np.sin(pitch)
    Ry = np.array([[cy_, 0, sy_], [0, 1, 0], [-sy_, 0, cy_]])
    Rx = np.array([[1, 0, 0], [0, cp, -sp], [0, sp, cp]])
    c2w = np.eye(4)
    c2w[:3, :3] = Ry @ Rx
    c2w[:3, 3] = centre
    return c2w.astype(np.float32)


def _depth(rng, H, W, lo, hi, holes=True):
    y, x = np.mgrid[0:H, 0:W]
    d = lo + (hi - lo) * (0.5 + 0.25 * np.sin(x / W * 7.0) + 0.25 * np.cos(y / H * 5.0))
    d = d + rng.rand(H, W) * 0.05
    if holes:
        d[rng.rand(H, W) < 0.03] = 0.0                          # invalid depth pixels (Mapper.py:166-168)
        d[H // 3:H // 3 + 20, W // 4:W // 4 + 40] = 0.0
    return d.astype(np.float32)


ROOM0_BOUND = np.array([[-2.9, 8.9], [-3.2, 5.5], [-3.5, 3.3]], np.float32)
ROOM0_CAM = (680, 1200, 600.0, 600.0, 599.5, 339.5)
RPG_BOUND = np.array([[-4.0, 4.0], [-4.0, 4.0], [-1.0, 5.0]], np.float32)
RPG_CAM = (260, 346, 196.7, 196.6, 173.7, 134.0)


def _grid_shape(bound, voxel):
    n = ((bound[:, 1] - bound[:, 0]) / voxel).astype(int)
    xyz = list(map(int, n))
    return (xyz[2], xyz[1], xyz[0])                             # val.shape[2:] = (Z, Y, X)


def mask_cases():
    rng = np.random.RandomState(11)
    out = {}
    out["room0"] = {"bound": ROOM0_BOUND, "cam": ROOM0_CAM, "c2w": _pose(rng, [2.5, 0.8, 0.2], 0.6, -0.2),
                    "depth": _depth(rng, 680, 1200, 1.0, 5.0),
                    "shapes": {"grid_middle": _grid_shape(ROOM0_BOUND, 0.32), "grid_fine": _grid_shape(ROOM0_BOUND, 0.16)}}
    out["rpg"] = {"bound": RPG_BOUND, "cam": RPG_CAM, "c2w": _pose(rng, [0.3, -0.5, 1.9], -2.2, 0.35),
                  "depth": _depth(rng, 260, 346, 0.5, 3.0),
                  "shapes": {"grid_middle": _grid_shape(RPG_BOUND, 0.32), "grid_color": _grid_shape(RPG_BOUND, 0.16)}}
    # camera near a wall of the bound looking out: most voxels behind / outside, depth image without holes
    out["edge"] = {"bound": RPG_BOUND, "cam": RPG_CAM, "c2w": _pose(rng, [3.7, 3.6, 4.6], 0.1, 1.2),
                   "depth": _depth(rng, 260, 346, 0.2, 0.9, holes=False),
                   "shapes": {"grid_fine": _grid_shape(RPG_BOUND, 0.16)}}
    return out


def overlap_cases():
    rng = np.random.RandomState(12)
    out = {}
    for name, cam, bound, n_kf in (("room0", ROOM0_CAM, ROOM0_BOUND, 12), ("rpg", RPG_CAM, RPG_BOUND, 30)):
        H, W = cam[0], cam[1]
        centre = bound.mean(1)
        kfs = [_pose(rng, centre + rng.uniform(-1, 1, 3), rng.uniform(-np.pi, np.pi), rng.uniform(-0.5, 0.5)) for _ in range(n_kf)]
        out[name] = {"cam": cam, "c2w": _pose(rng, centre, 0.4, 0.1), "kf_c2w": kfs, "k": 4, "pixels": 100, "seed": 77,
                     "depth": _depth(rng, H, W, 1.0, 4.0, holes=False), "color": rng.rand(H, W, 3).astype(np.float32)}
    return out
