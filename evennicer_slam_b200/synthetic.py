"""Deterministic synthetic scenes for parity tests, goldens and the bench.

Everything here is numpy-only and bit-reproducible on any host (integer hashing +
exactly-rounded IEEE arithmetic), because the golden fixtures under ``tests/golden``
store only OUTPUTS of the reference; the inputs are regenerated from these
functions on the GPU box.

Shapes and constants restate the reference's scene set-up (nothing is imported
from it):
  * bound padding            -- /root/reference/src/EvenNICER_SLAM.py:162-182
  * grid shapes/axis swap    -- /root/reference/src/EvenNICER_SLAM.py:217-275
  * decoder tensor shapes    -- /root/reference/src/conv_onet/models/decoder.py:91-166, 206-252, 277-310
  * Replica / RPG cameras    -- /root/reference/configs/Replica/replica.yaml:37-45, configs/rpg/rpg.yaml:62-71
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, List, Optional, Tuple

import numpy as np

# ----------------------------------------------------------------------------
# deterministic random numbers (splitmix64 on the element index)
# ----------------------------------------------------------------------------

_M64 = np.uint64(0xFFFFFFFFFFFFFFFF)


def _splitmix64(x: np.ndarray) -> np.ndarray:
    with np.errstate(over="ignore"):
        x = (x + np.uint64(0x9E3779B97F4A7C15)) & _M64
        z = x
        z = ((z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)) & _M64
        z = ((z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)) & _M64
        return z ^ (z >> np.uint64(31))


def det_uniform(shape, seed: int) -> np.ndarray:
    """U[0,1) float64, a pure function of (seed, flat index)."""
    n = int(np.prod(shape)) if len(tuple(np.atleast_1d(shape))) else 1
    idx = np.arange(n, dtype=np.uint64)
    with np.errstate(over="ignore"):
        h = _splitmix64(idx ^ _splitmix64(np.full(1, seed, dtype=np.uint64)))
    u = (h >> np.uint64(11)).astype(np.float64) * (1.0 / 9007199254740992.0)
    return u.reshape(shape)


def det_normal(shape, seed: int) -> np.ndarray:
    """Approximately N(0,1) float64 (Irwin-Hall of 4 uniforms), deterministic."""
    acc = np.zeros(shape, dtype=np.float64)
    for k in range(4):
        acc += det_uniform(shape, seed * 7919 + 104729 * (k + 1))
    return (acc - 2.0) * np.sqrt(3.0)


# ----------------------------------------------------------------------------
# scene geometry
# ----------------------------------------------------------------------------

ROOM0_BOUND = [[-2.9, 8.9], [-3.2, 5.5], [-3.5, 3.3]]          # configs/Replica/room0.yaml:3
RPG4_BOUND = [[-7.0, 9.4], [-6.5, 3.6], [-9.2, 9.5]]            # configs/rpg/recording4.yaml:4
TINY_BOUND = [[-1.0, 1.4], [-1.1, 0.9], [-0.9, 1.0]]            # small scene for full-tensor goldens

GRID_LEN = {"coarse": 2.0, "middle": 0.32, "fine": 0.16, "color": 0.16}   # configs/nice_slam.yaml:7-12
BOUND_DIVISIBLE = 0.32
COARSE_BOUND_ENLARGE = 2                                                 # configs/nice_slam.yaml:112
C_DIM = 32
HIDDEN = 32
EMBED = 93
LEVELS = ("coarse", "middle", "fine", "color")


def padded_bound(raw_bound, scale: float = 1.0) -> np.ndarray:
    """float64 (3,2) bound with the float32-rounded upper edge the reference produces.

    /root/reference/src/EvenNICER_SLAM.py:170-175: the ``int*0.32`` product is an
    int32 tensor times a python float -> float32, which is then added to the
    float64 lower edge.
    """
    b = np.array(raw_bound, dtype=np.float64) * scale
    n = ((b[:, 1] - b[:, 0]) / BOUND_DIVISIBLE).astype(np.int32) + 1
    span32 = n.astype(np.float32) * np.float32(BOUND_DIVISIBLE)
    b[:, 1] = span32.astype(np.float64) + b[:, 0]
    return b


def grid_shape(bound: np.ndarray, level: str) -> Tuple[int, int, int]:
    """(Z, Y, X) of ``grid_<level>`` -- EvenNICER_SLAM.py:241-273 (axis 0<->2 swap)."""
    xyz_len = bound[:, 1] - bound[:, 0]
    if level == "coarse":
        xyz_len = xyz_len * COARSE_BOUND_ENLARGE
    n = [int(v) for v in (xyz_len / GRID_LEN[level]).tolist()]
    return (n[2], n[1], n[0])


@dataclass
class Camera:
    H: int
    W: int
    fx: float
    fy: float
    cx: float
    cy: float


REPLICA_CAM = Camera(680, 1200, 600.0, 600.0, 599.5, 339.5)
RPG_CAM = Camera(260, 346, 196.71854278974607, 196.68898128242577, 172.5, 129.5)
TINY_CAM = Camera(48, 64, 40.0, 40.0, 31.5, 23.5)


def decoder_param_shapes(name: str) -> List[Tuple[str, Tuple[int, ...]]]:
    """state_dict keys (in registration order) and shapes of one decoder.

    MLP: decoder.py:108-164; MLP_no_xyz: decoder.py:224-250.
    """
    out: List[Tuple[str, Tuple[int, ...]]] = []
    if name == "coarse":
        dims = [HIDDEN, HIDDEN, HIDDEN, HIDDEN + C_DIM, HIDDEN]
        for i, k in enumerate(dims):
            out.append((f"pts_linears.{i}.weight", (HIDDEN, k)))
            out.append((f"pts_linears.{i}.bias", (HIDDEN,)))
        out.append(("output_linear.weight", (1, HIDDEN)))
        out.append(("output_linear.bias", (1,)))
        return out
    c_dim = 2 * C_DIM if name == "fine" else C_DIM
    n_out = 4 if name == "color" else 1
    for i in range(5):
        out.append((f"fc_c.{i}.weight", (HIDDEN, c_dim)))
        out.append((f"fc_c.{i}.bias", (HIDDEN,)))
    out.append(("embedder._B", (3, EMBED)))
    dims = [EMBED, HIDDEN, HIDDEN, HIDDEN + EMBED, HIDDEN]
    for i, k in enumerate(dims):
        out.append((f"pts_linears.{i}.weight", (HIDDEN, k)))
        out.append((f"pts_linears.{i}.bias", (HIDDEN,)))
    out.append(("output_linear.weight", (n_out, HIDDEN)))
    out.append(("output_linear.bias", (n_out,)))
    return out


def make_decoder_params(name: str, seed: int, occ_soften: bool = True) -> Dict[str, np.ndarray]:
    """Random-init decoder tensors with the reference's init distributions.

    Xavier-uniform weights with relu/linear gain for DenseLayer (decoder.py:70-79),
    U(+-1/sqrt(fan_in)) for the ``fc_c`` nn.Linear layers, N(0,25^2) for ``_B``
    (decoder.py:17-24).  Biases, zero in the reference's init, are drawn small and
    non-zero here so the parity tests exercise them.
    """
    params: Dict[str, np.ndarray] = {}
    for k, (key, shp) in enumerate(decoder_param_shapes(name)):
        s = seed * 1000 + k
        if key == "embedder._B":
            v = det_normal(shp, s) * 25.0
        elif key.endswith("bias"):
            v = (det_uniform(shp, s) - 0.5) * 0.2
        elif key.startswith("fc_c"):
            a = 1.0 / np.sqrt(shp[1])
            v = (det_uniform(shp, s) * 2.0 - 1.0) * a
        else:
            gain = 1.0 if key.startswith("output_linear") else np.sqrt(2.0)
            a = gain * np.sqrt(6.0 / (shp[0] + shp[1]))
            v = (det_uniform(shp, s) * 2.0 - 1.0) * a
        params[key] = np.ascontiguousarray(v.astype(np.float32))
    if occ_soften and name != "color":
        # keep sigmoid(10*occ) away from saturation so rays traverse many samples and every
        # gradient path carries signal (random output layers otherwise give |occ| ~ 1)
        params["output_linear.weight"] *= np.float32(0.03)
        params["output_linear.bias"][:] = np.float32({"coarse": -0.2, "middle": -0.15, "fine": -0.1}[name])
    return params


@dataclass
class Scene:
    """A synthetic NICE scene: bounds, four feature grids and four decoders (numpy)."""
    bound: np.ndarray                                   # (3,2) float64, fine/middle/color decoders
    coarse_bound: np.ndarray                            # bound * coarse_bound_enlarge
    grids: Dict[str, np.ndarray]                        # 'grid_<level>' -> (1,32,Z,Y,X) float32
    decoders: Dict[str, Dict[str, np.ndarray]]          # level -> state_dict (numpy)
    cam: Camera
    name: str = "scene"


def make_scene(raw_bound=ROOM0_BOUND, cam: Camera = REPLICA_CAM, seed: int = 20,
               name: str = "room0", grid_std: Optional[Dict[str, float]] = None,
               levels=LEVELS) -> Scene:
    """Grids ~ N(0, 0.01) (fine: N(0, 1e-4)) as in grid_init (EvenNICER_SLAM.py:248-272).

    ``grid_std`` overrides the per-level std (parity tests use larger values so
    occupancies are not all ~0 and every gradient path carries signal).
    """
    bound = padded_bound(raw_bound)
    std = {"coarse": 0.01, "middle": 0.01, "fine": 0.0001, "color": 0.01}
    if grid_std:
        std.update(grid_std)
    grids = {}
    for li, lv in enumerate(LEVELS):
        if lv not in levels:
            continue
        z, y, x = grid_shape(bound, lv)
        g = det_normal((1, C_DIM, z, y, x), seed * 31 + li) * std[lv]
        grids["grid_" + lv] = np.ascontiguousarray(g.astype(np.float32))
    decs = {lv: make_decoder_params(lv, seed * 17 + li) for li, lv in enumerate(LEVELS)}
    return Scene(bound=bound, coarse_bound=bound * COARSE_BOUND_ENLARGE, grids=grids,
                 decoders=decs, cam=cam, name=name)


# ----------------------------------------------------------------------------
# synthetic frames
# ----------------------------------------------------------------------------

def quat_to_c2w(cam_tensor: np.ndarray) -> np.ndarray:
    """[qw,qx,qy,qz,tx,ty,tz] -> (3,4) float32; restates common.py:189-228 (float32 math)."""
    q = cam_tensor[:4].astype(np.float32)
    t = cam_tensor[4:].astype(np.float32)
    qr, qi, qj, qk = q
    two_s = np.float32(2.0) / np.float32((q * q).sum())
    R = np.zeros((3, 3), dtype=np.float32)
    R[0, 0] = 1 - two_s * (qj ** 2 + qk ** 2)
    R[0, 1] = two_s * (qi * qj - qk * qr)
    R[0, 2] = two_s * (qi * qk + qj * qr)
    R[1, 0] = two_s * (qi * qj + qk * qr)
    R[1, 1] = 1 - two_s * (qi ** 2 + qk ** 2)
    R[1, 2] = two_s * (qj * qk - qi * qr)
    R[2, 0] = two_s * (qi * qk - qj * qr)
    R[2, 1] = two_s * (qj * qk + qi * qr)
    R[2, 2] = 1 - two_s * (qi ** 2 + qj ** 2)
    return np.concatenate([R, t[:, None]], axis=1).astype(np.float32)


def default_pose(raw_bound, jitter_seed: Optional[int] = None) -> np.ndarray:
    """Camera at the centre of the (un-padded) box, identity rotation (looking -z).

    For room0 this is the [1,0,0,0, 3.0,1.15,-0.1]-like pose of SURVEY 8(d); a small
    deterministic perturbation of quaternion and translation is added when
    ``jitter_seed`` is given (tracking / keyframe-window variants).
    """
    b = np.array(raw_bound, dtype=np.float64)
    c = b.mean(axis=1)
    cam = np.array([1.0, 0.0, 0.0, 0.0, c[0], c[1], c[2]], dtype=np.float64)
    if jitter_seed is not None:
        u = det_uniform((7,), 9001 + jitter_seed) - 0.5
        cam[:4] += u[:4] * 0.2
        cam[4:] += u[4:] * 0.3 * (b[:, 1] - b[:, 0]) / 4.0
    return cam.astype(np.float32)


def analytic_depth(raw_bound, cam: Camera, c2w: np.ndarray, shrink: float = 0.9,
                   zero_frac: float = 0.0, seed: int = 20) -> np.ndarray:
    """(H,W) float32 depth: ``shrink`` x distance along each pixel ray to the un-padded box.

    With ``zero_frac`` > 0 that fraction of pixels gets depth 0 (sensor holes) to
    exercise Renderer.py:144-151.
    """
    b = np.array(raw_bound, dtype=np.float64)
    jj, ii = np.meshgrid(np.arange(cam.H, dtype=np.float64), np.arange(cam.W, dtype=np.float64),
                         indexing="ij")
    dirs = np.stack([(ii - cam.cx) / cam.fx, -(jj - cam.cy) / cam.fy, -np.ones_like(ii)], -1)
    R = c2w[:3, :3].astype(np.float64)
    o = c2w[:3, 3].astype(np.float64)
    d = dirs @ R.T
    with np.errstate(divide="ignore", invalid="ignore"):
        t = (b[None, None, :, :] - o[None, None, :, None]) / d[..., None]
    far = np.min(np.max(t, axis=-1), axis=-1)
    depth = (far * shrink).astype(np.float32)
    if zero_frac > 0:
        u = det_uniform((cam.H, cam.W), seed + 77)
        depth = np.where(u < zero_frac, np.float32(0.0), depth).astype(np.float32)
    return depth


def synthetic_frame(raw_bound, cam: Camera, cam_tensor: np.ndarray, seed: int = 20,
                    zero_frac: float = 0.0):
    """(depth f32 (H,W), color f64 (H,W,3), event u8 (H,W,2)) -- dtypes of datasets.py:179-188."""
    c2w = quat_to_c2w(cam_tensor)
    depth = analytic_depth(raw_bound, cam, c2w, zero_frac=zero_frac, seed=seed)
    color = det_uniform((cam.H, cam.W, 3), seed + 1)
    ev = det_uniform((cam.H, cam.W, 2), seed + 2)
    event = (ev < 0.26).astype(np.uint8) + (ev < 0.04).astype(np.uint8)   # ~Poisson(0.3) clipped at 2
    return depth, color, event
