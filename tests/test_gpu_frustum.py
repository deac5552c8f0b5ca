"""ens_frustum_mask / ens_keyframe_overlap (mapper_ops.FrustumSelector) against the oracle (bitwise) and the reference's own
goldens (tests/golden/frustum.npz, made with cv2.remap + numpy); needs a B200 (``-m gpu``)."""
import os
import sys

import numpy as np
import pytest
import torch

import frustum_oracle as fo
from util import load_golden

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
import frustum_cases as fc  # noqa: E402

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _selector(case):
    from evennicer_slam_b200.mapper_ops import FrustumSelector
    return FrustumSelector(*case["cam"], case["bound"] if "bound" in case else fc.ROOM0_BOUND, DEV)


@pytest.mark.parametrize("name", ["room0", "rpg", "edge"])
def test_frustum_mask_matches_reference_golden_and_oracle(name):
    g = load_golden("frustum.npz")
    case = fc.mask_cases()[name]
    sel = _selector(case)
    for key, shape in case["shapes"].items():
        m = sel.get_mask_from_c2w(torch.from_numpy(case["c2w"]).to(DEV), key, list(shape), case["depth"])
        assert m.dtype == np.bool_ and m.shape == (shape[2], shape[1], shape[0])
        want = np.unpackbits(g[f"{name}.{key}.mask"])[:m.size].astype(bool).reshape(m.shape)
        assert np.array_equal(m, want), f"{key}: {int((m != want).sum())} voxels differ from the reference"
        orc = fo.get_mask_from_c2w(case["c2w"], shape, case["depth"], case["bound"], case["cam"],
                                   lambda a, b, n: torch.linspace(a, b, n).numpy())
        assert np.array_equal(m, orc)
        # device-resident form in the grids' layout: what FrustumGridAdam takes
        mz = sel.voxel_mask(case["c2w"], key, shape, torch.from_numpy(case["depth"]).to(DEV))
        assert mz.dtype == torch.bool and tuple(mz.shape) == tuple(shape)
        assert torch.equal(mz.cpu(), torch.from_numpy(want).permute(2, 1, 0))


def test_coarse_grid_selects_everything():
    case = fc.mask_cases()["rpg"]
    m = _selector(case).get_mask_from_c2w(case["c2w"], "grid_coarse", [3, 4, 5], case["depth"])
    assert m.shape == (5, 4, 3) and m.all()


def test_random_poses_match_oracle_bitwise():
    case = fc.mask_cases()["rpg"]
    sel = _selector(case)
    rng = np.random.RandomState(4)
    for _ in range(6):
        c2w = fc._pose(rng, rng.uniform(-3, 3, 3) + [0, 0, 2], rng.uniform(-3, 3), rng.uniform(-1.2, 1.2))
        shape = (13, 17, 19)
        m = sel.get_mask_from_c2w(c2w, "grid_fine", shape, case["depth"])
        orc = fo.get_mask_from_c2w(c2w, shape, case["depth"], case["bound"], case["cam"],
                                   lambda a, b, n: torch.linspace(a, b, n).numpy())
        assert np.array_equal(m, orc)


@pytest.mark.parametrize("name", ["room0", "rpg"])
def test_keyframe_overlap_matches_reference_golden(name):
    g = load_golden("frustum.npz")
    case = fc.overlap_cases()[name]
    sel = _selector(case)
    # the reference's draws, replayed: torch.randint on the CPU generator (the golden was made on the CPU)
    import render_oracle as ro
    cam = case["cam"]
    torch.manual_seed(case["seed"])
    idx = torch.randint(cam[0] * cam[1], (case["pixels"],)).numpy()
    i, j, d, _ = ro.select_pixels(idx, 0, cam[0], 0, cam[1], case["depth"], case["color"])
    o, dr = ro.rays_from_uv(i, j, case["c2w"], cam[2], cam[3], cam[4], cam[5])
    pts = fo.overlap_sample_points(o, dr, d, torch.linspace(0., 1., steps=16).numpy())
    cnt = sel.keyframe_overlap_counts(torch.from_numpy(pts).to(DEV), [torch.from_numpy(c).to(DEV) for c in case["kf_c2w"]])
    assert np.array_equal(cnt / pts.shape[0], g[f"{name}.percent_inside"])
    w2cs = np.stack([np.linalg.inv(c) for c in case["kf_c2w"]])
    assert np.array_equal(cnt, fo.keyframe_overlap(pts, w2cs, cam))


def test_keyframe_selection_overlap_drop_in():
    case = fc.overlap_cases()["rpg"]
    sel = _selector(case)
    kfd = [{"est_c2w": torch.from_numpy(c).to(DEV)} for c in case["kf_c2w"]]
    torch.manual_seed(3); np.random.seed(3)
    out = sel.keyframe_selection_overlap(torch.from_numpy(case["color"]).to(DEV), torch.from_numpy(case["depth"]).to(DEV),
                                         torch.from_numpy(case["c2w"]).to(DEV), kfd, case["k"], 16, case["pixels"])
    assert len(out) == case["k"] and len(set(int(x) for x in out)) == case["k"]
    assert all(0 <= int(x) < len(kfd) for x in out)
    assert sel.keyframe_selection_overlap(torch.from_numpy(case["color"]).to(DEV), torch.from_numpy(case["depth"]).to(DEV),
                                          torch.from_numpy(case["c2w"]).to(DEV), [], 4) == []
