"""oracle/frustum_oracle.py against the reference's own get_mask_from_c2w / keyframe_selection_overlap (goldens made by
tests/golden/make_frustum_golden.py with cv2.remap and numpy), CPU."""
import os
import sys

import numpy as np
import pytest
import torch

import frustum_oracle as fo
from util import load_golden

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
import frustum_cases as fc  # noqa: E402


def _linspace(a, b, n):
    return torch.linspace(a, b, n).numpy()                      # the reference's own call (Mapper.py:132-134)


def test_remap_restatement_matches_cv2_bitwise():
    cv2 = pytest.importorskip("cv2")
    rng = np.random.RandomState(0)
    img = rng.rand(37, 53).astype(np.float32) * 4
    img[rng.rand(37, 53) < 0.1] = 0
    n = 200000
    u = rng.uniform(-3, 56, n).astype(np.float32)
    v = rng.uniform(-3, 40, n).astype(np.float32)
    u[:8] = [0, 52, 52.5, -1, -0.99, 1e9, -1e9, 51.984375]
    v[:8] = [0, 36, 36.5, -1, 35.99, 3, 3, 1e30]
    want = np.concatenate([cv2.remap(img, u[i:i + 30000].reshape(-1, 1), v[i:i + 30000].reshape(-1, 1),
                                     interpolation=cv2.INTER_LINEAR)[:, 0] for i in range(0, n, 30000)])   # dst rows < SHRT_MAX
    got = fo.remap_linear(img, u, v)
    assert np.array_equal(got, want)


@pytest.mark.parametrize("name", ["room0", "rpg", "edge"])
def test_mask_oracle_matches_reference_golden(name):
    g = load_golden("frustum.npz")
    case = fc.mask_cases()[name]
    for key, shape in case["shapes"].items():
        m = fo.get_mask_from_c2w(case["c2w"], shape, case["depth"], case["bound"], case["cam"], _linspace)
        want = np.unpackbits(g[f"{name}.{key}.mask"])[:m.size].astype(bool).reshape(m.shape)
        assert int(m.sum()) == int(g[f"{name}.{key}.count"])
        assert np.array_equal(m, want)


@pytest.mark.parametrize("name", ["room0", "rpg"])
def test_overlap_oracle_matches_reference_golden(name):
    g = load_golden("frustum.npz")
    case = fc.overlap_cases()[name]
    import render_oracle as ro
    torch.manual_seed(case["seed"])
    cam = case["cam"]
    idx = torch.randint(cam[0] * cam[1], (case["pixels"],)).numpy()             # the draw of common.py:99
    i, j, d, _ = ro.select_pixels(idx, 0, cam[0], 0, cam[1], case["depth"], case["color"])
    ro_, rd_ = ro.rays_from_uv(i, j, case["c2w"], cam[2], cam[3], cam[4], cam[5])
    pts = fo.overlap_sample_points(ro_, rd_, d, torch.linspace(0., 1., steps=16).numpy())
    w2cs = np.stack([np.linalg.inv(c) for c in case["kf_c2w"]])
    cnt = fo.keyframe_overlap(pts, w2cs, cam)
    assert np.array_equal(cnt / pts.shape[0], g[f"{name}.percent_inside"])
