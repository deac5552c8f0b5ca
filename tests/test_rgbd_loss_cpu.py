"""Pins oracle/rgbd_loss_oracle.py against the reference's loss lines (Mapper.py:553-562, Tracker.py:180-196) run with
torch autograd on CPU.  No GPU, no product code."""
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle")]
import rgbd_loss_oracle as ro  # noqa: E402


def make(n, seed, zero_frac=0.1, outliers=True):
    rng = np.random.RandomState(seed)
    gt_depth = (rng.rand(n) * 3 + 0.5).astype(np.float32)
    gt_depth[rng.rand(n) < zero_frac] = 0.0
    depth = gt_depth.astype(np.float64) + rng.randn(n) * 0.05
    if outliers:
        depth[rng.rand(n) < 0.05] += 5.0                        # residuals beyond 10 x median: gated out by the tracker
    unc = rng.rand(n) * 0.02
    color = rng.rand(n, 3).astype(np.float32)
    gt_color = rng.rand(n, 3)                                   # float64, as the dataset delivers it (datasets.py:179-188)
    return gt_depth, gt_color, depth, unc, color


@pytest.mark.parametrize("n,use_color", [(1000, True), (1000, False), (7, True)])
def test_mapper_loss_oracle(n, use_color):
    gt_depth, gt_color, depth, unc, color = make(n, 20)
    d = torch.from_numpy(depth).requires_grad_(True); c = torch.from_numpy(color).requires_grad_(True)
    g, gc = torch.from_numpy(gt_depth), torch.from_numpy(gt_color)
    depth_mask = (g > 0)                                                            # Mapper.py:553
    loss = torch.abs(g[depth_mask] - d[depth_mask]).sum()                           # :557-558
    if use_color:
        loss = loss + 0.2 * torch.abs(gc - c).sum()                                 # :560-562
    loss.backward()
    o_loss, o_gd, o_gc = ro.mapper_loss(gt_depth, gt_color, depth, color, 0.2, use_color)
    assert abs(o_loss - loss.item()) < 1e-12 * abs(loss.item())
    assert np.array_equal(o_gd, d.grad.numpy())
    if use_color:
        assert np.array_equal(o_gc, c.grad.numpy())


@pytest.mark.parametrize("n,use_color,dyn", [(200, True, True), (200, True, False), (201, False, True), (5000, True, True),
                                             (2, True, True)])
def test_tracker_loss_oracle(n, use_color, dyn):
    gt_depth, gt_color, depth, unc, color = make(n, 3)
    d = torch.from_numpy(depth).requires_grad_(True); c = torch.from_numpy(color).requires_grad_(True)
    g, gc, u = torch.from_numpy(gt_depth), torch.from_numpy(gt_color), torch.from_numpy(unc).requires_grad_(True)
    uncertainty = u.detach()                                                        # Tracker.py:179
    if dyn:
        tmp = torch.abs(g - d) / torch.sqrt(uncertainty + 1e-10)                    # :181
        mask = (tmp < 10 * tmp.median()) & (g > 0)                                  # :182
    else:
        mask = g > 0
    loss = (torch.abs(g - d) / torch.sqrt(uncertainty + 1e-10))[mask].sum()         # :188-189
    if use_color:
        loss = loss + 0.2 * torch.abs(gc - c)[mask].sum()                           # :191-194
    loss.backward()
    o_loss, o_gd, o_gc = ro.tracker_loss(gt_depth, gt_color, depth, unc, color, 0.2, use_color, dyn)
    assert abs(o_loss - loss.item()) < 1e-12 * abs(loss.item())
    assert np.allclose(o_gd, d.grad.numpy(), rtol=1e-14, atol=0)
    if use_color:
        assert np.array_equal(o_gc, c.grad.numpy())
    if dyn and n >= 200:
        assert 0 < int((o_gd != 0).sum()) < int((gt_depth > 0).sum())               # the gate removed something
