"""Parity of the fused RGB-D losses (ens_rgbd_loss through losses.mapper_loss / tracker_loss) with the oracle
(oracle/rgbd_loss_oracle.py, pinned against torch autograd in tests/test_rgbd_loss_cpu.py).  ``-m gpu``.
float64 arithmetic: loss to 1e-12 relative (summation order), gradients exact / 1e-14."""
import numpy as np
import pytest
import torch

import rgbd_loss_oracle as ro
from test_rgbd_loss_cpu import make

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _t(x, grad=False):
    return torch.from_numpy(x).to(DEV).requires_grad_(grad)


@pytest.mark.parametrize("n,use_color,f64_color", [(1000, True, True), (1000, True, False), (1000, False, True),
                                                   (7, True, True), (65535, True, True)])
def test_mapper_loss(n, use_color, f64_color):
    from evennicer_slam_b200.losses import mapper_loss
    gt_depth, gt_color, depth, unc, color = make(n, 20)
    if not f64_color:
        gt_color = gt_color.astype(np.float32)
    d, c = _t(depth, True), _t(color, True)
    loss = mapper_loss(_t(gt_depth), _t(gt_color), d, c, 0.2, use_color)
    (2.0 * loss).backward()
    o_loss, o_gd, o_gc = ro.mapper_loss(gt_depth, gt_color, depth, color, 0.2, use_color)
    assert loss.dtype == torch.float64 and abs(loss.item() - o_loss) < 1e-12 * abs(o_loss)
    assert np.array_equal(d.grad.cpu().numpy(), 2.0 * o_gd)
    if use_color:
        assert np.array_equal(c.grad.cpu().numpy(), 2.0 * o_gc)
    else:
        assert c.grad is None


@pytest.mark.parametrize("n,use_color,dyn", [(200, True, True), (200, True, False), (201, False, True),
                                             (5000, True, True), (2, True, True), (1, True, True)])
def test_tracker_loss(n, use_color, dyn):
    from evennicer_slam_b200.losses import tracker_loss
    gt_depth, gt_color, depth, unc, color = make(n, 3)
    d, c, u = _t(depth, True), _t(color, True), _t(unc, True)
    loss = tracker_loss(_t(gt_depth), _t(gt_color), d, u, c, 0.2, use_color, dyn)
    loss.backward()
    o_loss, o_gd, o_gc = ro.tracker_loss(gt_depth, gt_color, depth, unc, color, 0.2, use_color, dyn)
    assert abs(loss.item() - o_loss) < 1e-12 * max(abs(o_loss), 1e-300)
    assert np.allclose(d.grad.cpu().numpy(), o_gd, rtol=1e-14, atol=0)
    assert u.grad is None                                       # the reference detaches the uncertainty (Tracker.py:179)
    if use_color:
        assert np.array_equal(c.grad.cpu().numpy(), o_gc)


def test_median_gate_with_ties_and_zeros():
    """many equal residuals (ties around the median) and exact zeros"""
    from evennicer_slam_b200.losses import tracker_loss
    n = 301
    gt_depth = np.full(n, 2.0, np.float32); gt_depth[::7] = 0.0
    depth = np.full(n, 2.0); depth[np.arange(n) % 5 < 3] += 0.25; depth[::11] -= 0.5; depth[10] += 50.0   # median inside a tie group
    unc = np.full(n, 0.01)
    color = np.zeros((n, 3), np.float32); gt_color = np.zeros((n, 3))
    d = _t(depth, True)
    loss = tracker_loss(_t(gt_depth), _t(gt_color), d, _t(unc), _t(color), 0.2, True, True)
    loss.backward()
    o_loss, o_gd, _ = ro.tracker_loss(gt_depth, gt_color, depth, unc, color, 0.2, True, True)
    assert o_loss > 0 and abs(loss.item() - o_loss) <= 1e-12 * abs(o_loss)
    assert np.allclose(d.grad.cpu().numpy(), o_gd, rtol=1e-14, atol=0)
    assert d.grad[10].item() == 0.0 and int((o_gd != 0).sum()) > n // 2
