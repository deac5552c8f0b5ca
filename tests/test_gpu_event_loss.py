"""Parity of the fused blurred-L2 event loss (ens_event_loss through losses.event_loss) with the oracle
(oracle/event_loss_oracle.py, pinned against torchvision autograd in tests/test_event_loss_cpu.py).  ``-m gpu``.
Tolerance 1e-5 relative on the loss and its parts, 1e-5 of the largest gradient entry (float32 81-tap sums)."""
import numpy as np
import pytest
import torch

import event_loss_oracle as eo

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.mark.parametrize("shape,ks,ws", [((102, 180, 2), [9], [1.0]), ((39, 51, 2), [9, 3], [1.0, 0.5]),
                                         ((17, 16, 3), [3], [1.0]), ((12, 33, 1), [5, 9, 3], [0.3, 1.0, 2.0]),
                                         ((8, 8, 2), [15], [1.0]), ((260, 346, 2), [9], [1.0])])
def test_event_loss_matches_oracle(shape, ks, ws):
    from evennicer_slam_b200.losses import event_loss
    rng = np.random.RandomState(20)
    gt = rng.poisson(0.3, size=shape).astype(np.float32)
    pred = (rng.rand(*shape) * 1.5).astype(np.float32)
    o_loss, o_parts, o_grad = eo.event_loss(gt, pred, ks, ws, 0.025)
    p = torch.from_numpy(pred).to(DEV).requires_grad_(True)
    loss, parts = event_loss(torch.from_numpy(gt).to(DEV), p, ks, ws, balancer=0.025)
    (3.0 * loss).backward()                                  # an upstream factor must scale the gradient
    assert loss.dtype == torch.float32 and abs(loss.item() - o_loss) < 1e-5 * abs(o_loss)
    assert np.allclose(parts.cpu().numpy()[1:], o_parts, rtol=1e-5)
    g = p.grad.cpu().numpy() / 3.0
    assert np.abs(g - o_grad).max() < 1e-5 * np.abs(o_grad).max()


def test_unblurred_branch_and_no_grad():
    from evennicer_slam_b200.losses import event_loss
    rng = np.random.RandomState(1)
    gt = rng.poisson(0.3, size=(20, 30, 2)).astype(np.float32)
    pred = rng.rand(20, 30, 2).astype(np.float32)
    loss, parts = event_loss(torch.from_numpy(gt).to(DEV), torch.from_numpy(pred).to(DEV), [9], [1.0], 0.5, blur=False)
    want = 0.5 * float(((gt.astype(np.float64) - pred) ** 2).sum())
    assert abs(loss.item() - want) < 1e-5 * want and parts.numel() == 2


def test_bad_arguments_raise():
    from evennicer_slam_b200.losses import event_loss
    a = torch.zeros(20, 30, 2, device=DEV)
    with pytest.raises(ValueError):
        event_loss(a, a, [4], [1.0])                         # even kernel size (torchvision raises too)
    with pytest.raises(RuntimeError):
        event_loss(a[:3], a[:3], [9], [1.0])                 # reflect pad needs pad < dim
    with pytest.raises(RuntimeError):
        event_loss(a, a, [17], [1.0])                        # larger than the kernel supports: loud, not silent
