"""Pins oracle/event_loss_oracle.py against the reference's event-loss lines (Tracker.py:204-224) executed with torch +
torchvision autograd on CPU.  No GPU, no product code."""
import os
import sys

import numpy as np
import pytest
import torch
from torchvision import transforms

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle")]
import event_loss_oracle as eo  # noqa: E402


def reference_lines(gt_event, full_event, kernel_sizes, kernel_weights, balancer, blur=True):
    loss_event = ((gt_event - full_event) ** 2).sum()                                       # Tracker.py:206
    parts = [loss_event.item()]
    if blur:
        for kernel_size, kernel_weight in zip(kernel_sizes, kernel_weights):                # :212
            gt_tmp = transforms.functional.gaussian_blur(gt_event.permute(2, 0, 1), kernel_size=kernel_size).permute(1, 2, 0)
            pr_tmp = transforms.functional.gaussian_blur(full_event.permute(2, 0, 1), kernel_size=kernel_size).permute(1, 2, 0)
            loss_tmp = ((gt_tmp - pr_tmp) ** 2).sum()                                       # :216
            loss_event = loss_event + kernel_weight * loss_tmp                              # :217
            parts.append(loss_tmp.item())
    return loss_event * balancer, parts                                                     # :228


@pytest.mark.parametrize("shape,ks,ws", [((102, 180, 2), [9], [1.0]), ((39, 51, 2), [9, 3], [1.0, 0.5]),
                                         ((17, 16, 3), [3], [1.0]), ((12, 33, 1), [5, 9, 3], [0.3, 1.0, 2.0])])
def test_oracle_matches_torchvision_autograd(shape, ks, ws):
    rng = np.random.RandomState(20)
    gt = rng.poisson(0.3, size=shape).astype(np.float32)
    pred = (rng.rand(*shape) * 1.5).astype(np.float32)
    p = torch.from_numpy(pred.copy()).requires_grad_(True)
    loss, parts = reference_lines(torch.from_numpy(gt), p, ks, ws, 0.025)
    loss.backward()
    o_loss, o_parts, o_grad = eo.event_loss(gt, pred, ks, ws, 0.025)
    assert abs(o_loss - loss.item()) < 1e-5 * abs(loss.item())
    assert np.allclose(o_parts, parts, rtol=1e-5)
    g = p.grad.numpy()
    assert np.abs(o_grad - g).max() < 1e-5 * np.abs(g).max()


def test_kernel1d_matches_torchvision():
    from torchvision.transforms import _functional_tensor as FT
    for ks in (3, 5, 9, 15):
        want = FT._get_gaussian_kernel1d(ks, ks * 0.15 + 0.35, torch.float32, torch.device("cpu")).numpy()
        assert np.abs(eo.gaussian_kernel1d(ks) - want).max() < 1e-7
