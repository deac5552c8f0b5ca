"""CPU oracle for the frustum-masked grid optimisation of the mapper (numpy restatement).

TEST INFRASTRUCTURE ONLY (same rules as render_oracle.py): imported by ``tests/`` and ``bench.py``'s checker legs,
never by the product package.

What it restates (paths under /root/reference/src):
  Mapper.py:343-361   the optimisable copy  ``val_grad = val[mask].clone()``, mask = voxel mask repeated over 32 channels
  Mapper.py:451-458   ``val[mask] = val_grad`` before every render
  Mapper.py:396-423   ``torch.optim.Adam`` over the grid groups with per-stage learning rates (463-486), default
                      betas (0.9, 0.999), eps 1e-8, no weight decay, no amsgrad
  Mapper.py:633-641   the write-back after ``optimizer.step()``
The Adam arithmetic follows the reference's pinned torch 1.11 (``torch/optim/_functional.py::adam``; environment.yaml:82):
    exp_avg.mul_(beta1).add_(grad, alpha=1 - beta1)
    exp_avg_sq.mul_(beta2).addcmul_(grad, grad, value=1 - beta2)
    denom = (exp_avg_sq.sqrt() / sqrt(1 - beta2**step)).add_(eps)
    param.addcdiv_(exp_avg, denom, value=-(lr / (1 - beta1**step)))
in float32 tensors with the scalars formed in float64 (python floats) and rounded to float32 where they meet a tensor.

Parity pinning: torch.optim.Adam itself is the third-party arithmetic; the reference holds no vectors
for it.  ``tests/test_adam_cpu.py`` pins this restatement against ``torch.optim.Adam`` run through the reference's
gather / index_put / write-back sequence in this container (torch 2.11: ``lerp_`` form of the first-moment update,
equal to 1.11's to one ulp), tolerance 2e-6 relative to the largest update.
"""
from __future__ import annotations

import numpy as np

F32 = np.float32


class MaskedAdamState:
    """Dense moments for one grid in the reference layout [1,32,Z,Y,X] (zero outside the mask, never read there)."""

    def __init__(self, shape):
        self.m = np.zeros(shape, F32)
        self.v = np.zeros(shape, F32)


def adam_step_masked(grid: np.ndarray, grad: np.ndarray, voxel_mask, state: MaskedAdamState, lr: float, step: int,
                     beta1: float = 0.9, beta2: float = 0.999, eps: float = 1e-8) -> None:
    """One optimizer.step() + write-back for one level, in place on `grid` ([1,32,Z,Y,X] float32).

    voxel_mask: bool [Z,Y,X] (``torch.from_numpy(mask).permute(2,1,0)``, Mapper.py:345) or None for every voxel."""
    assert grid.dtype == F32 and grad.dtype == F32 and grid.shape == grad.shape
    if voxel_mask is None:
        sel = np.ones(grid.shape, bool)
    else:
        sel = np.broadcast_to(np.asarray(voxel_mask, bool)[None, None], grid.shape)     # .repeat(1, 32, 1, 1, 1)
    g = grad[sel]
    m = state.m[sel]
    v = state.v[sel]
    p = grid[sel]
    b1, b2 = F32(beta1), F32(beta2)
    m = (m * b1 + g * F32(1.0 - beta1)).astype(F32)
    v = (v * b2 + (g * g) * F32(1.0 - beta2)).astype(F32)
    bc1 = 1.0 - beta1 ** step
    bc2 = 1.0 - beta2 ** step
    denom = (np.sqrt(v) / F32(np.sqrt(bc2)) + F32(eps)).astype(F32)
    p = (p - F32(lr / bc1) * (m / denom)).astype(F32)
    state.m[sel] = m
    state.v[sel] = v
    grid[sel] = p
