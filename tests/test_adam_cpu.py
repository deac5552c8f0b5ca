"""Pins oracle/adam_oracle.py against torch.optim.Adam driven through the reference's frustum-feature-selection sequence
(Mapper.py:343-361 gather, :451-458 index_put, :625 step, :633-641 write-back), on CPU.  No GPU, no product code."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle")]
import adam_oracle as ao  # noqa: E402


def reference_sequence(grids, masks, grads_per_iter, lrs_per_iter):
    """The reference's loop body around the renderer, with the render replaced by given dense gradients."""
    c = {k: torch.from_numpy(v.copy()) for k, v in grids.items()}
    masked, paras = {}, []
    for k, val in c.items():
        mask = torch.from_numpy(masks[k])[None, None].repeat(1, val.shape[1], 1, 1, 1)     # Mapper.py:345-346
        val_grad = val[mask].clone().requires_grad_(True)                                  # :348-350
        masked[k], masked[k + "mask"] = val_grad, mask
        paras.append({"params": [val_grad], "lr": 0})
    opt = torch.optim.Adam(paras)                                                          # :396-401
    for it, (dense, lrs) in enumerate(zip(grads_per_iter, lrs_per_iter)):
        for gi, k in enumerate(grids):
            opt.param_groups[gi]["lr"] = lrs[k]                                            # :469-473
        opt.zero_grad()
        loss = 0.0
        for k in grids:
            val = c[k].detach().clone()
            val[masked[k + "mask"]] = masked[k]                                            # :455-457
            loss = loss + (val * torch.from_numpy(dense[k])).sum()                         # stands for render + loss
        loss.backward()
        opt.step()                                                                         # :625
        for k in grids:
            val = c[k].detach()
            val[masked[k + "mask"]] = masked[k].clone().detach()                           # :637-640
            c[k] = val
    return {k: v.numpy() for k, v in c.items()}


def test_oracle_matches_torch_adam_through_the_reference_sequence():
    rng = np.random.RandomState(20)
    shapes = {"grid_middle": (1, 32, 5, 6, 7), "grid_fine": (1, 32, 9, 4, 11)}
    grids = {k: (rng.randn(*s) * 0.01).astype(np.float32) for k, s in shapes.items()}
    masks = {k: rng.rand(*s[2:]) < 0.6 for k, s in shapes.items()}
    iters = 7
    grads = [{k: (rng.randn(*s) * 10 ** rng.uniform(-4, 0)).astype(np.float32) for k, s in shapes.items()}
             for _ in range(iters)]
    lrs = [{"grid_middle": 0.1, "grid_fine": 0.0}] * 2 + [{"grid_middle": 0.005, "grid_fine": 0.005}] * 5
    want = reference_sequence(grids, masks, grads, lrs)

    got = {k: v.copy() for k, v in grids.items()}
    st = {k: ao.MaskedAdamState(s) for k, s in shapes.items()}
    for it in range(iters):
        for k in shapes:
            ao.adam_step_masked(got[k], grads[it][k], masks[k], st[k], lrs[it][k], it + 1)
    for k in shapes:
        sel = np.broadcast_to(masks[k][None, None], shapes[k])
        assert np.array_equal(got[k][~sel], grids[k][~sel]), "unselected voxels must keep their values"
        assert np.array_equal(want[k][~sel], grids[k][~sel])
        scale = np.abs(want[k] - grids[k]).max()
        assert scale > 1e-3
        assert np.abs(got[k] - want[k]).max() < 2e-6 * scale, k


def test_oracle_without_mask_is_plain_adam():
    rng = np.random.RandomState(3)
    shape = (1, 32, 3, 4, 5)
    p0 = rng.randn(*shape).astype(np.float32)
    p = torch.from_numpy(p0.copy()).requires_grad_(True)
    opt = torch.optim.Adam([p], lr=0.01)
    got, st = p0.copy(), ao.MaskedAdamState(shape)
    for it in range(4):
        g = rng.randn(*shape).astype(np.float32)
        opt.zero_grad(); p.grad = torch.from_numpy(g.copy()); opt.step()
        ao.adam_step_masked(got, g, None, st, 0.01, it + 1)
    want = p.detach().numpy()           # |p| ~ 1: allow the value's own rounding (2 ulp) on top of 2e-6 of the update
    assert np.all(np.abs(got - want) <= 2 * np.spacing(np.abs(want)) + 2e-6 * 0.04)
