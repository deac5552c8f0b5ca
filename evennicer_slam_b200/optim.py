"""Frustum-masked Adam on the feature grids, fused into one launch.

The reference's mapper optimises a gathered COPY of the grid features inside the current view frustum
(``frustum_feature_selection``, Mapper.py:343-361), scatters the copy into the grid before every render (:451-458),
steps ``torch.optim.Adam`` on it (:396-423, :625) and scatters it back (:633-641) -- three boolean-mask passes over up to
5.7 M elements per level and iteration, each with an implicit ``nonzero`` host sync.  ``FrustumGridAdam`` keeps the grids
whole, in the render kernels' layout, and applies the same update in place to the selected voxels with one kernel that
reads the dense gradient the render backward produced (``ens_grid_adam_step``).

    opt = FrustumGridAdam(c, masks)            # once per optimize_map call, like the reference's optimizer (:396)
    for joint_iter in range(n):
        ...render, loss.backward()...          # c[key].grad is dense, as in the reference without feature selection
        opt.step({'grid_middle': lr_m, 'grid_fine': lr_f, 'grid_color': lr_c})
        opt.zero_grad()

Semantics are those of the reference sequence: selected voxels (all 32 channels) take the Adam update with zero-initialised
moments; every other voxel keeps its value.  Learning rate 0 (a stage that freezes a level, configs/nice_slam.yaml:45-76)
still advances the moments, exactly as torch.optim.Adam does with lr = 0.  A level whose ``.grad`` is None at a step (the
middle stage renders without grid_fine / grid_color, Mapper.py:462-473) is SKIPPED and keeps its own step count, as
torch.optim.Adam keeps ``state['step']`` per parameter: its first update later uses step = 1.

Both optimisers write the parameters through raw device pointers; they bump the tensors' autograd version counters
afterwards (``torch.autograd.graph.increment_version``) so that version-keyed caches (``scene.SceneCache``) re-pack.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, Optional

import torch

from . import _lib
from .scene import is_native_strided


def _bump_version(t: torch.Tensor) -> None:
    """The kernels updated ``t`` behind autograd's back: make ``t._version`` say so."""
    torch.autograd.graph.increment_version(t)


def _native_base(t: torch.Tensor) -> torch.Tensor:
    """The contiguous [Z,Y,X,32] tensor a native-strided [1,32,Z,Y,X] view aliases."""
    return t.detach()[0].permute(1, 2, 3, 0)


class FrustumGridAdam:
    def __init__(self, c: Dict[str, torch.Tensor], masks: Optional[Dict[str, Optional[torch.Tensor]]] = None,
                 betas=(0.9, 0.999), eps: float = 1e-8, graph_safe: bool = False):
        """c: the scene dict; every grid to optimise must be in the native layout (``scene.as_native_layout``).
        masks: per key a bool/uint8 voxel mask [Z,Y,X] (``torch.from_numpy(mask).permute(2,1,0)`` of
        ``Mapper.get_mask_from_c2w``) or None = every voxel.  graph_safe: keep step and learning rates in device memory
        so a captured CUDA graph can replay ``step`` (advance them with ``set_dynamic`` outside the graph)."""
        self.keys = [k for k in c if masks is None or k in masks]
        if not 0 < len(self.keys) <= 4:
            raise ValueError("FrustumGridAdam optimises 1..4 grids")
        self.c = c
        self.betas, self.eps = betas, eps
        self.step_count = 0                     # calls of step(); the Adam step of a level is self.steps[level]
        self.steps = {k: 0 for k in self.keys}
        self.state = {}
        self.masks = {}
        for k in self.keys:
            g = c[k]
            if not (g.is_cuda and g.dtype == torch.float32 and is_native_strided(g)):
                raise ValueError(f"{k}: FrustumGridAdam needs a float32 CUDA grid in the native layout "
                                 "(scene.as_native_layout)")
            base = _native_base(g)
            self.state[k] = (torch.zeros_like(base), torch.zeros_like(base))
            m = None if masks is None else masks.get(k)
            if m is not None:
                if tuple(m.shape) != tuple(g.shape[2:]):
                    raise ValueError(f"{k}: voxel mask must be [Z,Y,X] = {tuple(g.shape[2:])}, got {tuple(m.shape)}")
                # the selection as an ascending voxel list, once per optimiser (= per optimize_map call); the
                # reference re-derives it inside every boolean-mask index_put (nonzero + host sync each time)
                m = torch.nonzero(m.to(device=g.device).reshape(-1) != 0).reshape(-1).to(torch.int32).contiguous()
            self.masks[k] = m
        dev = c[self.keys[0]].device
        self.dyn = torch.zeros(1 + len(self.keys), dtype=torch.float64, device=dev) if graph_safe else None

    def set_dynamic(self, step: int, lrs: Dict[str, float]) -> None:
        """graph_safe mode: write the step number and learning rates the next replayed ``step`` will use."""
        vals = [float(step)] + [float(lrs.get(k, 0.0)) for k in self.keys]
        self.dyn.copy_(torch.tensor(vals, dtype=torch.float64), non_blocking=True)

    def step(self, lrs: Dict[str, float], clear_grad: bool = False) -> None:
        L = _lib.lib()
        # levels that took part in this iteration's backward; the others are skipped (torch.optim.Adam skips p.grad is None)
        active = [k for k in self.keys if self.c[k].grad is not None]
        self.step_count += 1
        if not active:
            return
        if self.dyn is not None and len(active) != len(self.keys):
            raise RuntimeError("graph_safe FrustumGridAdam replays a fixed set of levels: capture one optimiser per stage")
        for k in active:
            self.steps[k] += 1
        from .functional import TIMER
        dev = self.c[self.keys[0]].device
        keep = []
        # one launch per distinct Adam step number (one in the steady state, at most one per stage of the schedule)
        for st in sorted({self.steps[k] for k in active}):
            ks = [k for k in active if self.steps[k] == st]
            n = len(ks)
            arr = (_lib.EnsAdamLevel * n)()
            for i, k in enumerate(ks):
                g = self.c[k]
                grad = g.grad
                if not is_native_strided(grad):       # a gradient that did not come from the fused backward
                    grad = grad.contiguous()[0].permute(1, 2, 3, 0).contiguous().permute(3, 0, 1, 2).unsqueeze(0)
                    if clear_grad:
                        raise RuntimeError("clear_grad needs the native-layout gradient of the fused backward")
                gb, base = _native_base(grad), _native_base(g)
                m, v = self.state[k]
                keep.append(gb)
                arr[i].grid, arr[i].grad = base.data_ptr(), gb.data_ptr()
                arr[i].exp_avg, arr[i].exp_avg_sq = m.data_ptr(), v.data_ptr()
                arr[i].n_voxels = base.shape[0] * base.shape[1] * base.shape[2]
                arr[i].voxel_index = None if self.masks[k] is None else self.masks[k].data_ptr()
                arr[i].n_selected = arr[i].n_voxels if self.masks[k] is None else self.masks[k].numel()
                arr[i].lr = float(lrs.get(k, 0.0))
            TIMER.launches += 1
            _lib.check(L.ens_grid_adam_step(arr, n, self.betas[0], self.betas[1], self.eps, st,
                                            _lib.ptr(self.dyn), 1 if clear_grad else 0, _lib.cur_stream(dev)),
                       "ens_grid_adam_step")
        for k in active:
            _bump_version(self.c[k])

    def zero_grad(self) -> None:
        for k in self.keys:
            self.c[k].grad = None


class FusedAdam:
    """``torch.optim.Adam`` for the small parameter groups of the reference's optimizers (decoder weights, camera tensors;
    Mapper.py:396-423, Tracker.py:335-342) as ONE launch per step (``ens_tensors_adam_step``).

    Mirrors the part of the torch API the reference uses: ``FusedAdam([{'params': [...], 'lr': 0}, ...])``,
    ``opt.param_groups[i]['lr'] = ...``, ``opt.step()``, ``opt.zero_grad()``; default betas / eps, no weight decay, no
    amsgrad.  Parameters whose ``.grad`` is None at a step are skipped, as torch does.  ``graph_safe``: step number and
    learning rates live in device memory (``set_dynamic``) so a captured CUDA graph can replay ``step``.
    """

    def __init__(self, params, lr: float = 1e-3, betas=(0.9, 0.999), eps: float = 1e-8, graph_safe: bool = False):
        groups = list(params)
        if groups and not isinstance(groups[0], dict):
            groups = [{"params": groups}]
        self.param_groups = [{"params": list(g["params"]), "lr": float(g.get("lr", lr))} for g in groups]
        if not 1 <= len(self.param_groups) <= 8:
            raise ValueError("FusedAdam takes 1..8 parameter groups")
        self.betas, self.eps = betas, eps
        self.step_count = 0                     # calls of step(); the Adam step of a tensor is self.steps[i] (torch keeps
        flat = [p for g in self.param_groups for p in g["params"]]      # state['step'] per parameter: None grads are skipped)
        self.steps = [0] * len(flat)
        if not flat:
            raise ValueError("FusedAdam got an empty parameter list")
        for p in flat:
            if not (p.is_cuda and p.dtype == torch.float32 and p.is_contiguous()):
                raise ValueError("FusedAdam needs contiguous float32 CUDA parameters (there is no CPU fallback)")
        dev = flat[0].device
        total = sum(p.numel() for p in flat)
        self.exp_avg = torch.zeros(total, dtype=torch.float32, device=dev)
        self.exp_avg_sq = torch.zeros(total, dtype=torch.float32, device=dev)
        self.dyn = torch.zeros(1 + len(self.param_groups), dtype=torch.float64, device=dev) if graph_safe else None
        self._dev = dev

    def set_dynamic(self, step: int, lrs=None) -> None:
        lrs = [g["lr"] for g in self.param_groups] if lrs is None else list(lrs)
        self.dyn.copy_(torch.tensor([float(step)] + [float(x) for x in lrs], dtype=torch.float64), non_blocking=True)

    def step(self) -> None:
        L = _lib.lib()
        ps, gs, sizes, grp, offs, sts, tens = [], [], [], [], [], [], []
        off = 0
        keep = []
        idx = 0
        for gi, g in enumerate(self.param_groups):
            for p in g["params"]:
                n = p.numel()
                if p.grad is not None:
                    gr = p.grad if (p.grad.is_contiguous() and p.grad.dtype == torch.float32) else p.grad.float().contiguous()
                    keep.append(gr)
                    self.steps[idx] += 1
                    ps.append(p.data_ptr()); gs.append(gr.data_ptr()); sizes.append(n); grp.append(gi); offs.append(off)
                    sts.append(self.steps[idx]); tens.append(p)
                off += n
                idx += 1
        self.step_count += 1
        if not ps:
            return
        if self.dyn is not None and len(set(sts)) != 1:
            raise RuntimeError("graph_safe FusedAdam replays one step number for all tensors: the set of tensors with a "
                               "gradient must not change between replays")
        from .functional import TIMER
        # the moments of a skipped tensor must not shift the others, and every tensor uses its OWN step number: one call
        # per run of consecutive state offsets with the same step
        i = 0
        lrs = (C.c_double * len(self.param_groups))(*[g["lr"] for g in self.param_groups])
        while i < len(ps):
            j = i + 1
            while j < len(ps) and offs[j] == offs[j - 1] + sizes[j - 1] and sts[j] == sts[i]:
                j += 1
            n = j - i
            TIMER.launches += 1
            _lib.check(L.ens_tensors_adam_step(
                (C.c_void_p * n)(*ps[i:j]), (C.c_void_p * n)(*gs[i:j]), (C.c_int64 * n)(*sizes[i:j]),
                (C.c_int * n)(*grp[i:j]), n, lrs, len(self.param_groups),
                C.c_void_p(self.exp_avg.data_ptr() + 4 * offs[i]), C.c_void_p(self.exp_avg_sq.data_ptr() + 4 * offs[i]),
                self.betas[0], self.betas[1], self.eps, sts[i], _lib.ptr(self.dyn),
                _lib.cur_stream(self._dev)), "ens_tensors_adam_step")
            i = j
        for p in tens:
            _bump_version(p)

    def zero_grad(self, set_to_none: bool = True) -> None:
        for g in self.param_groups:
            for p in g["params"]:
                if set_to_none:
                    p.grad = None
                elif p.grad is not None:
                    p.grad.zero_()
