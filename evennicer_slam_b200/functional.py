"""torch.autograd bridges over the C ABI (``include/ens_render.h``).

Only plumbing lives here: tensor allocation, stream selection, saved-for-backward state.
All arithmetic is in ``csrc/*.cu``.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Dict, List, Optional

import torch

from . import _ext, _lib
from ._lib import LEVELS, STAGES, STAGE_LEVELS, EnsGrads, EnsRenderCfg
from .scene import SceneCache, build_scene_struct, decoder_grad_views, is_native_strided


TC_MIN_POINTS = 1         # every forward-only call takes the placement + tcgen05 decode + compositing path (one variant for all
                          # batch sizes keeps the forward bit-independent of how a frame is cut into batches)
TC_POSE_FORWARD = os.environ.get("ENS_FWD_TC_MASKS", "1") != "0"   # tcgen05 forward when only masks are kept
TC_MAP = os.environ.get("ENS_MAP_TC", "1") != "0"   # tcgen05 forward + backward when decoder gradients are wanted (saved kind 3)
SAVE_FORWARD = True    # keep relu masks / activations from the forward kernel for the backward (False: it recomputes)
_DEBUG: Dict[str, object] = {}     # test hook: set _DEBUG["keep_workspace"]=True to inspect the backward scratch


class KernelTimer:
    """CUDA-event timing of individual C-ABI launches on the launching stream (bench.py uses it to get the
    dominant kernel's duration inside the timed region).  Also counts launches."""

    def __init__(self):
        self.enabled = False
        self.events = {}
        self.launches = 0

    def reset(self):
        self.events = {}
        self.launches = 0

    def launch(self, name, device, fn):
        self.launches += 1
        if not self.enabled:
            return fn()
        s = torch.cuda.current_stream(device)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(s)
        rc = fn()
        b.record(s)
        self.events.setdefault(name, []).append((a, b))
        return rc

    def summary(self):
        """name -> (count, mean ms); call after a synchronize."""
        return {k: (len(v), sum(a.elapsed_time(b) for a, b in v) / len(v)) for k, v in self.events.items()}


TIMER = KernelTimer()


def _f32c(t: torch.Tensor) -> torch.Tensor:
    t = t.detach()
    if t.dtype != torch.float32:
        t = t.float()
    return t if t.is_contiguous() else t.contiguous()


class RenderSetup:
    """Everything ens_render_fwd/bwd need besides the rays (plain python object, no grad)."""

    def __init__(self, stage: str, n_samples: int, n_surface: int, t_vals: torch.Tensor,
                 t_surf: torch.Tensor, bound: torch.Tensor, coarse_bound: torch.Tensor,
                 cache: SceneCache, n_importance: int = 0, lindisp: bool = False, perturb: float = 0.0,
                 occupancy: bool = True):
        self.stage = stage
        self.n_samples = n_samples
        self.n_surface = n_surface
        self.t_vals = t_vals
        self.t_surf = t_surf
        self.bound = bound
        self.coarse_bound = coarse_bound
        self.cache = cache
        self.n_importance = n_importance
        self.lindisp = lindisp
        self.perturb = perturb
        self.occupancy = occupancy

    def cfg_struct(self) -> EnsRenderCfg:
        cfg = EnsRenderCfg()
        cfg.n_samples = self.n_samples
        cfg.n_surface = self.n_surface
        cfg.n_importance = self.n_importance
        cfg.lindisp = int(bool(self.lindisp))
        cfg.perturb = float(self.perturb)
        cfg.occupancy = int(bool(self.occupancy))
        cfg.t_vals = self.t_vals.data_ptr()
        cfg.t_vals_surface = self.t_surf.data_ptr() if self.t_surf is not None else None
        return cfg


def depth_batch_max(gt_depth: torch.Tensor) -> torch.Tensor:
    """device double[2] = (max(gt_depth*1.2), max(gt_depth))  -- Renderer.py:110,145."""
    out = torch.empty(2, dtype=torch.float64, device=gt_depth.device)
    L = _lib.lib()
    TIMER.launches += 1
    _lib.check(L.ens_depth_max(_lib.ptr(gt_depth), gt_depth.numel(), _lib.ptr(out),
                               _lib.cur_stream(gt_depth.device)), "ens_depth_max")
    return out


class _RenderBatchRay(torch.autograd.Function):
    """Renderer.render_batch_ray as ONE fused forward kernel and ONE fused backward kernel.

    Tensor inputs (positional, after the python-object args): rays_o, rays_d, then the grids of the
    stage's levels, then every decoder parameter of those levels (state_dict order).
    """

    @staticmethod
    def forward(ctx, setup: RenderSetup, gt_depth, depth_max, n_grids, levels, want_aux, rays_o, rays_d, *tensors):
        L = _lib.lib()
        dev = rays_o.device
        grids = tensors[:n_grids]
        params = tensors[n_grids:]
        ro, rd = _f32c(rays_o), _f32c(rays_d)
        R = ro.shape[0]
        native, packed, per_level_params = {}, {}, {}
        off = 0
        for lv, g in zip(levels, grids):
            native[lv] = setup.cache.native_grid(lv, g)
            n = L.ens_decoder_num_tensors(LEVELS.index(lv))
            per_level_params[lv] = params[off:off + n]
            live = getattr(setup, "_cache_params", None)
            packed[lv] = setup.cache.packed_decoder(lv, live[off:off + n] if live is not None else per_level_params[lv])
            off += n
        sc = build_scene_struct(setup.bound, setup.coarse_bound, native, packed)
        cfg = setup.cfg_struct()
        has_depth = gt_depth is not None and setup.stage != "coarse"
        S = setup.n_samples + (setup.n_surface if has_depth else 0)
        depth = torch.empty(R, dtype=torch.float64, device=dev)
        var = torch.empty(R, dtype=torch.float64, device=dev)
        color = torch.empty((R, 3), dtype=torch.float32, device=dev)
        raw = torch.empty((R, S, 4), dtype=torch.float32, device=dev)
        z = torch.empty((R, S), dtype=torch.float64, device=dev) if want_aux else None
        w = torch.empty((R, S), dtype=torch.float32, device=dev) if want_aux else None
        gd = _f32c(gt_depth).reshape(-1) if has_depth else None
        # saved-for-backward state (what autograd keeps as saved tensors in the reference): relu masks, and the
        # hidden activations when a decoder parameter wants gradient -- the backward then skips the recompute
        needs = ctx.needs_input_grad
        want_bwd = any(needs[6:])
        want_dec = want_bwd and any(needs[8 + n_grids:])
        saved = None
        # 0: relu masks, 1: masks + activations (mma.sync kernels), 2: masks from the tcgen05 forward,
        # 3: relu outputs from the tcgen05 forward (tcgen05 backward with decoder gradients)
        saved_kind = 1 if want_dec else 0
        tc_map = want_dec and TC_MAP and SAVE_FORWARD and setup.stage != "coarse"
        if tc_map:
            nbytes = int(L.ens_fwd_saved_bytes_kind(R, S, STAGES[setup.stage], 3))
            if nbytes > 0:
                saved = torch.empty(nbytes // 4, dtype=torch.int32, device=dev)
                saved_kind = 3
            else:
                tc_map = False
        if want_bwd and SAVE_FORWARD and not tc_map:
            nbytes = int(L.ens_fwd_saved_bytes(R, S, STAGES[setup.stage], int(want_dec)))
            if nbytes > 0:
                saved = torch.empty(nbytes // 4, dtype=torch.int32, device=dev)
        scratch = None
        tc_forward = (not want_bwd) or tc_map or (saved is not None and not want_dec and TC_POSE_FORWARD)
        if tc_forward and R * S >= TC_MIN_POINTS:
            # no activations to keep (forward only: render_img, visualisation; or a backward without decoder gradients:
            # tracking, the tracker's event render): scratch for the placement -> tcgen05 decode -> compositing path
            nbytes = int(L.ens_fwd_scratch_bytes(R, S, STAGES[setup.stage]))
            if nbytes > 0:
                scratch = torch.empty(nbytes // 8, dtype=torch.float64, device=dev)
                if saved is not None and not tc_map:
                    saved_kind = 2
        stream = _lib.cur_stream(dev)
        _lib.check(TIMER.launch("render_fwd", dev, lambda: L.ens_render_fwd(
            C.byref(sc), C.byref(cfg), STAGES[setup.stage], _lib.ptr(ro), _lib.ptr(rd), _lib.ptr(gd),
            _lib.ptr(depth_max) if has_depth else None, R, _lib.ptr(depth), _lib.ptr(var), _lib.ptr(color),
            _lib.ptr(z), _lib.ptr(w), _lib.ptr(raw), _lib.ptr(saved), saved.numel() * 4 if saved is not None else 0,
            saved_kind, _lib.ptr(scratch), scratch.numel() * 8 if scratch is not None else 0,
            stream)), "ens_render_fwd")
        if _DEBUG.get("keep_saved"):
            _DEBUG["saved"] = (saved, saved_kind)
        ctx.saved_fwd = saved
        ctx.saved_has_h = bool(want_dec)
        ctx.saved_kind = saved_kind
        ctx.grid_native_strided = [is_native_strided(g) for g in grids]
        ctx.setup = setup
        ctx.levels = levels
        ctx.n_grids = n_grids
        ctx.S = S
        ctx.has_depth = has_depth
        ctx.native = native
        ctx.packed = packed
        ctx.param_shapes = {lv: per_level_params[lv] for lv in levels}
        ctx.save_for_backward(ro, rd, gd if gd is not None else torch.empty(0, device=dev),
                              depth_max if has_depth else torch.empty(0, device=dev), raw)
        if want_aux:
            ctx.mark_non_differentiable(raw, z, w)
            return depth, var, color, raw, z, w
        ctx.mark_non_differentiable(raw)
        return depth, var, color, raw

    @staticmethod
    def backward(ctx, g_depth, g_var, g_color, *unused):
        L = _lib.lib()
        setup: RenderSetup = ctx.setup
        ro, rd, gd, depth_max, raw = ctx.saved_tensors
        dev = ro.device
        R, S = ro.shape[0], ctx.S
        levels = ctx.levels
        n_grids = ctx.n_grids
        needs = ctx.needs_input_grad            # setup, gt_depth, depth_max, n_grids, levels, want_aux, ro, rd, *tensors
        need_ro, need_rd = needs[6], needs[7]
        need_grid = needs[8:8 + n_grids]
        need_param = needs[8 + n_grids:]
        grads = EnsGrads()
        g_native: Dict[str, torch.Tensor] = {}
        g_flat: Dict[str, torch.Tensor] = {}
        off = 0
        any_param = False
        for gi, lv in enumerate(levels):
            n = len(ctx.param_shapes[lv])
            if any(need_param[off:off + n]):
                any_param = True
            off += n
        # ONE zero-filled arena for every accumulate-into sink (native grid gradients, flat decoder gradients):
        # one memset launch instead of one per buffer
        sizes = []
        for gi, lv in enumerate(levels):
            if need_grid[gi]:
                sizes.append(("g", lv, ctx.native[lv].numel()))
        if any_param:   # the kernel computes decoder grads for all levels of the stage or none
            for lv in levels:
                sizes.append(("d", lv, int(L.ens_decoder_grad_floats(LEVELS.index(lv)))))
        arena = torch.zeros(sum((n + 3) & ~3 for _, _, n in sizes), dtype=torch.float32, device=dev) if sizes else None
        pos = 0
        for kind, lv, n in sizes:
            li = LEVELS.index(lv)
            view = arena[pos:pos + n]
            pos += (n + 3) & ~3                       # keep every sink 16-byte aligned (red.global.add.v4)
            if kind == "g":
                g_native[lv] = view.view(ctx.native[lv].shape)
                grads.grid[li] = view.data_ptr()
            else:
                g_flat[lv] = view
                grads.decoder[li] = view.data_ptr()
        g_ro = torch.empty_like(ro) if (need_ro or need_rd) else None
        g_rd = torch.empty_like(rd) if (need_ro or need_rd) else None
        grads.rays_o = g_ro.data_ptr() if g_ro is not None else None
        grads.rays_d = g_rd.data_ptr() if g_rd is not None else None
        saved = ctx.saved_fwd
        if saved is not None and any_param and not ctx.saved_has_h:
            saved = None                      # activations were not kept: let the kernel recompute
        # scratch: the recompute variant spills activations there; with a saved forward it carries the g_h tiles from
        # the data-gradient kernel to the weight-gradient kernel (split mapping backward)
        ws_bytes = int(L.ens_bwd_workspace_bytes(R, S, 1 if any_param else 0))
        ws = torch.empty(max(ws_bytes, 4) // 4, dtype=torch.float32, device=dev) if ws_bytes else None
        sc = build_scene_struct(setup.bound, setup.coarse_bound, ctx.native, ctx.packed)
        cfg = setup.cfg_struct()
        gdp = g_depth.detach().to(torch.float64).contiguous() if g_depth is not None else None
        gvp = g_var.detach().to(torch.float64).contiguous() if g_var is not None else None
        gcp = _f32c(g_color) if g_color is not None else None
        stream = _lib.cur_stream(dev)
        _lib.check(TIMER.launch("render_bwd", dev, lambda: L.ens_render_bwd(
            C.byref(sc), C.byref(cfg), STAGES[setup.stage], _lib.ptr(ro), _lib.ptr(rd),
            _lib.ptr(gd) if ctx.has_depth else None, _lib.ptr(depth_max) if ctx.has_depth else None, R,
            _lib.ptr(raw), _lib.ptr(gdp), _lib.ptr(gvp), _lib.ptr(gcp), C.byref(grads), _lib.ptr(ws), ws_bytes,
            _lib.ptr(saved), saved.numel() * 4 if saved is not None else 0, ctx.saved_kind if saved is not None else 0,
            stream)), "ens_render_bwd")
        if _DEBUG.get("keep_workspace"):
            _DEBUG["workspace"] = ws
        out: List[Optional[torch.Tensor]] = [None] * 6
        out.append(g_ro if need_ro else None)
        out.append(g_rd if need_rd else None)
        for gi, lv in enumerate(levels):
            if need_grid[gi] and ctx.grid_native_strided[gi]:
                # the grid lives in the kernels' layout: hand autograd a [1,32,Z,Y,X] view of the native gradient
                out.append(g_native[lv].permute(3, 0, 1, 2).unsqueeze(0))
            elif need_grid[gi]:
                nat = g_native[lv]
                Z, Y, X = nat.shape[:3]
                g_ref = torch.empty((1, 32, Z, Y, X), dtype=torch.float32, device=dev)
                TIMER.launches += 1
                _lib.check(L.ens_grid_from_native(_lib.ptr(nat), _lib.ptr(g_ref), Z * Y * X, 0,
                                                  _lib.cur_stream(dev)), "ens_grid_from_native")
                out.append(g_ref)
            else:
                out.append(None)
        off = 0
        for lv in levels:
            ps = ctx.param_shapes[lv]
            if any_param:
                views = decoder_grad_views(g_flat[lv], ps)
                for k, v in enumerate(views):
                    out.append(v if need_param[off + k] else None)
            else:
                out.extend([None] * len(ps))
            off += len(ps)
        return tuple(out)


_EMPTY: Dict[str, torch.Tensor] = {}


def _empty_on(dev) -> torch.Tensor:
    t = _EMPTY.get(str(dev))
    if t is None:
        t = torch.empty(0, dtype=torch.float32, device=dev)
        _EMPTY[str(dev)] = t
    return t


def _bounds12(bound: torch.Tensor, coarse_bound: torch.Tensor) -> List[float]:
    from .scene import _bound_floats
    return list(_bound_floats(bound)) + list(_bound_floats(coarse_bound))


def render_batch_ray(setup: RenderSetup, c: Dict[str, torch.Tensor], decoders, rays_d, rays_o, gt_depth=None,
                     want_aux: bool = False, depth_max: Optional[torch.Tensor] = None, freeze_params: bool = False):
    """Fused Renderer.render_batch_ray.  Returns (depth f64, var f64, color f32[, raw, z_vals, weights])."""
    from .scene import decoder_tensors
    levels = STAGE_LEVELS[setup.stage]
    grids = [c["grid_" + lv] for lv in levels]
    params: List[torch.Tensor] = []
    for lv in levels:
        params.extend(decoder_tensors(decoders, lv))
    cache_params = params              # SceneCache keys the packed blobs on the identity / version of the live Parameters
    if freeze_params:                  # Renderer.decoder_grads = False: the weights are constants of this render
        params = [p.detach() for p in params]
    has_depth = gt_depth is not None and setup.stage != "coarse"
    if has_depth:
        gt_depth = gt_depth.reshape(-1)
        if gt_depth.dtype != torch.float32:
            gt_depth = gt_depth.float()
        if depth_max is None:       # sharded callers pass the maxima of the WHOLE batch (sharding.py)
            depth_max = depth_batch_max(gt_depth.contiguous())
    else:
        depth_max = None
    ext = _ext.module() if (_ext.ENABLED and not TIMER.enabled) else None
    if ext is not None and setup.n_importance == 0 and not setup.lindisp and setup.perturb == 0 and setup.occupancy \
            and not _DEBUG:
        # the same sequence as _RenderBatchRay below with the autograd plumbing in C++ (csrc/ens_torch.cpp): the layouts the
        # kernels read are prepared here (SceneCache: version-checked), everything per call happens there
        L = _lib.lib()
        native, packed, n_params = [], [], []
        off = 0
        for lv, g in zip(levels, grids):
            n = L.ens_decoder_num_tensors(LEVELS.index(lv))
            native.append(setup.cache.native_grid(lv, g))
            packed.append(setup.cache.packed_decoder(lv, cache_params[off:off + n]))
            n_params.append(n)
            off += n
        dev = rays_o.device
        empty = _empty_on(dev)
        aux = [gt_depth if has_depth else empty, depth_max if has_depth else empty, setup.t_vals,
               setup.t_surf if setup.t_surf is not None else empty]
        out = ext.render(rays_o, rays_d, grids, params, aux, native, packed, STAGES[setup.stage], setup.n_samples,
                         setup.n_surface, _bounds12(setup.bound, setup.coarse_bound), n_params, want_aux, TC_MAP,
                         TC_POSE_FORWARD)
        if want_aux:
            return tuple(out)
        return out[0], out[1], out[2]
    setup._cache_params = cache_params
    try:
        out = _RenderBatchRay.apply(setup, gt_depth if has_depth else None, depth_max, len(grids), levels, want_aux,
                                    rays_o, rays_d, *grids, *params)
    finally:
        setup._cache_params = None
    if want_aux:
        return out
    return out[0], out[1], out[2]


def eval_points(setup: RenderSetup, c, decoders, p: torch.Tensor, apply_mask: bool = True) -> torch.Tensor:
    """Fused Renderer.eval_points / NICE.forward (no autograd; the differentiable path is render_batch_ray)."""
    from .scene import decoder_tensors
    L = _lib.lib()
    levels = STAGE_LEVELS[setup.stage]
    native, packed = {}, {}
    for lv in levels:
        native[lv] = setup.cache.native_grid(lv, c["grid_" + lv])
        packed[lv] = setup.cache.packed_decoder(lv, decoder_tensors(decoders, lv))
    sc = build_scene_struct(setup.bound, setup.coarse_bound, native, packed)
    p = p.detach().reshape(-1, 3)
    if p.dtype not in (torch.float32, torch.float64):
        p = p.float()
    p = p.contiguous()
    n = p.shape[0]
    out = torch.empty((n, 4), dtype=torch.float32, device=p.device)
    TIMER.launches += 1
    _lib.check(L.ens_eval_points(C.byref(sc), STAGES[setup.stage], _lib.ptr(p), int(p.dtype == torch.float64), n,
                                 int(apply_mask), _lib.ptr(out), _lib.cur_stream(p.device)), "ens_eval_points")
    return out


# ---------------------------------------------------------------------------------------------
# ray generation
# ---------------------------------------------------------------------------------------------
def _c2w_dev(c2w, device):
    import numpy as np
    if isinstance(c2w, np.ndarray):
        c2w = torch.from_numpy(c2w)
    return c2w.to(device)


class _SampleRays(torch.autograd.Function):
    """get_samples after the torch.randint draw (common.py:92-187); differentiable wrt c2w."""

    @staticmethod
    def forward(ctx, c2w, indices, crop, cam, depth, color):
        L = _lib.lib()
        dev = indices.device
        H0, H1, W0, W1 = crop
        H, W, fx, fy, cx, cy = cam
        n = indices.numel()
        m = _f32c(c2w)
        ro = torch.empty((n, 3), dtype=torch.float32, device=dev)
        rd = torch.empty((n, 3), dtype=torch.float32, device=dev)
        pi = torch.empty(n, dtype=torch.float32, device=dev)
        pj = torch.empty(n, dtype=torch.float32, device=dev)
        sd = torch.empty(n, dtype=torch.float32, device=dev)
        is64 = color.dtype == torch.float64
        sc = torch.empty((n, 3), dtype=color.dtype, device=dev)
        dep = depth if (depth.dtype == torch.float32 and depth.is_contiguous()) else depth.float().contiguous()
        col = color if color.is_contiguous() else color.contiguous()
        if col.dtype not in (torch.float32, torch.float64):
            raise ValueError("color must be float32 or float64")
        TIMER.launches += 1
        _lib.check(L.ens_sample_rays(_lib.ptr(indices), n, H0, H1, W0, W1, H, W, fx, fy, cx, cy, _lib.ptr(m),
                                     m.stride(0), _lib.ptr(dep), _lib.ptr(col), int(is64), _lib.ptr(pi),
                                     _lib.ptr(pj), _lib.ptr(ro), _lib.ptr(rd), _lib.ptr(sd), _lib.ptr(sc),
                                     _lib.cur_stream(dev)), "ens_sample_rays")
        ctx.cam = cam
        ctx.c2w_shape = c2w.shape
        ctx.save_for_backward(pi, pj)
        ctx.mark_non_differentiable(sd, sc)
        return ro, rd, sd, sc

    @staticmethod
    def backward(ctx, g_ro, g_rd, *unused):
        L = _lib.lib()
        pi, pj = ctx.saved_tensors
        H, W, fx, fy, cx, cy = ctx.cam
        dev = pi.device
        g = torch.zeros((3, 4), dtype=torch.float32, device=dev)
        gro = _f32c(g_ro) if g_ro is not None else None
        grd = _f32c(g_rd) if g_rd is not None else None
        TIMER.launches += 1
        _lib.check(L.ens_rays_bwd(_lib.ptr(pi), _lib.ptr(pj), pi.numel(), 0, fx, fy, cx, cy, _lib.ptr(gro),
                                  _lib.ptr(grd), _lib.ptr(g), _lib.cur_stream(dev)), "ens_rays_bwd")
        if ctx.c2w_shape[0] == 4:
            g = torch.cat([g, torch.zeros((1, 4), dtype=torch.float32, device=dev)], 0)
        return g, None, None, None, None, None


class _LatticeRays(torch.autograd.Function):
    """get_rays / get_rays_rescale (common.py:300-340); differentiable wrt c2w."""

    @staticmethod
    def forward(ctx, c2w, lin_w, lin_h, cam):
        L = _lib.lib()
        dev = lin_w.device
        H, W, fx, fy, cx, cy = cam
        nW, nH = lin_w.numel(), lin_h.numel()
        m = _f32c(c2w)
        ro = torch.empty((nH, nW, 3), dtype=torch.float32, device=dev)
        rd = torch.empty((nH, nW, 3), dtype=torch.float32, device=dev)
        TIMER.launches += 1
        _lib.check(L.ens_lattice_rays(_lib.ptr(lin_w), nW, _lib.ptr(lin_h), nH, fx, fy, cx, cy, _lib.ptr(m),
                                      m.stride(0), _lib.ptr(ro), _lib.ptr(rd), _lib.cur_stream(dev)),
                   "ens_lattice_rays")
        ctx.cam = cam
        ctx.c2w_shape = c2w.shape
        ctx.save_for_backward(lin_w, lin_h)
        return ro, rd

    @staticmethod
    def backward(ctx, g_ro, g_rd):
        L = _lib.lib()
        lin_w, lin_h = ctx.saved_tensors
        H, W, fx, fy, cx, cy = ctx.cam
        dev = lin_w.device
        n = lin_w.numel() * lin_h.numel()
        g = torch.zeros((3, 4), dtype=torch.float32, device=dev)
        gro = _f32c(g_ro).reshape(-1, 3) if g_ro is not None else None
        grd = _f32c(g_rd).reshape(-1, 3) if g_rd is not None else None
        TIMER.launches += 1
        _lib.check(L.ens_rays_bwd(_lib.ptr(lin_w), _lib.ptr(lin_h), n, lin_w.numel(), fx, fy, cx, cy,
                                  _lib.ptr(gro), _lib.ptr(grd), _lib.ptr(g), _lib.cur_stream(dev)), "ens_rays_bwd")
        if ctx.c2w_shape[0] == 4:
            g = torch.cat([g, torch.zeros((1, 4), dtype=torch.float32, device=dev)], 0)
        return g, None, None, None


class _PairRays(torch.autograd.Function):
    """get_rays_from_uv (common.py:74-89) for explicit pixel coordinate pairs."""

    @staticmethod
    def forward(ctx, c2w, pix_i, pix_j, cam):
        L = _lib.lib()
        dev = pix_i.device
        H, W, fx, fy, cx, cy = cam
        n = pix_i.numel()
        m = _f32c(c2w)
        ro = torch.empty((n, 3), dtype=torch.float32, device=dev)
        rd = torch.empty((n, 3), dtype=torch.float32, device=dev)
        TIMER.launches += 1
        _lib.check(L.ens_lattice_rays(_lib.ptr(pix_i), n, _lib.ptr(pix_j), 0, fx, fy, cx, cy, _lib.ptr(m),
                                      m.stride(0), _lib.ptr(ro), _lib.ptr(rd), _lib.cur_stream(dev)),
                   "ens_lattice_rays")
        ctx.cam = cam
        ctx.c2w_shape = c2w.shape
        ctx.save_for_backward(pix_i, pix_j)
        return ro, rd

    @staticmethod
    def backward(ctx, g_ro, g_rd):
        L = _lib.lib()
        pi, pj = ctx.saved_tensors
        H, W, fx, fy, cx, cy = ctx.cam
        dev = pi.device
        g = torch.zeros((3, 4), dtype=torch.float32, device=dev)
        gro = _f32c(g_ro) if g_ro is not None else None
        grd = _f32c(g_rd) if g_rd is not None else None
        TIMER.launches += 1
        _lib.check(L.ens_rays_bwd(_lib.ptr(pi), _lib.ptr(pj), pi.numel(), 0, fx, fy, cx, cy, _lib.ptr(gro),
                                  _lib.ptr(grd), _lib.ptr(g), _lib.cur_stream(dev)), "ens_rays_bwd")
        if ctx.c2w_shape[0] == 4:
            g = torch.cat([g, torch.zeros((1, 4), dtype=torch.float32, device=dev)], 0)
        return g, None, None, None


class _PoseToC2W(torch.autograd.Function):
    """get_camera_from_tensor (common.py:215-228) for (n,7) camera tensors: one kernel each way."""

    @staticmethod
    def forward(ctx, cam):
        L = _lib.lib()
        x = _f32c(cam)
        n = x.shape[0]
        out = torch.empty((n, 3, 4), dtype=torch.float32, device=x.device)
        TIMER.launches += 1
        _lib.check(L.ens_pose_fwd(_lib.ptr(x), n, _lib.ptr(out), _lib.cur_stream(x.device)), "ens_pose_fwd")
        ctx.save_for_backward(x)
        return out

    @staticmethod
    def backward(ctx, g):
        L = _lib.lib()
        (x,) = ctx.saved_tensors
        n = x.shape[0]
        gc = torch.empty((n, 7), dtype=torch.float32, device=x.device)
        TIMER.launches += 1
        _lib.check(L.ens_pose_bwd(_lib.ptr(x), n, _lib.ptr(_f32c(g)), _lib.ptr(gc), _lib.cur_stream(x.device)),
                   "ens_pose_bwd")
        return gc
