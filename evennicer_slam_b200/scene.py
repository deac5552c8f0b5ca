"""Device-side scene state for the fused renderer: native-layout grids and packed decoders.

The reference keeps the scene as ``c`` (dict ``'grid_<level>' -> float32 [1,32,Z,Y,X]``,
EvenNICER_SLAM.py:217-275) and ``decoders`` (an ``nn.Module`` with one ``nn.Parameter`` per
tensor, decoder.py:277-310).  Both are mutated in place between calls (Mapper.py:451-458,
633-641; Adam), and shared across processes.  The kernels want one 128-byte line per voxel
and one contiguous blob per decoder, so every call converts what changed:

  * a grid is re-laid-out only if the tensor OBJECT or its ``_version`` differs from the
    cached one (tracking: same clone for all 10 iterations -> converted once per frame;
    mapping: mutated every iteration -> converted every iteration);
  * a decoder is re-packed only if one of its parameters changed version.

Caches hold weak references, so a recycled ``data_ptr`` can never alias a dead tensor.
"""
from __future__ import annotations

import ctypes as C
import weakref
from typing import Dict, Sequence, Tuple

import torch

from . import _lib
from ._lib import LEVELS, EnsScene


_PARAM_SLOTS = weakref.WeakKeyDictionary()     # decoder module -> [(owning submodule's _parameters dict, name)], state_dict order


def decoder_tensors(decoders, level: str):
    """Parameters of one decoder in state_dict order (the order ens_pack_decoder expects).

    The module tree is walked once per decoder object (``named_parameters`` costs ~70 us per decoder, three times per
    render call); afterwards the live ``_parameters`` dicts are read directly, so a Parameter that was re-assigned, moved
    by ``.to()`` or loaded by ``load_state_dict`` is still picked up."""
    dec = getattr(decoders, level + "_decoder")
    slots = _PARAM_SLOTS.get(dec)
    if slots is None:
        slots = []
        for mod_name, mod in dec.named_modules():
            for name, p in mod._parameters.items():
                if p is not None:
                    slots.append((mod._parameters, name))
        # named_modules / _parameters iterate in registration order, which is named_parameters' (= state_dict) order
        ref = [p for _, p in dec.named_parameters()]
        assert len(ref) == len(slots) and all(d[n] is p for (d, n), p in zip(slots, ref))
        _PARAM_SLOTS[dec] = slots
    return [d[n] for d, n in slots]


def decoder_grad_views(flat: torch.Tensor, params: Sequence[torch.Tensor]):
    """Views of the flat per-decoder gradient buffer, one per parameter (reference shapes)."""
    sizes = [p.numel() for p in params]
    assert sum(sizes) == flat.numel(), (sum(sizes), flat.numel())
    # one split call instead of one slice per tensor (69 tensors per mapping step: host time, not GPU time)
    return [v.view(p.shape) for v, p in zip(flat.split_with_sizes(sizes), params)]


def is_native_strided(grid: torch.Tensor) -> bool:
    """True if ``grid`` ([1,32,Z,Y,X]) is a view of a contiguous [Z,Y,X,32] buffer -- the layout the kernels read."""
    if grid.dim() != 5 or grid.shape[0] != 1:
        return False
    _, Cc, Z, Y, X = grid.shape
    want = (1, Y * X * Cc, X * Cc, Cc)
    # the stride of a size-1 dimension is arbitrary (torch keeps whatever the producing op left there)
    return all(n == 1 or st == w for n, st, w in zip(grid.shape[1:], grid.stride()[1:], want))


def as_native_layout(grid: torch.Tensor) -> torch.Tensor:
    """Re-allocate a feature grid in the kernels' layout and return it as a [1,32,Z,Y,X] VIEW.

    Logically identical to the input (same shape, values, dtype): every caller of the reference keeps working
    (boolean-mask indexing, in-place index_put, clone, share_memory_, torch.save), but the renderer then reads the
    storage directly and returns gradients as views of its native gradient buffer -- no layout conversion per
    iteration.  One line after ``grid_init`` (EvenNICER_SLAM.py:217-275); see INTEGRATION.md.
    """
    if is_native_strided(grid):
        return grid
    g = grid.detach()
    native = g[0].permute(1, 2, 3, 0).contiguous()            # [Z,Y,X,32]
    out = native.permute(3, 0, 1, 2).unsqueeze(0)             # [1,32,Z,Y,X] view
    return out.requires_grad_(grid.requires_grad)


class _GridEntry:
    __slots__ = ("ref", "version", "native")


class _DecEntry:
    __slots__ = ("refs", "versions", "packed")


class SceneCache:
    """Per-Renderer cache of native grids / packed decoders, with version checks."""

    def __init__(self):
        self._grids: Dict[Tuple[str, int], _GridEntry] = {}
        self._decs: Dict[Tuple[str, int], _DecEntry] = {}
        self.stats = {"grid_convert": 0, "grid_hit": 0, "dec_pack": 0, "dec_hit": 0}

    def invalidate(self, grids: bool = True, decoders: bool = True):
        """Forget cached layouts (bench: model a scene that was mutated since the last call)."""
        if grids:
            self._grids.clear()
        if decoders:
            self._decs.clear()

    # -- grids ---------------------------------------------------------------------------
    def native_grid(self, level: str, grid: torch.Tensor) -> torch.Tensor:
        if grid.dim() != 5 or grid.shape[0] != 1 or grid.shape[1] != 32:
            raise ValueError(f"grid_{level} must be [1,32,Z,Y,X], got {tuple(grid.shape)}")
        if grid.dtype != torch.float32 or not grid.is_cuda:
            raise ValueError(f"grid_{level} must be a float32 CUDA tensor")
        if is_native_strided(grid):                 # already in the kernels' layout: no copy, nothing to cache
            self.stats["grid_hit"] += 1
            return grid.detach()[0].permute(1, 2, 3, 0)
        key = (level, grid.device.index)
        # inside a CUDA-graph capture the conversion must be PART of the graph (a cache hit would bake in a buffer that is
        # never refreshed on replay and may be freed): always convert into a fresh, graph-pool buffer and cache nothing
        capturing = torch.cuda.is_current_stream_capturing()
        e = self._grids.get(key)
        if not capturing and e is not None and e.ref() is grid and e.version == grid._version:
            self.stats["grid_hit"] += 1
            return e.native
        src = grid.detach()
        if not src.is_contiguous():
            src = src.contiguous()
        Z, Y, X = grid.shape[2:]
        # always a fresh buffer: an earlier forward may still hold the old one for its backward
        native = torch.empty((Z, Y, X, 32), dtype=torch.float32, device=grid.device)
        L = _lib.lib()
        from .functional import TIMER
        TIMER.launches += 1
        _lib.check(L.ens_grid_to_native(_lib.ptr(src), _lib.ptr(native), Z * Y * X,
                                        _lib.cur_stream(grid.device)), "ens_grid_to_native")
        self.stats["grid_convert"] += 1
        if capturing:
            return native
        e = _GridEntry()
        e.ref, e.version, e.native = weakref.ref(grid), grid._version, native
        self._grids[key] = e
        return native

    # -- decoders ------------------------------------------------------------------------
    def packed_decoder(self, level: str, params: Sequence[torch.Tensor]) -> torch.Tensor:
        dev = params[0].device
        key = (level, dev.index)
        capturing = torch.cuda.is_current_stream_capturing()      # see native_grid: pack inside the graph, cache nothing
        e = self._decs.get(key)
        if not capturing and e is not None and len(e.refs) == len(params) and all(
                r() is p and v == p._version for r, v, p in zip(e.refs, e.versions, params)):
            self.stats["dec_hit"] += 1
            return e.packed
        L = _lib.lib()
        li = LEVELS.index(level)
        n = L.ens_decoder_num_tensors(li)
        if len(params) != n:
            raise ValueError(f"{level} decoder: expected {n} tensors, got {len(params)}")
        keep = []
        arr = (C.c_void_p * n)()
        for i, p in enumerate(params):
            if p.dtype != torch.float32 or not p.is_cuda:
                raise ValueError("decoder parameters must be float32 CUDA tensors")
            t = p.detach()
            if not t.is_contiguous():
                t = t.contiguous()
            keep.append(t)
            arr[i] = t.data_ptr()
        packed = torch.empty(int(L.ens_packed_decoder_floats(li)), dtype=torch.float32, device=dev)
        from .functional import TIMER
        TIMER.launches += 1
        _lib.check(L.ens_pack_decoder(li, arr, n, _lib.ptr(packed), _lib.cur_stream(dev)), "ens_pack_decoder")
        self.stats["dec_pack"] += 1
        if capturing:
            return packed
        e = _DecEntry()
        e.refs = [weakref.ref(p) for p in params]
        e.versions = [p._version for p in params]
        e.packed = packed
        self._decs[key] = e
        return packed


_BOUND_CACHE: Dict[int, tuple] = {}


def _bound_floats(bound: torch.Tensor):
    """The six floats of a (3,2) bound tensor, read once per tensor object and version (two reads per render call before)."""
    e = _BOUND_CACHE.get(id(bound))
    if e is not None and e[0]() is bound and e[1] == bound._version:
        return e[2]
    vals = tuple(float(x) for x in bound.detach().to("cpu", torch.float64).reshape(-1))
    if len(_BOUND_CACHE) > 64:
        _BOUND_CACHE.clear()
    _BOUND_CACHE[id(bound)] = (weakref.ref(bound), bound._version, vals)
    return vals


def build_scene_struct(bound: torch.Tensor, coarse_bound: torch.Tensor,
                       native_grids: Dict[str, torch.Tensor], packed: Dict[str, torch.Tensor]) -> EnsScene:
    """Fill the C struct.  ``bound`` values are read on the host (they are CPU float64 in the
    reference: EvenNICER_SLAM.py:170-175)."""
    sc = EnsScene()
    b = _bound_floats(bound)
    cb = _bound_floats(coarse_bound)
    for k in range(3):
        sc.bound[k][0], sc.bound[k][1] = b[2 * k], b[2 * k + 1]
        sc.coarse_bound[k][0], sc.coarse_bound[k][1] = cb[2 * k], cb[2 * k + 1]
    for li, lv in enumerate(LEVELS):
        g = native_grids.get(lv)
        if g is not None:
            sc.grid[li] = g.data_ptr()
            sc.dims[li][0], sc.dims[li][1], sc.dims[li][2] = g.shape[0], g.shape[1], g.shape[2]
        w = packed.get(lv)
        if w is not None:
            sc.weights[li] = w.data_ptr()
    return sc
