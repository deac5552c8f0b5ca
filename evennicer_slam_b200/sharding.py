"""Ray / point sharding across the GPUs of one box (one process per GPU, torch.distributed).

The render path shards by independent units (rays, lattice points); SURVEY.md 8(e).  The only
cross-ray coupling inside ``render_batch_ray`` is the pair of batch maxima of ``gt_depth``
(Renderer.py:110,145), recovered exactly with one MAX all-reduce of two doubles.  Exchanges:

  full-frame render / mesh lattice : all-gather of per-ray (per-point) outputs, 28 (16) bytes each
  mapping step                     : SUM all-reduce of grid / decoder / pose gradients (replicated scene)
  tracking 200 px                  : replicas only (no collective)

NCCL over NVLink on the GPU box; the same code runs on gloo/CPU tensors in the tests.
"""
from __future__ import annotations

from typing import Optional, Sequence, Tuple

import torch
import torch.distributed as dist


def shard_range(n: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous, balanced [lo, hi) slice of n units for `rank` (first n % world ranks get one more)."""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def world(group=None) -> Tuple[int, int]:
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(group), dist.get_world_size(group)
    return 0, 1


def global_depth_max(local: torch.Tensor, group=None) -> torch.Tensor:
    """MAX all-reduce of the (max(gt*1.2), max(gt)) pair so every shard places samples as the whole batch would."""
    if world(group)[1] > 1:
        dist.all_reduce(local, op=dist.ReduceOp.MAX, group=group)
    return local


def _dense_span(t: torch.Tensor) -> Optional[Tuple[int, int]]:
    """[lo, hi) storage offsets (in elements) of a tensor that covers a contiguous piece of its storage exactly once --
    contiguous tensors and permuted views of contiguous buffers (the native-layout grid gradients) -- else None."""
    if t.numel() == 0:
        return None
    dims = sorted(((st, n) for st, n in zip(t.stride(), t.shape) if n > 1))
    expect = 1
    for st, n in dims:
        if st != expect:
            return None
        expect *= n
    return t.storage_offset(), t.storage_offset() + t.numel()


def allreduce_sum_(tensors: Sequence[Optional[torch.Tensor]], group=None, big_bytes: int = 1 << 20) -> None:
    """In-place SUM all-reduce of gradient tensors, with as few collectives as the memory layout allows.

    The render backward hands out its gradients as views of ONE zero-filled arena (native-layout grid gradients followed
    by the flat decoder gradients): tensors that are dense pieces of a shared storage are all-reduced as ONE span of that
    storage, in place (the sum is elementwise, so the views' permutations do not matter; only a few elements of alignment
    padding may separate the pieces).  Whatever is left (stand-alone tensors) is reduced in place when large, or
    coalesced into one flat buffer per dtype when small.
    """
    if world(group)[1] == 1:
        return
    by_store = {}
    loose = []
    for t in tensors:
        if t is None:
            continue
        sp = _dense_span(t)
        if sp is None:
            loose.append(t)
        else:
            by_store.setdefault((t.untyped_storage().data_ptr(), t.dtype, t.device), []).append((sp, t))
    small = {}
    for (_, dtype, dev), items in by_store.items():
        lo = min(sp[0] for sp, _ in items)
        hi = max(sp[1] for sp, _ in items)
        covered = sum(sp[1] - sp[0] for sp, _ in items)
        t0 = items[0][1]
        # only alignment padding may lie between the pieces (the arena pads every sink to 16 bytes): anything larger could
        # be somebody else's live data, which an in-place SUM over the span would corrupt
        if (hi - lo) - covered <= 16 * len(items) and ((hi - lo) * t0.element_size() >= big_bytes or len(items) > 1):
            span = torch.empty(0, dtype=dtype, device=dev).set_(t0.untyped_storage(), lo, (hi - lo,), (1,))
            dist.all_reduce(span, op=dist.ReduceOp.SUM, group=group)
        else:
            for _, t in items:
                small.setdefault((dtype, dev), []).append(t)
    for t in loose:
        small.setdefault((t.dtype, t.device), []).append(t)
    for (_, _), ts in small.items():
        flat = torch.cat([b.reshape(-1) for b in ts]) if len(ts) > 1 else ts[0].reshape(-1).contiguous()
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
        off = 0
        for b in ts:
            b.copy_(flat[off:off + b.numel()].view_as(b))
            off += b.numel()


def allgather_rows(local: torch.Tensor, counts: Sequence[int], group=None) -> torch.Tensor:
    """Concatenate per-rank row blocks (rank r holds counts[r] rows) on every rank."""
    r, w = world(group)
    if w == 1:
        return local
    mx = max(counts)
    pad = local
    if local.shape[0] < mx:
        pad = torch.cat([local, local.new_zeros((mx - local.shape[0],) + tuple(local.shape[1:]))], 0)
    out = [torch.empty_like(pad) for _ in range(w)]
    dist.all_gather(out, pad.contiguous(), group=group)
    return torch.cat([o[:n] for o, n in zip(out, counts)], 0)


def render_rays_sharded(renderer, c, decoders, rays_d, rays_o, device, stage, gt_depth=None, group=None):
    """render_batch_ray of ONE reference batch, rays split over the ranks; every rank gets all outputs.

    Identical to the unsharded result: the depth maxima are reduced over the whole batch first.
    """
    from .functional import depth_batch_max, render_batch_ray as _rbr
    r, w = world(group)
    n = rays_o.shape[0]
    lo, hi = shard_range(n, r, w)
    dmax = None
    gd = None
    if gt_depth is not None and stage != "coarse":
        gd = gt_depth.reshape(-1).float()
        if hi > lo:
            dmax = depth_batch_max(gd[lo:hi].contiguous())
        else:
            dmax = torch.full((2,), float("-inf"), dtype=torch.float64, device=rays_o.device)
        dmax = global_depth_max(dmax, group)
        gd = gd[lo:hi]
    setup = renderer._setup(stage, decoders, rays_o.device)
    if hi > lo:
        depth, var, color = _rbr(setup, c, decoders, rays_d[lo:hi], rays_o[lo:hi], gd, depth_max=dmax)
    else:
        depth = torch.empty(0, dtype=torch.float64, device=rays_o.device)
        var = torch.empty(0, dtype=torch.float64, device=rays_o.device)
        color = torch.empty((0, 3), dtype=torch.float32, device=rays_o.device)
    counts = [shard_range(n, q, w)[1] - shard_range(n, q, w)[0] for q in range(w)]
    return (allgather_rows(depth, counts, group), allgather_rows(var, counts, group),
            allgather_rows(color, counts, group))


_PERM_CACHE = {}


def _frame_permutation(n: int, batch: int, w: int, device) -> torch.Tensor:
    """Row index into the rank-major all-gathered block (every rank padded to the same length) for each ray of the
    frame in frame order, when every reference batch of `batch` rays is split over the ranks with shard_range."""
    key = (n, batch, w, str(device))
    perm = _PERM_CACHE.get(key)
    if perm is None:
        per_rank = [0] * w
        for i in range(0, n, batch):
            m = min(batch, n - i)
            for q in range(w):
                lo, hi = shard_range(m, q, w)
                per_rank[q] += hi - lo
        pad = max(per_rank)
        idx = torch.empty(n, dtype=torch.int64)
        cursor = [0] * w
        for i in range(0, n, batch):
            m = min(batch, n - i)
            for q in range(w):
                lo, hi = shard_range(m, q, w)
                idx[i + lo:i + hi] = torch.arange(q * pad + cursor[q], q * pad + cursor[q] + (hi - lo))
                cursor[q] += hi - lo
        perm = (idx.to(device), pad)
        _PERM_CACHE[key] = perm
    return perm


def render_frame_sharded(renderer, c, decoders, rays_d, rays_o, device, stage, gt_depth=None, group=None):
    """Renderer.render_img's batch loop over a whole frame on N ranks, no gradient: every reference batch
    (renderer.ray_batch_size rays, whose two depth maxima each rank takes from its own copy of the depth image, so no
    collective is needed for them) is split over the ranks; the ranks render their shards back to back and the frame
    is assembled with ONE all-gather per output at the end.  Bit-identical to the unsharded loop."""
    from .functional import depth_batch_max, render_batch_ray as _rbr
    r, w = world(group)
    n, B = rays_o.shape[0], renderer.ray_batch_size
    setup = renderer._setup(stage, decoders, rays_o.device)
    has_depth = gt_depth is not None and stage != "coarse"
    gd_all = gt_depth.reshape(-1).float() if has_depth else None
    outs = []
    with torch.no_grad():
        for i in range(0, n, B):
            m = min(B, n - i)
            lo, hi = shard_range(m, r, w)
            dmax = depth_batch_max(gd_all[i:i + m].contiguous()) if has_depth else None
            if hi > lo:
                outs.append(_rbr(setup, c, decoders, rays_d[i + lo:i + hi], rays_o[i + lo:i + hi],
                                 gd_all[i + lo:i + hi] if has_depth else None, depth_max=dmax))
        if w == 1:
            return tuple(torch.cat([o[k] for o in outs]) for k in range(3))
        perm, pad = _frame_permutation(n, B, w, rays_o.device)
        res = []
        for k, (dt, tail) in enumerate(((torch.float64, ()), (torch.float64, ()), (torch.float32, (3,)))):
            local = torch.cat([o[k] for o in outs]) if outs else torch.empty((0,) + tail, dtype=dt, device=rays_o.device)
            buf = torch.zeros((pad,) + tail, dtype=dt, device=rays_o.device)
            buf[:local.shape[0]] = local
            full = torch.empty((w * pad,) + tail, dtype=dt, device=rays_o.device)
            dist.all_gather_into_tensor(full, buf, group=group)
            res.append(full.index_select(0, perm))
        return tuple(res)


def eval_points_sharded(renderer, p, decoders, c, stage, device, group=None):
    """Renderer.eval_points over a point lattice split across ranks (mesh extraction, config 5)."""
    r, w = world(group)
    n = p.shape[0]
    lo, hi = shard_range(n, r, w)
    local = renderer.eval_points(p[lo:hi], decoders, c, stage, device)
    counts = [shard_range(n, q, w)[1] - shard_range(n, q, w)[0] for q in range(w)]
    return allgather_rows(local, counts, group)


def _coalesce(tensors) -> list:
    """Flat views covering `tensors` with as few pieces as their storages allow: tensors that are dense pieces of one storage
    (the flat decoder gradients of the backward's arena) become ONE span; the rest stay single."""
    by_store, out = {}, []
    for t in tensors:
        sp = _dense_span(t)
        if sp is None:
            out.append(t)
        else:
            by_store.setdefault((t.untyped_storage().data_ptr(), t.dtype, t.device), []).append((sp, t))
    for (_, dtype, dev), items in by_store.items():
        lo = min(sp[0] for sp, _ in items)
        hi = max(sp[1] for sp, _ in items)
        covered = sum(sp[1] - sp[0] for sp, _ in items)
        t0 = items[0][1]
        if len(items) > 1 and (hi - lo) - covered <= 16 * len(items):
            out.append(torch.empty(0, dtype=dtype, device=dev).set_(t0.untyped_storage(), lo, (hi - lo,), (1,)))
        else:
            out.extend(t for _, t in items)
    return out


class SparseGradAllReduce:
    """SUM all-reduce of the mapping step's gradients that exchanges only the grid voxels some rank touched.

    The render backward returns DENSE native-layout grid gradients (48 MB for room0, 5 % non-zero for a 1000-ray batch); an
    all-reduce of the whole arena is what limited the 1 -> 8 GPU curve of the sharded mapping step.  Per call:

      1. ``ens_grid_touched`` per level           -> int32 flag per voxel (one array for all levels)
      2. MAX all-reduce of the flags (1.3 MB)     -> the union of touched voxels, identical on every rank
      3. inclusive scan of the flags              -> row of every flagged voxel in the compact buffer
      4. ``ens_grid_compact`` per level           -> flagged rows copied into ``compact`` ([capacity][32]); the decoder /
                                                     camera gradients are appended behind them
      5. ONE SUM all-reduce of ``compact``        (capacity rows instead of every voxel)
      6. ``ens_grid_compact`` back, small tensors copied back

    All shapes are static (the step can be captured in a CUDA graph).  ``capacity_frac`` of the voxels fit in the compact
    buffer; rows beyond it are counted in ``self.overflow`` (device int32) -- check ``overflowed()`` outside the timed /
    captured region and fall back to ``allreduce_sum_`` if it ever trips.  Grid gradients must be native-layout views
    (what the backward returns for grids allocated with ``scene.as_native_layout``).
    """

    def __init__(self, grids, capacity_frac: float = 0.25, group=None):
        from .scene import is_native_strided
        self.group = group
        self.shapes = []
        tot = 0
        dev = None
        for g in grids:
            if not is_native_strided(g):
                raise ValueError("SparseGradAllReduce needs native-layout grids (scene.as_native_layout)")
            v = g.shape[2] * g.shape[3] * g.shape[4]
            self.shapes.append(v)
            tot += v
            dev = g.device
        self.n_vox = tot
        self.capacity = max(1, int(tot * capacity_frac))
        self.flags = torch.zeros(tot, dtype=torch.int32, device=dev)
        self.pos = torch.zeros(tot, dtype=torch.int32, device=dev)
        self.overflow = torch.zeros(1, dtype=torch.int32, device=dev)
        self.compact = None
        self.tail = 0

    def overflowed(self) -> bool:
        return bool(int(self.overflow.item()) != 0)

    def __call__(self, grid_grads, others) -> None:
        """grid_grads: the native-layout gradient views of the grids given to __init__ (same order); others: every other
        gradient tensor (decoder parameters, camera tensors).  In place."""
        from . import _lib
        if world(self.group)[1] == 1:
            return
        L = _lib.lib()
        dev = self.flags.device
        stream = _lib.cur_stream(dev)
        bases = [g.detach()[0].permute(1, 2, 3, 0) for g in grid_grads]          # contiguous [Z,Y,X,32]
        others = _coalesce([t for t in others if t is not None])
        tail = sum(t.numel() for t in others)
        if self.compact is None or self.tail != tail:
            self.tail = tail
            self.compact = torch.zeros(self.capacity * 32 + tail, dtype=torch.float32, device=dev)
        self.compact.zero_()          # rows no voxel maps to this step would otherwise keep (and keep summing) stale values
        # the backward lays the grid gradients out back to back in ONE arena: then all levels are one [n_vox][32] array and
        # every pass below is a single launch
        pieces = list(zip(bases, self.shapes))
        if all(b.is_contiguous() for b in bases) and all(
                bases[i].data_ptr() + bases[i].numel() * 4 == bases[i + 1].data_ptr() for i in range(len(bases) - 1)):
            pieces = [(bases[0], self.n_vox)]
        off = 0
        for b, v in pieces:
            assert b.is_contiguous()
            _lib.check(L.ens_grid_touched(_lib.ptr(b), v, C_ptr(self.flags, off), stream), "ens_grid_touched")
            off += v
        dist.all_reduce(self.flags, op=dist.ReduceOp.MAX, group=self.group)
        torch.cumsum(self.flags, 0, dtype=torch.int32, out=self.pos)
        off = 0
        for b, v in pieces:
            _lib.check(L.ens_grid_compact(_lib.ptr(b), C_ptr(self.flags, off), C_ptr(self.pos, off), v, _lib.ptr(self.compact),
                                          self.capacity, 1, _lib.ptr(self.overflow), stream), "ens_grid_compact")
            off += v
        tl = self.compact[self.capacity * 32:]
        o = 0
        for t in others:
            tl[o:o + t.numel()].copy_(t.reshape(-1))
            o += t.numel()
        dist.all_reduce(self.compact, op=dist.ReduceOp.SUM, group=self.group)
        off = 0
        for b, v in pieces:
            _lib.check(L.ens_grid_compact(_lib.ptr(b), C_ptr(self.flags, off), C_ptr(self.pos, off), v, _lib.ptr(self.compact),
                                          self.capacity, 0, _lib.ptr(self.overflow), stream), "ens_grid_compact")
            off += v
        o = 0
        for t in others:
            t.copy_(tl[o:o + t.numel()].view_as(t))
            o += t.numel()


def C_ptr(t: torch.Tensor, elem_offset: int):
    """device pointer of element `elem_offset` of a contiguous tensor, for the ctypes calls"""
    import ctypes
    return ctypes.c_void_p(t.data_ptr() + elem_offset * t.element_size())
