"""world_size-2 gloo tests of the host-side sharding logic (no GPU)."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from evennicer_slam_b200 import sharding as sh
    try:
        # depth maxima: MAX over shards == max of the whole batch
        full = torch.arange(11, dtype=torch.float32) * 0.37
        lo, hi = sh.shard_range(11, rank, world)
        local = torch.tensor([float(full[lo:hi].max() * 1.2), float(full[lo:hi].max())], dtype=torch.float64)
        g = sh.global_depth_max(local)
        assert g[1].item() == float(full.max()) and g[0].item() == float(full.max() * 1.2)
        # gradient all-reduce, several tensors / dtypes, bucketed
        a = torch.full((5, 3), float(rank + 1))
        b = torch.full((7,), float(10 * (rank + 1)), dtype=torch.float64)
        c = torch.full((2, 2), float(rank))
        sh.allreduce_sum_([a, None, b, c], big_bytes=64)
        tot = sum(range(1, world + 1))
        assert torch.all(a == tot) and torch.all(b == 10 * tot) and torch.all(c == sum(range(world)))
        # views of one flat buffer (how decoder gradients arrive) are reduced as one span, in place
        arena = torch.zeros(40)
        v1, v2, v3 = arena[4:10].view(2, 3), arena[10:22], arena[22:30].view(4, 2)
        v1.fill_(float(rank + 1)); v2.fill_(2.0 * (rank + 1)); v3.fill_(3.0 * (rank + 1))
        sh.allreduce_sum_([v1, v2, v3], big_bytes=1 << 20)
        assert torch.all(v1 == tot) and torch.all(v2 == 2 * tot) and torch.all(v3 == 3 * tot)
        assert torch.all(arena[:4] == 0) and torch.all(arena[30:] == 0)
        # native-layout grid gradients (permuted views) + flat decoder gradients in one arena -> ONE collective, in place
        arena2 = torch.zeros(2 * 3 * 4 * 32 + 8 + 20)
        gridg = arena2[:768].view(2, 3, 4, 32).permute(3, 0, 1, 2).unsqueeze(0)       # [1,32,Z,Y,X] view of [Z,Y,X,32]
        decg = arena2[776:796].view(4, 5)
        assert not gridg.is_contiguous() and sh._dense_span(gridg) == (0, 768)
        gridg.copy_(torch.arange(768.).view(1, 32, 2, 3, 4) * (rank + 1)); decg.fill_(float(rank + 1))
        calls = []
        real = dist.all_reduce
        dist.all_reduce = lambda *a, **k: (calls.append(1), real(*a, **k))[1]
        try:
            sh.allreduce_sum_([gridg, decg, None], big_bytes=1 << 20)
        finally:
            dist.all_reduce = real
        assert len(calls) == 1
        assert torch.equal(gridg, torch.arange(768.).view(1, 32, 2, 3, 4) * tot) and torch.all(decg == tot)
        assert torch.all(arena2[768:776] == 0) and torch.all(arena2[796:] == 0)
        # a strided (non-dense) tensor falls back to the coalesced copy path and is still reduced correctly
        base = torch.zeros(6, 4); col = base[:, 1]
        col.fill_(float(rank + 1))
        sh.allreduce_sum_([col])
        assert torch.all(base[:, 1] == tot) and torch.all(base[:, 0] == 0)
        # frame permutation: rank-major padded blocks -> frame order, ragged last batch
        perm, pad = sh._frame_permutation(23, 10, world, "cpu")
        blocks = []
        for rk in range(world):
            rows = []
            for i in range(0, 23, 10):
                m = min(10, 23 - i)
                a0, a1 = sh.shard_range(m, rk, world)
                rows.extend(range(i + a0, i + a1))
            blocks.append(rows + [-1] * (pad - len(rows)))
        flat = torch.tensor([x for b in blocks for x in b])
        assert torch.equal(flat[perm], torch.arange(23))
        # render_frame_sharded end to end on gloo, with the CUDA render replaced by a per-ray CPU stand-in: every
        # reference batch split over the ranks, the batch's depth maxima taken from the WHOLE batch, one gather per output
        from evennicer_slam_b200 import functional as fn

        class FakeRenderer:
            ray_batch_size = 10

            def _setup(self, stage, decoders, device):
                return stage

        def fake_depth_max(gd):
            return torch.stack([(gd * 1.2).max().double(), gd.max().double()])

        def fake_render(setup, c, decoders, rd, ro, gd, depth_max=None):
            # depends on the ray AND on the batch maxima, as sample placement does
            base = ro[:, 0].double() * 3.0 + rd[:, 1].double()
            far = depth_max[0] if depth_max is not None else torch.tensor(0.0, dtype=torch.float64)
            return base + far, base * 0.5 + (depth_max[1] if depth_max is not None else 0.0), \
                torch.stack([ro[:, 0], rd[:, 0], ro[:, 1] + rd[:, 2]], -1).float()
        real = (fn.depth_batch_max, fn.render_batch_ray)
        fn.depth_batch_max, fn.render_batch_ray = fake_depth_max, fake_render
        try:
            g0 = torch.Generator().manual_seed(11)
            n = 37                                                  # 3 full batches of 10 and a ragged one of 7
            ro_f, rd_f = torch.randn(n, 3, generator=g0), torch.randn(n, 3, generator=g0)
            gdepth = torch.rand(n, generator=g0) + 0.5
            got = sh.render_frame_sharded(FakeRenderer(), None, None, rd_f, ro_f, "cpu", "color", gt_depth=gdepth)
            want = [[], [], []]
            for i in range(0, n, 10):
                dm = fake_depth_max(gdepth[i:i + 10])
                out = fake_render(None, None, None, rd_f[i:i + 10], ro_f[i:i + 10], gdepth[i:i + 10], depth_max=dm)
                for k in range(3):
                    want[k].append(out[k])
            for k in range(3):
                assert torch.equal(got[k], torch.cat(want[k])), k
            got_nd = sh.render_frame_sharded(FakeRenderer(), None, None, rd_f, ro_f, "cpu", "coarse", gt_depth=gdepth)
            assert torch.equal(got_nd[2], torch.cat(want[2]))       # stage coarse ignores gt_depth (Renderer.py:89-93)
        finally:
            fn.depth_batch_max, fn.render_batch_ray = real
        # eval_points_sharded (mesh lattice, config 5): points split over the ranks, outputs gathered in order
        class FakeEval:
            def eval_points(self, p, decoders, c, stage, device):
                return torch.cat([p * 2.0, p.sum(-1, keepdim=True)], -1)
        pts = torch.arange(13 * 3, dtype=torch.float32).view(13, 3)
        out = sh.eval_points_sharded(FakeEval(), pts, None, None, "fine", "cpu")
        assert torch.equal(out, torch.cat([pts * 2.0, pts.sum(-1, keepdim=True)], -1))
        # ragged all-gather
        rows = torch.arange(lo, hi, dtype=torch.float32)[:, None].repeat(1, 3)
        counts = [sh.shard_range(11, r, world)[1] - sh.shard_range(11, r, world)[0] for r in range(world)]
        allr = sh.allgather_rows(rows, counts)
        assert torch.equal(allr[:, 0], torch.arange(11, dtype=torch.float32))
        q.put((rank, "ok"))
    except Exception as e:  # pragma: no cover
        q.put((rank, repr(e)))
    finally:
        dist.destroy_process_group()


def test_shard_range_partitions():
    from evennicer_slam_b200.sharding import shard_range
    for n in (0, 1, 7, 100000, 816000):
        for w in (1, 2, 3, 8):
            spans = [shard_range(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(w - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


def test_collectives_world2_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(60)
    assert all(r[1] == "ok" for r in res), res
