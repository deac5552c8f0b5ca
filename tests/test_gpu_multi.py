"""Multi-GPU parity under pytest (``-m gpu``; skipped unless >= 2 GPUs are visible): two ranks, one per GPU, NCCL.

Each check compares the sharded path with the single-GPU result on the same inputs:
  * sharding.render_frame_sharded  (ray batches split over the ranks, outputs all-gathered)      bit-identical
  * sharding.eval_points_sharded   (point lattice split over the ranks)                          bit-identical
  * the mapping step with rays sharded and gradients SUM all-reduced (sharding.allreduce_sum_)   1e-5 of the 1-GPU gradients
    (float32 sums in a different order; the relu decisions are per point and do not depend on the split)
"""
import os
import socket
import sys

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def _worker(rank, world, port, out):
    here = os.path.dirname(os.path.abspath(__file__))
    root = os.path.dirname(here)
    for p in (root, os.path.join(root, "oracle"), os.path.join(here, "golden")):
        if p not in sys.path:
            sys.path.insert(0, p)
    import torch.distributed as dist
    import cases
    from evennicer_slam_b200 import harness, sharding, common
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        scene = cases.tiny_scene()
        decoders, c, renderer, cfg = harness.build(scene, dev)
        renderer.ray_batch_size = 1000                     # several ragged reference batches in the tiny frame
        cam = scene.cam
        cam_t, depth, color, _ = cases.tiny_frame()
        depth_t = torch.from_numpy(depth).to(dev).reshape(-1)
        res = {}
        with torch.no_grad():
            c2w = common.get_camera_from_tensor(torch.from_numpy(cam_t.copy()).to(dev))
            ro, rd = common.get_rays(cam.H, cam.W, cam.fx, cam.fy, cam.cx, cam.cy, c2w, dev)
            ro, rd = ro.reshape(-1, 3), rd.reshape(-1, 3)
            sh = sharding.render_frame_sharded(renderer, c, decoders, rd, ro, dev, "color", gt_depth=depth_t)
            ref = renderer._render_rays_batched(c, decoders, ro, rd, dev, "color", depth_t)
            res["frame"] = all(bool(torch.equal(a, b)) for a, b in zip(sh, ref))
            pts = torch.from_numpy(cases.eval_points_lattice(scene)).to(dev)
            res["points"] = bool(torch.equal(sharding.eval_points_sharded(renderer, pts, decoders, c, "fine", dev),
                                             renderer.eval_points(pts, decoders, c, "fine", dev)))
        # mapping step: 96 rays, the ranks take halves; gradients all-reduced vs the whole batch on one GPU
        g = np.load(os.path.join(here, "golden", "tiny_render.npz"))
        rays_o = torch.from_numpy(np.concatenate([g["color.d.rays_o"], g["fine.d.rays_o"]])).to(dev)
        rays_d = torch.from_numpy(np.concatenate([g["color.d.rays_d"], g["fine.d.rays_d"]])).to(dev)
        sd = torch.from_numpy(np.concatenate([g["color.d.sample_depth"], g["fine.d.sample_depth"]])).to(dev)
        n = rays_o.shape[0]

        def grads(lo, hi, reduce):
            for p in decoders.parameters():
                p.grad = None
            cg = {k: v.clone().requires_grad_(True) for k, v in c.items()}
            from evennicer_slam_b200.functional import render_batch_ray, depth_batch_max
            setup = renderer._setup("color", decoders, dev)
            d, u, col = render_batch_ray(setup, cg, decoders, rays_d[lo:hi], rays_o[lo:hi], sd[lo:hi],
                                         depth_max=depth_batch_max(sd.contiguous()))
            (torch.abs(sd[lo:hi] - d).sum() + 0.2 * torch.abs(col).sum()).backward()
            gs = [cg[k].grad for k in ("grid_middle", "grid_fine", "grid_color")] + [p.grad for p in decoders.parameters() if p.grad is not None]
            if reduce:
                sharding.allreduce_sum_(gs)
            return [t.clone() for t in gs]
        lo, hi = sharding.shard_range(n, rank, world)
        part = grads(lo, hi, True)
        whole = grads(0, n, False)
        worst = max(float((a - b).abs().max() / b.abs().max().clamp_min(1e-30)) for a, b in zip(part, whole))
        res["grad_rel"] = worst
        # the same with native-layout grids and the touched-voxel all-reduce (sharding.SparseGradAllReduce)
        from evennicer_slam_b200.scene import as_native_layout
        from evennicer_slam_b200.functional import render_batch_ray, depth_batch_max
        cn = {k: as_native_layout(v.clone()).requires_grad_(True) for k, v in c.items()}
        keys = ["grid_middle", "grid_fine", "grid_color"]
        ar = sharding.SparseGradAllReduce([cn[k] for k in keys], capacity_frac=0.9)
        for p in decoders.parameters():
            p.grad = None
        setup = renderer._setup("color", decoders, dev)
        d, u, col = render_batch_ray(setup, cn, decoders, rays_d[lo:hi], rays_o[lo:hi], sd[lo:hi], depth_max=depth_batch_max(sd.contiguous()))
        (torch.abs(sd[lo:hi] - d).sum() + 0.2 * torch.abs(col).sum()).backward()
        ps = [p for p in decoders.parameters() if p.grad is not None]
        ar([cn[k].grad for k in keys], [p.grad for p in ps])
        got = [cn[k].grad for k in keys] + [p.grad for p in ps]
        res["sparse_rel"] = max(float((a - b).abs().max() / b.abs().max().clamp_min(1e-30)) for a, b in zip(got, whole))
        res["sparse_overflow"] = ar.overflowed()
        if rank == 0:
            out.put(res)
        dist.barrier()
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_two_rank_sharding_matches_one_gpu():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    res = out.get(timeout=300)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert res["frame"] and res["points"], res
    assert res["grad_rel"] < 1e-5, res
    assert res["sparse_rel"] < 1e-5 and not res["sparse_overflow"], res
