#!/usr/bin/env python
"""Benchmark of the fused ray-rendering hot path (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # this framework (CUDA, sm_100a)
    python bench.py --impl reference --gpus N --steps K ...   # CPU baseline: the oracle port on host cores

One JSON line on stdout.  A "step" is one pass of the hot path over one synthetic Replica mapping
batch (config C3/C1 of SURVEY.md 8(d)): 5 frames x 200 px drawn with get_samples (4 BA poses carry
gradient), colour-stage render_batch_ray over room0 grids (46 MiB), the Mapper loss
(Mapper.py:553-562) and backward into grid features, all decoder weights and the camera tensors.
The scene cache is invalidated every step (Mapper mutates the grids every iteration, Mapper.py:451-458),
so the layout conversions are inside the timed region.  L2 is flushed between timed steps.

  value  : rays/s, device-resident inputs, CUDA-event time summed over K steps (max over ranks)
  e2e    : rays/s through the C-ABI-level call with HOST (pinned) ray buffers in, outputs + ray
           gradients + loss out, copies inside the timed region
  N > 1  : weak scaling -- every rank runs its own 1000-ray batch on a replicated scene and the grid /
           decoder / pose gradients are SUM all-reduced over NCCL (inside the timed region)
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests", "golden")):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np  # noqa: E402

N_FRAMES = 5
PIX_PER_FRAME = 200
N_RAYS = N_FRAMES * PIX_PER_FRAME
S_TOTAL = 48
# algorithmic grid bytes: 1024 B per point per level; colour stage touches 3 levels
BYTES_PER_POINT_FWD = 1024 * 3
BYTES_PER_POINT_BWD = 1024 * 3 + 1024 * 3       # re-gather + scatter-add
BYTES_PER_RAY = (BYTES_PER_POINT_FWD + BYTES_PER_POINT_BWD) * S_TOTAL   # 442 368


_REAL_STDOUT = None


def emit(line):
    out = _REAL_STDOUT or sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def hbm_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def make_inputs():
    import cases
    import evennicer_slam_b200.synthetic as syn
    scene = cases.room0_scene()
    frames = []
    for f in range(N_FRAMES):
        cam_t = syn.default_pose(syn.ROOM0_BOUND, jitter_seed=10 + f)
        depth, color, _ = syn.synthetic_frame(syn.ROOM0_BOUND, syn.REPLICA_CAM, cam_t, seed=100 + f, zero_frac=0.02)
        frames.append((cam_t, depth, color))
    return scene, frames


# ---------------------------------------------------------------------------------------------
# clocks sampling (nvidia-smi during the timed region)
# ---------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.path = tempfile.mktemp(suffix=".csv")
        self.proc = None
        self.idx = gpu_index

    def start(self):
        try:
            self.f = open(self.path, "w")
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.idx), "-lms", "100"], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        self.f.close()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in open(self.path):
            parts = [x.strip() for x in line.split(",")]
            if len(parts) < 9:
                continue
            try:
                sm.append(float(parts[1])); mx.append(float(parts[2]))
            except ValueError:
                continue
            for nm, v in zip(names, parts[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        if sm:
            out = {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons),
                   "samples": len(sm)}
        try:
            os.remove(self.path)
        except OSError:
            pass
        return out


# ---------------------------------------------------------------------------------------------
# CPU arm: the oracle port on host cores
# ---------------------------------------------------------------------------------------------
def oracle_step(sc, frames, n_per_frame, seed, t32, t64):
    """get_samples -> render_batch_ray -> Mapper loss -> backward with the numpy oracle; returns rays done."""
    import evennicer_slam_b200.synthetic as syn
    import render_oracle as orc
    cam = syn.REPLICA_CAM
    rng = np.random.RandomState(seed)
    ros, rds, sds, scs = [], [], [], []
    for (cam_t, depth, color) in frames:
        idx = rng.randint(0, cam.H * cam.W, size=n_per_frame)
        i, j, sd, scol = orc.select_pixels(idx, 0, cam.H, 0, cam.W, depth, color)
        ro, rd = orc.rays_from_uv(i, j, syn.quat_to_c2w(cam_t), cam.fx, cam.fy, cam.cx, cam.cy)
        ros.append(ro); rds.append(rd); sds.append(sd); scs.append(scol)
    ro, rd, sd, scol = np.concatenate(ros), np.concatenate(rds), np.concatenate(sds), np.concatenate(scs)
    dep, var, col, cache = orc.render_batch_ray(sc, ro, rd, "color", sd, t32, t64)
    g_d = np.where(sd > 0, -np.sign(sd.astype(np.float64) - dep), 0.0)          # d/d depth of sum|gt-d|[gt>0]
    g_c = (-0.2 * np.sign(scol.astype(np.float32) - col)).astype(np.float32)     # 0.2 * sum|gt-c|
    grads = orc.render_batch_ray_backward(sc, cache, g_d, None, g_c)
    for f in range(len(frames)):
        sl = slice(f * n_per_frame, (f + 1) * n_per_frame)
        orc.rays_from_uv_backward(np.zeros(n_per_frame), np.zeros(n_per_frame), cam.fx, cam.fy, cam.cx, cam.cy,
                                  grads["rays_o"][sl], grads["rays_d"][sl])
    return ro.shape[0]


class ReferenceStepper:
    """The mapping step through the UNMODIFIED reference modules (baseline/install_ref.py -> baseline/_ref, or
    /root/reference): get_samples -> Renderer.render_batch_ray -> Mapper loss (Mapper.py:553-562, boolean mask) -> backward
    into grids, decoder weights and the camera tensors.  device 'cpu' = the reference arm / cpu_baseline; 'cuda:k' = the
    eager-PyTorch-on-B200 figure."""

    def __init__(self, device):
        import torch
        import ref_harness as rh
        self.torch, self.dev = torch, device
        scene, frames = make_inputs()
        self.ref = rh.load()
        self.model, c, self.renderer, self.cfg = rh.build_reference(scene, device=device)
        self.cam = scene.cam
        self.c = {k: v.clone().requires_grad_(True) for k, v in c.items()}
        self.depth = [torch.from_numpy(d).to(device) for (_, d, _) in frames]
        self.color = [torch.from_numpy(col).to(device) for (_, _, col) in frames]
        self.cams = [torch.from_numpy(ct.copy()).to(device).requires_grad_(f > 0) for f, (ct, _, _) in enumerate(frames)]

    def step(self, n_per_frame, seed):
        torch, cam = self.torch, self.cam
        torch.manual_seed(seed)
        for t in list(self.c.values()) + list(self.model.parameters()) + self.cams:
            t.grad = None
        ros, rds, sds, scs = [], [], [], []
        for f in range(len(self.cams)):
            c2w = self.ref.common.get_camera_from_tensor(self.cams[f])
            ro, rd, sd, sc = self.ref.common.get_samples(0, cam.H, 0, cam.W, n_per_frame, cam.H, cam.W, cam.fx, cam.fy,
                                                         cam.cx, cam.cy, c2w, self.depth[f], self.color[f], self.dev)
            ros.append(ro.float()); rds.append(rd.float()); sds.append(sd.float()); scs.append(sc.float())
        ro, rd, sd, sc = torch.cat(ros), torch.cat(rds), torch.cat(sds), torch.cat(scs)
        depth, unc, color = self.renderer.render_batch_ray(self.c, self.model, rd, ro, self.dev, "color", gt_depth=sd)
        mask = sd > 0
        loss = torch.abs(sd - depth)[mask].sum() + 0.2 * torch.abs(sc - color).sum()
        loss.backward()
        return ro.shape[0]


def reference_available():
    try:
        import ref_harness as rh
        return rh.available()
    except Exception:
        return False


def measure_gpu_reference(dev, steps=5):
    """The reference's own eager-PyTorch renderer on the SAME GPU, same inputs and step as the headline."""
    import torch
    if not reference_available():
        return {"unavailable": "reference modules not present (baseline/install_ref.py was not run in the build container)"}
    st = ReferenceStepper(str(dev))
    for w in range(2):
        st.step(PIX_PER_FRAME, w)
    torch.cuda.synchronize()
    evs = []
    for k in range(steps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); st.step(PIX_PER_FRAME, 100 + k); b.record()
        evs.append((a, b))
    torch.cuda.synchronize()
    ms = sum(a.elapsed_time(b) for a, b in evs) / steps
    launches = None
    try:
        from torch.profiler import profile, ProfilerActivity
        with profile(activities=[ProfilerActivity.CUDA]) as prof:
            st.step(PIX_PER_FRAME, 999)
            torch.cuda.synchronize()
        launches = sum(1 for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA and "memcpy" not in e.name.lower()
                       and "memset" not in e.name.lower())
    except Exception:
        pass
    return {"rays_per_s": N_RAYS / (ms * 1e-3), "ms_per_step": ms, "launches": launches, "steps": steps,
            "what": "unmodified reference (src/common.py, src/conv_onet, src/utils/Renderer.py) in eager PyTorch on this GPU: "
                    "get_samples + render_batch_ray + Mapper loss + backward, same scene / frames / 1000-ray colour-stage step"}


def blas_threads():
    try:
        from threadpoolctl import threadpool_info
        n = [p.get("num_threads", 1) for p in threadpool_info() if p.get("user_api") == "blas"]
        return int(max(n)) if n else 1
    except Exception:
        return 1


def use_all_host_threads():
    """torchrun exports OMP_NUM_THREADS=1 to its workers; the CPU arm is meant to use every host core it can."""
    n = os.cpu_count() or 1
    try:
        from threadpoolctl import threadpool_limits
        threadpool_limits(limits=n)            # BLAS / OpenMP pools numpy already initialised
    except Exception:
        pass
    try:
        import torch
        torch.set_num_threads(n)
    except Exception:
        pass


def cpu_reference_sample(budget_s, steps=None, warmup=0):
    """Bounded sample of the mapping step through the REAL reference on the host cores (kind 'reference')."""
    import torch
    use_all_host_threads()
    st = ReferenceStepper("cpu")
    st.step(4, 0)                                        # one-time costs (allocator, thread pools) stay out of the probe
    t0 = time.perf_counter()
    n = st.step(20, 1)
    rate = n / (time.perf_counter() - t0)
    if steps is None:
        n_pf = int(max(10, min(PIX_PER_FRAME, rate * budget_s / N_FRAMES)))
        steps = int(max(1, min(40, budget_s * rate / max(n_pf * N_FRAMES, 1))))
    else:
        n_pf = int(max(4, min(PIX_PER_FRAME, rate * (budget_s / max(steps + warmup, 1)) / N_FRAMES)))
    for w in range(warmup):
        st.step(n_pf, 10 + w)
    t0 = time.perf_counter()
    n = 0
    for r in range(steps):
        n += st.step(n_pf, 100 + r)
    dt = time.perf_counter() - t0
    return n / dt, n, steps, n_pf, dt, torch.get_num_threads()


def cpu_baseline(budget_s=12.0):
    """Bounded sample of the mapping workload on the host cores: the reference itself when its modules are present
    (kind 'reference'), else the numpy oracle port (kind 'port')."""
    import torch
    if reference_available():
        val, n, reps, n_pf, dt, thr = cpu_reference_sample(budget_s)
        return {"value": val, "unit": "rays/s", "cores": thr, "kind": "reference",
                "sample": f"{n} rays = {reps} x ({N_FRAMES} frames x {n_pf} px) of the {N_RAYS}-ray colour-stage mapping batch, "
                          f"fwd+bwd, the reference's own PyTorch CPU renderer (baseline/_ref), {dt:.1f} s", "host_cpus": os.cpu_count()}
    use_all_host_threads()
    import render_oracle as orc
    scene, frames = make_inputs()
    sc = orc.OracleScene.from_synthetic(scene)
    t32 = torch.linspace(0., 1., 32).numpy()
    t64 = torch.linspace(0., 1., 16).double().numpy()
    n_pf = 20                                            # 100-ray probe to size the sample
    t0 = time.perf_counter()
    n = oracle_step(sc, frames, n_pf, 0, t32, t64)
    probe = time.perf_counter() - t0
    rate = n / probe
    n_pf = int(max(10, min(PIX_PER_FRAME, rate * budget_s / N_FRAMES)))
    # whole batches (new pixel draw each) until about `budget_s` seconds of CPU work have been timed
    reps = int(max(1, min(40, budget_s * rate / max(n_pf * N_FRAMES, 1))))
    t0 = time.perf_counter()
    n = 0
    for r in range(reps):
        n += oracle_step(sc, frames, n_pf, 1 + r, t32, t64)
    dt = time.perf_counter() - t0
    return {"value": n / dt, "unit": "rays/s", "cores": blas_threads(), "kind": "port",
            "sample": f"{n} rays = {reps} x ({N_FRAMES} frames x {n_pf} px) of the {N_RAYS}-ray colour-stage mapping batch, "
                      f"fwd+bwd, numpy oracle, {dt:.1f} s", "host_cpus": os.cpu_count()}


def run_reference_arm(args):
    import torch
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    if reference_available():
        val, n, steps, n_pf, dt, thr = cpu_reference_sample(150.0, steps=args.steps, warmup=args.warmup)
        emit({
            "impl": "reference", "metric": "rays/sec (fwd+bwd) on the Replica mapping batch", "value": val,
            "unit": "rays/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * dt / max(args.steps, 1), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32 (f64 sample placement / depth sums)", "data": "synthetic",
            "config": workload_config(),
            "cpu_baseline": {"value": val, "unit": "rays/s", "cores": thr, "kind": "reference",
                             "sample": f"{n_pf * N_FRAMES} rays per step of the {N_RAYS}-ray batch, the reference's own PyTorch "
                                       "CPU renderer (unmodified modules under baseline/_ref)", "host_cpus": os.cpu_count()},
            "e2e": {"value": val, "unit": "rays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0,
        })
        return
    use_all_host_threads()
    import render_oracle as orc
    scene, frames = make_inputs()
    sc = orc.OracleScene.from_synthetic(scene)
    t32 = torch.linspace(0., 1., 32).numpy()
    t64 = torch.linspace(0., 1., 16).double().numpy()
    t0 = time.perf_counter()
    n = oracle_step(sc, frames, 10, 0, t32, t64)
    rate = n / (time.perf_counter() - t0)
    total = args.steps + args.warmup
    n_pf = int(max(4, min(PIX_PER_FRAME, rate * (150.0 / max(total, 1)) / N_FRAMES)))
    for w in range(args.warmup):
        oracle_step(sc, frames, n_pf, 10 + w, t32, t64)
    t0 = time.perf_counter()
    done = 0
    for k in range(args.steps):
        done += oracle_step(sc, frames, n_pf, 100 + k, t32, t64)
    dt = time.perf_counter() - t0
    val = done / dt
    line = {
        "impl": "reference", "metric": "rays/sec (fwd+bwd) on the Replica mapping batch", "value": val,
        "unit": "rays/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * dt / max(args.steps, 1), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32 (f64 sample placement / depth sums)", "data": "synthetic",
        "config": workload_config(),
        "cpu_baseline": {"value": val, "unit": "rays/s", "cores": blas_threads(), "kind": "port",
                         "sample": f"{n_pf * N_FRAMES} rays per step of the {N_RAYS}-ray batch, numpy oracle port "
                                   "of the reference renderer (the Python reference cannot travel to the GPU box)",
                         "host_cpus": os.cpu_count()},
        "e2e": {"value": val, "unit": "rays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


def workload_config():
    return {"workload": "Replica room0 mapping step, colour stage: 5 frames x 200 px = 1000 rays x (32 stratified + "
                        "16 surface) samples, get_samples + render_batch_ray + Mapper loss + backward into grids "
                        "(46 MiB, 3 levels), decoder weights and 4 BA camera tensors",
            "rays_per_step_per_gpu": N_RAYS, "samples_per_ray": S_TOTAL, "grids": "room0 [1,32,Z,Y,X] x 4 levels",
            "decoders": "random-init NICE (pretrained blobs absent from the reference tree)",
            "cache": "decoders re-packed every step (Adam changes them every iteration); grids allocated once in the "
                     "kernels' layout and exposed as [1,32,Z,Y,X] views (scene.as_native_layout; --reference-grid-layout "
                     "times contiguous grids that are re-laid-out every step instead)",
            "l2": "flushed between timed steps (256 MiB write)", "sharding": "rays (weak): 1000 rays per GPU, "
            "gradients SUM all-reduced"}


# ---------------------------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------------------------
def run_gpu_arm(args):
    import torch
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device (there is no CPU fallback for the product path)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    from evennicer_slam_b200 import common, harness, sharding, functional, _lib
    from evennicer_slam_b200.functional import TIMER

    if hasattr(torch.autograd.graph, "set_warn_on_accumulate_grad_stream_mismatch"):
        # the camera tensors are created on the default stream and used on the per-keyframe streams (intended)
        torch.autograd.graph.set_warn_on_accumulate_grad_stream_mismatch(False)
    scene, frames = make_inputs()
    decoders, c, renderer, cfg = harness.build(scene, dev, native_layout=not args.reference_grid_layout)
    cam = scene.cam
    # config 4 sharded over the ranks -- measured first, before any collective has been captured into a CUDA graph
    # (eager NCCL calls issued after such a capture were seen to run several times slower)
    sharded = measure_sharded_frame(dev, renderer, decoders, c, frames, scene, world) if (world > 1 and not args.no_other_configs) else None
    depth_t = [torch.from_numpy(d).to(dev) for (_, d, _) in frames]
    color_t = [torch.from_numpy(col).to(dev) for (_, _, col) in frames]
    cams = [torch.from_numpy(ct.copy()).to(dev) for (ct, _, _) in frames]
    cam_params = [t.clone().requires_grad_(True) for t in cams[1:]]            # oldest frame fixed (Mapper.py:375)
    grids = {k: v.clone().requires_grad_(True) for k, v in c.items()}
    params = [p for p in decoders.parameters()]
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    torch.manual_seed(20 + rank)

    def zero_grads():
        for t in cam_params + list(grids.values()) + params:
            t.grad = None

    # The per-keyframe chains (camera tensor -> c2w -> pixel draw -> rays, and their backward into the camera tensors)
    # are independent of each other: each runs on its own stream, so the captured graph has five parallel branches
    # instead of ~35 tiny kernels in a row (autograd replays every backward node on its forward stream).
    side = [torch.cuda.Stream(dev) for _ in range(N_FRAMES)] if args.frame_streams else None

    def step():
        renderer._cache.invalidate()
        ros, rds, sds, scs = [None] * N_FRAMES, [None] * N_FRAMES, [None] * N_FRAMES, [None] * N_FRAMES
        cur = torch.cuda.current_stream(dev)
        for f in range(N_FRAMES):
            if side is not None:
                side[f].wait_stream(cur)
            with torch.cuda.stream(side[f] if side is not None else cur):
                ct = cams[0] if f == 0 else cam_params[f - 1]
                c2w = common.get_camera_from_tensor(ct)
                ro, rd, sd, sc_ = common.get_samples(0, cam.H, 0, cam.W, PIX_PER_FRAME, cam.H, cam.W, cam.fx, cam.fy,
                                                     cam.cx, cam.cy, c2w, depth_t[f], color_t[f], dev)
                ros[f], rds[f], sds[f], scs[f] = ro.float(), rd.float(), sd.float(), sc_.float()
        if side is not None:
            for f in range(N_FRAMES):
                cur.wait_stream(side[f])
                for t in (ros[f], rds[f], sds[f], scs[f]):
                    t.record_stream(cur)
        ro, rd, sd, sc_ = torch.cat(ros), torch.cat(rds), torch.cat(sds), torch.cat(scs)
        depth, unc, color = renderer.render_batch_ray(grids, decoders, rd, ro, dev, "color", gt_depth=sd)
        # Mapper.py:553-562: sum|gt-d|[gt>0] + 0.2*sum|gt_c-c| ; the mask is applied by where() instead of
        # boolean indexing (same sum, no host sync -> the step can be captured in a CUDA graph)
        loss = torch.where(sd > 0, torch.abs(sd - depth), 0.0).sum() + 0.2 * torch.abs(sc_ - color).sum()
        loss.backward()
        if world > 1:
            if sparse_ar is not None:
                # only the voxels some rank touched cross NVLink (sharding.SparseGradAllReduce)
                sparse_ar([grids[k].grad for k in ar_keys],
                          [grids[k].grad for k in grids if k not in ar_keys] + [t.grad for t in params + cam_params])
            else:
                sharding.allreduce_sum_([t.grad for t in list(grids.values()) + params + cam_params])
        return loss

    ar_keys = [k for k in grids if "coarse" not in k]
    sparse_ar = None
    if world > 1 and not args.reference_grid_layout and not args.dense_allreduce:
        sparse_ar = sharding.SparseGradAllReduce([grids[k] for k in ar_keys], capacity_frac=0.2)

    # ---- warm-up ----
    for _ in range(max(args.warmup, 3)):
        zero_grads(); step()
    torch.cuda.synchronize()
    run_step = step
    use_graph = not args.no_graph
    if use_graph:
        from evennicer_slam_b200.graph import GraphedStep
        zero_grads()                      # grads are then (re)allocated from the graph's private pool
        run_step = GraphedStep(step, warmup=2, device=dev, before_capture=zero_grads)
        for _ in range(3):
            run_step()
        torch.cuda.synchronize()

    # ---- timed region: K steps, CUDA events per step, L2 flushed between steps ----
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    TIMER.reset(); TIMER.enabled = not use_graph
    evs = []
    t_wall0 = time.perf_counter()
    for _ in range(args.steps):
        if not use_graph:
            zero_grads()
        flush.fill_(1)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); run_step(); b.record()
        evs.append((a, b))
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t_wall = time.perf_counter() - t_wall0
    TIMER.enabled = False
    dev_ms = sum(a.elapsed_time(b) for a, b in evs)
    eager_ms = None
    if use_graph:
        # per-kernel durations and the launch count come from an eager pass of the same K steps (CUDA events
        # cannot be queried inside a captured graph); the kernels and their inputs are identical
        TIMER.reset(); TIMER.enabled = True
        evs2 = []
        for _ in range(args.steps):
            zero_grads()
            flush.fill_(1)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); step(); b.record()
            evs2.append((a, b))
        torch.cuda.synchronize()
        TIMER.enabled = False
        eager_ms = sum(a.elapsed_time(b) for a, b in evs2) / args.steps
    launches_timed = TIMER.launches
    ksum = TIMER.summary()
    # ---- the drop-in path an UNCHANGED Mapper.py takes: eager launches, the callers' own lines (inside_mask pre-filter with
    #      boolean indexing, boolean-mask loss: Mapper.py:523-578), autograd plumbing in the C++ extension ----
    bound_dev = torch.from_numpy(scene.bound.copy()).to(dev)

    def step_caller():
        renderer._cache.invalidate()
        ros, rds, sds, scs = [], [], [], []
        for f in range(N_FRAMES):
            ct = cams[0] if f == 0 else cam_params[f - 1]
            c2w = common.get_camera_from_tensor(ct)
            ro, rd, sd, sc_ = common.get_samples(0, cam.H, 0, cam.W, PIX_PER_FRAME, cam.H, cam.W, cam.fx, cam.fy, cam.cx, cam.cy,
                                                 c2w, depth_t[f], color_t[f], dev)
            ros.append(ro.float()); rds.append(rd.float()); sds.append(sd.float()); scs.append(sc_.float())
        batch_rays_o, batch_rays_d, batch_gt_depth, batch_gt_color = torch.cat(ros), torch.cat(rds), torch.cat(sds), torch.cat(scs)
        with torch.no_grad():
            det_rays_o = batch_rays_o.clone().detach().unsqueeze(-1)
            det_rays_d = batch_rays_d.clone().detach().unsqueeze(-1)
            t = (bound_dev.unsqueeze(0) - det_rays_o) / det_rays_d
            t, _ = torch.min(torch.max(t, dim=2)[0], dim=1)
            inside_mask = t >= batch_gt_depth
        batch_rays_d = batch_rays_d[inside_mask]; batch_rays_o = batch_rays_o[inside_mask]
        batch_gt_depth = batch_gt_depth[inside_mask]; batch_gt_color = batch_gt_color[inside_mask]
        depth, unc, color = renderer.render_batch_ray(grids, decoders, batch_rays_d, batch_rays_o, dev, "color", gt_depth=batch_gt_depth)
        depth_mask = (batch_gt_depth > 0)
        loss = torch.abs(batch_gt_depth[depth_mask] - depth[depth_mask]).sum() + 0.2 * torch.abs(batch_gt_color - color).sum()
        loss.backward()
        return loss

    from evennicer_slam_b200 import _ext
    ext = _ext.module() if _ext.ENABLED else None
    for _ in range(3):
        zero_grads(); step_caller()
    torch.cuda.synchronize()
    if ext is not None:
        ext.reset_launches()
    l0 = TIMER.launches
    evs3 = []
    tw0 = time.perf_counter()
    for _ in range(args.steps):
        zero_grads()
        flush.fill_(1)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); step_caller(); b.record()
        evs3.append((a, b))
    torch.cuda.synchronize()
    caller_wall_ms = (time.perf_counter() - tw0) * 1e3 / args.steps
    caller_ms = sum(a.elapsed_time(b) for a, b in evs3) / args.steps
    caller_launches = (TIMER.launches - l0 + (ext.launches() if ext is not None else 0)) / args.steps
    # the timed region is ~K x 0.6 ms, shorter than one nvidia-smi sampling period: keep the same load (untimed runs of the
    # same step) going for ~0.5 s more so the clocks line holds several samples taken UNDER this load.  A fixed count on
    # EVERY rank: the step contains collectives when world > 1.
    for _ in range(600 if use_graph else 60):
        run_step()
    torch.cuda.synchronize()
    clocks = sampler.stop() if rank == 0 else None
    t = torch.tensor([dev_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_ms = float(t.item())
    ms_per_step = dev_ms / args.steps
    value = world * N_RAYS * args.steps / (dev_ms * 1e-3)

    # ---- e2e: C-ABI-level call with host buffers (pinned), copies inside the timed region ----
    e2e = measure_e2e(args, dev, renderer, decoders, grids, frames, scene, world)

    # ---- tracking iteration (config C2), reported beside the headline ----
    try:
        track_ms, track_graph_ms, track_fused = measure_tracking(dev, renderer, decoders, c, frames, scene, flush)
    except Exception as e:          # pragma: no cover - reported, never fatal for the headline
        track_ms, track_graph_ms, track_fused = None, None, {"error": repr(e)}
        sys.stderr.write(f"tracking measurement failed: {e!r}\n")
        for p in decoders.parameters():
            p.requires_grad_(True)
    other = None
    if rank == 0 and not args.no_other_configs:
        try:
            other = measure_other_configs(dev, renderer, decoders, c, frames, scene)
        except Exception as e:      # pragma: no cover - reported, never fatal for the headline
            other = {"error": repr(e)}
            sys.stderr.write(f"other configs failed: {e!r}\n")
    if other is not None and sharded is not None:
        other["full_frame_sharded"] = sharded

    if rank == 0:
        peak, peak_src = hbm_peak()
        n_b, t_b = ksum.get("render_bwd", (0, float("nan")))
        n_f, t_f = ksum.get("render_fwd", (0, float("nan")))
        achieved = BYTES_PER_POINT_BWD * N_RAYS * S_TOTAL / (t_b * 1e-3) / 1e9
        roof = {"bound": "hbm", "kernel": "ens_render_bwd: bwd_tc_wg_kernel<color> (tcgen05 data + weight gradients) with its placement / compositing-backward / unfold / ray-reduce launches" if functional.TC_MAP else "ens_render_bwd: render_bwd_mma_kernel<color, split> + wgrad_split_kernel<color>", "achieved": achieved, "peak": peak,
                "unit": "GB/s", "frac": achieved / peak, "traffic": load_traffic(),
                "peak_source": peak_src, "kernel_ms": t_b, "kernel_share_of_step": t_b / ms_per_step,
                "fwd_kernel_ms": t_f, "fwd_achieved_gbs": BYTES_PER_POINT_FWD * N_RAYS * S_TOTAL / (t_f * 1e-3) / 1e9,
                "step_algorithmic_gbs": BYTES_PER_RAY * N_RAYS / (ms_per_step * 1e-3) / 1e9,
                "step_frac": BYTES_PER_RAY * N_RAYS / (ms_per_step * 1e-3) / 1e9 / peak}
        cpu = cpu_baseline() if (world == 1 and not args.no_cpu_baseline) else None
        gpu_ref = None
        if world == 1 and not args.no_gpu_reference:
            try:
                gpu_ref = measure_gpu_reference(dev)
            except Exception as e:          # pragma: no cover - reported, never fatal for the headline
                gpu_ref = {"error": repr(e)}
        line = {
            "metric": "rays/sec (fwd+bwd) on the Replica mapping batch", "value": value, "unit": "rays/s",
            "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32 (f64 sample placement / depth sums)", "data": "synthetic",
            "config": workload_config(), "clocks": clocks, "e2e": e2e,
            "gpu_launches": launches_timed, "gpu_launches_per_step": launches_timed / args.steps,
            "roofline": roof, "cpu_baseline": cpu, "gpu_reference": gpu_ref,
            "tracking_ms_per_iter": track_ms, "tracking_ms_per_iter_graph": track_graph_ms,
            "tracking_ms_per_iter_fused_loss": track_fused, "other_configs": other, "wall_s_timed_region": t_wall,
            "allreduce": (None if world == 1 else ("touched voxels only (sharding.SparseGradAllReduce, capacity 20 % of the voxels; "
                          f"overflowed: {sparse_ar.overflowed()})" if sparse_ar is not None else "dense gradient arena")),
            "launch_mode": "cuda-graph replay of the whole step" if use_graph else "eager",
            "ms_per_step_eager": caller_ms,
            "eager": {"ms_per_step": caller_ms, "wall_ms_per_step": caller_wall_ms, "c_abi_calls_per_step": caller_launches,
                      "plumbing": "torch C++ extension (csrc/ens_torch.cpp)" if ext is not None else "python autograd.Function + ctypes",
                      "what": "the step as an unchanged Mapper.py runs it: eager launches, per-keyframe get_samples, inside_mask "
                              "pre-filter and boolean-mask loss (Mapper.py:523-578), backward; L2 flushed between steps",
                      "ms_per_step_python_plumbing": eager_ms},
        }
        emit(line)
    if world > 1:
        # Tear-down of a process group whose collectives were captured in a CUDA graph has been seen to hang in
        # destroy_process_group(); the measurement is complete and printed, so synchronise and leave hard.
        dist.barrier()
        torch.cuda.synchronize()
        sys.stdout.flush(); sys.stderr.flush()
        # release the captured graphs (they hold NCCL work) before the communicator goes away, then tear down properly; a
        # watchdog keeps a wedged destroy_process_group from hanging the launcher (the line is already printed)
        import gc
        import threading
        run_step = None
        gc.collect()
        torch.cuda.synchronize()
        threading.Timer(20.0, lambda: os._exit(0)).start()
        dist.destroy_process_group()
        os._exit(0)


def load_traffic():
    """dram bytes per launch of the dominant kernel from the committed ncu capture, if any."""
    path = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(path):
        try:
            with open(path) as f:
                t = json.load(f)
                from evennicer_slam_b200 import functional
                # the round-1 capture belongs to the mma.sync backward (ENS_MAP_TC=0), the round-2 one to the tcgen05 backward
                return t.get("render_bwd_dram_bytes_per_launch" if functional.TC_MAP else "r01_render_bwd_dram_bytes_per_launch")
        except Exception:
            return None
    return None


def measure_e2e(args, dev, renderer, decoders, grids, frames, scene, world):
    """Host-buffer call: pinned rays/depth/colour in, depth/var/colour + ray gradients + loss out."""
    import torch
    import render_oracle as orc
    import evennicer_slam_b200.synthetic as syn
    cam = scene.cam
    rng = np.random.RandomState(7)
    ros, rds, sds, scs = [], [], [], []
    for (cam_t, depth, color) in frames:
        idx = rng.randint(0, cam.H * cam.W, size=PIX_PER_FRAME)
        i, j, sd, scol = orc.select_pixels(idx, 0, cam.H, 0, cam.W, depth, color)
        ro, rd = orc.rays_from_uv(i, j, syn.quat_to_c2w(cam_t), cam.fx, cam.fy, cam.cx, cam.cy)
        ros.append(ro); rds.append(rd); sds.append(sd); scs.append(scol.astype(np.float32))
    h_in = [torch.from_numpy(np.concatenate(x)).pin_memory() for x in (ros, rds, sds, scs)]
    h2d = sum(t.numel() * t.element_size() for t in h_in)
    h_out = {"depth": torch.empty(N_RAYS, dtype=torch.float64).pin_memory(),
             "var": torch.empty(N_RAYS, dtype=torch.float64).pin_memory(),
             "color": torch.empty((N_RAYS, 3), dtype=torch.float32).pin_memory(),
             "g_ro": torch.empty((N_RAYS, 3), dtype=torch.float32).pin_memory(),
             "g_rd": torch.empty((N_RAYS, 3), dtype=torch.float32).pin_memory(),
             "loss": torch.empty(1, dtype=torch.float64).pin_memory()}
    d2h = sum(t.numel() * t.element_size() for t in h_out.values())
    params = list(decoders.parameters())
    # static device buffers: every step copies the pinned host batch into them, then replays the captured
    # render + loss + backward (same kernels as the eager call; the copies and the final sync are inside the timing)
    d_in = [torch.empty_like(t, device=dev) for t in h_in]
    d_ro = d_in[0].requires_grad_(True)
    d_rd = d_in[1].requires_grad_(True)
    outs = {}

    def compute():
        renderer._cache.invalidate()
        for t in list(grids.values()) + params + [d_ro, d_rd]:
            t.grad = None
        sd, sc_ = d_in[2], d_in[3]
        depth, unc, color = renderer.render_batch_ray(grids, decoders, d_rd, d_ro, dev, "color", gt_depth=sd)
        loss = torch.where(sd > 0, torch.abs(sd - depth), 0.0).sum() + 0.2 * torch.abs(sc_ - color).sum()
        loss.backward()
        if world > 1:      # sharded mapping: the replicated scene's gradients are summed over the ranks, as in the main step
            from evennicer_slam_b200 import sharding
            sharding.allreduce_sum_([t.grad for t in list(grids.values()) + params])
        outs.update(depth=depth.detach(), var=unc.detach(), color=color.detach(), g_ro=d_ro.grad, g_rd=d_rd.grad,
                    loss=loss.detach().reshape(1))
        return loss

    use_graph = not args.no_graph
    if use_graph:
        from evennicer_slam_b200.graph import GraphedStep
        for h, d in zip(h_in, d_in):
            d.detach().copy_(h, non_blocking=True)
        run = GraphedStep(compute, warmup=2, device=dev)
    else:
        run = compute

    def step():
        for h, d in zip(h_in, d_in):
            d.detach().copy_(h, non_blocking=True)
        run()
        for k, v in outs.items():
            h_out[k].copy_(v, non_blocking=True)
        torch.cuda.current_stream().synchronize()

    for _ in range(3):
        step()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    t = torch.tensor([dt], dtype=torch.float64, device=dev)
    if world > 1:
        import torch.distributed as dist
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return {"value": world * N_RAYS * args.steps / float(t.item()), "unit": "rays/s", "h2d_bytes_per_step": h2d,
            "d2h_bytes_per_step": d2h,
            "timing": "host wall clock around K synchronous steps: pinned H2D copies, "
                      + ("CUDA-graph replay of" if use_graph else "eager") +
                      " render_batch_ray + loss + backward" + (" + gradient all-reduce" if world > 1 else "") +
                      ", D2H copies, stream sync"}


def measure_tracking(dev, renderer, decoders, c, frames, scene, flush, iters=20):
    """Config C2: 200 px from the crop [100:580, 100:1100], colour stage, pose gradient only (Tracker.py:159-197)."""
    import torch
    from evennicer_slam_b200 import common
    cam = scene.cam
    cam_t, depth, color = frames[-1]
    depth_t = torch.from_numpy(depth).to(dev)
    color_t = torch.from_numpy(color).to(dev)
    ct = torch.from_numpy(cam_t.copy()).to(dev).requires_grad_(True)
    req = [p.requires_grad for p in decoders.parameters()]
    for p in decoders.parameters():
        p.requires_grad_(False)

    bound_dev = torch.from_numpy(scene.bound.copy()).to(dev)

    def rgbd_branch():
        # Tracker.py:159-197 as the unchanged caller runs it: pixel draw, inside_mask pre-filter, render, boolean-mask loss
        c2w = common.get_camera_from_tensor(ct)
        ro, rd, sd, sc_ = common.get_samples(100, cam.H - 100, 100, cam.W - 100, 200, cam.H, cam.W, cam.fx, cam.fy,
                                             cam.cx, cam.cy, c2w, depth_t, color_t, dev)
        with torch.no_grad():
            t = (bound_dev.unsqueeze(0) - ro.clone().detach().unsqueeze(-1)) / rd.clone().detach().unsqueeze(-1)
            t, _ = torch.min(torch.max(t, dim=2)[0], dim=1)
            inside_mask = t >= sd
        rd, ro, sd, sc_ = rd[inside_mask], ro[inside_mask], sd[inside_mask], sc_[inside_mask]
        d, u, col = renderer.render_batch_ray(c, decoders, rd, ro, dev, "color", gt_depth=sd)
        u = u.detach()
        tmp = torch.abs(sd - d) / torch.sqrt(u + 1e-10)
        mask = (tmp < 10 * tmp.median()) & (sd > 0)
        loss = (torch.abs(sd - d) / torch.sqrt(u + 1e-10))[mask].sum() + 0.5 * torch.abs(sc_ - col)[mask].sum()
        return c2w, loss

    def it():
        ct.grad = None
        _, loss = rgbd_branch()
        loss.backward()

    ev_w = torch.rand((int(cam.H * 0.15), int(cam.W * 0.15), 3), device=dev) - 0.5
    topt = torch.optim.Adam([ct], lr=1e-3)

    def it_full():
        # one WHOLE tracking iteration (Tracker.py:104-245) without the UNet: the event branch's down-scaled full-frame
        # render with gradient (18 360 rays, :150) -- the UNet and its blurred loss replaced by a fixed linear functional of
        # the rendered colour, so the render backward runs as in the real iteration -- plus the RGB-D branch and Adam.step
        topt.zero_grad()
        c2w, loss_rgbd = rgbd_branch()
        _, _, full_color = renderer.render_img_rescale(c, decoders, c2w, dev, "color", gt_depth=depth_t, scale_factor=0.15)
        loss_rgbd.backward(retain_graph=True)
        (full_color * ev_w).sum().backward()
        topt.step()

    def it_capturable():
        # the same iteration without host synchronisation: the boolean-mask selections become where() sums
        ct.grad = None
        c2w = common.get_camera_from_tensor(ct)
        ro, rd, sd, sc_ = common.get_samples(100, cam.H - 100, 100, cam.W - 100, 200, cam.H, cam.W, cam.fx, cam.fy,
                                             cam.cx, cam.cy, c2w, depth_t, color_t, dev)
        d, u, col = renderer.render_batch_ray(c, decoders, rd, ro, dev, "color", gt_depth=sd)
        u = u.detach()
        tmp = torch.abs(sd - d) / torch.sqrt(u + 1e-10)
        mask = (tmp < 10 * tmp.median()) & (sd > 0)
        loss = torch.where(mask, tmp, 0.0).sum() + 0.5 * torch.where(mask[:, None], torch.abs(sc_ - col), 0.0).sum()
        loss.backward()
        return loss

    def timed(fn):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        tot = 0.0
        for _ in range(iters):
            flush.fill_(1)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); fn(); b.record()
            torch.cuda.synchronize()
            tot += a.elapsed_time(b)
        return tot / iters

    def it_fused():
        # Tracker.py:180-196 through losses.tracker_loss (one launch for the loss and its gradients)
        from evennicer_slam_b200.losses import tracker_loss
        ct.grad = None
        c2w = common.get_camera_from_tensor(ct)
        ro, rd, sd, sc_ = common.get_samples(100, cam.H - 100, 100, cam.W - 100, 200, cam.H, cam.W, cam.fx, cam.fy,
                                             cam.cx, cam.cy, c2w, depth_t, color_t, dev)
        d, u, col = renderer.render_batch_ray(c, decoders, rd, ro, dev, "color", gt_depth=sd)
        loss = tracker_loss(sd, sc_, d, u, col, 0.5, True, True)
        loss.backward()
        return loss

    eager_ms = timed(it)
    graph_ms = None
    fused = {}
    try:
        fused["full_iteration_eager_ms"] = timed(it_full)
        fused["full_iteration_what"] = ("RGB-D branch (200 px) + event-branch render_img_rescale fwd+bwd (18 360 rays, pose "
                                        "gradient) + torch Adam.step on the camera tensor; UNet excluded")
    except Exception as e:      # pragma: no cover
        fused["full_iteration_error"] = repr(e)
    try:
        from evennicer_slam_b200.graph import GraphedStep
        graph_ms = timed(GraphedStep(it_capturable, warmup=2, device=dev))
        fused["eager_ms"] = timed(it_fused)
        fused["graph_ms"] = timed(GraphedStep(it_fused, warmup=2, device=dev))
    except Exception as e:      # pragma: no cover - reported, not fatal
        sys.stderr.write(f"tracking graph capture failed: {e!r}\n")
    for p, r in zip(decoders.parameters(), req):
        p.requires_grad_(r)
    return eager_ms, graph_ms, fused


def measure_sharded_frame(dev, renderer, decoders, c, frames, scene, world):
    """Config 4 on N GPUs (all ranks call this): the 816 000-ray frame in the reference's 100 000-ray batches, every
    batch split over the ranks (replicated scene, the two batch-global depth maxima MAX-reduced, outputs all-gathered,
    sharding.render_rays_sharded).  Also checks, on real NCCL, that one sharded batch equals the unsharded render
    bit for bit."""
    import torch
    import torch.distributed as dist
    from evennicer_slam_b200 import common, sharding
    cam = scene.cam
    cam_t, depth, color = frames[-1]
    depth_t = torch.from_numpy(depth).to(dev).reshape(-1)
    req = [p.requires_grad for p in decoders.parameters()]
    for p in decoders.parameters():
        p.requires_grad_(False)
    with torch.no_grad():
        c2w = common.get_camera_from_tensor(torch.from_numpy(cam_t.copy()).to(dev))
        ro, rd = common.get_rays(cam.H, cam.W, cam.fx, cam.fy, cam.cx, cam.cy, c2w, dev)
        ro, rd = ro.reshape(-1, 3), rd.reshape(-1, 3)
        B = renderer.ray_batch_size

        def frame():
            return sharding.render_frame_sharded(renderer, c, decoders, rd, ro, dev, "color", gt_depth=depth_t)
        outs = frame()
        # parity on real NCCL: two reference batches (the first and the ragged last) rendered unsharded
        same = True
        for i in (0, (ro.shape[0] // B) * B if ro.shape[0] % B else ro.shape[0] - B):
            ref = renderer.render_batch_ray(c, decoders, rd[i:i + B], ro[i:i + B], dev, "color", gt_depth=depth_t[i:i + B])
            same = same and all(bool(torch.equal(a[i:i + B], b)) for a, b in zip(outs, ref))
        one = sharding.render_rays_sharded(renderer, c, decoders, rd[:B], ro[:B], dev, "color", gt_depth=depth_t[:B])
        same = same and all(bool(torch.equal(a[:B], b)) for a, b in zip(outs, one))
        torch.cuda.synchronize(); dist.barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(2):
            frame()
        b.record(); torch.cuda.synchronize()
        t = torch.tensor([a.elapsed_time(b) / 2], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    for p, r in zip(decoders.parameters(), req):
        p.requires_grad_(r)
    ms = float(t.item())
    n = cam.H * cam.W
    res = {"rays": n, "n_gpus": world, "ms": ms, "rays_per_s": n / ms * 1e3, "equals_unsharded_bitwise": same}
    try:
        res["mesh_lattice_sharded"] = measure_sharded_mesh(dev, renderer, decoders, c, scene, world)
    except Exception as e:          # pragma: no cover - reported in place
        res["mesh_lattice_sharded"] = {"error": repr(e)}
    try:
        res["mapping_large_batch_sharded"] = measure_sharded_large_batch(dev, renderer, decoders, c, frames, scene, world)
    except Exception as e:          # pragma: no cover - reported in place
        res["mapping_large_batch_sharded"] = {"error": repr(e)}
    return res


def measure_sharded_large_batch(dev, renderer, decoders, c, frames, scene, world):
    """STRONG scaling of large-batch mapping (north_star: "large-batch mapping also shards rays across GPUs, replicating the
    grids and all-reducing their gradients"): the SAME 65 535-ray colour-stage batch on N GPUs, every rank renders
    and back-propagates its contiguous shard (the two depth maxima taken over the whole batch, so sample placement is that
    of the unsharded batch), the gradients are SUM all-reduced (touched voxels only).  All ranks call this."""
    import torch
    import torch.distributed as dist
    from evennicer_slam_b200 import sharding
    from evennicer_slam_b200.functional import depth_batch_max, render_batch_ray
    from evennicer_slam_b200.scene import as_native_layout
    n_big = 65536 // N_FRAMES * N_FRAMES
    ro, rd, sd, sc_ = _mapping_batch(dev, scene, frames, n_big, 32)
    rank = dist.get_rank()
    lo, hi = sharding.shard_range(n_big, rank, world)
    grids = {k: as_native_layout(v.detach().clone()).requires_grad_(True) for k, v in c.items()}
    params = [p for p in decoders.parameters()]
    keys = [k for k in grids if "coarse" not in k]
    ar = sharding.SparseGradAllReduce([grids[k] for k in keys], capacity_frac=0.6)
    dmax = depth_batch_max(sd.contiguous())
    setup = renderer._setup("color", decoders, dev)
    ro_l = ro[lo:hi].clone().requires_grad_(True); rd_l = rd[lo:hi].clone().requires_grad_(True)

    def step():
        for t in [ro_l, rd_l] + list(grids.values()) + params:
            t.grad = None
        depth, unc, color = render_batch_ray(setup, grids, decoders, rd_l, ro_l, sd[lo:hi], depth_max=dmax)
        loss = torch.where(sd[lo:hi] > 0, torch.abs(sd[lo:hi] - depth), 0.0).sum() + 0.2 * torch.abs(sc_[lo:hi] - color).sum()
        loss.backward()
        ar([grids[k].grad for k in keys], [p.grad for p in params])
    step(); step()
    torch.cuda.synchronize(); dist.barrier()
    a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(3):
        step()
    e.record(); torch.cuda.synchronize()
    t = torch.tensor([a.elapsed_time(e) / 3], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    return {"rays": n_big, "n_gpus": world, "ms": ms, "rays_per_s": n_big / ms * 1e3, "scaling": "strong",
            "allreduce_overflowed": ar.overflowed(),
            "frac_of_hbm_roofline_per_gpu": n_big / ms * 1e3 * BYTES_PER_RAY / (hbm_peak()[0] * 1e9) / world}


def measure_sharded_mesh(dev, renderer, decoders, c, scene, world):
    """Config 5 on N GPUs (all ranks call this): eval_points('fine') over the 256^3 marching-cubes lattice of
    Mesher.get_grid_uniform (Mesher.py:321-347), z-slabs of the lattice sharded over the ranks; every rank evaluates its
    slab in the reference's 500 000-point chunks and the occupancy column is assembled by one all-gather
    (sharding.eval_points_sharded).  Checked on real NCCL against the unsharded evaluation of two chunks, bit for bit."""
    import torch
    import torch.distributed as dist
    from evennicer_slam_b200 import sharding
    b = scene.bound
    lin = [torch.linspace(float(b[k, 0]), float(b[k, 1]), 256, device=dev, dtype=torch.float64) for k in range(3)]
    total, chunk = 256 ** 3, 4000000
    rank = dist.get_rank()
    lo_r, hi_r = sharding.shard_range(total, rank, world)

    def points(lo, hi):
        idx = torch.arange(lo, hi, device=dev)
        return torch.stack([lin[0][idx % 256], lin[1][(idx // 256) % 256], lin[2][idx // 65536]], -1)

    def lattice():
        outs = []
        with torch.no_grad():
            for lo in range(lo_r, hi_r, chunk):
                outs.append(renderer.eval_points(points(lo, min(hi_r, lo + chunk)), decoders, c, "fine", dev)[:, 3].contiguous())
            local = torch.cat(outs)
            per = [sharding.shard_range(total, q, world) for q in range(world)]
            pad = max(h - l for l, h in per)
            buf = torch.zeros(pad, dtype=torch.float32, device=dev)
            buf[:local.numel()] = local
            full = torch.empty(world * pad, dtype=torch.float32, device=dev)
            dist.all_gather_into_tensor(full, buf)
            return torch.cat([full[q * pad:q * pad + (h - l)] for q, (l, h) in enumerate(per)])
    occ = lattice()
    same = True
    with torch.no_grad():
        for lo in (0, total - 500000):
            ref = renderer.eval_points(points(lo, lo + 500000), decoders, c, "fine", dev)[:, 3]
            same = same and bool(torch.equal(occ[lo:lo + 500000], ref))
    torch.cuda.synchronize(); dist.barrier()
    a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(2):
        lattice()
    e.record(); torch.cuda.synchronize()
    t = torch.tensor([a.elapsed_time(e) / 2], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    return {"points": total, "n_gpus": world, "ms": ms, "points_per_s": total / ms * 1e3, "equals_unsharded_bitwise": same,
            "frac_of_hbm_roofline_per_gpu": total / ms * 1e3 * 2048 / (hbm_peak()[0] * 1e9) / world}


def measure_other_configs(dev, renderer, decoders, c, frames, scene):
    """Reported beside the headline (BASELINE.json configs 2, 4, 5): CUDA-event times, inputs resident.

      event_render   : Tracker.py:150 -- render_img_rescale(0.15) = 102 x 180 = 18 360 rays, colour stage, forward +
                       backward into the camera tensor only
      full_frame     : Renderer.render_img, 680 x 1200 = 816 000 rays in the reference's 100 000-ray batches, no grad
      mesh_lattice   : eval_points('fine') over 256^3 = 16 777 216 lattice points in 500 000-point chunks (Mesher.py)
    """
    import torch
    from evennicer_slam_b200 import common
    cam = scene.cam
    cam_t, depth, color = frames[-1]
    depth_t = torch.from_numpy(depth).to(dev)
    req = [p.requires_grad for p in decoders.parameters()]
    for p in decoders.parameters():
        p.requires_grad_(False)
    out = {}

    def timed(fn, reps):
        """median CUDA-event time of `reps` calls after two warm-up calls (lazy module loads, allocator growth)"""
        fn(); fn(); torch.cuda.synchronize()
        ts = []
        for _ in range(reps):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); fn(); b.record(); torch.cuda.synchronize()
            ts.append(a.elapsed_time(b))
        return float(np.median(ts))

    ct = torch.from_numpy(cam_t.copy()).to(dev).requires_grad_(True)

    def event_render():
        ct.grad = None
        c2w = common.get_camera_from_tensor(ct)
        d, u, col = renderer.render_img_rescale(c, decoders, c2w, dev, "color", gt_depth=depth_t, scale_factor=0.15)
        (col.sum() + d.sum()).backward()
    ms = timed(event_render, 5)
    n = int(cam.H * 0.15) * int(cam.W * 0.15)
    out["event_render"] = {"rays": n, "ms": ms, "rays_per_s": n / ms * 1e3,
                           "frac_of_hbm_roofline": n / ms * 1e3 * 294912 / (hbm_peak()[0] * 1e9)}

    c2w = common.get_camera_from_tensor(ct).detach()

    def full_frame():
        renderer.render_img(c, decoders, c2w, dev, "color", gt_depth=depth_t)
    ms = timed(full_frame, 3)
    n = cam.H * cam.W
    out["full_frame"] = {"rays": n, "ms": ms, "rays_per_s": n / ms * 1e3,
                         "frac_of_hbm_roofline": n / ms * 1e3 * 147456 / (hbm_peak()[0] * 1e9)}

    b = scene.bound
    lin = [torch.linspace(float(b[k, 0]), float(b[k, 1]), 256, device=dev, dtype=torch.float64) for k in range(3)]
    chunk = 500000

    def mesh():
        # the 256^3 lattice of Mesher.get_grid_uniform, generated chunk by chunk (x fastest, as np.meshgrid ravel)
        total = 256 ** 3
        for lo in range(0, total, chunk * 8):
            idx = torch.arange(lo, min(total, lo + chunk * 8), device=dev)
            pts = torch.stack([lin[0][idx % 256], lin[1][(idx // 256) % 256], lin[2][idx // 65536]], -1)
            renderer.eval_points(pts, decoders, c, "fine", dev)
    ms = timed(mesh, 3)
    n = 256 ** 3
    out["mesh_lattice"] = {"points": n, "ms": ms, "points_per_s": n / ms * 1e3,
                           "frac_of_hbm_roofline": n / ms * 1e3 * 2048 / (hbm_peak()[0] * 1e9)}
    for p, r in zip(decoders.parameters(), req):
        p.requires_grad_(r)
    # the side measurements must never cost the headline line: a failure is reported in place of the entry
    for name, fn in (("mapping_variants", lambda: measure_mapping_variants(dev, renderer, decoders, c, frames, scene)),
                     ("grid_adam", lambda: measure_grid_adam(dev)), ("event_loss", lambda: measure_event_loss(dev)),
                     ("mapper_ops", lambda: measure_mapper_ops(dev))):
        try:
            out.update(fn())
        except Exception as e:      # pragma: no cover
            out[name + "_error"] = repr(e)
            sys.stderr.write(f"{name} failed: {e!r}\n")
            for p, r in zip(decoders.parameters(), req):
                p.requires_grad_(r)
            torch.cuda.synchronize()
            torch.cuda.empty_cache()
    return out


def measure_event_loss(dev):
    """blurred-L2 event loss, value + gradient (Tracker.py:204-224; kernel 9, weight 1, balancer
    0.025: configs/Replica/replica.yaml:8-13) on the 102 x 180 x 2 event image -- the fused launch against the reference's
    torchvision lines run eagerly on the same GPU."""
    import torch
    from torchvision import transforms
    from evennicer_slam_b200.losses import event_loss
    gen = torch.Generator(device=dev); gen.manual_seed(3)
    gt = torch.poisson(torch.full((102, 180, 2), 0.3, device=dev), generator=gen)
    pred = (torch.rand((102, 180, 2), device=dev, generator=gen) * 1.5).requires_grad_(True)

    def fused():
        pred.grad = None
        loss, _ = event_loss(gt, pred, [9], [1.0], 0.025)
        loss.backward()

    def ref():
        pred.grad = None
        loss = ((gt - pred) ** 2).sum()
        g = transforms.functional.gaussian_blur(gt.permute(2, 0, 1), kernel_size=9).permute(1, 2, 0)
        p = transforms.functional.gaussian_blur(pred.permute(2, 0, 1), kernel_size=9).permute(1, 2, 0)
        loss = (loss + 1.0 * ((g - p) ** 2).sum()) * 0.025
        loss.backward()

    def timed(fn):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        ts = []
        for _ in range(20):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); fn(); b.record(); torch.cuda.synchronize()
            ts.append(a.elapsed_time(b))
        return float(np.median(ts))
    return {"event_loss_102x180": {"ms": timed(fused), "torchvision_reference_lines_ms": timed(ref),
                                   "note": "eager, loss + backward to the predicted event image"}}


def measure_mapper_ops(dev):
    """SURVEY.md 8(f) rank 3 and the rest of rank 2: frustum feature selection (Mapper.py:115-186) on room0's fine grid,
    keyframe overlap (Mapper.py:222-241) for 30 keyframes x 1600 points, UNet input assembly (event_net.py:74-87) at
    102 x 180 -- wall clock per call including the host side (4x4 inverse, one pose copy), eager."""
    import torch
    import frustum_cases as fc
    from evennicer_slam_b200.event_net import assemble_input
    from evennicer_slam_b200.mapper_ops import FrustumSelector

    def wall(fn, reps=20):
        fn(); torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(reps):
            fn()
        torch.cuda.synchronize()
        return (time.perf_counter() - t0) / reps * 1e3
    case = fc.mask_cases()["room0"]
    sel = FrustumSelector(*case["cam"], case["bound"], dev)
    depth = torch.from_numpy(case["depth"]).to(dev)
    c2w = torch.from_numpy(case["c2w"]).to(dev)
    shape = case["shapes"]["grid_fine"]
    out = {"frustum_mask_room0_fine": {"voxels": int(np.prod(shape)), "ms": wall(lambda: sel.voxel_mask(c2w, "grid_fine", shape, depth)),
                                       "note": "device bool [Z,Y,X] mask; the reference's numpy + cv2.remap code takes 22 ms for the same grid (oracle timing, tools/time_frustum.py)"}}
    oc = fc.overlap_cases()["rpg"]
    sel2 = FrustumSelector(*oc["cam"], fc.RPG_BOUND, dev)
    kf = [torch.from_numpy(c).to(dev) for c in oc["kf_c2w"]]
    pts = torch.rand(1600, 3, device=dev) * 4 - 2
    out["keyframe_overlap_30x1600"] = {"ms": wall(lambda: sel2.keyframe_overlap_counts(pts, kf))}
    a = torch.rand(102, 180, 3, device=dev, dtype=torch.float64)
    b = torch.rand(102, 180, 3, device=dev, requires_grad=True)

    def unet_in():
        b.grad = None
        assemble_input(a, b, 1.0).sum().backward()
    out["unet_input_102x180"] = {"ms": wall(unet_in), "note": "assemble + backward into the rendered image, eager"}
    return out


def measure_grid_adam(dev):
    """the fused frustum-masked Adam step (ens_grid_adam_step) on the middle/fine/colour grids,
    the voxels inside one camera frustum selected, against the reference's own sequence on the same GPU (boolean-mask index_put into the
    grid, torch.optim.Adam.step on the gathered copies, boolean-mask write-back: Mapper.py:451-458, 625, 633-641).
    HBM-bound: a selected voxel moves 7 lines of 128 B (value, gradient, two moments in; value, two moments out)
    plus its 4-byte index."""
    import torch
    import evennicer_slam_b200.synthetic as syn
    from evennicer_slam_b200 import scene as scn
    from evennicer_slam_b200.optim import FrustumGridAdam
    import cases
    peak = hbm_peak()[0] * 1e9
    out = {}
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    keys = ("grid_middle", "grid_fine", "grid_color")
    lrs = {k: 0.005 for k in keys}

    def timed(fn, reps=10):
        fn(); fn(); torch.cuda.synchronize()
        ts = []
        for _ in range(reps):
            # cold AND clean L2: the write flush alone would leave ~120 MB of dirty lines whose write-back lands in the
            # timed kernel's HBM traffic; a read pass over the same buffer replaces them with clean ones
            flush.fill_(1); flush.view(torch.int64).sum()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); fn(); b.record(); torch.cuda.synchronize()
            ts.append(a.elapsed_time(b))
        return float(np.median(ts))

    def frustum_mask(sc, shape):
        """voxel centres inside the view frustum of the scene's default pose (the shape of Mapper.get_mask_from_c2w's
        selection, without its depth test): spatially coherent, unlike a random mask"""
        b = sc.bound
        Z, Y, X = shape
        zs, ys, xs = torch.meshgrid(torch.linspace(float(b[2, 0]), float(b[2, 1]), Z, device=dev),
                                    torch.linspace(float(b[1, 0]), float(b[1, 1]), Y, device=dev),
                                    torch.linspace(float(b[0, 0]), float(b[0, 1]), X, device=dev), indexing="ij")
        pts = torch.stack([xs, ys, zs], -1).reshape(-1, 3)
        c2w = torch.from_numpy(syn.quat_to_c2w(syn.default_pose(np.asarray(b).tolist(), jitter_seed=10))).float().to(dev)
        pc = (pts - c2w[:3, 3]) @ c2w[:3, :3]                   # camera coordinates (x right, y up, z backward)
        cam = sc.cam
        z = -pc[:, 2]
        u = cam.fx * pc[:, 0] / z.clamp_min(1e-5) + cam.cx
        v = -cam.fy * pc[:, 1] / z.clamp_min(1e-5) + cam.cy
        m = (z > 0) & (u > 0) & (u < cam.W) & (v > 0) & (v < cam.H)
        m |= ((pts - c2w[:3, 3]) ** 2).sum(-1) < 0.25
        return m.reshape(Z, Y, X)

    for name, sc in (("room0", cases.room0_scene()),
                     ("rpg_recording4", syn.make_scene(syn.RPG4_BOUND, syn.RPG_CAM, seed=20, name="rpg4"))):
        gen = torch.Generator(device=dev); gen.manual_seed(5)
        c = {k: scn.as_native_layout(torch.from_numpy(sc.grids[k]).to(dev)).requires_grad_(True) for k in keys}
        masks = {k: frustum_mask(sc, tuple(c[k].shape[2:])) for k in keys}
        for k in keys:
            c[k].grad = scn.as_native_layout(torch.randn(c[k].shape, device=dev, generator=gen) * 1e-3)
        opt = FrustumGridAdam(c, masks)
        ms = timed(lambda: opt.step(lrs))
        n_sel = sum(int(masks[k].sum()) for k in keys)
        n_vox = sum(masks[k].numel() for k in keys)
        bytes_ = n_sel * (7 * 128 + 4)
        opt_all = FrustumGridAdam(c, None)                          # every voxel: the selection-independent upper case
        ms_all = timed(lambda: opt_all.step(lrs))
        out[f"grid_adam_{name}_all_voxels"] = {"voxels": n_vox, "selected": n_vox, "ms": ms_all,
                                               "gbs": n_vox * 7 * 128 / ms_all / 1e6,
                                               "frac_of_hbm_peak": n_vox * 7 * 128 / (ms_all * 1e-3) / peak}
        del opt_all
        # the reference sequence, eager torch on the same GPU
        cr = {k: c[k].detach().clone().contiguous() for k in keys}
        full = {k: masks[k][None, None].repeat(1, 32, 1, 1, 1) for k in keys}
        vg = {k: cr[k][full[k]].clone().requires_grad_(True) for k in keys}
        topt = torch.optim.Adam([{"params": [vg[k]], "lr": lrs[k]} for k in keys])
        dense = {k: c[k].grad.contiguous() for k in keys}

        def ref_seq():
            for k in keys:
                val = cr[k]; val[full[k]] = vg[k].detach()          # Mapper.py:455-457
                vg[k].grad = dense[k][full[k]]                      # what index_put's backward hands the copy
            topt.step()                                             # :625
            for k in keys:
                cr[k][full[k]] = vg[k].detach().clone()             # :637-640
        ms_ref = timed(ref_seq, 5)
        out[f"grid_adam_{name}"] = {"voxels": n_vox, "selected": n_sel, "ms": ms, "gbs": bytes_ / ms / 1e6,
                                    "frac_of_hbm_peak": bytes_ / (ms * 1e-3) / peak,
                                    "torch_reference_sequence_ms": ms_ref}
        del c, cr, full, vg, topt, dense, opt, masks
        torch.cuda.empty_cache()
    return out


def _mapping_batch(dev, scene, frames, n_rays, seed):
    """`n_rays` mapping rays drawn from the keyframes with the reference's pixel draw (common.get_samples), fixed."""
    import torch
    from evennicer_slam_b200 import common
    cam = scene.cam
    torch.manual_seed(seed)
    per = n_rays // len(frames)
    ros, rds, sds, scs = [], [], [], []
    for (cam_t, depth, color) in frames:
        c2w = common.get_camera_from_tensor(torch.from_numpy(cam_t.copy()).to(dev))
        ro, rd, sd, sc_ = common.get_samples(0, cam.H, 0, cam.W, per, cam.H, cam.W, cam.fx, cam.fy, cam.cx, cam.cy, c2w,
                                             torch.from_numpy(depth).to(dev), torch.from_numpy(color).to(dev), dev)
        ros.append(ro.float()); rds.append(rd.float()); sds.append(sd.float()); scs.append(sc_.float())
    return [torch.cat(x).detach() for x in (ros, rds, sds, scs)]


def measure_mapping_variants(dev, renderer, decoders, c, frames, scene):
    """SURVEY.md 8(d) C3 beside the headline: the mapping step per stage and as the reference's 60-iteration schedule
    (Mapper.py:462-467: 25 middle, 12 fine, 23 colour), a large batch (65 536 rays: the throughput regime, where the
    1000-ray step is latency-bound), and the same 1000-ray step on the RPG recording4 scene (197 MiB of grids > L2).
    Rays are fixed; each entry is render_batch_ray + Mapper loss + backward into grids, decoders and rays."""
    import torch
    import evennicer_slam_b200.synthetic as syn
    from evennicer_slam_b200 import harness
    from evennicer_slam_b200.graph import GraphedStep
    peak = hbm_peak()[0] * 1e9
    out = {}
    grids = {k: v.clone().requires_grad_(True) for k, v in c.items()}
    params = list(decoders.parameters())
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def make_step(rend, dec, g, batch, stage):
        ro, rd, sd, sc_ = batch
        ro = ro.clone().requires_grad_(True); rd = rd.clone().requires_grad_(True)

        def step():
            rend._cache.invalidate()
            depth, unc, color = rend.render_batch_ray(g, dec, rd, ro, dev, stage, gt_depth=sd)
            loss = torch.where(sd > 0, torch.abs(sd - depth), 0.0).sum()
            if stage == "color":
                loss = loss + 0.2 * torch.abs(sc_ - color).sum()
            loss.backward()
        return step, [ro, rd]

    def clear(ts):
        for t in ts:
            t.grad = None

    def timed(fn, reps):
        ts = []
        for _ in range(reps):
            flush.fill_(1)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); fn(); b.record(); torch.cuda.synchronize()
            ts.append(a.elapsed_time(b))
        return float(np.median(ts))

    # -- per stage, graph-replayed, 1000 rays -------------------------------------------------------------------
    batch = _mapping_batch(dev, scene, frames, N_RAYS, 31)
    lf = {"middle": 1, "fine": 2, "color": 3}
    replay = {}
    for stage in ("middle", "fine", "color"):
        step, leaves = make_step(renderer, decoders, grids, batch, stage)
        every = leaves + list(grids.values()) + params
        for _ in range(3):
            clear(every); step()
        torch.cuda.synchronize(); clear(every)
        replay[stage] = GraphedStep(step, warmup=2, device=dev, before_capture=lambda: clear(every))
        replay[stage](); replay[stage](); torch.cuda.synchronize()
        ms = timed(replay[stage], 15)
        bytes_per_ray = 3 * 1024 * lf[stage] * S_TOTAL
        out[f"mapping_stage_{stage}"] = {"rays": N_RAYS, "ms": ms, "rays_per_s": N_RAYS / ms * 1e3,
                                         "frac_of_hbm_roofline": N_RAYS / ms * 1e3 * bytes_per_ray / peak}
    # the colour-stage step with the mapper's loss as ONE launch (losses.mapper_loss)
    from evennicer_slam_b200.losses import mapper_loss
    ro_f = batch[0].clone().requires_grad_(True); rd_f = batch[1].clone().requires_grad_(True)

    def step_fused():
        renderer._cache.invalidate()
        depth, unc, color = renderer.render_batch_ray(grids, decoders, rd_f, ro_f, dev, "color", gt_depth=batch[2])
        mapper_loss(batch[2], batch[3], depth, color, 0.2, True).backward()
    every = [ro_f, rd_f] + list(grids.values()) + params
    for _ in range(3):
        clear(every); step_fused()
    torch.cuda.synchronize(); clear(every)
    gf = GraphedStep(step_fused, warmup=2, device=dev, before_capture=lambda: clear(every))
    gf(); gf(); torch.cuda.synchronize()
    ms = timed(gf, 15)
    out["mapping_stage_color_fused_loss"] = {"rays": N_RAYS, "ms": ms, "rays_per_s": N_RAYS / ms * 1e3,
                                             "frac_of_hbm_roofline": N_RAYS / ms * 1e3 * BYTES_PER_RAY / peak}
    del gf
    # -- a whole colour-stage mapping iteration INCLUDING the optimiser, as one replayed graph: render + fused loss +
    #    backward + FrustumGridAdam on the three grids (one launch, step / learning rates read from device memory) +
    #    FusedAdam on the 69 decoder tensors (one launch; the reference's optimizer.step(), Mapper.py:625)
    from evennicer_slam_b200.optim import FrustumGridAdam
    it_grids = {k: v.detach().clone().requires_grad_(True) for k, v in grids.items()}
    it_keys = [k for k in it_grids if "coarse" not in k]
    gopt = FrustumGridAdam({k: it_grids[k] for k in it_keys}, None, graph_safe=True)
    from evennicer_slam_b200.optim import FusedAdam
    dopt = FusedAdam(params, lr=0.005, graph_safe=True)
    dopt.set_dynamic(1)
    ro_i = batch[0].clone().requires_grad_(True); rd_i = batch[1].clone().requires_grad_(True)
    every_i = [ro_i, rd_i] + list(it_grids.values()) + params
    w0 = [p.detach().clone() for p in params]

    def iteration():
        # no cache invalidation needed: the optimisers bump the parameter versions, and under capture SceneCache packs
        # inside the graph
        depth, unc, color = renderer.render_batch_ray(it_grids, decoders, rd_i, ro_i, dev, "color", gt_depth=batch[2])
        mapper_loss(batch[2], batch[3], depth, color, 0.2, True).backward()
        gopt.step({})
        dopt.step()
    gopt.set_dynamic(1, {k: 0.005 for k in it_keys})
    for _ in range(3):
        clear(every_i); iteration()
    torch.cuda.synchronize(); clear(every_i)
    gi = GraphedStep(iteration, warmup=2, device=dev, before_capture=lambda: clear(every_i))
    gi(); gi(); torch.cuda.synchronize()
    ms = timed(gi, 15)
    out["mapping_iteration_with_optimizer"] = {"rays": N_RAYS, "ms": ms, "rays_per_s": N_RAYS / ms * 1e3,
                                               "note": "colour stage, fixed rays: render + mapper_loss + backward + fused grid "
                                                       "Adam (every voxel) + FusedAdam on the 69 decoder tensors; 0.905 ms "
                                                       "with torch.optim.Adam(capturable, foreach) on the decoders instead"}
    with torch.no_grad():                       # the decoders are shared with the rest of the bench: restore them
        for p, w in zip(params, w0):
            p.copy_(w)
    del gi, gopt, dopt, it_grids
    sched = [("middle", 25), ("fine", 12), ("color", 23)]
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for stage, k in sched:
        for _ in range(k):
            replay[stage]()
    b.record(); torch.cuda.synchronize()
    out["mapping_schedule_60_iters"] = {"iters": {s_: k for s_, k in sched}, "ms_total": a.elapsed_time(b),
                                        "note": "render + loss + backward only (no optimizer step), no L2 flush between iterations"}
    del replay

    # -- large batch, eager (launch overhead is negligible at this size) ------------------------------------------
    n_big = 65536 // N_FRAMES * N_FRAMES
    big = _mapping_batch(dev, scene, frames, n_big, 32)
    step, leaves = make_step(renderer, decoders, grids, big, "color")
    every = leaves + list(grids.values()) + params

    def big_step():
        clear(every); step()
    big_step(); big_step(); torch.cuda.synchronize()
    ms = timed(big_step, 5)
    out["mapping_large_batch"] = {"rays": n_big, "ms": ms, "rays_per_s": n_big / ms * 1e3,
                                  "frac_of_hbm_roofline": n_big / ms * 1e3 * BYTES_PER_RAY / peak}
    del big, step, leaves, every
    clear(list(grids.values()) + params)
    torch.cuda.empty_cache()

    # -- the mapper's event branch (Mapper.py:591): 102 x 180 rescaled frame rendered WITH gradients into grids,
    #    decoders and the camera tensor (the tracker's variant, pose only, is `event_render` above)
    from evennicer_slam_b200 import common
    cam_t, depth, _ = frames[-1]
    depth_t = torch.from_numpy(depth).to(dev)
    ct = torch.from_numpy(cam_t.copy()).to(dev).requires_grad_(True)
    every = [ct] + list(grids.values()) + params

    def event_map():
        clear(every)
        renderer._cache.invalidate()
        c2w = common.get_camera_from_tensor(ct)
        d, u, col = renderer.render_img_rescale(grids, decoders, c2w, dev, "color", gt_depth=depth_t, scale_factor=0.15)
        col.sum().backward()
    event_map(); event_map(); torch.cuda.synchronize()
    ms = timed(event_map, 5)
    n_ev = int(scene.cam.H * 0.15) * int(scene.cam.W * 0.15)
    out["mapping_event_render"] = {"rays": n_ev, "ms": ms, "rays_per_s": n_ev / ms * 1e3,
                                   "frac_of_hbm_roofline": n_ev / ms * 1e3 * BYTES_PER_RAY / peak}
    clear(every)

    # -- RPG recording4 scene: grids larger than L2 ---------------------------------------------------------------
    rscene = syn.make_scene(syn.RPG4_BOUND, syn.RPG_CAM, seed=20, name="rpg4", grid_std={"fine": 0.01})
    rdec, rc, rrend, _ = harness.build(rscene, dev, native_layout=True)
    rframes = []
    for f in range(N_FRAMES):
        cam_t = syn.default_pose(syn.RPG4_BOUND, jitter_seed=10 + f)
        depth, color, _ = syn.synthetic_frame(syn.RPG4_BOUND, syn.RPG_CAM, cam_t, seed=100 + f, zero_frac=0.02)
        rframes.append((cam_t, depth, color))
    rgrids = {k: v.clone().requires_grad_(True) for k, v in rc.items()}
    rbatch = _mapping_batch(dev, rscene, rframes, N_RAYS, 33)
    step, leaves = make_step(rrend, rdec, rgrids, rbatch, "color")
    every = leaves + list(rgrids.values()) + list(rdec.parameters())
    for _ in range(3):
        clear(every); step()
    torch.cuda.synchronize(); clear(every)
    g = GraphedStep(step, warmup=2, device=dev, before_capture=lambda: clear(every))
    g(); g(); torch.cuda.synchronize()
    ms = timed(g, 15)
    out["mapping_rpg_recording4"] = {"rays": N_RAYS, "grid_mib": sum(v.numel() * 4 for v in rc.values()) / 2 ** 20, "ms": ms,
                                     "rays_per_s": N_RAYS / ms * 1e3,
                                     "frac_of_hbm_roofline": N_RAYS / ms * 1e3 * BYTES_PER_RAY / peak}
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--dense-allreduce", action="store_true", help="N > 1: all-reduce the whole gradient arena instead of the touched voxels")
    ap.add_argument("--no-gpu-reference", action="store_true", help="skip the reference's eager-PyTorch renderer on this GPU")
    ap.add_argument("--no-graph", action="store_true", help="time eager launches instead of a CUDA-graph replay")
    ap.add_argument("--no-other-configs", action="store_true", help="skip the event-render / full-frame / mesh timings")
    ap.add_argument("--no-frame-streams", dest="frame_streams", action="store_false",
                    help="keep the five per-keyframe get_samples chains on one stream (default: five streams = parallel "
                         "branches of the captured graph)")
    ap.add_argument("--reference-grid-layout", action="store_true",
                    help="keep the grids as contiguous [1,32,Z,Y,X] tensors (re-laid-out every step) instead of "
                         "[1,32,Z,Y,X] views of native [Z,Y,X,32] storage (scene.as_native_layout)")
    args = ap.parse_args()
    # The contract is ONE JSON line on stdout.  Libraries (NCCL's version banner, torch warnings) also write to fd 1,
    # so for the duration of the run fd 1 points at stderr and the JSON line goes to the real stdout at the end.
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_gpu_arm(args)


if __name__ == "__main__":
    main()
