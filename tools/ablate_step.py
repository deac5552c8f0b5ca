"""Where the graph-replayed mapping step's time goes: replays of the step with pieces removed (GPU box)."""
import os, sys, time, cProfile, pstats, io
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT]
import torch
import bench
from evennicer_slam_b200 import common, harness
from evennicer_slam_b200.graph import GraphedStep
dev = torch.device('cuda', 0)
torch.cuda.set_device(0)
scene, frames = bench.make_inputs()
decoders, c, renderer, cfg = harness.build(scene, dev, native_layout=True)
cam = scene.cam
NF, PIX = bench.N_FRAMES, bench.PIX_PER_FRAME
depth_t = [torch.from_numpy(d).to(dev) for (_, d, _) in frames]
color_t = [torch.from_numpy(col).to(dev) for (_, _, col) in frames]
cams = [torch.from_numpy(ct.copy()).to(dev) for (ct, _, _) in frames]
cam_params = [t.clone().requires_grad_(True) for t in cams[1:]]
grids = {k: v.clone().requires_grad_(True) for k, v in c.items()}
params = list(decoders.parameters())
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
side = [torch.cuda.Stream(dev) for _ in range(NF)]
torch.manual_seed(20)

def zero_grads():
    for t in cam_params + list(grids.values()) + params:
        t.grad = None

def draw():
    ros, rds, sds, scs = [None] * NF, [None] * NF, [None] * NF, [None] * NF
    cur = torch.cuda.current_stream(dev)
    for f in range(NF):
        side[f].wait_stream(cur)
        with torch.cuda.stream(side[f]):
            ct = cams[0] if f == 0 else cam_params[f - 1]
            c2w = common.get_camera_from_tensor(ct)
            ro, rd, sd, sc_ = common.get_samples(0, cam.H, 0, cam.W, PIX, cam.H, cam.W, cam.fx, cam.fy, cam.cx, cam.cy,
                                                 c2w, depth_t[f], color_t[f], dev)
            ros[f], rds[f], sds[f], scs[f] = ro.float(), rd.float(), sd.float(), sc_.float()
    for f in range(NF):
        cur.wait_stream(side[f])
        for t in (ros[f], rds[f], sds[f], scs[f]):
            t.record_stream(cur)
    return torch.cat(ros), torch.cat(rds), torch.cat(sds), torch.cat(scs)

fixed = [t.detach().clone() for t in draw()]
gd = torch.ones(NF * PIX, dtype=torch.float64, device=dev)
gc = torch.ones(NF * PIX, 3, dtype=torch.float32, device=dev)

def full():
    renderer._cache.invalidate()
    ro, rd, sd, sc_ = draw()
    depth, unc, color = renderer.render_batch_ray(grids, decoders, rd, ro, dev, "color", gt_depth=sd)
    loss = torch.where(sd > 0, torch.abs(sd - depth), 0.0).sum() + 0.2 * torch.abs(sc_ - color).sum()
    loss.backward()

def trivial_loss():
    renderer._cache.invalidate()
    ro, rd, sd, sc_ = draw()
    depth, unc, color = renderer.render_batch_ray(grids, decoders, rd, ro, dev, "color", gt_depth=sd)
    torch.autograd.backward([depth, color], [gd, gc])

def fixed_rays():
    renderer._cache.invalidate()
    ro, rd, sd, sc_ = fixed
    depth, unc, color = renderer.render_batch_ray(grids, decoders, rd, ro, dev, "color", gt_depth=sd)
    loss = torch.where(sd > 0, torch.abs(sd - depth), 0.0).sum() + 0.2 * torch.abs(sc_ - color).sum()
    loss.backward()

def kernels_only():
    renderer._cache.invalidate()
    ro, rd, sd, sc_ = fixed
    depth, unc, color = renderer.render_batch_ray(grids, decoders, rd, ro, dev, "color", gt_depth=sd)
    torch.autograd.backward([depth, color], [gd, gc])

def kernels_only_nopack():
    ro, rd, sd, sc_ = fixed
    depth, unc, color = renderer.render_batch_ray(grids, decoders, rd, ro, dev, "color", gt_depth=sd)
    torch.autograd.backward([depth, color], [gd, gc])

def draw_only():
    ro, rd, sd, sc_ = draw()
    (ro.sum() + rd.sum()).backward()

def timed(fn, n=30):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(n):
        flush.fill_(1)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    return ts[len(ts) // 2]

for name, fn in [("full", full), ("trivial_loss", trivial_loss), ("fixed_rays", fixed_rays), ("kernels_only", kernels_only),
                 ("kernels_only_nopack", kernels_only_nopack), ("draw_only", draw_only)]:
    for _ in range(3):
        zero_grads(); fn()
    torch.cuda.synchronize()
    zero_grads()
    g = GraphedStep(fn, warmup=2, device=dev, before_capture=zero_grads)
    print(f"{name:22s} graph {timed(g):.4f} ms", flush=True)

# eager: wall clock per step and a host profile
for _ in range(5):
    zero_grads(); full()
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(50):
    zero_grads(); full()
torch.cuda.synchronize()
print("eager full step wall ms", (time.perf_counter() - t0) / 50 * 1e3)
pr = cProfile.Profile(); pr.enable()
for _ in range(50):
    zero_grads(); full()
torch.cuda.synchronize()
pr.disable()
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(45); print(s.getvalue()[:9000])
