"""cProfile of the eager drop-in mapping step (the callers' own lines on the drop-ins): where the host time goes."""
import cProfile, os, pstats, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, ROOT + '/tests/golden', ROOT + '/oracle', ROOT + '/tools']
import numpy as np, torch
import bench
from evennicer_slam_b200 import common, harness
dev = torch.device('cuda', 0)
scene, frames = bench.make_inputs()
decoders, c, renderer, cfg = harness.build(scene, dev, native_layout=True)
cam = scene.cam
depth_t = [torch.from_numpy(d).to(dev) for (_, d, _) in frames]
color_t = [torch.from_numpy(col).to(dev) for (_, _, col) in frames]
cams = [torch.from_numpy(ct.copy()).to(dev) for (ct, _, _) in frames]
cam_params = [t.clone().requires_grad_(True) for t in cams[1:]]
grids = {k: v.clone().requires_grad_(True) for k, v in c.items()}
params = list(decoders.parameters())
bound_dev = torch.from_numpy(scene.bound.copy()).to(dev)

def step():
    for t in cam_params + list(grids.values()) + params:
        t.grad = None
    renderer._cache.invalidate()
    ros, rds, sds, scs = [], [], [], []
    for f in range(5):
        ct = cams[0] if f == 0 else cam_params[f - 1]
        c2w = common.get_camera_from_tensor(ct)
        ro, rd, sd, sc_ = common.get_samples(0, cam.H, 0, cam.W, 200, cam.H, cam.W, cam.fx, cam.fy, cam.cx, cam.cy, c2w, depth_t[f], color_t[f], dev)
        ros.append(ro.float()); rds.append(rd.float()); sds.append(sd.float()); scs.append(sc_.float())
    ro, rd, sd, sc_ = torch.cat(ros), torch.cat(rds), torch.cat(sds), torch.cat(scs)
    with torch.no_grad():
        t = (bound_dev.unsqueeze(0) - ro.clone().detach().unsqueeze(-1)) / rd.clone().detach().unsqueeze(-1)
        t, _ = torch.min(torch.max(t, dim=2)[0], dim=1)
        inside = t >= sd
    rd, ro, sd, sc_ = rd[inside], ro[inside], sd[inside], sc_[inside]
    d, u, col = renderer.render_batch_ray(grids, decoders, rd, ro, dev, 'color', gt_depth=sd)
    m = sd > 0
    loss = torch.abs(sd[m] - d[m]).sum() + 0.2 * torch.abs(sc_ - col).sum()
    loss.backward()

for _ in range(5): step()
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(50): step()
torch.cuda.synchronize()
print('wall ms per step', (time.perf_counter() - t0) * 20)
pr = cProfile.Profile(); pr.enable()
for _ in range(50): step()
torch.cuda.synchronize()
pr.disable()
pstats.Stats(pr).sort_stats('tottime').print_stats(22)
