"""Host-side profile of the eager tracking iteration (where the 1.5 ms go when nothing is graph-captured)."""
import os, sys, time, cProfile, pstats, io
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT]
import torch
import bench
from evennicer_slam_b200 import common, harness
dev = torch.device('cuda', 0)
scene, frames = bench.make_inputs()
decoders, c, renderer, cfg = harness.build(scene, dev, native_layout=True)
for p in decoders.parameters():
    p.requires_grad_(False)
cam = scene.cam
cam_t, depth, color = frames[-1]
depth_t = torch.from_numpy(depth).to(dev); color_t = torch.from_numpy(color).to(dev)
ct = torch.from_numpy(cam_t.copy()).to(dev).requires_grad_(True)
def it():
    ct.grad = None
    c2w = common.get_camera_from_tensor(ct)
    ro, rd, sd, sc_ = common.get_samples(100, cam.H - 100, 100, cam.W - 100, 200, cam.H, cam.W, cam.fx, cam.fy, cam.cx, cam.cy, c2w, depth_t, color_t, dev)
    d, u, col = renderer.render_batch_ray(c, decoders, rd, ro, dev, "color", gt_depth=sd)
    u = u.detach()
    tmp = torch.abs(sd - d) / torch.sqrt(u + 1e-10)
    mask = (tmp < 10 * tmp.median()) & (sd > 0)
    loss = (torch.abs(sd - d) / torch.sqrt(u + 1e-10))[mask].sum() + 0.5 * torch.abs(sc_ - col)[mask].sum()
    loss.backward()
for _ in range(20): it()
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(200): it()
torch.cuda.synchronize()
print("eager tracking iteration wall ms", (time.perf_counter() - t0) / 200 * 1e3)
pr = cProfile.Profile(); pr.enable()
for _ in range(200): it()
torch.cuda.synchronize()
pr.disable()
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats("tottime").print_stats(28); print(s.getvalue()[:6000])
