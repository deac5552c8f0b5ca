"""Per-kernel CUDA-event times of one tracking iteration (200 px, pose only), eager."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT]
import torch
import bench
from evennicer_slam_b200 import common, harness
from evennicer_slam_b200.functional import TIMER
from evennicer_slam_b200.losses import tracker_loss
dev = torch.device('cuda', 0)
scene, frames = bench.make_inputs()
decoders, c, renderer, cfg = harness.build(scene, dev, native_layout=True)
for p in decoders.parameters():
    p.requires_grad_(False)
cam = scene.cam
cam_t, depth, color = frames[-1]
depth_t = torch.from_numpy(depth).to(dev); color_t = torch.from_numpy(color).to(dev)
ct = torch.from_numpy(cam_t.copy()).to(dev).requires_grad_(True)
def it(n=200):
    ct.grad = None
    c2w = common.get_camera_from_tensor(ct)
    ro, rd, sd, sc_ = common.get_samples(100, cam.H - 100, 100, cam.W - 100, n, cam.H, cam.W, cam.fx, cam.fy, cam.cx, cam.cy, c2w, depth_t, color_t, dev)
    d, u, col = renderer.render_batch_ray(c, decoders, rd, ro, dev, "color", gt_depth=sd)
    tracker_loss(sd, sc_, d, u, col, 0.5, True, True).backward()
for n in (200, 1000):
    for _ in range(3): it(n)
    torch.cuda.synchronize()
    TIMER.reset(); TIMER.enabled = True
    for _ in range(20): it(n)
    torch.cuda.synchronize(); TIMER.enabled = False
    print(n, {k: (v[0], round(v[1], 4)) for k, v in TIMER.summary().items()})
