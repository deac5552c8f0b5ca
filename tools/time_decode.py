"""CUDA-event time of the forward-only tcgen05 decode: eval_points on a 128^3 / 256^3 lattice and a full-frame render."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, ROOT + '/tests/golden', ROOT + '/oracle', ROOT + '/tools']
import numpy as np, torch
import cases
from evennicer_slam_b200 import harness
dev = 'cuda:0'
scene = cases.room0_scene()
decoders, c, renderer, cfg = harness.build(scene, dev, requires_grad=False)
b = scene.bound
n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
ax = [torch.linspace(float(b[k, 0]) + 1e-3, float(b[k, 1]) - 1e-3, n, device=dev) for k in range(3)]
pts = torch.stack(torch.meshgrid(*ax, indexing='ij'), -1).reshape(-1, 3).contiguous()
def timed(fn, reps=5):
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); out = fn(); e.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(e))
    return float(np.median(ts)), out
with torch.no_grad():
    for stage in ('middle', 'fine', 'color'):
        ms, out = timed(lambda: renderer.eval_points(pts, decoders, c, stage, dev))
        print(f"eval_points {n}^3 stage {stage}: {ms:.3f} ms  checksum {float(out.double().sum()):.6f}")
