"""CUDA-event time of render fwd / bwd kernels for a 1000-ray room0 colour batch, with and without decoder/grid grads."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, ROOT + '/tests/golden', ROOT + '/oracle', ROOT + '/scratch']
import numpy as np, torch
import cases
from evennicer_slam_b200 import harness, functional
from util_prof import mapping_batch
dev = 'cuda:0'
nrays = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
scene = cases.room0_scene()
for wgrad, ggrid, split in ((True, True, '1'), (True, True, '0'), (False, True, '1')):
    os.environ['ENS_BWD_SPLIT'] = split
    decoders, c, renderer, cfg = harness.build(scene, dev, requires_grad=wgrad)
    ro, rd, sd, sc = mapping_batch(scene, nrays, dev)
    cg = {k: v.clone().requires_grad_(ggrid) for k, v in c.items()}
    for it in range(9):
        if it == 3:
            torch.cuda.synchronize(); functional.TIMER.reset(); functional.TIMER.enabled = True
        ro_ = ro.clone().requires_grad_(True); rd_ = rd.clone().requires_grad_(True)
        d, u, col = renderer.render_batch_ray(cg, decoders, rd_, ro_, dev, 'color', gt_depth=sd)
        loss = torch.where(sd > 0, torch.abs(sd - d), 0.0).sum() + 0.2 * torch.abs(sc - col).sum()
        loss.backward()
    torch.cuda.synchronize()
    print('split', split, 'decoder grads', wgrad, 'grid grads', ggrid, {k: (n, round(ms, 4)) for k, (n, ms) in functional.TIMER.summary().items()}, flush=True)
    functional.TIMER.enabled = False
