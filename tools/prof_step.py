"""Short driver for ncu: a few colour-stage mapping fwd+bwd passes (1000 rays, room0) with no bench scaffolding."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, ROOT + '/tests/golden', ROOT + '/oracle']
import numpy as np, torch
import cases
from evennicer_slam_b200 import harness
from util_prof import mapping_batch
dev = 'cuda:0'
nrays = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
wgrad = (sys.argv[2] != 'nowg') if len(sys.argv) > 2 else True
scene = cases.room0_scene()
decoders, c, renderer, cfg = harness.build(scene, dev, requires_grad=wgrad)
ro, rd, sd, sc = mapping_batch(scene, nrays, dev)
cg = {k: v.clone().requires_grad_(wgrad) for k, v in c.items()}
for it in range(3):
    ro_ = ro.clone().requires_grad_(True); rd_ = rd.clone().requires_grad_(True)
    d, u, col = renderer.render_batch_ray(cg, decoders, rd_, ro_, dev, 'color', gt_depth=sd)
    loss = torch.where(sd > 0, torch.abs(sd - d), 0.0).sum() + 0.2 * torch.abs(sc - col).sum()
    loss.backward()
torch.cuda.synchronize()
print('done', float(loss))
