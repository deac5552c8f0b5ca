"""Short driver for ncu: fused masked Adam steps on RPG-recording4-sized grids (every voxel, then a frustum selection)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, ROOT + '/tests/golden']
import torch
import evennicer_slam_b200.synthetic as syn
from evennicer_slam_b200 import scene as scn
from evennicer_slam_b200.optim import FrustumGridAdam
dev = 'cuda:0'
sc = syn.make_scene(syn.RPG4_BOUND, syn.RPG_CAM, seed=20, name="rpg4")
keys = ("grid_middle", "grid_fine", "grid_color")
c = {k: scn.as_native_layout(torch.from_numpy(sc.grids[k]).to(dev)).requires_grad_(True) for k in keys}
for k in keys:
    c[k].grad = scn.as_native_layout(torch.randn(c[k].shape, device=dev) * 1e-3)
opt = FrustumGridAdam(c, None)
for _ in range(3):
    opt.step({k: 0.005 for k in keys})
torch.cuda.synchronize()
print('done')
