"""In-kernel timestamps of the tcgen05 backward (bwd_tc_wg_kernel): second tile of CTA 0 of every role, E thread 0 and H thread 128."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, ROOT + '/tests/golden', ROOT + '/oracle', ROOT + '/tools']
import numpy as np, torch
import cases
from evennicer_slam_b200 import harness
from util_prof import mapping_batch
dev = 'cuda:0'
nrays = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
scene = cases.room0_scene()
decoders, c, renderer, cfg = harness.build(scene, dev, requires_grad=True)
ro, rd, sd, sc = mapping_batch(scene, nrays, dev)
GRID = 'nogrid' not in sys.argv; RAYS = 'norays' not in sys.argv
cg = {k: v.clone().requires_grad_(GRID) for k, v in c.items()}
dbg = torch.zeros(4 * 10 * 2 * 8, dtype=torch.int64, device=dev)
for it in range(4):
    if it == 3: os.environ['ENS_BWD_TC_DBG'] = str(dbg.data_ptr())
    ro_ = ro.clone().requires_grad_(RAYS); rd_ = rd.clone().requires_grad_(RAYS)
    d, u, col = renderer.render_batch_ray(cg, decoders, rd_, ro_, dev, 'color', gt_depth=sd)
    loss = torch.where(sd > 0, torch.abs(sd - d), 0.0).sum() + 0.2 * torch.abs(sc - col).sum()
    loss.backward()
torch.cuda.synchronize()
t = dbg.cpu().numpy().reshape(4, 10, 2, 8)
for role in range(4):
    base = t[role, 8, 0, 0]
    if base == 0: continue
    print('role', role, '(cycles relative to the tile start of the E thread)')
    for who in (0, 1):
        print('  ', 'E' if who == 0 else 'H', 'tile start/loop start/loop end/tail end:', [int(x - base) if x else None for x in t[role, 8, who, :4]])
        print('     tail stamps', [int(x - base) if x else None for x in t[role, 9, who, :4]])
        for ps in range(8):
            row = t[role, ps, who]
            if row[0] == 0: continue
            print('     pass', ps, [int(x - base) if x else None for x in row[:7]])
