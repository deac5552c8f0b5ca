"""CUDA-event time of render fwd / bwd launches for a room0 colour batch: tcgen05 path (saved kind 3 / 2) vs mma.sync path."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, ROOT + '/tests/golden', ROOT + '/oracle', ROOT + '/tools']
import numpy as np, torch
import cases
from evennicer_slam_b200 import harness, functional
from util_prof import mapping_batch
dev = 'cuda:0'
nrays = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
scene = cases.room0_scene()
ref = {}
for wgrad, ggrid, tc in ((True, True, True), (True, False, True), (True, True, False), (True, False, False), (False, True, True), (False, False, True), (False, False, False)):
    functional.TC_MAP = tc
    os.environ['ENS_BWD_TC'] = '1' if tc else '0'
    decoders, c, renderer, cfg = harness.build(scene, dev, requires_grad=wgrad)
    ro, rd, sd, sc = mapping_batch(scene, nrays, dev)
    cg = {k: v.clone().requires_grad_(ggrid) for k, v in c.items()}
    for it in range(9):
        if it == 3:
            torch.cuda.synchronize(); functional.TIMER.reset(); functional.TIMER.enabled = True
        ro_ = ro.clone().requires_grad_(True); rd_ = rd.clone().requires_grad_(True)
        for p in decoders.parameters(): p.grad = None
        for v in cg.values(): v.grad = None
        d, u, col = renderer.render_batch_ray(cg, decoders, rd_, ro_, dev, 'color', gt_depth=sd)
        loss = torch.where(sd > 0, torch.abs(sd - d), 0.0).sum() + 0.2 * torch.abs(sc - col).sum()
        loss.backward()
    torch.cuda.synchronize()
    print('tc', tc, 'decoder grads', wgrad, 'grid grads', ggrid, {k: (n, round(ms, 4)) for k, (n, ms) in functional.TIMER.summary().items()}, flush=True)
    functional.TIMER.enabled = False
    key = (wgrad, ggrid)
    cur = {'ro': ro_.grad.clone(), 'rd': rd_.grad.clone()}
    if ggrid: cur.update({k: v.grad.clone() for k, v in cg.items() if v.grad is not None})
    if wgrad: cur.update({n: p.grad.clone() for n, p in decoders.named_parameters() if p.grad is not None})
    if key in ref:
        worst = max((float((cur[k] - ref[key][k]).abs().max() / ref[key][k].abs().max().clamp_min(1e-30)), k) for k in cur)
        print('   vs tc path: worst rel diff', worst)
    else:
        ref[key] = cur
