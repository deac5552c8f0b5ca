"""Frustum feature selection / keyframe overlap: GPU kernels vs. the CPU oracle (numpy restatement of Mapper.py:115-250)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, ROOT + '/tests/golden', ROOT + '/oracle']
import numpy as np, torch
import frustum_cases as fc, frustum_oracle as fo
from evennicer_slam_b200.mapper_ops import FrustumSelector
dev = 'cuda:0'
case = fc.mask_cases()['room0']
sel = FrustumSelector(*case['cam'], case['bound'], dev)
depth = torch.from_numpy(case['depth']).to(dev)
c2w = torch.from_numpy(case['c2w']).to(dev)
for key, shape in (('grid_middle', (21, 27, 36)), ('grid_fine', (42, 54, 73)), ('grid_fine_8cm', (85, 108, 147))):
    sel.voxel_mask(c2w, 'grid_fine', shape, depth); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(20): m = sel.voxel_mask(c2w, 'grid_fine', shape, depth)
    torch.cuda.synchronize(); gpu = (time.perf_counter() - t0) / 20
    t0 = time.perf_counter()
    mo = fo.get_mask_from_c2w(case['c2w'], shape, case['depth'], case['bound'], case['cam'], lambda a, b, n: torch.linspace(a, b, n).numpy())
    cpu = time.perf_counter() - t0
    same = bool(np.array_equal(m.cpu().numpy(), mo.transpose(2, 1, 0)))
    print(f"{key} {shape}: {int(np.prod(shape))} voxels  gpu {gpu*1e3:.3f} ms (incl. host 4x4 inverse + pose D2H)  cpu oracle {cpu*1e3:.1f} ms  identical {same}")
oc = fc.overlap_cases()['rpg']
sel2 = FrustumSelector(*oc['cam'], fc.RPG_BOUND, dev)
kf = [torch.from_numpy(c).to(dev) for c in oc['kf_c2w']] * 10          # 300 keyframes
pts = torch.rand(1600, 3, device=dev) * 4 - 2
sel2.keyframe_overlap_counts(pts, kf); torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(20): cnt = sel2.keyframe_overlap_counts(pts, kf)
gpu = (time.perf_counter() - t0) / 20
t0 = time.perf_counter()
w2cs = np.stack([np.linalg.inv(c.cpu().numpy()) for c in kf])
cnt_o = fo.keyframe_overlap(pts.cpu().numpy(), w2cs, oc['cam'])
cpu = time.perf_counter() - t0
print(f"keyframe overlap, 300 keyframes x 1600 points: gpu {gpu*1e3:.3f} ms  cpu oracle {cpu*1e3:.1f} ms  identical {bool(np.array_equal(cnt, cnt_o))}")
