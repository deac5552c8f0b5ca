import numpy as np, torch
import evennicer_slam_b200.synthetic as syn
import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'oracle'))
import render_oracle as orc

def mapping_batch(scene, nrays, dev, seed=7):
    cam = scene.cam
    rng = np.random.RandomState(seed)
    nf = 5; per = nrays // nf
    ros, rds, sds, scs = [], [], [], []
    for f in range(nf):
        cam_t = syn.default_pose(syn.ROOM0_BOUND, jitter_seed=10 + f)
        depth, color, _ = syn.synthetic_frame(syn.ROOM0_BOUND, cam, cam_t, seed=100 + f, zero_frac=0.02)
        idx = rng.randint(0, cam.H * cam.W, size=per)
        i, j, sd, scol = orc.select_pixels(idx, 0, cam.H, 0, cam.W, depth, color)
        ro, rd = orc.rays_from_uv(i, j, syn.quat_to_c2w(cam_t), cam.fx, cam.fy, cam.cx, cam.cy)
        ros.append(ro); rds.append(rd); sds.append(sd); scs.append(scol.astype(np.float32))
    t = lambda xs: torch.from_numpy(np.concatenate(xs)).to(dev)
    return t(ros), t(rds), t(sds), t(scs)
