import sys, os, time
ROOT=os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0]=[ROOT, ROOT+'/tests', ROOT+'/tests/golden', ROOT+'/oracle']
import numpy as np, torch
import cases, render_oracle as orc
import evennicer_slam_b200.synthetic as syn
N=int(sys.argv[1]); MARGIN=float(sys.argv[2])
scene=cases.tiny_scene(); cam=scene.cam
cam_t,depth,color,event=cases.tiny_frame()
sc=orc.OracleScene.from_synthetic(scene)
t32=torch.linspace(0.,1.,32).numpy(); t64=torch.linspace(0.,1.,16).double().numpy()
c2w=syn.quat_to_c2w(cam_t)
for stage in ('coarse','middle','fine','color'):
  for ud in (True,False):
    t0=time.time()
    for s in range(cases.SEED, cases.SEED+20000):
        torch.manual_seed(s)
        idx=torch.randint(cam.H*cam.W,(N,)).numpy()
        i,j,sd,scol=orc.select_pixels(idx,0,cam.H,0,cam.W,depth,color)
        ro,rd=orc.rays_from_uv(i,j,c2w,cam.fx,cam.fy,cam.cx,cam.cy)
        _,_,_,cache=orc.render_batch_ray(sc,ro,rd,stage,sd if ud else None,t32,t64)
        m=min(float(np.abs(u).min()) for dc in cache['caches'].values() for (x,u) in dc['mlp']['acts'])
        if m>MARGIN: break
    print(stage,ud,'seed',s,'margin %.2e'%m,'tries',s-cases.SEED+1,'%.1fs'%(time.time()-t0),flush=True)
