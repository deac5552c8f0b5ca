import sys, os
ROOT=os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0]=[ROOT, ROOT+'/tests', ROOT+'/tests/golden', ROOT+'/oracle']
import numpy as np, torch
import cases, render_oracle as orc
from util import load_golden
g=load_golden('tiny_render.npz')
scene=cases.tiny_scene()
sc=orc.OracleScene.from_synthetic(scene)
t32=torch.linspace(0.,1.,32).numpy(); t64=torch.linspace(0.,1.,16).double().numpy()
for stage in ('coarse','middle','fine','color'):
  for ud in (True,False):
    tag=f"{stage}.{'d' if ud else 'n'}"
    _,_,_,cache=orc.render_batch_ray(sc,g[f'{tag}.rays_o'],g[f'{tag}.rays_d'],stage,g[f'{tag}.sample_depth'] if ud else None,t32,t64)
    for name,dc in cache['caches'].items():
        us=np.concatenate([np.abs(u).ravel() for (x,u) in dc['mlp']['acts']])
        print(tag,name,'n=',us.size,'min=%.2e'%us.min(),'<1e-6:',int((us<1e-6).sum()),'<3e-6:',int((us<3e-6).sum()),'<1e-5:',int((us<1e-5).sum()),'<1e-4:',int((us<1e-4).sum()))
