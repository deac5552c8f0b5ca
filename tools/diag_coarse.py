import sys, os
ROOT=os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0]=[ROOT, ROOT+'/tests', ROOT+'/tests/golden', ROOT+'/oracle']
import numpy as np, torch
import cases, render_oracle as orc
from util import load_golden, rel_err
from evennicer_slam_b200 import harness
DEV='cuda:0'
scene=cases.tiny_scene()
decoders,c,renderer,cfg=harness.build(scene,DEV)
sc=orc.OracleScene.from_synthetic(scene)
g=load_golden('tiny_render.npz')
stage=sys.argv[1] if len(sys.argv)>1 else 'coarse'
tag=f'{stage}.n'
t32=torch.linspace(0.,1.,32,device=DEV).cpu().numpy(); t64=torch.linspace(0.,1.,16).double().numpy()
for mode in ('d','v','all'):
    for p in decoders.parameters(): p.grad=None; p.requires_grad_(True)
    cg={k:v.clone().requires_grad_(True) for k,v in c.items()}
    ro=torch.from_numpy(g[f'{tag}.rays_o']).to(DEV).requires_grad_(True)
    rd=torch.from_numpy(g[f'{tag}.rays_d']).to(DEV).requires_grad_(True)
    depth,var,color,raw,z,w=renderer.render_batch_ray_aux(cg,decoders,rd,ro,DEV,stage,gt_depth=None)
    g_d,g_v,g_c=cases.upstream_grads(ro.shape[0])
    if mode=='d': g_v=g_v*0; g_c=g_c*0
    if mode=='v': g_d=g_d*0; g_c=g_c*0
    loss=(depth*torch.from_numpy(g_d).to(DEV)).sum()+(var*torch.from_numpy(g_v).to(DEV)).sum()+(color.double()*torch.from_numpy(g_c).double().to(DEV)).sum()
    loss.backward()
    od,ov,oc,cache=orc.render_batch_ray(sc,g[f'{tag}.rays_o'],g[f'{tag}.rays_d'],stage,None,t32,t64)
    og=orc.render_batch_ray_backward(sc,cache,g_d,g_v,g_c)
    print('mode',mode,'weights', rel_err(w.cpu().numpy(), cache['weights']))
    for name in orc.STAGE_DECODERS[stage]:
        print('  grid', rel_err(cg['grid_'+name].grad.cpu().numpy(), og['grids']['grid_'+name]), 'g_ro', rel_err(ro.grad.cpu().numpy(), og['rays_o']))
        for key,p in getattr(decoders,name+'_decoder').named_parameters():
            print('   ',key, f"{rel_err(p.grad.cpu().numpy(), og['decoders'][name][key]):.2e}")
