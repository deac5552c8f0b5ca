import sys, os
ROOT=os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0]=[ROOT, ROOT+'/tests', ROOT+'/tests/golden', ROOT+'/oracle']
import numpy as np, torch
import cases, render_oracle as orc
from util import load_golden, rel_err
from evennicer_slam_b200 import harness, functional
functional._DEBUG['keep_workspace']=True
DEV='cuda:0'
scene=cases.tiny_scene()
decoders,c,renderer,cfg=harness.build(scene,DEV)
sc=orc.OracleScene.from_synthetic(scene)
g=load_golden('tiny_render.npz')
stage='coarse'; tag='coarse.n'
t32=torch.linspace(0.,1.,32,device=DEV).cpu().numpy(); t64=torch.linspace(0.,1.,16).double().numpy()
cg={k:v.clone().requires_grad_(True) for k,v in c.items()}
ro=torch.from_numpy(g[f'{tag}.rays_o']).to(DEV).requires_grad_(True)
rd=torch.from_numpy(g[f'{tag}.rays_d']).to(DEV).requires_grad_(True)
depth,var,color,raw,z,w=renderer.render_batch_ray_aux(cg,decoders,rd,ro,DEV,stage,gt_depth=None)
g_d,g_v,g_c=cases.upstream_grads(ro.shape[0])
loss=(depth*torch.from_numpy(g_d).to(DEV)).sum()+(var*torch.from_numpy(g_v).to(DEV)).sum()
loss.backward()
ws=functional._DEBUG['workspace'].cpu().numpy()
P=96*32
H=ws[:P*160].reshape(P,5,32)
od,ov,oc,cache=orc.render_batch_ray(sc,g[f'{tag}.rays_o'],g[f'{tag}.rays_d'],stage,None,t32,t64)
mc=cache['caches']['coarse']['mlp']
for i in range(5):
    x,u=mc['acts'][i]
    href=np.maximum(u,0)
    print('layer',i,'h err',np.abs(H[:,i]-href).max(),'mask mismatches',((H[:,i]>0)!=(u>0)).sum())
og=orc.render_batch_ray_backward(sc,cache,g_d,g_v,None)
db4_cuda=decoders.coarse_decoder.pts_linears[4].bias.grad.cpu().numpy()
db4_ref=og['decoders']['coarse']['pts_linears.4.bias']
print('db4 cuda',db4_cuda[:8]); print('db4 ref ',db4_ref[:8])
print('ratio', (db4_cuda/db4_ref)[:16])
gocc,_=orc.composite_backward(cache['comp'],g_d,g_v,None)
go=np.where(cache['mask'],gocc.reshape(-1),0)
Wo=sc.decoders['coarse']['output_linear.weight'][0]
db4_from_cudaH=((H[:,4]>0)*go[:,None]*Wo[None]).sum(0)
print('db4 from cuda masks+oracle g_occ', db4_from_cudaH[:8])
print('dbo cuda', decoders.coarse_decoder.output_linear.bias.grad.item(), 'ref', og['decoders']['coarse']['output_linear.bias'], 'sum|go|', np.abs(go).sum())
