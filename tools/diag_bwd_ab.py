"""A/B: tensor-core backward vs FFMA backward on the same inputs; locate the largest differences."""
import sys, os
ROOT=os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0]=[ROOT, ROOT+'/tests', ROOT+'/tests/golden', ROOT+'/oracle']
import numpy as np, torch
import cases
from util import load_golden, rel_err
from evennicer_slam_b200 import harness
DEV='cuda:0'
scene=cases.tiny_scene()
decoders,c,renderer,cfg=harness.build(scene,DEV)
g=load_golden('tiny_render.npz')
def run(tag, stage, use_depth, variant):
    os.environ['ENS_BWD_VARIANT']=variant
    for p in decoders.parameters(): p.grad=None; p.requires_grad_(True)
    cg={k:v.clone().requires_grad_(True) for k,v in c.items()}
    ro=torch.from_numpy(g[f'{tag}.rays_o']).to(DEV).requires_grad_(True)
    rd=torch.from_numpy(g[f'{tag}.rays_d']).to(DEV).requires_grad_(True)
    sd=torch.from_numpy(g[f'{tag}.sample_depth']).to(DEV)
    depth,var,color,raw,z,w=renderer.render_batch_ray_aux(cg,decoders,rd,ro,DEV,stage,gt_depth=sd if use_depth else None)
    g_d,g_v,g_c=cases.upstream_grads(ro.shape[0])
    loss=(depth*torch.from_numpy(g_d).to(DEV)).sum()+(var*torch.from_numpy(g_v).to(DEV)).sum()+(color.double()*torch.from_numpy(g_c).double().to(DEV)).sum()
    loss.backward()
    out={'ro':ro.grad.cpu().numpy(),'rd':rd.grad.cpu().numpy()}
    for k,v in cg.items():
        if v.grad is not None: out[k]=v.grad.cpu().numpy()
    for n,p in decoders.named_parameters():
        if p.grad is not None: out[n]=p.grad.cpu().numpy()
    return out, z.cpu().numpy()
for stage in (sys.argv[1:] or ['fine']):
  for use_depth in (False, True):
    tag=f"{stage}.{'d' if use_depth else 'n'}"
    a,z=run(tag,stage,use_depth,'mma'); b,_=run(tag,stage,use_depth,'fma')
    print('=====',tag, 'R,S=',z.shape)
    for k in a:
        e=rel_err(a[k],b[k])
        flag = ' <<<<' if e>1e-4 else ''
        print(f"{k:45s} {e:.2e} max|ref|={np.abs(b[k]).max():.3e}{flag}")
        if e>1e-4:
            d=np.abs(a[k]-b[k]); idx=np.unravel_index(np.argsort(d.ravel())[-5:], d.shape)
            for q in zip(*idx):
                print('      at',q,'mma',a[k][q],'fma',b[k][q])
            print('      n(|diff|>1e-4*max)=', int((d>1e-4*np.abs(b[k]).max()).sum()), 'of', d.size)
