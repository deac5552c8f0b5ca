"""room0 1000-ray colour-stage case: every gradient tensor vs the reference golden, for both backward variants."""
import sys, os
ROOT=os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0]=[ROOT, ROOT+'/tests', ROOT+'/tests/golden', ROOT+'/oracle']
import numpy as np, torch
import cases
from util import load_golden, rel_err
from evennicer_slam_b200 import harness
DEV='cuda:0'
scene=cases.room0_scene()
decoders,c,renderer,cfg=harness.build(scene,DEV)
g=load_golden('room0_color_1000.npz')
res={}
for variant in ('mma','fma'):
    os.environ['ENS_BWD_VARIANT']=variant
    for p in decoders.parameters(): p.grad=None
    cg={k:v.clone().requires_grad_(True) for k,v in c.items()}
    ro=torch.from_numpy(g['rays_o']).to(DEV).requires_grad_(True)
    rd=torch.from_numpy(g['rays_d']).to(DEV).requires_grad_(True)
    sd=torch.from_numpy(g['sample_depth']).to(DEV)
    depth,var,color,raw,z,w=renderer.render_batch_ray_aux(cg,decoders,rd,ro,DEV,'color',gt_depth=sd)
    g_d,g_v,g_c=cases.upstream_grads(cases.N_ROOM0_RAYS)
    ((depth*torch.from_numpy(g_d).to(DEV)).sum()+(var*torch.from_numpy(g_v).to(DEV)).sum()+(color.double()*torch.from_numpy(g_c).double().to(DEV)).sum()).backward()
    out={'g_rays_o':ro.grad.cpu().numpy(),'g_rays_d':rd.grad.cpu().numpy()}
    for name in ('fine','color','middle'):
        for key,p in getattr(decoders,name+'_decoder').named_parameters():
            out[f'gdec.{name}.{key}']=p.grad.cpu().numpy()
    res[variant]=out
for k in res['mma']:
    ref=g[k]
    if np.abs(ref).max()==0: continue
    em,ef,emf=rel_err(res['mma'][k],ref),rel_err(res['fma'][k],ref),rel_err(res['mma'][k],res['fma'][k])
    print(f"{k:42s} mma-vs-ref {em:.1e}  fma-vs-ref {ef:.1e}  mma-vs-fma {emf:.1e}")
for k in ('g_rays_o','g_rays_d'):
    for v in ('mma','fma'):
        per=np.abs(res[v][k]-g[k]).max(1)/np.abs(g[k]).max()
        print(k,v,'rays >1e-3:',int((per>1e-3).sum()),'>1e-4:',int((per>1e-4).sum()),'max %.1e'%per.max())
