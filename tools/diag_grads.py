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
g=load_golden('tiny_render.npz')
stages=sys.argv[1:] or ['coarse','middle','fine','color']
for stage in stages:
  for use_depth in (True,False):
    for wp in (True, False):
        tag=f"{stage}.{'d' if use_depth else 'n'}"
        for p in decoders.parameters(): p.grad=None; p.requires_grad_(wp)
        cg={k:v.clone().requires_grad_(True) for k,v in c.items()}
        ro=torch.from_numpy(g[f'{tag}.rays_o']).to(DEV).requires_grad_(True)
        rd=torch.from_numpy(g[f'{tag}.rays_d']).to(DEV).requires_grad_(True)
        sd=torch.from_numpy(g[f'{tag}.sample_depth']).to(DEV)
        depth,var,color,raw,z,w=renderer.render_batch_ray_aux(cg,decoders,rd,ro,DEV,stage,gt_depth=sd if use_depth else None)
        g_d,g_v,g_c=cases.upstream_grads(ro.shape[0])
        loss=(depth*torch.from_numpy(g_d).to(DEV)).sum()+(var*torch.from_numpy(g_v).to(DEV)).sum()+(color.double()*torch.from_numpy(g_c).double().to(DEV)).sum()
        loss.backward()
        msg=[f"{tag} wp={int(wp)} z_exact={np.array_equal(z.cpu().numpy(),g[f'{tag}.z_vals'])} raw={rel_err(raw.cpu().numpy(),g[f'{tag}.raw']):.1e} depth={rel_err(depth.detach().cpu().numpy(),g[f'{tag}.depth']):.1e} var={rel_err(var.detach().cpu().numpy(),g[f'{tag}.var']):.1e}"]
        msg.append(f"g_ro={rel_err(ro.grad.cpu().numpy(),g[f'{tag}.g_rays_o']):.1e} g_rd={rel_err(rd.grad.cpu().numpy(),g[f'{tag}.g_rays_d']):.1e}")
        for name in orc.STAGE_DECODERS[stage]:
            gk='grid_'+name
            msg.append(f"{gk}={rel_err(cg[gk].grad.cpu().numpy(),g[f'{tag}.ggrid.{gk}']):.1e}")
            if wp:
                worst=0;wk=None
                for key,p in getattr(decoders,name+'_decoder').named_parameters():
                    ref=g[f'{tag}.gdec.{name}.{key}']
                    if np.abs(ref).max()>0:
                        e=rel_err(p.grad.cpu().numpy(),ref)
                        if e>worst: worst,wk=e,key
                msg.append(f"dec[{name}] worst={worst:.1e}@{wk}")
        print(' | '.join(msg))
