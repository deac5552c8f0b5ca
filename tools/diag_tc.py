"""tcgen05 eval_points vs the mma.sync variant and the oracle; timing on a large point set."""
import sys, os, time
ROOT=os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0]=[ROOT, ROOT+'/tests', ROOT+'/tests/golden', ROOT+'/oracle']
import numpy as np, torch
import cases, render_oracle as orc
from util import load_golden, rel_err
from evennicer_slam_b200 import harness
DEV='cuda:0'
scene=cases.tiny_scene()
decoders,c,renderer,cfg=harness.build(scene,DEV,requires_grad=False)
g=load_golden('tiny_eval_points.npz')
pts=cases.eval_points_lattice(scene)
for stage in ('middle','fine','color'):
    for dt in (np.float64, np.float32):
        p=torch.from_numpy(pts.astype(dt)).to(DEV)
        os.environ['ENS_EVAL_VARIANT']='mma'; a=renderer.eval_points(p,decoders,c,stage,DEV).cpu().numpy()
        os.environ['ENS_EVAL_VARIANT']='tc'; b=renderer.eval_points(p,decoders,c,stage,DEV); torch.cuda.synchronize(); b=b.cpu().numpy()
        ref=g[f"{stage}.{'f64' if dt==np.float64 else 'f32'}"]
        print(stage, dt.__name__, 'n',len(pts),'tc-vs-mma %.2e'%rel_err(b,a),'tc-vs-ref %.2e'%rel_err(b,ref),'mma-vs-ref %.2e'%rel_err(a,ref),'mask eq',np.array_equal(b[:,3]==100,ref[:,3]==100), flush=True)
# timing, room0
scene=cases.room0_scene()
decoders,c,renderer,cfg=harness.build(scene,DEV,requires_grad=False)
b=scene.bound
N=int(sys.argv[1]) if len(sys.argv)>1 else 2_000_000
torch.manual_seed(0)
p=(torch.rand(N,3,device=DEV,dtype=torch.float64)*torch.tensor(b[:,1]-b[:,0],device=DEV)+torch.tensor(b[:,0],device=DEV)).float()
for stage in ('fine','color'):
    for var in ('mma','tc'):
        os.environ['ENS_EVAL_VARIANT']=var
        for _ in range(2): out=renderer.eval_points(p,decoders,c,stage,DEV)
        torch.cuda.synchronize()
        e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3): out=renderer.eval_points(p,decoders,c,stage,DEV)
        e1.record(); torch.cuda.synchronize()
        ms=e0.elapsed_time(e1)/3
        print(f"{stage} {var}: {ms:.2f} ms for {N} points -> {N/ms/1e3:.1f} Mpts/s", flush=True)
        if var=='mma': ref=out.clone()
        else: print('   tc-vs-mma on room0: %.2e'%rel_err(out.cpu().numpy(),ref.cpu().numpy()))
