"""Per-phase timing of a sharded 100k-ray batch (torchrun, N ranks)."""
import os, sys, time
ROOT=os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0]=[ROOT, ROOT+'/tests/golden', ROOT+'/oracle', ROOT]
import numpy as np, torch, torch.distributed as dist
import cases
from evennicer_slam_b200 import harness, common, sharding, functional
lr=int(os.environ.get('LOCAL_RANK','0')); torch.cuda.set_device(lr); dev=torch.device('cuda',lr)
dist.init_process_group('nccl', device_id=dev)
rank, world = dist.get_rank(), dist.get_world_size()
import bench
scene, frames = bench.make_inputs()
decoders,c,renderer,cfg=harness.build(scene,dev,requires_grad=False,native_layout=True)
cam=scene.cam; cam_t,depth,color=frames[-1]
depth_t=torch.from_numpy(depth).to(dev).reshape(-1)
with torch.no_grad():
    c2w=common.get_camera_from_tensor(torch.from_numpy(cam_t.copy()).to(dev))
    ro,rd=common.get_rays(cam.H,cam.W,cam.fx,cam.fy,cam.cx,cam.cy,c2w,dev); ro,rd=ro.reshape(-1,3),rd.reshape(-1,3)
    B=100000
    def ev(): e=torch.cuda.Event(enable_timing=True); e.record(); return e
    for rep in range(3):
        lo,hi=sharding.shard_range(B,rank,world)
        t=[ev()]
        gd=depth_t[:B].float()
        dmax=functional.depth_batch_max(gd[lo:hi].contiguous()); t.append(ev())
        dmax=sharding.global_depth_max(dmax); t.append(ev())
        setup=renderer._setup('color',decoders,dev)
        d,v,col=functional.render_batch_ray(setup,c,decoders,rd[lo:hi],ro[lo:hi],gd[lo:hi],depth_max=dmax); t.append(ev())
        counts=[sharding.shard_range(B,q,world)[1]-sharding.shard_range(B,q,world)[0] for q in range(world)]
        a=sharding.allgather_rows(d,counts); t.append(ev())
        b=sharding.allgather_rows(v,counts); cc=sharding.allgather_rows(col,counts); t.append(ev())
        full=functional.render_batch_ray(setup,c,decoders,rd[:B],ro[:B],gd,depth_max=None); t.append(ev())
        torch.cuda.synchronize()
        if rank==0 and rep==2:
            names=['depth_max','allreduce_max','render_shard','allgather_depth','allgather_var_color','render_full_100k']
            print({n: round(t[i].elapsed_time(t[i+1]),3) for i,n in enumerate(names)}, 'rays in shard', hi-lo, flush=True)
dist.barrier(); torch.cuda.synchronize(); os._exit(0)
