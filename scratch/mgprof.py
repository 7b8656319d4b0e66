import os, sys, time, json
sys.path.insert(0, '.')
import torch, torch.distributed as dist
import numpy as np
from gridfluidsim3d_b200 import capi, slabs, synth
import bench
rank=int(os.environ['RANK']); world=int(os.environ['WORLD_SIZE']); local=int(os.environ['LOCAL_RANK'])
torch.cuda.set_device(local); dev=torch.device('cuda',local)
dist.init_process_group('nccl', device_id=dev)
wl='splash256'; dims,dx,_=synth.CONFIGS[wl]
owned=slabs.slab_ranges(dims[2],world)[rank]
sc=bench.make_scene_device(wl,dev,seed=12345+rank,k_range=owned)
stream=torch.cuda.Stream(device=dev); torch.cuda.set_stream(stream)
ctx=capi.Context(local,stream=stream.cuda_stream); ctx.domain_init(dims,dx); ctx.set_material(sc['material'])
ctx.set_particles_aos(sc['aos'].cpu().numpy()); ctx.set_field(0,*[t.cpu().numpy() for t in sc['new']]); ctx.set_field(1,*[t.cpu().numpy() for t in sc['saved']])
b=slabs.CudaSlabBackend(ctx,dims,owned,0,migrate_cap=max(4096,sc['aos'].shape[0]//8),shared_stream=True)
drv=slabs.SlabDriver(b,rank,world,halo=2); tr=slabs.DistTransport(); dt=sc['dt']
T={}
def tick(name,t0):
    torch.cuda.synchronize(); T[name]=T.get(name,0)+time.perf_counter()-t0
def step(timed):
    t=time.perf_counter(); b.sort(); b.p2g_begin()
    if timed: tick('sort+splat',t); t=time.perf_counter()
    tr.layers(drv,('partials','halos'))
    if timed: tick('C1+C2',t); t=time.perf_counter()
    b.p2g_end(); b.g2p_advect(dt)
    if timed: tick('p2g_end+g2p',t); t=time.perf_counter()
    tr.migrate(drv)
    if timed: tick('C3 migrate',t)
for _ in range(3): step(False)
dist.barrier(); torch.cuda.synchronize()
n=10
t0=time.perf_counter()
for _ in range(n): step(False)
torch.cuda.synchronize(); untimed=(time.perf_counter()-t0)/n
for _ in range(n): step(True)
if rank==0: print('untimed step ms',untimed*1e3, {k:round(v/n*1e3,3) for k,v in T.items()})
dist.barrier(); dist.destroy_process_group()
