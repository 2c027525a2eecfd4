"""probe: does torch symmetric memory + NVLS multicast work on this box? (torchrun, N ranks)"""
import os, sys, time
import torch, torch.distributed as dist
rank=int(os.environ["RANK"]); world=int(os.environ["WORLD_SIZE"]); lr=int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr); dev=torch.device("cuda",lr)
dist.init_process_group("nccl", device_id=dev)
import torch.distributed._symmetric_memory as symm_mem
n=14*1000*1000
try:
    t=symm_mem.empty(n,dtype=torch.float32,device=dev)
    hdl=symm_mem.rendezvous(t, dist.group.WORLD.group_name)
    if rank==0:
        print("symm ok: world",hdl.world_size,"mc_ptr",hex(hdl.multicast_ptr) if hdl.multicast_ptr else None, "has signal pad", bool(hdl.signal_pad_ptrs), flush=True)
        print([a for a in dir(hdl) if not a.startswith('_')], flush=True)
    t.fill_(rank+1.0)
    hdl.barrier(channel=0)
    if hdl.multicast_ptr:
        torch.ops.symm_mem.multimem_all_reduce_(t,"sum",dist.group.WORLD.group_name)
        torch.cuda.synchronize()
        exp=sum(range(1,world+1))
        ok=bool((t==exp).all())
        e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
        dist.barrier(); torch.cuda.synchronize(); e0.record()
        for _ in range(20): torch.ops.symm_mem.multimem_all_reduce_(t,"sum",dist.group.WORLD.group_name)
        e1.record(); torch.cuda.synchronize()
        if rank==0: print("multimem_all_reduce_ ok",ok,"ms",e0.elapsed_time(e1)/20, flush=True)
    g=torch.ones(n,device=dev)
    dist.all_reduce(g); torch.cuda.synchronize()
    e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    dist.barrier(); torch.cuda.synchronize(); e0.record()
    for _ in range(20): dist.all_reduce(g)
    e1.record(); torch.cuda.synchronize()
    if rank==0: print("nccl all_reduce 56MB ms",e0.elapsed_time(e1)/20, flush=True)
except Exception as ex:
    print("rank",rank,"symm FAILED:",repr(ex)[:500], flush=True)
dist.destroy_process_group()
