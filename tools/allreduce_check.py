"""torchrun: the in-switch gradient all-reduce (csrc/collective.cu) vs torch.distributed.all_reduce
-- exactness on integers-as-floats, closeness on random data, and device time of both."""
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from jaxsplat_b200.multiview import SwitchReducedBuffer  # noqa: E402

rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dev = torch.device("cuda", lr)
dist.init_process_group("nccl", device_id=dev)
n = 14 * int(sys.argv[1]) if len(sys.argv) > 1 else 14 * 1000 * 1000
os.environ["JSB_SWITCH_ALLREDUCE"] = "1"  # the kernel under test, also at two ranks (NCCL is the default there)
buf = SwitchReducedBuffer(n, dev)
out = {"world": world, "floats": n, "mode": buf.mode}
g = torch.Generator(device=dev).manual_seed(100 + rank)
ok = True
for trial in range(4):
    if trial % 2 == 0:   # small integers: every summation order gives the same float
        x = torch.randint(-1000, 1000, (n,), generator=g, device=dev).float()
    else:
        x = torch.randn(n, generator=g, device=dev)
    buf.tensor.copy_(x)
    ref = x.clone()
    dist.all_reduce(ref)
    buf.all_reduce()
    torch.cuda.synchronize()
    if trial % 2 == 0:
        ok = ok and bool(torch.equal(buf.tensor, ref))
    else:
        err = float((buf.tensor - ref).abs().max())
        ok = ok and err <= 1e-5 * max(1.0, float(ref.abs().max()))
        out["max_abs_err_random"] = err
out["ok"] = ok


def timed(fn, iters=50):
    for _ in range(5):
        fn()
    dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / iters], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t)


other = torch.zeros(n, device=dev)
out["ms_switch_kernel"] = timed(buf.all_reduce)
out["ms_nccl"] = timed(lambda: dist.all_reduce(other))
bytes_ = n * 4
out["busbw_GBs_switch_kernel"] = 2 * (world - 1) / world * bytes_ / (out["ms_switch_kernel"] * 1e-3) / 1e9
out["busbw_GBs_nccl"] = 2 * (world - 1) / world * bytes_ / (out["ms_nccl"] * 1e-3) / 1e9
if rank == 0:
    print(json.dumps(out), flush=True)
ok_all = torch.tensor([1 if ok else 0], device=dev)
dist.all_reduce(ok_all, op=dist.ReduceOp.MIN)
dist.destroy_process_group()
sys.exit(0 if int(ok_all) else 1)
