"""One-GPU exercise of the in-switch all-reduce kernel (csrc/collective.cu): a multicast object with a
team of ONE device, created through the driver API (torch's symmetric memory skips multicast for a
single rank), its memory mapped once as ordinary (unicast) memory and once at the multicast address.
The kernel then runs with world = 1: multimem.ld_reduce returns the rank's own data (a sum over one
member), multimem.st writes it back through the multicast mapping, and the arrive / depart flags
are exchanged with itself -- every instruction of the data path, on the box the driver's tests run on.

    python tools/multicast_single.py [floats=1048576]     ->  one JSON line, exit code 0 = ok
"""
import ctypes
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from jaxsplat_b200 import _lib  # noqa: E402


def run(n_floats: int = 1 << 20):
    from cuda.bindings import driver as drv

    def ck(res, what):
        err = res[0]
        if err != drv.CUresult.CUDA_SUCCESS:
            raise RuntimeError(f"{what}: {err}")
        return res[1:] if len(res) > 2 else (res[1] if len(res) == 2 else None)

    torch.cuda.init()
    torch.zeros(1, device="cuda:0")  # primary context current
    dev = ck(drv.cuDeviceGet(0), "cuDeviceGet")
    sup = ck(drv.cuDeviceGetAttribute(drv.CUdevice_attribute.CU_DEVICE_ATTRIBUTE_MULTICAST_SUPPORTED, dev), "attr")
    if not sup:
        return {"skipped": "device reports no multicast support"}
    flag_bytes = 4096
    want = n_floats * 4 + flag_bytes
    prop = drv.CUmulticastObjectProp()
    prop.numDevices = 1
    prop.size = want
    prop.handleTypes = 0
    prop.flags = 0
    gran = ck(drv.cuMulticastGetGranularity(prop, drv.CUmulticastGranularity_flags.CU_MULTICAST_GRANULARITY_RECOMMENDED), "granularity")
    size = (want + gran - 1) // gran * gran
    prop.size = size
    mc = ck(drv.cuMulticastCreate(prop), "cuMulticastCreate")
    ck(drv.cuMulticastAddDevice(mc, dev), "cuMulticastAddDevice")
    aprop = drv.CUmemAllocationProp()
    aprop.type = drv.CUmemAllocationType.CU_MEM_ALLOCATION_TYPE_PINNED
    aprop.location.type = drv.CUmemLocationType.CU_MEM_LOCATION_TYPE_DEVICE
    aprop.location.id = 0
    mem_gran = ck(drv.cuMemGetAllocationGranularity(aprop, drv.CUmemAllocationGranularity_flags.CU_MEM_ALLOC_GRANULARITY_RECOMMENDED), "mem granularity")
    size = (size + mem_gran - 1) // mem_gran * mem_gran
    mem = ck(drv.cuMemCreate(size, aprop, 0), "cuMemCreate")
    ck(drv.cuMulticastBindMem(mc, 0, mem, 0, size, 0), "cuMulticastBindMem")
    acc = drv.CUmemAccessDesc()
    acc.location.type = drv.CUmemLocationType.CU_MEM_LOCATION_TYPE_DEVICE
    acc.location.id = 0
    acc.flags = drv.CUmemAccess_flags.CU_MEM_ACCESS_FLAGS_PROT_READWRITE
    uc_va = ck(drv.cuMemAddressReserve(size, max(gran, mem_gran), 0, 0), "reserve unicast")
    ck(drv.cuMemMap(uc_va, size, 0, mem, 0), "map unicast")
    ck(drv.cuMemSetAccess(uc_va, size, [acc], 1), "access unicast")
    mc_va = ck(drv.cuMemAddressReserve(size, max(gran, mem_gran), 0, 0), "reserve multicast")
    ck(drv.cuMemMap(mc_va, size, 0, mc, 0), "map multicast")
    ck(drv.cuMemSetAccess(mc_va, size, [acc], 1), "access multicast")
    uc, mcp = int(uc_va), int(mc_va)
    out = {"floats": n_floats, "granularity": int(gran), "bytes": int(size)}
    try:
        ck(drv.cuMemsetD8(uc_va, 0, size), "memset")
        x = torch.randn(n_floats, generator=torch.Generator(device="cuda:0").manual_seed(7), device="cuda:0")
        stream = torch.cuda.current_stream().cuda_stream
        ok = True
        for epoch in (1, 2, 3):
            ck(drv.cuMemcpyDtoD(uc_va, x.data_ptr(), n_floats * 4), "copy in")
            torch.cuda.synchronize()
            peers = (ctypes.c_void_p * 1)(uc)
            rc = _lib.lib().jsb200_multimem_allreduce_f32(ctypes.c_void_p(stream), ctypes.c_void_p(mcp), peers,
                                                          n_floats, n_floats * 4, 0, 1, epoch)
            if rc != 0:
                _lib.check_error("multimem_allreduce")
                raise RuntimeError("multimem_allreduce failed")
            torch.cuda.synchronize()
            y = torch.empty_like(x)
            ck(drv.cuMemcpyDtoD(y.data_ptr(), uc_va, n_floats * 4), "copy out")
            torch.cuda.synchronize()
            ok = ok and bool(torch.equal(x, y))  # a sum over one member: the data itself, bit for bit
            x = x * 1.5 + 0.25
        out["ok"] = ok
    finally:
        drv.cuMemUnmap(mc_va, size)
        drv.cuMemUnmap(uc_va, size)
        drv.cuMemAddressFree(mc_va, size)
        drv.cuMemAddressFree(uc_va, size)
        drv.cuMemRelease(mem)
        drv.cuMemRelease(mc)
    return out


if __name__ == "__main__":
    try:
        r = run(int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20)
    except RuntimeError as e:  # the driver refused a step of the set-up (a one-GPU slice of a node
        # cannot create multicast objects: cuMulticastCreate -> CUDA_ERROR_INVALID_VALUE)
        r = {"skipped": str(e)}
    print(json.dumps(r))
    sys.exit(0 if r.get("ok") or r.get("skipped") else 1)
