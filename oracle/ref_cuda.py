"""Loader for the UNMODIFIED reference native library built by oracle/Makefile into oracle/_ref/.

TEST / BENCH INFRASTRUCTURE ONLY.  The product (jaxsplat_b200/) never imports this module.
It drives the reference's four XLA custom-call targets without JAX by pulling the function
pointers out of the PyCapsules that ``_jaxsplat.registrations()`` returns (lib/ffi.cu:4-13) and
calling them through the same ``CustomCallTable`` harness used for our own library.  It also
reaches the reference's internal ``ops::sort_and_bin`` (lib/ops.h:31-41) through its exported
mangled symbol so the sorted ids / tile bins (scratch in the reference) can be compared.
"""
from __future__ import annotations

import ctypes
import glob
import importlib.util
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
_SORT_AND_BIN = "_ZN3ops12sort_and_binEP11CUstream_stRKNS_10DescriptorEjPK6float2PKfPKiSB_PiP4int2"
_CUMSUM = "_ZN3ops6cumsumEP11CUstream_stPKiPii"


def ref_so_path():
    hits = sorted(glob.glob(os.path.join(_HERE, "_ref", "_jaxsplat*.so")))
    return hits[0] if hits else None


def available() -> bool:
    return ref_so_path() is not None


_mod = None


def module():
    global _mod
    if _mod is None:
        path = ref_so_path()
        if path is None:
            raise FileNotFoundError("oracle/_ref/_jaxsplat*.so missing: run `make -C oracle` where "
                                    "/root/reference exists")
        spec = importlib.util.spec_from_file_location("_jaxsplat", path)
        _mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(_mod)
    return _mod


def table():
    """CustomCallTable over the reference's capsules, using the reference's own make_descriptor."""
    from jaxsplat_b200._ffi import CustomCallTable

    m = module()

    def make_descriptor(num_points, img_shape, f, c, glob_scale, clip_thresh, block_width):
        return m.make_descriptor(num_points=int(num_points), img_shape=tuple(int(x) for x in img_shape),
                                 f=tuple(float(x) for x in f), c=tuple(float(x) for x in c),
                                 glob_scale=float(glob_scale), clip_thresh=float(clip_thresh),
                                 block_width=int(block_width))

    return CustomCallTable(m.registrations(), "reference", make_descriptor=make_descriptor)


def sort_and_bin(desc: bytes, num_intersects: int, xys, depths, radii, cum_tiles_hit, stream=None):
    """reference ops::sort_and_bin on torch CUDA tensors -> (ids_sorted int32[I], tile_bins int32[T,2])"""
    import struct

    import torch

    lib = ctypes.CDLL(ref_so_path())
    fn = getattr(lib, _SORT_AND_BIN)
    fn.restype = None
    fn.argtypes = [ctypes.c_void_p, ctypes.c_char_p, ctypes.c_uint] + [ctypes.c_void_p] * 6
    tiles_x, tiles_y = struct.unpack_from("<2I", desc, 52)
    ids = torch.zeros(num_intersects, dtype=torch.int32, device=xys.device)
    bins = torch.zeros((tiles_x * tiles_y, 2), dtype=torch.int32, device=xys.device)
    if stream is None:
        stream = torch.cuda.current_stream().cuda_stream
    fn(ctypes.c_void_p(stream), desc, num_intersects, xys.data_ptr(), depths.data_ptr(),
       radii.data_ptr(), cum_tiles_hit.data_ptr(), ids.data_ptr(), bins.data_ptr())
    torch.cuda.synchronize()
    return ids, bins
