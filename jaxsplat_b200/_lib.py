"""ctypes loader for the C-ABI library ``lib/libjaxsplat_b200.so`` (include/jaxsplat_b200.h).

There is deliberately NO fallback: if the CUDA library has not been built the import of any
compute entry point fails loudly (``python -c "import __graft_entry__ as g; g.build()"`` or
``make -C jaxsplat_b200/csrc`` builds it in-tree).
"""
from __future__ import annotations

import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# JSB200_LIB: load another build of the same library (A/B experiments: tools/gpu_ab_builds.sh)
LIB_PATH = os.environ.get("JSB200_LIB") or os.path.join(_HERE, "lib", "libjaxsplat_b200.so")

# every symbol include/jaxsplat_b200.h declares
EXPORTS = (
    "jsb200_project_fwd",
    "jsb200_project_bwd",
    "jsb200_rasterize_fwd",
    "jsb200_rasterize_bwd",
    "jsb200_project_fwd_status",
    "jsb200_project_bwd_status",
    "jsb200_rasterize_fwd_status",
    "jsb200_rasterize_bwd_status",
    "jsb200_render_fwd",
    "jsb200_render_bwd",
    "jsb200_render_bwd_precleared",
    "jsb200_clear_async",
    "jsb200_render_views_fwd",
    "jsb200_render_views_bwd",
    "jsb200_make_descriptor",
    "jsb200_set_descriptor_band",
    "jsb200_sort_and_bin",
    "jsb200_cumsum",
    "jsb200_last_error",
    "jsb200_clear_error",
    "jsb200_error_count",
    "jsb200_set_option",
    "jsb200_debug_binning",
    "jsb200_debug_hit_words",
    "jsb200_get_stage_times",
    "jsb200_get_stats",
    "jsb200_get_list_stats",
    "jsb200_release",
    "jsb200_multimem_allreduce_f32",
    "jsb200_adam_step",
    "jsb200_mse_loss_grad",
    "jsb200_mse_loss_grad_u8",
)

_lib = None


class NativeLibraryMissing(ImportError):
    pass


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise NativeLibraryMissing(
            f"{LIB_PATH} not found: build it with `make -C jaxsplat_b200/csrc` "
            "(or __graft_entry__.build()). jaxsplat_b200 has no CPU or PyTorch fallback."
        )
    L = ctypes.CDLL(LIB_PATH)
    vp, cp, sz, u, i, f = (ctypes.c_void_p, ctypes.c_char_p, ctypes.c_size_t, ctypes.c_uint,
                           ctypes.c_int, ctypes.c_float)
    for name in ("jsb200_project_fwd", "jsb200_project_bwd", "jsb200_rasterize_fwd",
                 "jsb200_rasterize_bwd"):
        fn = getattr(L, name)
        fn.argtypes = [vp, ctypes.POINTER(vp), cp, sz]
        fn.restype = None
    for name in ("jsb200_project_fwd_status", "jsb200_project_bwd_status",
                 "jsb200_rasterize_fwd_status", "jsb200_rasterize_bwd_status"):
        fn = getattr(L, name)
        fn.argtypes = [vp, ctypes.POINTER(vp), cp, sz, vp]
        fn.restype = None
    L.jsb200_render_views_fwd.argtypes = [vp, ctypes.POINTER(vp), cp, sz, u,
                                          ctypes.POINTER(ctypes.c_ulonglong)]
    L.jsb200_render_views_fwd.restype = i
    L.jsb200_render_views_bwd.argtypes = [vp, ctypes.POINTER(vp), cp, sz, u,
                                          ctypes.POINTER(ctypes.c_ulonglong)]
    L.jsb200_render_views_bwd.restype = i
    L.jsb200_render_fwd.argtypes = [vp, ctypes.POINTER(vp), cp, sz, ctypes.POINTER(ctypes.c_ulonglong)]
    L.jsb200_render_fwd.restype = i
    L.jsb200_render_bwd.argtypes = [vp, ctypes.POINTER(vp), cp, sz, ctypes.c_ulonglong]
    L.jsb200_render_bwd.restype = i
    L.jsb200_render_bwd_precleared.argtypes = [vp, ctypes.POINTER(vp), cp, sz, ctypes.c_ulonglong]
    L.jsb200_render_bwd_precleared.restype = i
    L.jsb200_clear_async.argtypes = [vp, vp, sz]
    L.jsb200_clear_async.restype = i
    L.jsb200_make_descriptor.argtypes = [u, u, u, f, f, f, f, f, f, u, ctypes.c_char_p]
    L.jsb200_make_descriptor.restype = i
    L.jsb200_set_descriptor_band.argtypes = [ctypes.c_char_p, u, u]
    L.jsb200_set_descriptor_band.restype = i
    L.jsb200_sort_and_bin.argtypes = [vp, cp, sz, u, vp, vp, vp, vp, vp, vp]
    L.jsb200_sort_and_bin.restype = i
    L.jsb200_cumsum.argtypes = [vp, vp, vp, i]
    L.jsb200_cumsum.restype = i
    L.jsb200_last_error.argtypes = []
    L.jsb200_last_error.restype = ctypes.c_char_p
    L.jsb200_clear_error.argtypes = []
    L.jsb200_clear_error.restype = None
    L.jsb200_error_count.argtypes = []
    L.jsb200_error_count.restype = ctypes.c_ulonglong
    L.jsb200_set_option.argtypes = [cp, i]
    L.jsb200_set_option.restype = i
    L.jsb200_debug_binning.argtypes = [vp, vp, vp, vp, vp]
    L.jsb200_debug_binning.restype = ctypes.c_longlong
    L.jsb200_debug_hit_words.argtypes = [vp, vp, sz]
    L.jsb200_debug_hit_words.restype = ctypes.c_longlong
    L.jsb200_get_stage_times.argtypes = [vp, ctypes.POINTER(ctypes.c_double),
                                         ctypes.POINTER(ctypes.c_ulonglong), i]
    L.jsb200_get_stage_times.restype = i
    L.jsb200_get_stats.argtypes = [ctypes.POINTER(ctypes.c_ulonglong)]
    L.jsb200_get_stats.restype = None
    L.jsb200_get_list_stats.argtypes = [ctypes.POINTER(ctypes.c_ulonglong)]
    L.jsb200_get_list_stats.restype = None
    L.jsb200_release.argtypes = []
    L.jsb200_release.restype = None
    L.jsb200_multimem_allreduce_f32.argtypes = [vp, vp, ctypes.POINTER(vp), sz, sz, i, i, u]
    L.jsb200_multimem_allreduce_f32.restype = i
    L.jsb200_adam_step.argtypes = [vp, vp, vp, vp, vp, sz, f, f, f, f, ctypes.c_ulonglong]
    L.jsb200_adam_step.restype = i
    L.jsb200_mse_loss_grad.argtypes = [vp, vp, vp, sz, f, vp, vp]
    L.jsb200_mse_loss_grad.restype = i
    L.jsb200_mse_loss_grad_u8.argtypes = [vp, vp, vp, sz, f, vp, vp]
    L.jsb200_mse_loss_grad_u8.restype = i
    _lib = L
    return L


def check_error(where: str) -> None:
    """Raise if the calling thread's last native call recorded an error."""
    L = lib()
    msg = L.jsb200_last_error()
    if msg:
        text = msg.decode("utf-8", "replace")
        L.jsb200_clear_error()
        raise RuntimeError(f"jaxsplat_b200 {where}: {text}")


def clear_async(stream: int, ptr: int, nbytes: int) -> None:
    """zero `nbytes` at device address `ptr` on the raw stream handle `stream` (cudaMemsetAsync)"""
    if lib().jsb200_clear_async(ctypes.c_void_p(stream), ctypes.c_void_p(ptr), ctypes.c_size_t(nbytes)) != 0:
        check_error("clear_async")
        raise RuntimeError("jaxsplat_b200 clear_async failed")


def stats() -> dict:
    out = (ctypes.c_ulonglong * 8)()
    lib().jsb200_get_stats(out)
    keys = ("rasterize_fwd_calls", "rasterize_bwd_calls", "binning_runs", "binning_cache_hits",
            "last_num_intersects", "workspace_bytes", "kernel_launches", "last_row_entries")
    return dict(zip(keys, [int(x) for x in out]))


def list_stats() -> dict:
    out = (ctypes.c_ulonglong * 4)()
    lib().jsb200_get_list_stats(out)
    return {"intersections": int(out[0]), "row_entries": int(out[1]), "super_entries": int(out[2]),
            "super_shift": int(out[3])}


def raw_stream(device=None) -> int:
    """the current CUDA stream's handle of `device` (a torch.device / index; default: the current
    device) without building a torch.cuda.Stream object (three attribute look-ups and a device query
    per call: ~15 us of a small scene's 350 us step)"""
    import torch
    if device is None:
        idx = torch.cuda.current_device()
    elif isinstance(device, int):
        idx = device
    else:
        idx = device.index if device.index is not None else torch.cuda.current_device()
    get = getattr(torch._C, "_cuda_getCurrentRawStream", None)
    if get is None:  # (a torch without the private accessor)
        return torch.cuda.current_stream(idx).cuda_stream
    return get(idx)


def set_option(key: str, value: int) -> None:
    if lib().jsb200_set_option(key.encode(), int(value)) != 0:
        check_error("set_option")


STAGES = ("project_fwd", "bin_prep", "depth_sort", "emit", "tile_scatter", "tile_offsets", "raster_fwd",
          "raster_bwd", "project_bwd")


def stage_times(stream: int, reset: bool = True) -> dict:
    """accumulated per-stage device milliseconds / span counts (option "timing" must be on)"""
    ms = (ctypes.c_double * 9)()
    calls = (ctypes.c_ulonglong * 9)()
    if lib().jsb200_get_stage_times(ctypes.c_void_p(stream), ms, calls, int(reset)) != 0:
        check_error("get_stage_times")
    return {k: dict(ms=float(ms[i]), calls=int(calls[i])) for i, k in enumerate(STAGES)}
