"""Host-side mirror of the reference's ``_jaxsplat`` extension module (lib/ffi.cu:4-68).

``registrations()`` returns the four XLA custom-call targets as PyCapsules named
``"xla._CUSTOM_CALL_TARGET"`` and ``make_descriptor(...)`` returns the 80-byte opaque blob --
same names, same argument meaning as the reference, so ``jaxsplat/_project/impl.py:12-13`` style
registration code works unchanged against this module.

``CustomCallTable`` is the framework-neutral way to invoke such targets on torch CUDA tensors
(``fn(stream, void** buffers, opaque, len)``); tests use the very same class to drive the
reference's own capsules, so both libraries run under one harness.
"""
from __future__ import annotations

import ctypes
import functools
from typing import Dict, Sequence

from . import _lib

TARGETS = ("project_fwd", "project_bwd", "rasterize_fwd", "rasterize_bwd")
_CAPSULE_NAME = b"xla._CUSTOM_CALL_TARGET"
_CALL_T = ctypes.CFUNCTYPE(None, ctypes.c_void_p, ctypes.POINTER(ctypes.c_void_p), ctypes.c_char_p,
                           ctypes.c_size_t)


def make_descriptor(num_points: int, img_shape, f, c, glob_scale: float, clip_thresh: float,
                    block_width: int) -> bytes:
    """reference lib/ffi.cu:15-52; img_shape is (H, W).  (Cached: a training loop asks for the same
    80 bytes every step.)"""
    return _make_descriptor_cached(int(num_points), (int(img_shape[0]), int(img_shape[1])),
                                   (float(f[0]), float(f[1])), (float(c[0]), float(c[1])),
                                   float(glob_scale), float(clip_thresh), int(block_width))


@functools.lru_cache(maxsize=256)
def _make_descriptor_cached(num_points, img_shape, f, c, glob_scale, clip_thresh, block_width) -> bytes:
    buf = ctypes.create_string_buffer(80)
    rc = _lib.lib().jsb200_make_descriptor(int(num_points), int(img_shape[0]), int(img_shape[1]),
                                           float(f[0]), float(f[1]), float(c[0]), float(c[1]),
                                           float(glob_scale), float(clip_thresh), int(block_width),
                                           buf)
    if rc != 0:
        _lib.check_error("make_descriptor")
    return buf.raw


def with_band(descriptor: bytes, tile_row_begin: int, tile_row_end: int) -> bytes:
    """Extension: restrict a rasterize descriptor to tile rows [begin, end) (screen-tile-row
    sharding; include/jaxsplat_b200.h: jsb200_set_descriptor_band)."""
    buf = ctypes.create_string_buffer(descriptor, 80)
    if _lib.lib().jsb200_set_descriptor_band(buf, int(tile_row_begin), int(tile_row_end)) != 0:
        _lib.check_error("set_descriptor_band")
    return buf.raw


def registrations() -> Dict[str, object]:
    """reference lib/ffi.cu:4-13: name -> PyCapsule("xla._CUSTOM_CALL_TARGET")."""
    L = _lib.lib()
    new = ctypes.pythonapi.PyCapsule_New
    new.restype = ctypes.py_object
    new.argtypes = [ctypes.c_void_p, ctypes.c_char_p, ctypes.c_void_p]
    out = {}
    for name in TARGETS:
        addr = ctypes.cast(getattr(L, "jsb200_" + name), ctypes.c_void_p).value
        out[name] = new(addr, _CAPSULE_NAME, None)
    return out


def capsule_pointer(capsule) -> int:
    get = ctypes.pythonapi.PyCapsule_GetPointer
    get.restype = ctypes.c_void_p
    get.argtypes = [ctypes.py_object, ctypes.c_char_p]
    return get(capsule, _CAPSULE_NAME)


class CustomCallTable:
    """Calls XLA-style custom-call targets on torch CUDA tensors."""

    def __init__(self, capsules: Dict[str, object], name: str, check=None, make_descriptor=None):
        self.name = name
        self.make_descriptor = make_descriptor or globals()["make_descriptor"]
        self._fns = {k: _CALL_T(capsule_pointer(v)) for k, v in capsules.items()}
        self._check = check

    def call(self, target: str, operands: Sequence, results: Sequence, opaque: bytes,
             stream=None) -> None:
        import torch

        tensors = list(operands) + list(results)
        for t in tensors:
            if not t.is_cuda or not t.is_contiguous():
                raise ValueError(f"{target}: every buffer must be a contiguous CUDA tensor")
        if stream is None:
            from . import _lib
            stream = _lib.raw_stream()
        bufs = (ctypes.c_void_p * len(tensors))(*[t.data_ptr() for t in tensors])
        self._fns[target](ctypes.c_void_p(stream), bufs, opaque, len(opaque))
        if self._check is not None:
            self._check(target)


_native = None


def native_table() -> CustomCallTable:
    global _native
    if _native is None:
        _native = CustomCallTable(registrations(), "jaxsplat_b200", check=_lib.check_error)
    return _native
