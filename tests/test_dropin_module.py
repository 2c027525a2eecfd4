"""CPU: the compiled ``_jaxsplat`` module (jaxsplat_b200/dropin/, replaces the pybind module of
reference lib/ffi.cu:54-68) and the status-returning custom-call variants.  Nothing here launches a
kernel; the GPU twin is tests/test_gpu_views_and_dropin.py."""
import ctypes
import importlib.util
import os
import subprocess
import sysconfig

import pytest

from conftest import ROOT, load_golden
from jaxsplat_b200 import _ffi, _lib

DROPIN = os.path.join(ROOT, "jaxsplat_b200", "dropin",
                      "_jaxsplat" + sysconfig.get_config_var("EXT_SUFFIX"))


def load_dropin():
    assert os.path.exists(DROPIN), "build it: make -C jaxsplat_b200/csrc (or __graft_entry__.build())"
    spec = importlib.util.spec_from_file_location("_jaxsplat", DROPIN)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def test_module_has_the_reference_surface():
    m = load_dropin()
    regs = m.registrations()
    assert sorted(regs) == ["project_bwd", "project_fwd", "rasterize_bwd", "rasterize_fwd"]
    L = _lib.lib()
    for name, cap in regs.items():
        ptr = _ffi.capsule_pointer(cap)  # raises unless named xla._CUSTOM_CALL_TARGET
        assert ptr == ctypes.cast(getattr(L, "jsb200_" + name), ctypes.c_void_p).value
    for name, cap in m.registrations(status_returning=True).items():
        ptr = _ffi.capsule_pointer(cap)
        assert ptr == ctypes.cast(getattr(L, "jsb200_" + name + "_status"), ctypes.c_void_p).value


def test_module_make_descriptor_matches_the_reference_bytes():
    m = load_dropin()
    for name in ("g_small16", "g_small8", "g_inside16"):
        g = load_golden(name)
        n = g["in_means3d"].shape[0]
        kw = g["kw"]
        # keyword names of the reference's pybind signature (lib/ffi.cu:57-67)
        ours = m.make_descriptor(num_points=n, img_shape=kw["img_shape"], f=kw["f"], c=kw["c"],
                                 glob_scale=kw["glob_scale"], clip_thresh=kw["clip_thresh"],
                                 block_width=kw["block_width"])
        pos = m.make_descriptor(n, kw["img_shape"], kw["f"], kw["c"], kw["glob_scale"],
                                kw["clip_thresh"], kw["block_width"])
        ref = g["descriptor"].tobytes()  # written by the reference's own make_descriptor
        assert ours == pos and len(ours) == 80
        assert ours[:76] == ref[:76] and ours[76:] == b"\0\0\0\0"
        assert ours == _ffi.make_descriptor(n, kw["img_shape"], kw["f"], kw["c"], kw["glob_scale"],
                                            kw["clip_thresh"], kw["block_width"])
    with pytest.raises(TypeError):
        m.make_descriptor(1, (16, 16), (1.0, 1.0), (0.0, 0.0), 1.0, 0.01)  # block_width missing
    with pytest.raises(ValueError):
        m.make_descriptor(1, (16, 16), (1.0, 1.0), (0.0, 0.0), 1.0, 0.01, 0)
    m.clear_error()
    assert m.last_error() is None


FAKE_STATUS_C = r"""
#include <stddef.h>
#include <string.h>
static char g_msg[512];
static void *g_status;
__attribute__((visibility("default")))
void XlaCustomCallStatusSetFailure(void *status, const char *msg, size_t len) {
    g_status = status;
    if (len > sizeof g_msg - 1) len = sizeof g_msg - 1;
    memcpy(g_msg, msg, len);
    g_msg[len] = 0;
}
__attribute__((visibility("default"))) const char *fake_status_message(void) { return g_msg; }
__attribute__((visibility("default"))) void *fake_status_pointer(void) { return g_status; }
"""


def test_status_variant_reports_through_xla_status(tmp_path):
    """a failing call hands its message to XlaCustomCallStatusSetFailure (here: a stand-in for the
    symbol jaxlib exports) instead of throwing through the C frame like the reference"""
    src = tmp_path / "fake_status.c"
    src.write_text(FAKE_STATUS_C)
    so = tmp_path / "libfake_status.so"
    subprocess.check_call(["gcc", "-shared", "-fPIC", str(src), "-o", str(so)])
    fake = ctypes.CDLL(str(so), mode=ctypes.RTLD_GLOBAL)
    fake.fake_status_message.restype = ctypes.c_char_p
    fake.fake_status_pointer.restype = ctypes.c_void_p
    L = _lib.lib()
    L.jsb200_clear_error()
    status = ctypes.create_string_buffer(64)  # opaque to us: only its address travels
    bufs = (ctypes.c_void_p * 1)(None)
    for name in ("project_fwd", "project_bwd", "rasterize_fwd", "rasterize_bwd"):
        getattr(L, "jsb200_" + name + "_status")(None, bufs, b"x" * 12, 12,
                                                 ctypes.cast(status, ctypes.c_void_p))
        assert b"descriptor" in fake.fake_status_message()
        assert fake.fake_status_pointer() == ctypes.addressof(status)
        assert b"descriptor" in L.jsb200_last_error()
        L.jsb200_clear_error()
    # a null status is tolerated (legacy callers)
    L.jsb200_project_fwd_status(None, bufs, b"x" * 12, 12, None)
    assert L.jsb200_last_error() is not None
    L.jsb200_clear_error()


def test_views_entry_points_validate_before_touching_the_gpu():
    L = _lib.lib()
    L.jsb200_clear_error()
    bufs = (ctypes.c_void_p * 1)(None)
    assert L.jsb200_render_views_fwd(None, bufs, b"x" * 3, 3, 2, None) == 1
    assert b"descriptor" in L.jsb200_last_error()
    L.jsb200_clear_error()
    d = _ffi.make_descriptor(4, (32, 32), (16.0, 16.0), (16.0, 16.0), 1.0, 0.01, 16)
    assert L.jsb200_render_views_bwd(None, None, d, len(d), 2, None) == 1
    assert b"null buffer list" in L.jsb200_last_error()
    L.jsb200_clear_error()
    with pytest.raises(RuntimeError):
        _lib.set_option("view_slots", 0)
    _lib.set_option("view_slots", 8)


def test_jax_binding_is_importable_and_guarded_without_jax():
    """the JAX host layer cannot run here (no JAX in the image) but must import, hand out the
    same capsules as the compiled module, and fail loudly when asked to register"""
    from jaxsplat_b200 import jax_binding as jb
    L = _lib.lib()
    for status in (False, True):
        regs = jb.registrations(status_returning=status)
        assert sorted(regs) == sorted(_ffi.TARGETS)
        for name, cap in regs.items():
            sym = "jsb200_" + name + ("_status" if status else "")
            assert _ffi.capsule_pointer(cap) == ctypes.cast(getattr(L, sym), ctypes.c_void_p).value
    d = jb.make_descriptor(7, (32, 48), (24.0, 16.0), (24.0, 16.0), 1.0, 0.01, 16)
    assert d == _ffi.make_descriptor(7, (32, 48), (24.0, 16.0), (24.0, 16.0), 1.0, 0.01, 16)
    if not jb.HAVE_JAX:
        with pytest.raises(ImportError):
            jb.register_targets()
    for fn in ("render", "project", "rasterize"):
        assert callable(getattr(jb, fn))
