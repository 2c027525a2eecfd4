"""CPU: the C-ABI library loads, exports every symbol include/jaxsplat_b200.h declares, and the
host-side logic (descriptor layout, registrations, argument validation, synthetic scenes) works
without a GPU.  No compute entry point is called here."""
import ctypes
import os
import re
import struct

import pytest
import torch

from conftest import ROOT, load_golden
from jaxsplat_b200 import _ffi, _lib, _ops, synthetic


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "jaxsplat_b200.h")).read()
    return sorted(set(re.findall(r"\b(jsb200_[a-z_0-9]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    L = _lib.lib()
    declared = _declared_symbols()
    assert len(declared) >= 14
    for name in declared:
        assert hasattr(L, name), name
    assert sorted(_lib.EXPORTS) == declared


def test_descriptor_layout_matches_reference_bytes():
    """the reference's own make_descriptor output (stored in the golden fixtures) equals ours
    except for its 4 uninitialised pad bytes, which we zero"""
    for name in ("g_small16", "g_small8", "g_inside16"):
        g = load_golden(name)
        n = g["in_means3d"].shape[0]
        ours = _ffi.make_descriptor(n, g["kw"]["img_shape"], g["kw"]["f"], g["kw"]["c"],
                                    g["kw"]["glob_scale"], g["kw"]["clip_thresh"],
                                    g["kw"]["block_width"])
        ref = g["descriptor"].tobytes()
        assert len(ours) == 80 and len(ref) == 80
        assert ours[:76] == ref[:76]
        assert ours[76:] == b"\0\0\0\0"


def test_descriptor_fields():
    d = _ffi.make_descriptor(1000, (1080, 1920), (960.0, 540.0), (959.5, 539.5), 1.5, 0.01, 16)
    (n, w, h, one, fx, fy, cx, cy, gs, clip, bw, g1, b1, tx, ty, tz, bx, by, bz,
     pad) = struct.unpack("<4I4f2f3I3I3II", d)
    assert (n, w, h, one) == (1000, 1920, 1080, 1)
    assert (fx, fy, cx, cy) == (960.0, 540.0, 959.5, 539.5)
    assert gs == 1.5 and abs(clip - 0.01) < 1e-9
    assert (bw, g1, b1) == (16, 4, 256)
    assert (tx, ty, tz) == (120, 68, 1) and (bx, by, bz) == (16, 16, 1) and pad == 0


def test_registrations_are_xla_capsules():
    regs = _ffi.registrations()
    assert sorted(regs) == ["project_bwd", "project_fwd", "rasterize_bwd", "rasterize_fwd"]
    L = _lib.lib()
    for name, cap in regs.items():
        ptr = _ffi.capsule_pointer(cap)  # raises if the capsule name is not xla._CUSTOM_CALL_TARGET
        assert ptr == ctypes.cast(getattr(L, "jsb200_" + name), ctypes.c_void_p).value


def test_bad_opaque_is_an_error_not_a_crash():
    L = _lib.lib()
    L.jsb200_clear_error()
    before = L.jsb200_error_count()
    bufs = (ctypes.c_void_p * 1)(None)
    L.jsb200_project_fwd(None, bufs, b"x" * 10, 10)  # wrong descriptor size (reference: throws)
    assert L.jsb200_error_count() == before + 1
    assert b"descriptor" in L.jsb200_last_error()
    with pytest.raises(RuntimeError):
        _lib.check_error("project_fwd")
    assert L.jsb200_last_error() is None


def test_unknown_option_rejected():
    with pytest.raises(RuntimeError):
        _lib.set_option("no_such_option", 1)


def test_every_option_is_documented_in_the_header_and_accepted():
    """the keys jsb200_set_option compares against == the keys include/jaxsplat_b200.h documents;
    every one is accepted with its documented default (options only set host-side switches: no GPU)"""
    src = open(os.path.join(ROOT, "jaxsplat_b200", "csrc", "capi.cu")).read()
    body = src[src.index("int jsb200_set_option("):]
    body = body[:body.index("\n}\n")]
    keys = set(re.findall(r'strcmp\(key, "(\w+)"\)', body))
    hdr = open(os.path.join(ROOT, "include", "jaxsplat_b200.h")).read()
    doc = hdr[hdr.index("/* key:"):hdr.index("JSB200_API int jsb200_set_option")]
    documented = dict(re.findall(r'"(\w+)" \((\d+) default', doc))
    assert keys == set(documented), (sorted(keys), sorted(documented))
    for key, default in documented.items():
        _lib.set_option(key, int(default))


def test_clear_async_argument_checks_need_no_gpu():
    """jsb200_clear_async: zero bytes is a no-op, a null buffer is an error that is raised and
    cleared (the host layer's fall-back to a framework fill relies on both)"""
    _lib.clear_async(0, 0, 0)
    with pytest.raises(RuntimeError, match="null buffer"):
        _lib.clear_async(0, 0, 16)
    assert _lib.lib().jsb200_last_error() is None


def test_ops_validate_shapes_before_touching_the_gpu():
    n = 8
    z = torch.zeros
    with pytest.raises(TypeError):
        _ops.project_fwd(z(n, 3), z(n, 3), z(n, 3), z(4, 4), num_points=n, img_shape=(32, 32),
                         f=(16, 16), c=(16, 16), glob_scale=1.0, clip_thresh=0.01, block_width=16)
    with pytest.raises(TypeError):
        _ops.rasterize_fwd(z(n, 3), z(n, 1), z(3), z(n, 2), z(n, 1), z(n, 1), z(n, 3),
                           z(n, 1, dtype=torch.int32), num_points=n, img_shape=(32, 32),
                           block_width=16)  # radii must be int32
    with pytest.raises(ValueError):  # CPU tensors are refused: there is no CPU path
        _ops.project_fwd(z(n, 3), z(n, 3), z(n, 4), z(4, 4), num_points=n, img_shape=(32, 32),
                         f=(16, 16), c=(16, 16), glob_scale=1.0, clip_thresh=0.01, block_width=16)


def test_synthetic_scenes_are_deterministic():
    a = synthetic.make_config("C1")
    b = synthetic.make_config("C1")
    assert torch.equal(a.means3d, b.means3d) and torch.equal(a.quats, b.quats)
    assert a.means3d.shape == (10_000, 3) and a.img_shape == (256, 256)
    c2 = synthetic.make_config("C2", num_points=1000)
    assert c2.means3d.shape == (1000, 3) and c2.img_shape == (512, 512)
    assert torch.allclose(c2.quats.norm(dim=1), torch.ones(1000), atol=1e-6)
    cams = synthetic.ring_cameras(4)
    for v in cams:
        R = v[:3, :3].double()
        assert torch.allclose(R @ R.T, torch.eye(3, dtype=torch.float64), atol=1e-6)
        # the origin is in front of the camera at distance ~8
        assert 7.5 < float(v[2, 3]) < 8.6


def test_product_never_imports_the_oracle():
    """the oracle is test infrastructure: nothing under jaxsplat_b200/ may reference it"""
    pkg = os.path.join(ROOT, "jaxsplat_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in text and "from oracle" not in text, f
                assert "liboracle" not in text and "oracle/" not in text, f


def test_descriptor_band_extension():
    d = _ffi.make_descriptor(10, (1080, 1920), (1.0, 1.0), (0.0, 0.0), 1.0, 0.01, 16)
    b = _ffi.with_band(d, 17, 34)
    # the band lives in the 4 pad bytes and is marked by a magic block_dim_2d.z (1 in a reference
    # descriptor, whose pad bytes are uninitialised and must therefore never be read as a band)
    assert b[:72] == d[:72]
    assert struct.unpack("<I", d[72:76])[0] == 1
    assert struct.unpack("<I", b[72:76])[0] == 0x4A534231
    assert struct.unpack("<I", b[76:])[0] == (17 | (34 << 16))
    with pytest.raises(RuntimeError):
        _ffi.with_band(d, 5, 5)


def test_error_text_is_thread_local():
    """the sticky error belongs to the calling thread (XLA drives several devices from several
    host threads): a failure on one thread is invisible to another, the global counter sees both"""
    import threading
    L = _lib.lib()
    L.jsb200_clear_error()
    before = L.jsb200_error_count()
    seen = {}

    def worker(tag, fail):
        L.jsb200_clear_error()
        if fail:
            bufs = (ctypes.c_void_p * 1)(None)
            L.jsb200_rasterize_fwd(None, bufs, b"y" * 7, 7)
        msg = L.jsb200_last_error()
        seen[tag] = None if msg is None else bytes(msg)
        L.jsb200_clear_error()

    threads = [threading.Thread(target=worker, args=("bad", True)),
               threading.Thread(target=worker, args=("good", False))]
    for th in threads:
        th.start()
    for th in threads:
        th.join()
    assert seen["bad"] is not None and b"descriptor" in seen["bad"]
    assert seen["good"] is None
    assert L.jsb200_last_error() is None            # nor on this thread
    assert L.jsb200_error_count() == before + 1


def test_tile_row_bands_properties():
    """bands are contiguous, cover every tile row exactly once, are non-empty whenever there are
    at least as many rows as ranks, and follow the weights"""
    from hypothesis import given, settings, strategies as st
    from jaxsplat_b200 import multiview

    @settings(max_examples=200, deadline=None)
    @given(st.integers(1, 300), st.integers(1, 16), st.randoms(use_true_random=False))
    def check(tiles_y, world, rnd):
        weights = [rnd.random() * 10 + 0.01 for _ in range(tiles_y)]
        for w in (None, weights):
            bands = multiview.tile_row_bands(tiles_y, world, w)
            assert len(bands) == world
            assert bands[0][0] == 0 and bands[-1][1] == tiles_y
            for (a0, a1), (b0, b1) in zip(bands, bands[1:]):
                assert a1 == b0 and a0 <= a1
            if tiles_y >= world:
                assert all(b1 > b0 for b0, b1 in bands)
    check()
    # heavy rows in the middle: the band holding them is the narrowest
    w = [1.0] * 30
    w[14] = w[15] = 100.0
    bands = multiview.tile_row_bands(30, 3, w)
    widths = [b1 - b0 for b0, b1 in bands]
    assert widths[1] == min(widths)


def test_adam_rejects_host_tensors_and_mismatched_gradient_buffers():
    """optim.Adam is CUDA-only like the rest of the product (no CPU fallback): host tensors and
    ill-sized fixed gradient buffers are refused before anything native is called"""
    import torch

    from jaxsplat_b200.optim import Adam
    with pytest.raises(TypeError):
        Adam([torch.zeros(4)])
    with pytest.raises((TypeError, ValueError)):
        Adam([], grads=[torch.zeros(4)])
