"""GPU: CUDA path vs the CPU oracle on seeded scenes the oracle finishes in seconds, plus the
edge cases the domain has (empty frames, ragged sizes, culled Gaussians, every block size) and
the library's own invariants (culling and caching never change a result)."""
import ctypes
import os

import numpy as np
import pytest
import torch

from conftest import grad_close
import gpu_util as G
from jaxsplat_b200 import _lib, _ops, synthetic
from oracle import cpu

pytestmark = pytest.mark.gpu


def _full_gpu(inp, kw, rkw, target, v_alpha=None):
    P = G.project_fwd(inp, kw)
    R = G.rasterize_fwd(inp, P, rkw)
    H, W = rkw["img_shape"]
    v_img = ((2.0 / (H * W * 3)) * (R["img"] - G.t(target))).contiguous()
    va = torch.zeros((H, W, 1), device=G.DEV) if v_alpha is None else G.t(v_alpha)
    B = G.rasterize_bwd(inp, P, R, v_img, va, rkw)
    n = inp["means3d"].shape[0]
    z = torch.zeros((n, 1), device=G.DEV)
    Gd = G.project_bwd(inp, P, z, B["v_xy"], z.clone(), B["v_conic"], kw)
    out = {}
    for d in (P, R, B, Gd):
        out.update({k: G.npy(v) for k, v in d.items()})
    out["v_img"] = G.npy(v_img)
    return out


def _check_vs_oracle(scene, seed=5, proj_bwd_rel=1e-4):
    inp = G.scene_inputs(scene)
    kw, rkw = G.scene_kw(scene)
    H, W = scene.img_shape
    target = np.random.RandomState(seed).rand(H, W, 3).astype(np.float32)
    gpu = _full_gpu(inp, kw, rkw, target)
    # --- projection: bit-exact
    P = cpu.project_fwd(inp["means3d"], inp["scales"], inp["quats"], inp["viewmat"], **kw)
    for k in ("radii", "num_tiles_hit", "cum_tiles_hit"):
        assert np.array_equal(P[k], gpu[k]), k
    for k in ("cov3ds", "xys", "depths", "conics", "compensation"):
        assert G.bits_equal(P[k], gpu[k]) == 0, k
    # --- binning: bit-exact
    b = cpu.sort_and_bin(P["xys"], P["depths"], P["radii"], P["cum_tiles_hit"], **rkw)
    I = b["num_intersects"]
    if I:
        # the last rasterize call on the stream was the backward: same cached binning
        got, ids, bins, keys = G.debug_binning(I, b["tile_bins"].shape[0], depths=G.t(P["depths"]))
        assert got == I
        assert np.array_equal(G.npy(ids), b["ids_sorted"])
        assert np.array_equal(G.npy(bins), b["tile_bins"])
        assert np.array_equal(G.npy(keys), b["keys_sorted"])
    # --- rasteriser forward (oracle uses exp2f, GPU MUFU.EX2: bounded decision flips)
    R = cpu.rasterize_fwd(inp["colors"], inp["opacities"], inp["background"], P["xys"],
                          P["depths"], P["radii"], P["conics"], P["cum_tiles_hit"], binning=b, **rkw)
    same = (R["final_idx"] == gpu["final_idx"])[..., 0]
    assert 1.0 - same.mean() <= 2e-3
    assert np.abs(R["img"][same] - gpu["img"][same]).max(initial=0) <= 1e-5
    # --- backward, op level: feed the GPU's own forward results to the oracle
    B = cpu.rasterize_bwd(inp["colors"], inp["opacities"], inp["background"], P["xys"],
                          P["depths"], P["radii"], P["conics"], P["cum_tiles_hit"],
                          gpu["final_Ts"], gpu["final_idx"], gpu["v_img"],
                          np.zeros((H, W, 1), np.float32), binning=b, **rkw)
    for k in ("v_colors", "v_opacity", "v_xy", "v_xy_abs", "v_conic"):
        ok, err, scale = grad_close(gpu[k], B[k], rel=1e-4)
        assert ok, (k, err, scale)
    n = inp["means3d"].shape[0]
    z = np.zeros((n, 1), np.float32)
    Gd = cpu.project_bwd(inp["means3d"], inp["scales"], inp["quats"], inp["viewmat"], P["cov3ds"],
                         P["xys"], P["radii"], P["conics"], P["compensation"], z, gpu["v_xy"], z,
                         gpu["v_conic"], **kw)
    for k in ("v_mean3d", "v_scale", "v_quat", "v_cov2d", "v_cov3d"):
        ok, err, scale = grad_close(gpu[k], Gd[k], rel=proj_bwd_rel)
        assert ok, (k, err, scale)
    return gpu


def test_config1_vs_oracle():
    _check_vs_oracle(synthetic.make_config("C1"))


def test_config2_vs_oracle():
    _check_vs_oracle(synthetic.make_config("C2"))


@pytest.mark.parametrize("bw", [2, 4, 8, 16])
def test_block_sizes_vs_oracle(bw):
    s = synthetic.benchmark_scene(1200, (50, 70), scale_max=0.3, seed=21 + bw)
    s.block_size = bw
    _check_vs_oracle(s)


@pytest.mark.parametrize("n", [1, 255, 256, 257, 4097])
def test_ragged_counts(n):
    _check_vs_oracle(synthetic.benchmark_scene(n, (33, 47), scale_max=0.4, seed=n))


def test_single_gaussian_on_axis_closed_form():
    """one isotropic Gaussian on the optical axis: xy = c, conic = 1/(f^2 s^2/z^2 + 0.3) I"""
    s = synthetic.benchmark_scene(1, (64, 64), seed=0)
    s.means3d[:] = 0
    s.scales[:] = 0.2
    s.quats[:] = torch.tensor([1.0, 0, 0, 0])
    inp = G.scene_inputs(s)
    kw, rkw = G.scene_kw(s)
    P = G.project_fwd(inp, kw)
    z, f = 8.0, 32.0
    var = (f * 0.2 / z) ** 2 + 0.3
    assert np.allclose(G.npy(P["xys"])[0], [32.0, 32.0], atol=1e-4)
    assert np.allclose(G.npy(P["conics"])[0], [1 / var, 0, 1 / var], rtol=1e-5, atol=1e-7)
    # helpers.h:51-55: eigenvalue spread is floored at sqrt(0.1) even for an isotropic splat
    assert int(P["radii"][0]) == int(np.ceil(3 * np.sqrt(var + np.sqrt(0.1))))
    assert float(P["depths"][0]) == 8.0


def test_empty_frame():
    """all Gaussians behind the camera -> image = background, final_Ts = 0, zero grads (Q5)"""
    s = synthetic.benchmark_scene(500, (40, 56), seed=3)
    s.viewmat = synthetic._viewmat_identity(-30.0)
    s.background = torch.tensor([0.25, 0.5, 0.75])
    inp = G.scene_inputs(s)
    kw, rkw = G.scene_kw(s)
    _ops.poison_outputs(True)
    try:
        out = _full_gpu(inp, kw, rkw, np.zeros((40, 56, 3), np.float32))
    finally:
        _ops.poison_outputs(False)
    assert not out["radii"].any() and not out["cum_tiles_hit"].any() and not out["cov3ds"].any()
    assert np.array_equal(out["img"], np.broadcast_to(inp["background"], (40, 56, 3)))
    assert not out["final_Ts"].any() and not out["final_idx"].any()
    for k in ("v_colors", "v_opacity", "v_xy", "v_xy_abs", "v_conic", "v_mean3d", "v_scale", "v_quat"):
        assert not out[k].any(), k


def _adversarial_scene():
    """thin, long, diagonal splats + tiny/huge opacities: the hard cases for the conservative
    warp-rectangle pre-filter"""
    g = torch.Generator().manual_seed(77)
    n = 6000
    s = synthetic.benchmark_scene(n, (160, 208), scale_max=0.3, seed=78)
    s.scales[:, 0] = torch.rand(n, generator=g) * 1.5 + 0.05      # long axis
    s.scales[:, 1] = torch.rand(n, generator=g) * 0.004 + 0.0002  # needle thin
    s.scales[:, 2] = torch.rand(n, generator=g) * 0.004 + 0.0002
    s.opacities[: n // 3] = torch.rand((n // 3, 1), generator=g) * 0.01  # around the 1/255 cut
    s.opacities[n // 3: n // 2] = 1.0
    return s


@pytest.mark.parametrize("scene_fn", [_adversarial_scene, lambda: synthetic.make_config("C2", num_points=30000)])
def test_culling_never_changes_a_bit(scene_fn):
    s = scene_fn()
    inp = G.scene_inputs(s)
    kw, rkw = G.scene_kw(s)
    H, W = s.img_shape
    target = np.random.RandomState(1).rand(H, W, 3).astype(np.float32)
    try:
        _lib.set_option("cull", 1)
        a = _full_gpu(inp, kw, rkw, target)
        _lib.set_option("cull", 0)
        b = _full_gpu(inp, kw, rkw, target)
    finally:
        _lib.set_option("cull", 1)
    for k in ("final_idx",):
        assert np.array_equal(a[k], b[k]), k
    for k in ("final_Ts", "img"):
        assert G.bits_equal(a[k], b[k]) == 0, k
    for k in ("v_colors", "v_opacity", "v_xy", "v_xy_abs", "v_conic"):
        ok, err, scale = grad_close(a[k], b[k], rel=2e-5)
        assert ok, (k, err, scale)


def test_adversarial_scene_vs_oracle():
    # needle splats make the EWA VJP ill-conditioned (large cancelling terms): two float32
    # evaluation orders of the same formula differ by up to ~1e-2 of the largest gradient (v_quat:
    # 0.5e-2 .. 1.1e-2 measured, depending on the last bits of the incoming v_conic)
    _check_vs_oracle(_adversarial_scene(), proj_bwd_rel=2e-2)


def test_backward_cache_on_off_identical_binning():
    s = synthetic.make_config("C1")
    inp = G.scene_inputs(s)
    kw, rkw = G.scene_kw(s)
    target = np.zeros((256, 256, 3), np.float32)
    before = _lib.stats()
    a = _full_gpu(inp, kw, rkw, target)
    mid = _lib.stats()
    assert mid["binning_cache_hits"] == before["binning_cache_hits"] + 1  # bwd reused the fwd sort
    assert mid["binning_runs"] == before["binning_runs"] + 1
    try:
        _lib.set_option("cache", 0)
        b = _full_gpu(inp, kw, rkw, target)
    finally:
        _lib.set_option("cache", 1)
    after = _lib.stats()
    assert after["binning_runs"] == mid["binning_runs"] + 2  # fwd and bwd both binned
    for k in ("v_colors", "v_opacity", "v_xy", "v_conic", "v_mean3d"):
        ok, err, scale = grad_close(a[k], b[k], rel=2e-5)
        assert ok, (k, err, scale)


def test_backward_rebins_when_inputs_changed():
    """a backward whose inputs do not match the cached forward must not reuse its binning"""
    s1 = synthetic.benchmark_scene(3000, (64, 64), seed=1)
    s2 = synthetic.benchmark_scene(3000, (64, 64), seed=2)
    kw, rkw = G.scene_kw(s1)
    i1, i2 = G.scene_inputs(s1), G.scene_inputs(s2)
    P1, P2 = G.project_fwd(i1, kw), G.project_fwd(i2, kw)
    R1 = G.rasterize_fwd(i1, P1, rkw)
    R2 = G.rasterize_fwd(i2, P2, rkw)          # cache now holds scene 2
    v = torch.ones((64, 64, 3), device=G.DEV) * 1e-3
    va = torch.zeros((64, 64, 1), device=G.DEV)
    B1 = G.rasterize_bwd(i1, P1, R1, v, va, rkw)  # must re-bin for scene 1
    ref = cpu.rasterize_bwd(i1["colors"], i1["opacities"], i1["background"], G.npy(P1["xys"]),
                            G.npy(P1["depths"]), G.npy(P1["radii"]), G.npy(P1["conics"]),
                            G.npy(P1["cum_tiles_hit"]), G.npy(R1["final_Ts"]),
                            G.npy(R1["final_idx"]), G.npy(v), G.npy(va), **rkw)
    for k in ("v_colors", "v_opacity", "v_xy", "v_conic"):
        ok, err, scale = grad_close(G.npy(B1[k]), ref[k], rel=1e-4)
        assert ok, (k, err, scale)


def test_cumsum_and_sort_and_bin_entry_points():
    L = _lib.lib()
    x = torch.randint(0, 50, (100_003,), dtype=torch.int32, device=G.DEV)
    y = torch.empty_like(x)
    st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    assert L.jsb200_cumsum(st, x.data_ptr(), y.data_ptr(), x.numel()) == 0
    assert torch.equal(y, torch.cumsum(x, 0).to(torch.int32))
    s = synthetic.make_config("C1")
    inp = G.scene_inputs(s)
    kw, rkw = G.scene_kw(s)
    P = G.project_fwd(inp, kw)
    I = int(P["cum_tiles_hit"][-1])
    from jaxsplat_b200 import _ffi
    desc = _ffi.make_descriptor(10000, s.img_shape, (0, 0), (0, 0), 0, 0, 16)
    ids = torch.empty(I, dtype=torch.int32, device=G.DEV)
    bins = torch.empty((256, 2), dtype=torch.int32, device=G.DEV)
    rc = L.jsb200_sort_and_bin(st, desc, 80, I, P["xys"].data_ptr(), P["depths"].data_ptr(),
                               P["radii"].data_ptr(), P["cum_tiles_hit"].data_ptr(),
                               ids.data_ptr(), bins.data_ptr())
    assert rc == 0
    torch.cuda.synchronize()
    b = cpu.sort_and_bin(G.npy(P["xys"]), G.npy(P["depths"]), G.npy(P["radii"]),
                         G.npy(P["cum_tiles_hit"]), **rkw)
    assert np.array_equal(G.npy(ids), b["ids_sorted"]) and np.array_equal(G.npy(bins), b["tile_bins"])


def _binning_of(scene, band=None):
    """sorted ids + tile bins the library builds for `scene` (forward rasterisation only)"""
    inp = G.scene_inputs(scene)
    kw, rkw = G.scene_kw(scene)
    P = G.project_fwd(inp, kw)
    n = inp["colors"].shape[0]
    outs = _ops.rasterize_fwd(G.t(inp["colors"]), G.t(inp["opacities"]), G.t(inp["background"]),
                              P["xys"], P["depths"], P["radii"], P["conics"], P["cum_tiles_hit"],
                              num_points=n, band=band, **rkw)
    I = _lib.stats()["last_num_intersects"]
    H, W = scene.img_shape
    bw = scene.block_size
    tiles = ((W + bw - 1) // bw) * ((H + bw - 1) // bw)
    if I == 0:  # nothing visible: no binning is kept (the forward only fills the background)
        return 0, np.zeros(0, np.int32), np.zeros((tiles, 2), np.int32), [G.npy(o) for o in outs]
    got, ids, bins, _ = G.debug_binning(I, tiles)
    return I, G.npy(ids), G.npy(bins), [G.npy(o) for o in outs]


@pytest.mark.parametrize("case", ["C1", "wide_boxes", "bw8", "ragged", "band", "tall_boxes"])
def test_row_column_lists_equal_the_counting_distribution(case):
    """the two list builders (rowcol_lists.cu, tile_lists.cu) must agree bit for bit"""
    band = None
    if case == "C1":
        s = synthetic.make_config("C1")
    elif case == "wide_boxes":   # many tiles per Gaussian: staging overflow + small chunks
        s = synthetic.benchmark_scene(20000, (272, 400), scale_max=1.5, seed=9)
    elif case == "bw8":
        s = synthetic.benchmark_scene(30000, (200, 312), scale_max=0.3, seed=4)
        s.block_size = 8
    elif case == "ragged":
        s = synthetic.benchmark_scene(777, (90, 50), scale_max=0.4, seed=6)
    elif case == "band":
        s = synthetic.benchmark_scene(20000, (320, 256), scale_max=0.4, seed=8)
        band = (5, 13)
    else:  # tall boxes: more row entries per chunk than the level-1 staging area holds
        s = synthetic.benchmark_scene(4000, (1024, 64), scale_max=2.0, seed=10)
    try:
        _lib.set_option("rowcol", 1)
        Ia, ids_a, bins_a, outs_a = _binning_of(s, band)
        _lib.set_option("rowcol", 0)
        Ib, ids_b, bins_b, outs_b = _binning_of(s, band)
    finally:
        _lib.set_option("rowcol", 1)
    assert Ia == Ib and Ia > 0
    assert np.array_equal(bins_a, bins_b)
    assert np.array_equal(ids_a, ids_b)
    if band is None:
        for a, b in zip(outs_a, outs_b):
            assert np.array_equal(a, b)


@pytest.mark.parametrize("case", ["C1", "C2_30k", "one_depth", "wide_depth_range", "all_culled",
                                  "ragged", "band", "two_tiles_per_cta", "many_equal_depths",
                                  "skewed_depths", "half_culled"])
def test_resident_depth_sort_builds_the_same_lists(case):
    """option "sort1k" (default): the depth order from ONE cooperative launch over the live bits of
    `key - min` (resident_sort.cu) against
    four onesweep passes over all 32 key bits: sorted ids,
    tile bins and every forward output must be equal bit for bit -- including one depth for
    everything (no pass at all), a depth range with > 27 live bits (four passes), nothing visible
    (only sentinel keys), and more Gaussians than one tile per CTA holds"""
    band = None
    if case == "C1":
        s = synthetic.make_config("C1")
    elif case == "C2_30k":
        s = synthetic.make_config("C2", num_points=30000)
    elif case == "one_depth":
        s = synthetic.benchmark_scene(5000, (160, 208), scale_max=0.3, seed=12)
        s.means3d[:, 2] = 0.25
    elif case == "wide_depth_range":  # depths 0.02 .. 3000: 28+ live bits
        s = synthetic.benchmark_scene(8000, (160, 208), scale_max=0.3, seed=13)
        z = torch.logspace(-1.7, 3.5, 8000)[torch.randperm(8000, generator=torch.Generator().manual_seed(1))]
        s.means3d[:, 2] = z - 8.0
        s.means3d[:, :2] *= (z / 8.0).unsqueeze(1)
        s.scales *= (z / 8.0).unsqueeze(1)
    elif case == "all_culled":      # everything behind the camera: only sentinel keys
        s = synthetic.benchmark_scene(3000, (96, 128), scale_max=0.3, seed=14)
        s.means3d[:, 2] = -20.0
    elif case == "ragged":
        s = synthetic.benchmark_scene(777, (90, 50), scale_max=0.4, seed=6)
    elif case == "band":
        s = synthetic.benchmark_scene(20000, (320, 256), scale_max=0.4, seed=8)
        band = (5, 13)
    elif case == "two_tiles_per_cta":  # > 148 x 8192 Gaussians: the kernel re-reads its tiles
        s = synthetic.benchmark_scene(1_300_000, (128, 160), scale_max=0.01, seed=15)
    elif case == "many_equal_depths":  # depth ties across many Gaussians: stability decides the order
        s = synthetic.benchmark_scene(40000, (160, 208), scale_max=0.2, seed=16)
        s.means3d[:, 2] = torch.round(s.means3d[:, 2] * 4.0) / 4.0
    elif case == "skewed_depths":  # 90 % of the keys inside 1/4000 of the depth range
        s = synthetic.benchmark_scene(60000, (160, 208), scale_max=0.05, seed=18)
        g = torch.Generator().manual_seed(19)
        s.means3d[:54000, 2] = 1.0 + torch.rand(54000, generator=g) * 0.002
    else:  # half of the scene behind the camera: a large sentinel bucket next to the real ones
        s = synthetic.benchmark_scene(50000, (160, 208), scale_max=0.1, seed=20)
        s.means3d[::2, 2] = -30.0
    try:
        _lib.set_option("sort1k", 0)           # four onesweep passes over all 32 key bits
        Ib, ids_b, bins_b, outs_b = _binning_of(s, band)
        _lib.set_option("sort1k", 1)
        Ia, ids_a, bins_a, outs_a = _binning_of(s, band)
        assert Ia == Ib and (Ia > 0 or case == "all_culled")
        assert np.array_equal(bins_a, bins_b)
        assert np.array_equal(ids_a, ids_b)
        if band is None:
            for a, b in zip(outs_a, outs_b):
                assert np.array_equal(a, b)
    finally:
        _lib.set_option("sort1k", 1)


def _long_list_scene():
    """tile lists longer than one span of 32 hit words (1024 entries) that do not saturate: small
    opacities, many splats per tile"""
    s = synthetic.benchmark_scene(60000, (64, 96), scale_max=0.25, seed=31)
    s.opacities *= 0.02
    return s


@pytest.mark.parametrize("scene_fn", [_adversarial_scene, _long_list_scene,
                                      lambda: synthetic.make_config("C2", num_points=30000),
                                      lambda: synthetic.benchmark_scene(9000, (200, 120), scale_max=0.2, seed=21),
                                      lambda: synthetic.benchmark_scene(60, (40, 40), scale_max=0.3, seed=3),
                                      lambda: synthetic.benchmark_scene(3000, (70, 90), scale_max=0.01, seed=5)])
def test_hit_driven_backward_equals_the_batch_kernel(scene_fn):
    """option "bwdhits" (default): one warp per 8x8 quadrant replaying only the list entries the
    forward marked as accepted (raster_bwd_hits_kernel) against the per-tile batch kernel that
    stages and cull-tests every entry: the same pairs with the same arithmetic, so only the order of
    the atomics differs (the reference's own run-to-run noise) -- all five outputs, v_xy_abs too"""
    s = scene_fn()
    inp = G.scene_inputs(s)
    kw, rkw = G.scene_kw(s)
    H, W = s.img_shape
    target = np.random.RandomState(2).rand(H, W, 3).astype(np.float32)
    v_alpha = np.random.RandomState(3).rand(H, W, 1).astype(np.float32) * 1e-3
    try:
        _lib.set_option("bwdhits", 1)
        a = _full_gpu(inp, kw, rkw, target, v_alpha)
        _lib.set_option("bwdhits", 0)
        b = _full_gpu(inp, kw, rkw, target, v_alpha)
    finally:
        _lib.set_option("bwdhits", 1)
    for k in ("v_colors", "v_opacity", "v_xy", "v_xy_abs", "v_conic"):
        ok, err, scale = grad_close(a[k], b[k], rel=2e-5)
        assert ok, (k, err, scale)


def test_hit_driven_backward_through_the_fused_render_and_bands():
    """the fused render (no v_xy_abs, token-validated binning) and a tile-row band with hit words"""
    import jaxsplat_b200 as jsb
    s = synthetic.make_config("C2", num_points=40000)
    sc = s.to(G.DEV)
    H, W = s.img_shape
    target = torch.rand((H, W, 3), generator=torch.Generator().manual_seed(9)).to(G.DEV)
    res = {}
    # (hit-driven, raw moments finished by the projection backward) / (hit-driven, finished
    # gradients from the rasteriser) / (batch kernel)
    for mode in ((1, 1), (1, 0), (0, 1)):
        try:
            _lib.set_option("bwdhits", mode[0])
            _lib.set_option("bwdmom", mode[1])
            for rows in (None, (5, 19)):
                params = [p.clone().requires_grad_(True) for p in sc.params()]
                img = jsb.render(*params, tile_rows=rows, **sc.render_kwargs())
                y0, y1 = (0, H) if rows is None else (rows[0] * 16, min(H, rows[1] * 16))
                (((img[y0:y1] - target[y0:y1]) ** 2).sum() / (H * W * 3)).backward()
                res[(mode, rows)] = [G.npy(p.grad) for p in params]
        finally:
            _lib.set_option("bwdhits", 1)
            _lib.set_option("bwdmom", 1)
    for rows in (None, (5, 19)):
        for mode in ((1, 1), (1, 0)):
            for a, b in zip(res[(mode, rows)], res[((0, 1), rows)]):
                ok, err, scale = grad_close(a, b, rel=2e-5)
                assert ok, (mode, rows, err, scale)


@pytest.mark.parametrize("scene_fn", [_adversarial_scene, _long_list_scene,
                                      lambda: synthetic.make_config("C2", num_points=30000),
                                      lambda: synthetic.benchmark_scene(9000, (200, 120), scale_max=0.2, seed=21),
                                      lambda: synthetic.benchmark_scene(60, (40, 40), scale_max=0.3, seed=3),
                                      lambda: synthetic.make_config("C1")])
def test_barrier_free_forward_is_bit_identical_to_the_batch_kernel(scene_fn):
    """option "fwdwarp" (default): four independent quadrant warps per tile, each gathering,
    cull-testing and compacting the list on its own (raster_fwd_warp_kernel), against the per-CTA
    batch kernel: the same pairs in the same rounding order, so final_idx, final_Ts and the image
    may not differ in a bit -- with and without culling -- and the hit words they leave drive the
    backward to the same gradients"""
    s = scene_fn()
    inp = G.scene_inputs(s)
    kw, rkw = G.scene_kw(s)
    H, W = s.img_shape
    P = G.project_fwd(inp, kw)
    target = np.random.RandomState(2).rand(H, W, 3).astype(np.float32)
    for cull in (1, 0):
        try:
            _lib.set_option("cull", cull)
            _lib.set_option("fwdwarp", 1)
            a = G.rasterize_fwd(inp, P, rkw)
            _lib.set_option("fwdwarp", 0)
            b = G.rasterize_fwd(inp, P, rkw)
        finally:
            _lib.set_option("fwdwarp", 1)
            _lib.set_option("cull", 1)
        assert torch.equal(a["final_idx"], b["final_idx"])
        assert torch.equal(a["final_Ts"].view(torch.int32), b["final_Ts"].view(torch.int32))
        assert torch.equal(a["img"].view(torch.int32), b["img"].view(torch.int32))
    try:
        _lib.set_option("fwdwarp", 1)
        ga = _full_gpu(inp, kw, rkw, target)
        _lib.set_option("fwdwarp", 0)
        gb = _full_gpu(inp, kw, rkw, target)
    finally:
        _lib.set_option("fwdwarp", 1)
    for k in ("v_colors", "v_opacity", "v_xy", "v_xy_abs", "v_conic"):
        ok, err, scale = grad_close(ga[k], gb[k], rel=2e-5)
        assert ok, (k, err, scale)


@pytest.mark.parametrize("case", ["C2_30k", "wide_boxes", "ragged", "band", "narrow_grid", "tiny",
                                  "long_lists", "adversarial"])
def test_super_column_lists_render_the_same_image_and_gradients(case):
    """option "supercols" (default 2): the fused render's level 2 distributes 8-byte row entries
    into super-columns of 2^k tiles and the rasterisers keep the entries whose tile columns cover
    their tile -- the tile's own list in the same order, so the image may not differ in a bit from
    the exact per-tile lists (supercols 0) for any k, and the gradients only by atomic order; also
    when the backward has to rebuild the lists itself (cache off: it replays every entry)"""
    import jaxsplat_b200 as jsb
    rows = None
    if case == "C2_30k":
        s = synthetic.make_config("C2", num_points=30000)
    elif case == "wide_boxes":   # boxes spanning several super-columns
        s = synthetic.benchmark_scene(20000, (272, 400), scale_max=1.5, seed=9)
    elif case == "ragged":
        s = synthetic.benchmark_scene(777, (90, 50), scale_max=0.4, seed=6)
    elif case == "band":
        s = synthetic.benchmark_scene(20000, (320, 256), scale_max=0.4, seed=8)
        rows = (5, 13)
    elif case == "narrow_grid":  # four tiles per row: one super-column at k >= 2
        s = synthetic.benchmark_scene(4000, (1024, 64), scale_max=2.0, seed=10)
    elif case == "tiny":
        s = synthetic.benchmark_scene(60, (40, 40), scale_max=0.3, seed=3)
    elif case == "long_lists":
        s = _long_list_scene()
    else:
        s = _adversarial_scene()
    sc = s.to(G.DEV)
    H, W = s.img_shape
    target = torch.rand((H, W, 3), generator=torch.Generator().manual_seed(9)).to(G.DEV)
    y0, y1 = (0, H) if rows is None else (rows[0] * 16, min(H, rows[1] * 16))

    def run():
        params = [p.clone().requires_grad_(True) for p in sc.params()]
        img = jsb.render(*params, tile_rows=rows, **sc.render_kwargs())
        (((img[y0:y1] - target[y0:y1]) ** 2).sum() / (H * W * 3)).backward()
        return img[y0:y1].detach().clone(), [G.npy(p.grad) for p in params]

    try:
        _lib.set_option("supercols", 0)
        img0, g0 = run()
        # atomic-order noise of the exact lists themselves (needle-thin splats amplify it through
        # the conic VJP of the projection backward): the yardstick for the comparison below
        noise = [0.0] * len(g0)
        for _ in range(3):
            _img0b, g0b = run()
            noise = [max(nz, float(np.abs(a.astype(np.float64) - b.astype(np.float64)).max()))
                     for nz, a, b in zip(noise, g0, g0b)]
        for k in (1, 2, 3, 5):
            _lib.set_option("supercols", k)
            for cache in (1, 0):
                _lib.set_option("cache", cache)
                img, g = run()
                assert torch.equal(img.view(torch.int32), img0.view(torch.int32)), (k, cache)
                for a, b, nz in zip(g, g0, noise):
                    ok, err, scale = grad_close(a, b, rel=2e-5)
                    # (needle-thin splats: the conic VJP amplifies any change of summation order, not
                    # only the run-to-run one the yardstick sees)
                    loose = 5e-3 * scale if case == "adversarial" else 0.0
                    assert ok or err <= max(16.0 * nz, loose), (k, cache, err, scale, nz)
    finally:
        _lib.set_option("supercols", 2)
        _lib.set_option("cache", 1)


@pytest.mark.parametrize("case", ["C2_30k", "wide_boxes", "tiny", "block8"])
def test_programmatic_dependent_launch_changes_nothing(case):
    """option "pdl": the step's kernels may become resident while their predecessor drains and wait
    on it in the kernel (griddepcontrol.wait before the first global access).  Several steps
    enqueued back to back without a host sync, each on perturbed parameters so that a kernel that
    ran ahead of its inputs would see the previous step's buffers: images bit for bit, gradients to
    atomic-order noise, against the plain launches."""
    import os
    import jaxsplat_b200 as jsb
    if case == "C2_30k":
        s = synthetic.make_config("C2", num_points=30000)
    elif case == "wide_boxes":
        s = synthetic.benchmark_scene(20000, (272, 400), scale_max=1.5, seed=9)
    elif case == "block8":
        s = synthetic.benchmark_scene(5000, (200, 120), scale_max=0.5, seed=12)
        s.block_size = 8
    else:
        s = synthetic.benchmark_scene(60, (40, 40), scale_max=0.3, seed=3)
    sc = s.to(G.DEV)
    H, W = s.img_shape
    target = torch.rand((H, W, 3), generator=torch.Generator().manual_seed(9)).to(G.DEV)
    base = [p.clone() for p in sc.params()]
    shifts = [torch.tensor([0.01 * k, -0.02 * k, 0.015 * k], device=G.DEV) for k in range(4)]

    def run():
        out = []
        for sh in shifts:  # (no sync inside the loop)
            params = [p.clone().requires_grad_(True) for p in base]
            with torch.no_grad():
                params[0] += sh
            img = jsb.render(*params, **sc.render_kwargs())
            (((img - target) ** 2).sum() / (H * W * 3)).backward()
            out.append((img.detach(), [p.grad for p in params]))
        torch.cuda.synchronize()
        return [(i.clone(), [G.npy(g) for g in gs]) for i, gs in out]

    try:
        _lib.set_option("pdl", 0)
        ref = run()
        ref2 = run()
        _lib.set_option("pdl", 1)
        for _ in range(3):
            got = run()
            for (img, g), (img0, g0), (_i, g0b) in zip(got, ref, ref2):
                assert torch.equal(img.view(torch.int32), img0.view(torch.int32))
                for a, b, b2 in zip(g, g0, g0b):
                    nz = float(np.abs(b.astype(np.float64) - b2.astype(np.float64)).max())
                    ok, err, scale = grad_close(a, b, rel=2e-5)
                    assert ok or err <= 16.0 * nz, (err, scale, nz)
    finally:
        _lib.set_option("pdl", int(os.environ.get("JSB_PDL", "1")))


@pytest.mark.parametrize("case", ["C2_30k", "wide_boxes", "tiny"])
def test_accumulators_cleared_on_a_side_stream_during_the_forward(case, monkeypatch):
    """render()'s backward accumulates v_colors / v_opacity / v_xy / v_conic with atomics; the host
    layer clears them on a side stream while the forward runs and calls
    jsb200_render_bwd_precleared (no memset in front of the rasteriser backward).  Same gradients
    as with the library's own memset (atomic-order noise), also for several steps in flight, a
    second backward through the same graph (the accumulators are dirty then: ordinary path) and a
    forward whose backward never runs."""
    import jaxsplat_b200 as jsb
    from jaxsplat_b200 import _render
    if case == "C2_30k":
        s = synthetic.make_config("C2", num_points=30000)
    elif case == "wide_boxes":
        s = synthetic.benchmark_scene(20000, (272, 400), scale_max=1.5, seed=9)
    else:
        s = synthetic.benchmark_scene(60, (40, 40), scale_max=0.3, seed=3)
    monkeypatch.setattr(_render, "_PRECLEAR_MIN_BYTES", 0)
    sc = s.to(G.DEV)
    H, W = s.img_shape
    target = torch.rand((H, W, 3), generator=torch.Generator().manual_seed(9)).to(G.DEV)

    def run(steps=3, twice=False):
        out = []
        for k in range(steps):  # (no sync inside the loop)
            params = [p.clone().requires_grad_(True) for p in sc.params()]
            img = jsb.render(*params, **sc.render_kwargs())
            loss = ((img - target) ** 2).sum() / (H * W * 3)
            if twice:
                loss.backward(retain_graph=True)
                for p in params:
                    p.grad = None
            loss.backward()
            out.append([p.grad for p in params])
            # a forward that is never differentiated, and garbage written over freed memory
            dead = jsb.render(*[p.clone().requires_grad_(True) for p in sc.params()], **sc.render_kwargs())
            del dead
            torch.full((9 * params[0].shape[0],), float("nan"), device=G.DEV)
        torch.cuda.synchronize()
        return [[G.npy(g) for g in gs] for gs in out]

    try:
        jsb.set_preclear(False)
        ref = run()
        ref2 = run()
        jsb.set_preclear(True)
        for twice in (False, True):
            for gs in run(twice=twice):
                for a, b, b2 in zip(gs, ref[0], ref2[0]):
                    assert np.isfinite(a).all()
                    nz = float(np.abs(b.astype(np.float64) - b2.astype(np.float64)).max())
                    ok, err, scale = grad_close(a, b, rel=2e-5)
                    assert ok or err <= 16.0 * nz, (twice, err, scale, nz)
    finally:
        jsb.set_preclear(True)


@pytest.mark.parametrize("case", ["C2_30k", "ragged", "long_lists", "adversarial", "tiny"])
def test_block_size_8_tuned_kernels_equal_the_generic_ones(case):
    """block_size 8 (reference lib/ffi.cu:29-36 runs one kernel for any block width): a tile is one
    8x8 quadrant, so the barrier-free forward and the hit-driven backward run it with one warp per
    tile.  Against the generic one-thread-per-pixel kernels (options "fwdwarp" 0 / "bwdhits" 0):
    final_idx, final_Ts and the image bit for bit, all five rasteriser gradients to 1e-4; and the
    fused render (no v_xy_abs, raw moments) with and without the forward's hit
    words (cache off: the backward rebuilds the lists and replays every entry)."""
    import jaxsplat_b200 as jsb
    if case == "C2_30k":
        s = synthetic.make_config("C2", num_points=30000)
    elif case == "ragged":
        s = synthetic.benchmark_scene(777, (90, 50), scale_max=0.4, seed=6)
    elif case == "long_lists":
        s = _long_list_scene()
    elif case == "adversarial":
        s = _adversarial_scene()
    else:
        s = synthetic.benchmark_scene(60, (40, 40), scale_max=0.3, seed=3)
    s.block_size = 8
    inp = G.scene_inputs(s)
    kw, rkw = G.scene_kw(s)
    H, W = s.img_shape
    target = np.random.RandomState(2).rand(H, W, 3).astype(np.float32)
    v_alpha = np.random.RandomState(3).rand(H, W, 1).astype(np.float32) * 1e-3
    try:
        a = _full_gpu(inp, kw, rkw, target, v_alpha)
        _lib.set_option("fwdwarp", 0)
        _lib.set_option("bwdhits", 0)
        b = _full_gpu(inp, kw, rkw, target, v_alpha)
    finally:
        _lib.set_option("fwdwarp", 1)
        _lib.set_option("bwdhits", 1)
    assert np.array_equal(a["final_idx"], b["final_idx"])
    assert G.bits_equal(a["final_Ts"], b["final_Ts"]) == 0
    assert G.bits_equal(a["img"], b["img"]) == 0
    for k in ("v_colors", "v_opacity", "v_xy", "v_xy_abs", "v_conic"):
        # (the generic kernel differentiates pixel by pixel in scalar arithmetic, the hit-driven one
        # in regrouped packed arithmetic: the stated gradient tolerance, not atomic-order noise)
        ok, err, scale = grad_close(a[k], b[k], rel=1e-4)
        assert ok, (k, err, scale)
    # fused render
    sc = s.to(G.DEV)
    tgt = torch.from_numpy(target).to(G.DEV)
    res = {}
    try:
        for mode in ("tuned", "tuned_no_cache", "generic"):
            _lib.set_option("cache", 0 if mode == "tuned_no_cache" else 1)
            _lib.set_option("fwdwarp", 0 if mode == "generic" else 1)
            _lib.set_option("bwdhits", 0 if mode == "generic" else 1)
            params = [p.clone().requires_grad_(True) for p in sc.params()]
            img = jsb.render(*params, **sc.render_kwargs())
            ((img - tgt) ** 2).mean().backward()
            res[mode] = (img.detach().clone(), [G.npy(p.grad) for p in params])
    finally:
        _lib.set_option("cache", 1)
        _lib.set_option("fwdwarp", 1)
        _lib.set_option("bwdhits", 1)
    for mode in ("tuned", "tuned_no_cache"):
        assert torch.equal(res[mode][0].view(torch.int32), res["generic"][0].view(torch.int32))
        for x, y in zip(res[mode][1], res["generic"][1]):
            ok, err, scale = grad_close(x, y, rel=1e-4 if case != "adversarial" else 3e-2)  # (needle-thin splats amplify atomic-order noise)
            assert ok, (mode, err, scale)
