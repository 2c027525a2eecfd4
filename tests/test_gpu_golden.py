"""GPU: the CUDA library (through its C ABI) against the fixtures produced by the reference's
CUDA kernels.  Integer / index outputs bit-exact; forward float outputs bit-identical (we mirror
the reference's compiled rounding order); gradients within rel 1e-4 (atomic ordering)."""
import numpy as np
import pytest
import torch

from conftest import grad_close
import gpu_util as G

pytestmark = pytest.mark.gpu

PROJ_FLOAT = ["cov3ds", "xys", "depths", "conics", "compensation"]
PROJ_INT = ["radii", "num_tiles_hit", "cum_tiles_hit"]


def test_project_fwd(golden):
    g = golden
    from jaxsplat_b200 import _ops
    _ops.poison_outputs(True)  # results arrive uninitialised from XLA: everything must be written
    try:
        P = G.project_fwd(G.golden_inputs(g), g["kw"])
    finally:
        _ops.poison_outputs(False)
    for k in PROJ_INT:
        assert np.array_equal(G.npy(P[k]), g[k]), k
    for k in PROJ_FLOAT:
        assert G.bits_equal(G.npy(P[k]), g[k]) == 0, k


def test_rasterize_fwd_and_binning(golden):
    g = golden
    from jaxsplat_b200 import _ops
    P = G.to_dev(g, PROJ_FLOAT + PROJ_INT)
    _ops.poison_outputs(True)
    try:
        R = G.rasterize_fwd(G.golden_inputs(g), P, g["rkw"])
    finally:
        _ops.poison_outputs(False)
    assert np.array_equal(G.npy(R["final_idx"]), g["final_idx"])
    assert np.abs(G.npy(R["img"]) - g["img"]).max() <= 1e-5
    assert np.abs(G.npy(R["final_Ts"]) - g["final_Ts"]).max() <= 1e-5
    # stronger than the 1e-5 contract: the forward is bit-identical
    assert G.bits_equal(G.npy(R["img"]), g["img"]) == 0
    assert G.bits_equal(G.npy(R["final_Ts"]), g["final_Ts"]) == 0
    I = int(g["cum_tiles_hit"].reshape(-1)[-1])
    if I == 0:
        return
    got, ids, bins, keys = G.debug_binning(I, g["tile_bins"].shape[0], depths=P["depths"])
    assert got == I
    assert np.array_equal(G.npy(ids), g["gaussian_ids_sorted"])
    assert np.array_equal(G.npy(bins), g["tile_bins"])
    k = G.npy(keys)
    assert np.all(k[1:] >= k[:-1])
    depth_bits = g["depths"].reshape(-1).view(np.int32).astype(np.int64)
    assert np.array_equal(k & 0xFFFFFFFF, depth_bits[g["gaussian_ids_sorted"]])


def test_rasterize_bwd(golden):
    g = golden
    P = G.to_dev(g, PROJ_FLOAT + PROJ_INT)
    R = G.to_dev(g, ["final_Ts", "final_idx"])
    B = G.rasterize_bwd(G.golden_inputs(g), P, R, G.t(g["v_img"]), G.t(g["v_img_alpha"]), g["rkw"])
    for k in ("v_colors", "v_opacity", "v_xy", "v_xy_abs", "v_conic"):
        ok, err, scale = grad_close(G.npy(B[k]), g[k], rel=1e-4)
        assert ok, (k, err, scale)


def test_project_bwd(golden):
    g = golden
    P = G.to_dev(g, PROJ_FLOAT + PROJ_INT)
    out = G.project_bwd(G.golden_inputs(g), P, G.t(g["v_compensation"]), G.t(g["v_xy"]),
                        G.t(g["v_depth"]), G.t(g["v_conic"]), g["kw"])
    for k in ("v_mean3d", "v_scale", "v_quat", "v_cov2d", "v_cov3d"):
        ok, err, scale = grad_close(G.npy(out[k]), g[k], rel=1e-4)
        assert ok, (k, err, scale)
