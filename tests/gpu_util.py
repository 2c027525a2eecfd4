"""helpers shared by the GPU tests: run the four primitives of a given native library on torch
CUDA tensors built from numpy inputs."""
import ctypes

import numpy as np
import torch

from jaxsplat_b200 import _ffi, _lib, _ops

DEV = "cuda:0"


def t(a, dtype=None):
    x = torch.from_numpy(np.ascontiguousarray(a)).to(DEV)
    return x if dtype is None else x.to(dtype)


def npy(x):
    return x.detach().cpu().numpy()


def bits_equal(a, b):
    """bit equality of float arrays, treating +0 and -0 as equal"""
    a = np.ascontiguousarray(a, np.float32).reshape(-1)
    b = np.ascontiguousarray(b, np.float32).reshape(-1)
    bad = (a.view(np.int32) != b.view(np.int32)) & ~((a == 0) & (b == 0))
    return int(bad.sum())


def project_fwd(inp, kw, table=None):
    n = inp["means3d"].shape[0]
    names = ("cov3ds", "xys", "depths", "radii", "conics", "compensation", "num_tiles_hit",
             "cum_tiles_hit")
    outs = _ops.project_fwd(t(inp["means3d"]), t(inp["scales"]), t(inp["quats"]),
                            t(inp["viewmat"]), num_points=n, table=table, **kw)
    return dict(zip(names, outs))


def rasterize_fwd(inp, P, rkw, table=None):
    n = inp["colors"].shape[0]
    outs = _ops.rasterize_fwd(t(inp["colors"]), t(inp["opacities"]), t(inp["background"]),
                              P["xys"], P["depths"], P["radii"], P["conics"], P["cum_tiles_hit"],
                              num_points=n, table=table, **rkw)
    return dict(zip(("final_Ts", "final_idx", "img"), outs))


def rasterize_bwd(inp, P, R, v_img, v_alpha, rkw, table=None):
    n = inp["colors"].shape[0]
    outs = _ops.rasterize_bwd(t(inp["colors"]), t(inp["opacities"]), t(inp["background"]),
                              P["xys"], P["depths"], P["radii"], P["conics"], P["cum_tiles_hit"],
                              R["final_Ts"], R["final_idx"], v_img, v_alpha, num_points=n,
                              table=table, **rkw)
    return dict(zip(("v_colors", "v_opacity", "v_xy", "v_xy_abs", "v_conic"), outs))


def project_bwd(inp, P, v_comp, v_xy, v_depth, v_conic, kw, table=None):
    n = inp["means3d"].shape[0]
    outs = _ops.project_bwd(t(inp["means3d"]), t(inp["scales"]), t(inp["quats"]),
                            t(inp["viewmat"]), P["cov3ds"], P["xys"], P["radii"], P["conics"],
                            P["compensation"], v_comp, v_xy, v_depth, v_conic, num_points=n,
                            table=table, **kw)
    return dict(zip(("v_mean3d", "v_scale", "v_quat", "v_cov2d", "v_cov3d"), outs))


def golden_inputs(g):
    return {k: g["in_" + k] for k in ("means3d", "scales", "quats", "colors", "opacities",
                                      "viewmat", "background")}


def to_dev(d, keys):
    return {k: t(d[k]) for k in keys}


def debug_binning(num_intersects, num_tiles, depths=None):
    """sorted ids / tile bins / reference-style 64-bit keys of the last rasterize call"""
    ids = torch.zeros(max(num_intersects, 1), dtype=torch.int32, device=DEV)
    bins = torch.zeros((num_tiles, 2), dtype=torch.int32, device=DEV)
    keys = torch.zeros(max(num_intersects, 1), dtype=torch.int64, device=DEV) if depths is not None else None
    got = _lib.lib().jsb200_debug_binning(
        ctypes.c_void_p(torch.cuda.current_stream().cuda_stream),
        None if depths is None else ctypes.c_void_p(depths.data_ptr()),
        ctypes.c_void_p(ids.data_ptr()), ctypes.c_void_p(bins.data_ptr()),
        None if keys is None else ctypes.c_void_p(keys.data_ptr()))
    torch.cuda.synchronize()
    if got < 0:
        _lib.check_error("debug_binning")
    return int(got), ids[:num_intersects], bins, (None if keys is None else keys[:num_intersects])


def scene_inputs(scene):
    return {k: getattr(scene, k).numpy() for k in ("means3d", "scales", "quats", "colors",
                                                   "opacities", "viewmat", "background")}


def scene_kw(scene):
    kw = dict(img_shape=scene.img_shape, f=scene.f, c=scene.c, glob_scale=scene.glob_scale,
              clip_thresh=scene.clip_thresh, block_width=scene.block_size)
    rkw = dict(img_shape=scene.img_shape, block_width=scene.block_size)
    return kw, rkw
