"""numpy/ctypes front end of the CPU oracle (oracle/oracle.c) -- TEST INFRASTRUCTURE ONLY.

Mirrors the four primitives' operand/result lists (include/jaxsplat_b200.h) on numpy arrays so
tests can compare the CUDA library, the reference CUDA build and this oracle op by op.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "liboracle.so")
_lib = None


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "liboracle.so"], stdout=subprocess.DEVNULL)
    return _SO


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        _lib = ctypes.CDLL(_SO)
        _lib.orc_sort_end_bit.restype = ctypes.c_int
        _lib.orc_num_threads.restype = ctypes.c_int
    return _lib


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def _f(a, shape=None):
    a = np.ascontiguousarray(a, dtype=np.float32)
    return a if shape is None else a.reshape(shape)


def _i(a, shape=None):
    a = np.ascontiguousarray(a, dtype=np.int32)
    return a if shape is None else a.reshape(shape)


def tiles_of(img_shape, bw):
    H, W = img_shape
    return (W + bw - 1) // bw, (H + bw - 1) // bw


def project_fwd(mean3ds, scales, quats, viewmat, *, img_shape, f, c, glob_scale, clip_thresh,
                block_width):
    n = mean3ds.shape[0]
    H, W = img_shape
    tx, ty = tiles_of(img_shape, block_width)
    out = dict(cov3ds=np.empty((n, 6), np.float32), xys=np.empty((n, 2), np.float32),
               depths=np.empty((n, 1), np.float32), radii=np.empty((n, 1), np.int32),
               conics=np.empty((n, 3), np.float32), compensation=np.empty((n, 1), np.float32),
               num_tiles_hit=np.empty((n, 1), np.int32), cum_tiles_hit=np.empty((n, 1), np.int32))
    cf = ctypes.c_float
    lib().orc_project_fwd(ctypes.c_int(n), _p(_f(mean3ds)), _p(_f(scales)), cf(glob_scale),
                          _p(_f(quats)), _p(_f(viewmat)), cf(f[0]), cf(f[1]), cf(c[0]), cf(c[1]),
                          ctypes.c_uint(W), ctypes.c_uint(H), ctypes.c_uint(tx), ctypes.c_uint(ty),
                          ctypes.c_int(block_width), cf(clip_thresh), _p(out["cov3ds"]),
                          _p(out["xys"]), _p(out["depths"]), _p(out["radii"]), _p(out["conics"]),
                          _p(out["compensation"]), _p(out["num_tiles_hit"]))
    lib().orc_cumsum(ctypes.c_int(n), _p(out["num_tiles_hit"]), _p(out["cum_tiles_hit"]))
    return out


def sort_and_bin(xys, depths, radii, cum_tiles_hit, *, img_shape, block_width):
    """ops.cu:645-793 -> dict(keys_sorted int64[I], ids_sorted int32[I], tile_bins int32[T,2])"""
    n = xys.shape[0]
    tx, ty = tiles_of(img_shape, block_width)
    cum = _i(cum_tiles_hit).reshape(-1)
    I = int(cum[-1]) if n else 0
    keys = np.zeros(max(I, 1), np.int64)
    ids = np.zeros(max(I, 1), np.int32)
    lib().orc_map_gaussian_to_intersects(ctypes.c_int(n), _p(_f(xys)), _p(_f(depths)),
                                         _p(_i(radii)), _p(cum), ctypes.c_uint(tx),
                                         ctypes.c_uint(ty), ctypes.c_int(block_width), _p(keys),
                                         _p(ids))
    ks, vs = np.zeros_like(keys), np.zeros_like(ids)
    end_bit = lib().orc_sort_end_bit(ctypes.c_uint(tx * ty))
    lib().orc_sort_pairs(ctypes.c_int64(I), _p(keys), _p(ids), _p(ks), _p(vs),
                         ctypes.c_int(end_bit))
    bins = np.zeros((tx * ty, 2), np.int32)
    lib().orc_tile_bin_edges(ctypes.c_int64(I), _p(ks), ctypes.c_uint(tx * ty), _p(bins))
    return dict(num_intersects=I, keys_unsorted=keys[:I], ids_unsorted=ids[:I], keys_sorted=ks[:I],
                ids_sorted=vs[:I], tile_bins=bins)


def rasterize_fwd(colors, opacities, background, xys, depths, radii, conics, cum_tiles_hit, *,
                  img_shape, block_width, binning=None):
    H, W = img_shape
    tx, ty = tiles_of(img_shape, block_width)
    b = binning or sort_and_bin(xys, depths, radii, cum_tiles_hit, img_shape=img_shape,
                                block_width=block_width)
    out = dict(final_Ts=np.empty((H, W, 1), np.float32), final_idx=np.empty((H, W, 1), np.int32),
               img=np.empty((H, W, 3), np.float32), binning=b)
    ids = np.ascontiguousarray(b["ids_sorted"]) if b["num_intersects"] else np.zeros(1, np.int32)
    lib().orc_rasterize_fwd(ctypes.c_uint(W), ctypes.c_uint(H), ctypes.c_uint(tx),
                            ctypes.c_uint(ty), ctypes.c_int(block_width),
                            ctypes.c_int64(b["num_intersects"]), _p(ids), _p(b["tile_bins"]),
                            _p(_f(xys)), _p(_f(conics)), _p(_f(colors)), _p(_f(opacities)),
                            _p(_f(background)), _p(out["final_Ts"]), _p(out["final_idx"]),
                            _p(out["img"]))
    return out


def rasterize_bwd(colors, opacities, background, xys, depths, radii, conics, cum_tiles_hit,
                  final_Ts, final_idx, v_img, v_img_alpha, *, img_shape, block_width,
                  binning=None):
    H, W = img_shape
    n = colors.shape[0]
    tx, ty = tiles_of(img_shape, block_width)
    b = binning or sort_and_bin(xys, depths, radii, cum_tiles_hit, img_shape=img_shape,
                                block_width=block_width)
    d = dict(v_xy=np.zeros((n, 2)), v_xy_abs=np.zeros((n, 2)), v_conic=np.zeros((n, 3)),
             v_colors=np.zeros((n, 3)), v_opacity=np.zeros((n, 1)))
    ids = np.ascontiguousarray(b["ids_sorted"]) if b["num_intersects"] else np.zeros(1, np.int32)
    lib().orc_rasterize_bwd(ctypes.c_uint(W), ctypes.c_uint(H), ctypes.c_uint(tx),
                            ctypes.c_uint(ty), ctypes.c_int(block_width), ctypes.c_int(n),
                            ctypes.c_int64(b["num_intersects"]), _p(ids), _p(b["tile_bins"]),
                            _p(_f(xys)), _p(_f(conics)), _p(_f(colors)), _p(_f(opacities)),
                            _p(_f(background)), _p(_f(final_Ts)), _p(_i(final_idx)),
                            _p(_f(v_img)), _p(_f(v_img_alpha)), _p(d["v_xy"]), _p(d["v_xy_abs"]),
                            _p(d["v_conic"]), _p(d["v_colors"]), _p(d["v_opacity"]))
    return {k: v.astype(np.float32) for k, v in d.items()}


def project_bwd(mean3ds, scales, quats, viewmat, cov3ds, xys, radii, conics, compensation,
                v_compensation, v_xy, v_depth, v_conic, *, img_shape, f, c, glob_scale,
                clip_thresh, block_width):
    n = mean3ds.shape[0]
    out = dict(v_mean3d=np.empty((n, 3), np.float32), v_scale=np.empty((n, 3), np.float32),
               v_quat=np.empty((n, 4), np.float32), v_cov2d=np.empty((n, 3), np.float32),
               v_cov3d=np.empty((n, 6), np.float32))
    cf = ctypes.c_float
    lib().orc_project_bwd(ctypes.c_int(n), _p(_f(mean3ds)), _p(_f(scales)), cf(glob_scale),
                          _p(_f(quats)), _p(_f(viewmat)), cf(f[0]), cf(f[1]), _p(_f(cov3ds)),
                          _p(_i(radii)), _p(_f(conics)), _p(_f(compensation)), _p(_f(v_xy)),
                          _p(_f(v_depth)), _p(_f(v_conic)), _p(_f(v_compensation)),
                          _p(out["v_cov2d"]), _p(out["v_cov3d"]), _p(out["v_mean3d"]),
                          _p(out["v_scale"]), _p(out["v_quat"]))
    return out


def render_fwd_bwd(scene_np: dict, target: np.ndarray):
    """full hot path for loss = mean((img - target)^2); scene_np holds numpy arrays + statics."""
    s = scene_np
    kw = dict(img_shape=s["img_shape"], f=s["f"], c=s["c"], glob_scale=s["glob_scale"],
              clip_thresh=s["clip_thresh"], block_width=s["block_size"])
    rkw = dict(img_shape=s["img_shape"], block_width=s["block_size"])
    P = project_fwd(s["means3d"], s["scales"], s["quats"], s["viewmat"], **kw)
    R = rasterize_fwd(s["colors"], s["opacities"], s["background"], P["xys"], P["depths"],
                      P["radii"], P["conics"], P["cum_tiles_hit"], **rkw)
    H, W = s["img_shape"]
    v_img = ((2.0 / (H * W * 3)) * (R["img"] - target)).astype(np.float32)
    loss = float(np.mean((R["img"].astype(np.float64) - target) ** 2))
    B = rasterize_bwd(s["colors"], s["opacities"], s["background"], P["xys"], P["depths"],
                      P["radii"], P["conics"], P["cum_tiles_hit"], R["final_Ts"], R["final_idx"],
                      v_img, np.zeros((H, W, 1), np.float32), binning=R["binning"], **rkw)
    n = s["means3d"].shape[0]
    z = np.zeros((n, 1), np.float32)
    G = project_bwd(s["means3d"], s["scales"], s["quats"], s["viewmat"], P["cov3ds"], P["xys"],
                    P["radii"], P["conics"], P["compensation"], z, B["v_xy"], z, B["v_conic"], **kw)
    out = {}
    out.update(P)
    out.update({k: v for k, v in R.items() if k != "binning"})
    out.update(B)
    out.update(G)
    out["loss"] = loss
    out["binning"] = R["binning"]
    return out


def scene_to_numpy(scene) -> dict:
    d = {}
    for k in ("means3d", "scales", "quats", "colors", "opacities", "viewmat", "background"):
        d[k] = getattr(scene, k).detach().cpu().numpy()
    for k in ("img_shape", "f", "c", "glob_scale", "clip_thresh", "block_size"):
        d[k] = getattr(scene, k)
    return d


def num_threads() -> int:
    return int(lib().orc_num_threads())
