"""The four primitives (``project_fwd/bwd``, ``rasterize_fwd/bwd``) on torch CUDA tensors.

Shapes, dtypes and operand/result orders follow the reference's abstract-eval contracts
(jaxsplat/_project/abstract.py:7-22,57-77 and jaxsplat/_rasterize/abstract.py:7-24,62-85) and
its buffer unpacking (lib/ops.cu:237-605).  ``table`` selects the native library that executes
the call; by default ours.  Result buffers are allocated uninitialised (optionally poisoned),
exactly what XLA hands a custom call.
"""
from __future__ import annotations

from typing import Optional

import torch

from . import _ffi

_POISON = False
_DEFAULT_TABLE = None


def set_default_table(table) -> None:
    """Benchmark/test hook: run the public API on another library exposing the same four
    custom-call targets (the reference's capsules).  None restores this package's own library."""
    global _DEFAULT_TABLE
    _DEFAULT_TABLE = table


def _resolve(table):
    return table or _DEFAULT_TABLE or _ffi.native_table()


def poison_outputs(enable: bool) -> None:
    """Test hook: pre-fill result buffers with NaN / 0x7f7f7f7f like an adversarial XLA."""
    global _POISON
    _POISON = bool(enable)


def _out(shape, dtype, device):
    t = torch.empty(shape, dtype=dtype, device=device)
    if _POISON:
        if dtype.is_floating_point:
            t.fill_(float("nan"))
        else:
            t.fill_(0x7F7F7F7F)
    return t


def _f32(t: torch.Tensor, shape, name: str) -> torch.Tensor:
    if t.dtype != torch.float32 or tuple(t.shape) != tuple(shape):
        raise TypeError(f"{name}: expected float32{tuple(shape)}, got {t.dtype}{tuple(t.shape)}")
    return t.contiguous()


def _f32n(t, shape, name: str):
    """a float32 tensor of `shape`, or None (lean residuals: not produced / recomputed)"""
    return None if t is None else _f32(t, shape, name)


def _i32(t: torch.Tensor, shape, name: str) -> torch.Tensor:
    if t.dtype != torch.int32 or tuple(t.shape) != tuple(shape):
        raise TypeError(f"{name}: expected int32{tuple(shape)}, got {t.dtype}{tuple(t.shape)}")
    return t.contiguous()


def _desc(table, num_points, img_shape, f, c, glob_scale, clip_thresh, block_width) -> bytes:
    return table.make_descriptor(num_points, tuple(img_shape), tuple(f), tuple(c), glob_scale,
                                 clip_thresh, block_width)


def project_fwd(mean3ds, scales, quats, viewmat, *, num_points, img_shape, f, c, glob_scale,
                clip_thresh, block_width, table: Optional[_ffi.CustomCallTable] = None):
    table = _resolve(table)
    n, dev = num_points, mean3ds.device
    ins = [_f32(mean3ds, (n, 3), "mean3ds"), _f32(scales, (n, 3), "scales"),
           _f32(quats, (n, 4), "quats"), _f32(viewmat, (4, 4), "viewmat")]
    outs = [_out((n, 6), torch.float32, dev), _out((n, 2), torch.float32, dev),
            _out((n, 1), torch.float32, dev), _out((n, 1), torch.int32, dev),
            _out((n, 3), torch.float32, dev), _out((n, 1), torch.float32, dev),
            _out((n, 1), torch.int32, dev), _out((n, 1), torch.int32, dev)]
    table.call("project_fwd", ins, outs,
               _desc(table, n, img_shape, f, c, glob_scale, clip_thresh, block_width))
    # (cov3ds, xys, depths, radii, conics, compensation, num_tiles_hit, cum_tiles_hit)
    return tuple(outs)


def project_bwd(mean3ds, scales, quats, viewmat, cov3ds, xys, radii, conics, compensation,
                v_compensation, v_xy, v_depth, v_conic, *, num_points, img_shape, f, c, glob_scale,
                clip_thresh, block_width, table: Optional[_ffi.CustomCallTable] = None):
    table = _resolve(table)
    n, dev = num_points, mean3ds.device
    ins = [_f32(mean3ds, (n, 3), "mean3ds"), _f32(scales, (n, 3), "scales"),
           _f32(quats, (n, 4), "quats"), _f32(viewmat, (4, 4), "viewmat"),
           _f32(cov3ds, (n, 6), "cov3ds"), _f32(xys, (n, 2), "xys"), _i32(radii, (n, 1), "radii"),
           _f32(conics, (n, 3), "conics"), _f32(compensation, (n, 1), "compensation"),
           _f32(v_compensation, (n, 1), "v_compensation"), _f32(v_xy, (n, 2), "v_xy"),
           _f32(v_depth, (n, 1), "v_depth"), _f32(v_conic, (n, 3), "v_conic")]
    outs = [_out((n, 3), torch.float32, dev), _out((n, 3), torch.float32, dev),
            _out((n, 4), torch.float32, dev), _out((n, 3), torch.float32, dev),
            _out((n, 6), torch.float32, dev)]
    table.call("project_bwd", ins, outs,
               _desc(table, n, img_shape, f, c, glob_scale, clip_thresh, block_width))
    # (v_mean3d, v_scale, v_quat, v_cov2d, v_cov3d)
    return tuple(outs)


def _raster_desc(table, n, img_shape, block_width, band):
    # the reference's rasterize lowerings pass zeros for f, c, glob_scale, clip_thresh
    # (jaxsplat/_rasterize/lowering.py:28-36)
    d = _desc(table, n, img_shape, (0.0, 0.0), (0.0, 0.0), 0.0, 0.0, block_width)
    if band is not None:
        d = _ffi.with_band(d, band[0], band[1])  # extension, this library only
    return d


def rasterize_fwd(colors, opacities, background, xys, depths, radii, conics, cum_tiles_hit, *,
                  num_points, img_shape, block_width, band=None, out=None,
                  table: Optional[_ffi.CustomCallTable] = None):
    table = _resolve(table)
    n, dev = num_points, colors.device
    H, W = img_shape
    ins = [_f32(colors, (n, 3), "colors"), _f32(opacities, (n, 1), "opacities"),
           _f32(background, (3,), "background"), _f32(xys, (n, 2), "xys"),
           _f32(depths, (n, 1), "depths"), _i32(radii, (n, 1), "radii"),
           _f32(conics, (n, 3), "conics"), _i32(cum_tiles_hit, (n, 1), "cum_tiles_hit")]
    # `out`: caller-provided (final_Ts, final_idx, img) so several tile-row bands can fill one frame
    if out is not None:
        outs = list(out)
    elif band is not None:
        # a band call only writes the band's pixel rows: the rest of a fresh frame is zero, not
        # uninitialised memory
        outs = [torch.zeros((H, W, 1), dtype=torch.float32, device=dev),
                torch.zeros((H, W, 1), dtype=torch.int32, device=dev),
                torch.zeros((H, W, 3), dtype=torch.float32, device=dev)]
    else:
        outs = [_out((H, W, 1), torch.float32, dev), _out((H, W, 1), torch.int32, dev),
                _out((H, W, 3), torch.float32, dev)]
    table.call("rasterize_fwd", ins, outs, _raster_desc(table, n, img_shape, block_width, band))
    # (final_Ts, final_idx, img)
    return tuple(outs)


def rasterize_bwd(colors, opacities, background, xys, depths, radii, conics, cum_tiles_hit,
                  final_Ts, final_idx, v_img, v_img_alpha, *, num_points, img_shape, block_width,
                  band=None, table: Optional[_ffi.CustomCallTable] = None):
    table = _resolve(table)
    n, dev = num_points, colors.device
    H, W = img_shape
    ins = [_f32(colors, (n, 3), "colors"), _f32(opacities, (n, 1), "opacities"),
           _f32(background, (3,), "background"), _f32(xys, (n, 2), "xys"),
           _f32(depths, (n, 1), "depths"), _i32(radii, (n, 1), "radii"),
           _f32(conics, (n, 3), "conics"), _i32(cum_tiles_hit, (n, 1), "cum_tiles_hit"),
           _f32(final_Ts, (H, W, 1), "final_Ts"), _i32(final_idx, (H, W, 1), "final_idx"),
           _f32(v_img, (H, W, 3), "v_img"), _f32(v_img_alpha, (H, W, 1), "v_img_alpha")]
    outs = [_out((n, 3), torch.float32, dev), _out((n, 1), torch.float32, dev),
            _out((n, 2), torch.float32, dev), _out((n, 2), torch.float32, dev),
            _out((n, 3), torch.float32, dev)]
    table.call("rasterize_bwd", ins, outs, _raster_desc(table, n, img_shape, block_width, band))
    # (v_colors, v_opacity, v_xy, v_xy_abs, v_conic)
    return tuple(outs)


# ---- fused render: this library only, through include/jaxsplat_b200.h ------------
def _call_native(fn_name, tensors, opaque, *extra):
    import ctypes

    from . import _lib
    for t in tensors:
        if t is not None and (not t.is_cuda or not t.is_contiguous()):
            raise ValueError(f"{fn_name}: every buffer must be a contiguous CUDA tensor")
    bufs = (ctypes.c_void_p * len(tensors))(*[None if t is None else t.data_ptr() for t in tensors])
    stream = ctypes.c_void_p(_lib.raw_stream())  # the current device's current stream
    rc = getattr(_lib.lib(), fn_name)(stream, bufs, opaque, len(opaque), *extra)
    if rc != 0:
        _lib.check_error(fn_name)
        raise RuntimeError(f"{fn_name} failed")
    _lib.check_error(fn_name)


def render_fwd(mean3ds, scales, quats, viewmat, colors, opacities, background, *, num_points,
               img_shape, f, c, glob_scale, clip_thresh, block_width, band=None, lean=False):
    """project_fwd + binning + rasterize_fwd in one native call (jsb200_render_fwd).  Returns
    ((cov3ds, xys, depths, radii, conics, compensation, final_Ts, final_idx, img), token);
    outputs are bit-identical to the two-op chain.  lean=True: cov3ds / compensation are not
    produced (None) -- ``render_bwd`` given None for cov3ds recomputes the covariance."""
    import ctypes

    n, dev = num_points, mean3ds.device
    H, W = img_shape
    ins = [_f32(mean3ds, (n, 3), "mean3ds"), _f32(scales, (n, 3), "scales"),
           _f32(quats, (n, 4), "quats"), _f32(viewmat, (4, 4), "viewmat"),
           _f32(colors, (n, 3), "colors"), _f32(opacities, (n, 1), "opacities"),
           _f32(background, (3,), "background")]
    outs = [None if lean else _out((n, 6), torch.float32, dev), _out((n, 2), torch.float32, dev),
            _out((n, 1), torch.float32, dev), _out((n, 1), torch.int32, dev),
            _out((n, 3), torch.float32, dev),
            None if lean else _out((n, 1), torch.float32, dev)]
    if band is not None:  # only the band's pixel rows are written: the rest of the frame is zero
        outs += [torch.zeros((H, W, 1), dtype=torch.float32, device=dev),
                 torch.zeros((H, W, 1), dtype=torch.int32, device=dev),
                 torch.zeros((H, W, 3), dtype=torch.float32, device=dev)]
    else:
        outs += [_out((H, W, 1), torch.float32, dev), _out((H, W, 1), torch.int32, dev),
                 _out((H, W, 3), torch.float32, dev)]
    d = _ffi.make_descriptor(n, tuple(img_shape), tuple(f), tuple(c), glob_scale, clip_thresh,
                             block_width)
    if band is not None:
        d = _ffi.with_band(d, band[0], band[1])
    token = ctypes.c_ulonglong(0)
    _call_native("jsb200_render_fwd", ins + outs, d, ctypes.byref(token))
    return tuple(outs), int(token.value)


def render_bwd(mean3ds, scales, quats, viewmat, colors, opacities, background, cov3ds, xys, depths,
               radii, conics, compensation, final_Ts, final_idx, v_img, v_img_alpha,
               *, token, num_points, img_shape, f, c, glob_scale, clip_thresh, block_width,
               band=None, grad_out=None, cleared_acc=None):
    """rasterize_bwd + project_bwd in one native call (jsb200_render_bwd).  Returns
    (v_mean3d, v_scale, v_quat, v_colors, v_opacity, v_xy, v_conic, None, None): the two
    intermediate cotangents v_cov2d / v_cov3d of the projection VJP are not materialised (the C
    entry point writes them when given buffers).
    grad_out: optional caller-owned buffers for the first five results (see
    ``jaxsplat_b200.set_gradient_sink``); a sixth tensor of 5*N floats placed right behind
    v_opacity lets the library clear its four accumulators with one memset.
    cleared_acc: a float32 tensor of 9*N zeros (ordered before this call on the current stream)
    for the four atomically accumulated results; the library then issues no memset
    (jsb200_render_bwd_precleared)."""
    import ctypes

    n, dev = num_points, mean3ds.device
    H, W = img_shape
    ins = [_f32(mean3ds, (n, 3), "mean3ds"), _f32(scales, (n, 3), "scales"),
           _f32(quats, (n, 4), "quats"), _f32(viewmat, (4, 4), "viewmat"),
           _f32(colors, (n, 3), "colors"), _f32(opacities, (n, 1), "opacities"),
           _f32(background, (3,), "background"), _f32n(cov3ds, (n, 6), "cov3ds"),
           _f32(xys, (n, 2), "xys"), _f32(depths, (n, 1), "depths"), _i32(radii, (n, 1), "radii"),
           _f32(conics, (n, 3), "conics"), _f32n(compensation, (n, 1), "compensation"),
           _f32(final_Ts, (H, W, 1), "final_Ts"),
           _i32(final_idx, (H, W, 1), "final_idx"), _f32(v_img, (H, W, 3), "v_img"),
           # None = zero cotangent for the alpha image (16x16 and 8x8 tiles: the tuned kernels)
           None if (v_img_alpha is None and block_width in (8, 16)) else _f32(
               torch.zeros((H, W, 1), dtype=torch.float32, device=dev) if v_img_alpha is None
               else v_img_alpha, (H, W, 1), "v_img_alpha")]
    if grad_out is not None:
        g = [_f32(t, shape, "grad_out") for t, shape in zip(grad_out[:5], ((n, 3), (n, 3), (n, 4),
                                                                         (n, 3), (n, 1)))]
        tail = grad_out[5] if len(grad_out) > 5 else _out((5 * n,), torch.float32, dev)
        tail = _f32(tail, (5 * n,), "grad_out tail")
        # fresh view objects: autograd may then adopt them as .grad without a copy
        # (v_cov2d / v_cov3d, the last two: None = not wanted, the kernel skips their 36 B/Gaussian)
        outs = [t.view(t.shape) for t in g] + [tail[:2 * n].view(n, 2), tail[2 * n:].view(n, 3),
                                                None, None]
    else:
        # the four atomically accumulated gradients back to back: the library clears them in one go
        acc = (_out((n, 9), torch.float32, dev).reshape(-1) if cleared_acc is None
               else _f32(cleared_acc, (9 * n,), "cleared_acc"))
        outs = [_out((n, 3), torch.float32, dev), _out((n, 3), torch.float32, dev),
                _out((n, 4), torch.float32, dev), acc[:3 * n].view(n, 3),
                acc[3 * n:4 * n].view(n, 1), acc[4 * n:6 * n].view(n, 2),
                acc[6 * n:9 * n].view(n, 3), None, None]
    d = _ffi.make_descriptor(n, tuple(img_shape), tuple(f), tuple(c), glob_scale, clip_thresh,
                             block_width)
    if band is not None:
        d = _ffi.with_band(d, band[0], band[1])
    precleared = cleared_acc is not None and grad_out is None
    _call_native("jsb200_render_bwd_precleared" if precleared else "jsb200_render_bwd", ins + outs,
                 d, ctypes.c_ulonglong(int(token)))
    return tuple(outs)


# ---- batched cameras: this library only ------------------------------------------
def render_views_fwd(mean3ds, scales, quats, viewmats, colors, opacities, background, *,
                     num_points, img_shape, f, c, glob_scale, clip_thresh, block_width, lean=False):
    """B cameras of the same Gaussians in one native call (jsb200_render_views_fwd).  Returns
    ((cov3ds, xys, depths, radii, conics, compensation, final_Ts, final_idx, imgs), tokens), every
    array with a leading view axis; view b is bit-identical to ``render_fwd`` with viewmats[b]."""
    import ctypes

    n, dev = num_points, mean3ds.device
    H, W = img_shape
    if viewmats.dim() != 3:
        raise TypeError(f"viewmats: expected float32(B, 4, 4), got {tuple(viewmats.shape)}")
    B = int(viewmats.shape[0])
    ins = [_f32(mean3ds, (n, 3), "mean3ds"), _f32(scales, (n, 3), "scales"),
           _f32(quats, (n, 4), "quats"), _f32(viewmats, (B, 4, 4), "viewmats"),
           _f32(colors, (n, 3), "colors"), _f32(opacities, (n, 1), "opacities"),
           _f32(background, (3,), "background")]
    outs = [None if lean else _out((B, n, 6), torch.float32, dev), _out((B, n, 2), torch.float32, dev),
            _out((B, n, 1), torch.float32, dev), _out((B, n, 1), torch.int32, dev),
            _out((B, n, 3), torch.float32, dev),
            None if lean else _out((B, n, 1), torch.float32, dev),
            _out((B, H, W, 1), torch.float32, dev), _out((B, H, W, 1), torch.int32, dev),
            _out((B, H, W, 3), torch.float32, dev)]
    d = _ffi.make_descriptor(n, tuple(img_shape), tuple(f), tuple(c), glob_scale, clip_thresh,
                             block_width)
    tokens = (ctypes.c_ulonglong * max(B, 1))()
    _call_native("jsb200_render_views_fwd", ins + outs, d, ctypes.c_uint(B), tokens)
    return tuple(outs), [int(t) for t in tokens[:B]]


def render_views_bwd(mean3ds, scales, quats, viewmats, colors, opacities, background, cov3ds, xys,
                     depths, radii, conics, compensation, final_Ts, final_idx, v_imgs, v_img_alphas,
                     *, tokens, num_points, img_shape, f, c, glob_scale, clip_thresh, block_width):
    """Backward of ``render_views_fwd`` in one native call (jsb200_render_views_bwd).  Returns
    (v_mean3d, v_scale, v_quat, v_colors, v_opacity) summed over the views."""
    import ctypes

    n, dev = num_points, mean3ds.device
    H, W = img_shape
    B = int(viewmats.shape[0])
    if v_img_alphas is None and block_width not in (8, 16):
        v_img_alphas = torch.zeros((B, H, W, 1), dtype=torch.float32, device=dev)
    ins = [_f32(mean3ds, (n, 3), "mean3ds"), _f32(scales, (n, 3), "scales"),
           _f32(quats, (n, 4), "quats"), _f32(viewmats, (B, 4, 4), "viewmats"),
           _f32(colors, (n, 3), "colors"), _f32(opacities, (n, 1), "opacities"),
           _f32(background, (3,), "background"), _f32n(cov3ds, (B, n, 6), "cov3ds"),
           _f32(xys, (B, n, 2), "xys"), _f32(depths, (B, n, 1), "depths"),
           _i32(radii, (B, n, 1), "radii"), _f32(conics, (B, n, 3), "conics"),
           _f32n(compensation, (B, n, 1), "compensation"),
           _f32(final_Ts, (B, H, W, 1), "final_Ts"), _i32(final_idx, (B, H, W, 1), "final_idx"),
           _f32(v_imgs, (B, H, W, 3), "v_imgs"),
           None if v_img_alphas is None else _f32(v_img_alphas, (B, H, W, 1), "v_img_alphas")]
    outs = [_out((n, 3), torch.float32, dev), _out((n, 3), torch.float32, dev),
            _out((n, 4), torch.float32, dev), _out((n, 3), torch.float32, dev),
            _out((n, 1), torch.float32, dev)]
    d = _ffi.make_descriptor(n, tuple(img_shape), tuple(f), tuple(c), glob_scale, clip_thresh,
                             block_width)
    toks = (ctypes.c_ulonglong * max(B, 1))(*[int(t) for t in tokens])
    _call_native("jsb200_render_views_bwd", ins + outs, d, ctypes.c_uint(B), toks)
    return tuple(outs)
