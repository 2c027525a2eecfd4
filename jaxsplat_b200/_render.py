"""Fused ``render``: one native call forward, one backward.

Same values and gradients as ``project`` followed by ``rasterize`` (reference
jaxsplat/__init__.py:54-70), without the intermediate autograd nodes: the forward keeps the
projected arrays and the sorted lists on the device, the backward reuses the forward's binning
through a token (no fingerprint pass, no host synchronisation) and feeds v_xy / v_conic straight
into the projection VJP (the round trip of reference _rasterize/__init__.py:187 ->
_project/__init__.py:170-198).
"""
from __future__ import annotations

import torch

from . import _lib, _ops
from ._project import ProjectDescriptor


_GRAD_SINK = None


def set_gradient_sink(buffers) -> None:
    """Where the fused ``render`` backward writes the five parameter gradients (extension).

    ``buffers`` = [v_means3d (N,3), v_scales (N,3), v_quats (N,4), v_colors (N,3), v_opacities (N,1)]
    (+ optionally a 5*N-float scratch tensor that directly follows v_opacities in memory), typically
    views into ONE flat buffer that a data-parallel driver all-reduces in place: the gradients then
    never get packed or copied -- the same idea as DDP's ``gradient_as_bucket_view``.  autograd
    adopts the returned views as ``.grad`` when ``.grad`` is None.  ``None`` restores private
    result buffers.  One render backward per step may use the sink (a second one would overwrite
    it); multi-view accumulation goes through ``render_views`` or ordinary ``.grad`` accumulation."""
    global _GRAD_SINK
    _GRAD_SINK = None if buffers is None else list(buffers)


# The rasteriser backward accumulates four of its results with atomics (36 bytes per Gaussian) and
# needs them zero.  Cleared in front of it that is a 36 MB memset on the critical path at 1M
# Gaussians; cleared on a side stream while the forward's list build and rasteriser run (both
# issue-bound, the memory system idle) it costs nothing.  The buffer is allocated BEFORE the forward
# is enqueued and the side stream waits for that point of the main stream, so whatever owned the
# memory before is done with it; the backward waits for the side stream's event.
_PRECLEAR = True
_PRECLEAR_MIN_BYTES = 4 << 20
_side_streams = {}


def set_preclear(enabled: bool) -> None:
    """Clear the backward's accumulators on a side stream during the forward (default) or leave it
    to the backward's own memset."""
    global _PRECLEAR
    _PRECLEAR = bool(enabled)


def _side_stream(dev):
    s = _side_streams.get(dev)
    if s is None:
        s = _side_streams[dev] = torch.cuda.Stream(device=dev)
    return s


class _Render(torch.autograd.Function):
    @staticmethod
    def forward(ctx, desc: ProjectDescriptor, band, means3d, scales, quats, colors, opacities,
                viewmat, background):
        n = means3d.shape[0]
        ctx.acc = ctx.acc_ready = None
        preclear = (_PRECLEAR and _GRAD_SINK is None and any(ctx.needs_input_grad[2:7])
                    and n * 36 >= _PRECLEAR_MIN_BYTES)
        if preclear:
            main = torch.cuda.current_stream(means3d.device)
            acc = torch.empty(9 * n, dtype=torch.float32, device=means3d.device)
            fork = main.record_event()
        (cov3ds, xys, depths, radii, conics, compensation, final_Ts, final_idx,
         img), token = _ops.render_fwd(means3d, scales, quats, viewmat, colors, opacities,
                                       background, band=band, lean=True, **desc.kwargs())
        # (lean: cov3ds / compensation are None -- the backward recomputes the covariance from scales
        # and quaternions instead of the forward writing and the backward reading 24 bytes per Gaussian)
        ctx.desc, ctx.band, ctx.token = desc, band, token
        ctx.save_for_backward(means3d, scales, quats, viewmat, colors, opacities, background,
                              cov3ds, xys, depths, radii, conics, compensation, final_Ts,
                              final_idx)
        if preclear:
            # (enqueued now, i.e. while the forward's later kernels run; ordered after `fork`)
            side = _side_stream(means3d.device)
            side.wait_event(fork)
            try:  # the driver's memset on the side stream: no kernel of torch's in the step
                _lib.clear_async(side.cuda_stream, acc.data_ptr(), acc.numel() * 4)
            except Exception:  # (e.g. a side stream of another device than the current one)
                with torch.cuda.stream(side):
                    acc.zero_()
            acc.record_stream(side)
            ctx.acc, ctx.acc_ready = acc, side.record_event()
        return img

    @staticmethod
    def backward(ctx, v_img):
        saved = ctx.saved_tensors
        # (one use: a second backward through the same graph finds the accumulators dirty)
        acc, ready, ctx.acc, ctx.acc_ready = ctx.acc, ctx.acc_ready, None, None
        if acc is not None and _GRAD_SINK is None:
            torch.cuda.current_stream(acc.device).wait_event(ready)
        else:
            acc = None
        if _GRAD_SINK is not None:
            # autograd ACCUMULATES a returned gradient into an existing .grad; if that .grad already
            # is the sink buffer, `grad += out` with out aliasing grad would silently double it
            for q, sink in zip((saved[0], saved[1], saved[2], saved[4], saved[5]), _GRAD_SINK):
                if q.grad is not None and q.grad.data_ptr() == sink.data_ptr():
                    raise RuntimeError(
                        "jaxsplat_b200.set_gradient_sink: a parameter's .grad already aliases its sink "
                        "buffer; set .grad = None before the backward (the sink views are adopted as "
                        ".grad), or accumulation would double the gradient")
        # render() exposes no alpha image: its cotangent is identically zero (None = not read)
        (v_mean3d, v_scale, v_quat, v_colors, v_opacity, *_rest) = _ops.render_bwd(
            *saved, v_img.contiguous(), None, token=ctx.token, band=ctx.band,
            grad_out=_GRAD_SINK, cleared_acc=acc, **ctx.desc.kwargs())
        return None, None, v_mean3d, v_scale, v_quat, v_colors, v_opacity, None, None


def render_fused(means3d, scales, quats, colors, opacities, *, viewmat, background, img_shape, f, c,
                 glob_scale, clip_thresh, block_size=16, tile_rows=None):
    desc = ProjectDescriptor(num_points=means3d.shape[0], img_shape=tuple(img_shape),
                             f=(float(f[0]), float(f[1])), c=(float(c[0]), float(c[1])),
                             glob_scale=float(glob_scale), clip_thresh=float(clip_thresh),
                             block_width=int(block_size))
    band = None if tile_rows is None else (int(tile_rows[0]), int(tile_rows[1]))
    return _Render.apply(desc, band, means3d, scales, quats, colors, opacities, viewmat, background)


class _RenderViews(torch.autograd.Function):
    @staticmethod
    def forward(ctx, desc: ProjectDescriptor, means3d, scales, quats, colors, opacities, viewmats,
                background):
        (cov3ds, xys, depths, radii, conics, compensation, final_Ts, final_idx,
         imgs), tokens = _ops.render_views_fwd(means3d, scales, quats, viewmats, colors, opacities,
                                               background, lean=True, **desc.kwargs())
        ctx.desc, ctx.tokens = desc, tokens
        ctx.save_for_backward(means3d, scales, quats, viewmats, colors, opacities, background,
                              cov3ds, xys, depths, radii, conics, compensation, final_Ts,
                              final_idx)
        return imgs

    @staticmethod
    def backward(ctx, v_imgs):
        grads = _ops.render_views_bwd(*ctx.saved_tensors, v_imgs.contiguous(), None,
                                      tokens=ctx.tokens, **ctx.desc.kwargs())
        return (None,) + tuple(grads) + (None, None)


def render_views(means3d, scales, quats, colors, opacities, *, viewmats, background, img_shape, f,
                 c, glob_scale, clip_thresh, block_size=16):
    """``render`` for a batch of cameras (an addition to the reference API, which
    renders one camera per call: jaxsplat/__init__.py:6-70).

    Same arguments as ``render`` except ``viewmats`` (B,4,4).  Returns (B,H,W,3); view b is
    bit-identical to ``render(..., viewmat=viewmats[b])``.  Differentiable w.r.t. the five
    positional arrays: their gradients are the sums over the views, computed in one native call.
    """
    desc = ProjectDescriptor(num_points=means3d.shape[0], img_shape=tuple(img_shape),
                             f=(float(f[0]), float(f[1])), c=(float(c[0]), float(c[1])),
                             glob_scale=float(glob_scale), clip_thresh=float(clip_thresh),
                             block_width=int(block_size))
    return _RenderViews.apply(desc, means3d, scales, quats, colors, opacities, viewmats, background)
