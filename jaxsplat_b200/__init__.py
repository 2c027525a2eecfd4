"""jaxsplat_b200 -- B200-native (sm_100a) differentiable Gaussian splatting hot path behind the
jaxsplat API (reference jaxsplat/__init__.py:6-73).

``render`` / ``project`` / ``rasterize`` keep the reference's names, argument order, keyword
names, shapes and gradient structure; arrays are torch CUDA tensors here because JAX is not
available in this image (the JAX binding that registers the same four custom-call targets lives
in ``jaxsplat_b200.jax_binding`` and is import-guarded).  All compute runs in
``lib/libjaxsplat_b200.so`` through the C ABI of ``include/jaxsplat_b200.h``; there is no
fallback path.
"""
from . import _ops
from ._project import project
from ._rasterize import rasterize
from ._render import render_fused, render_views, set_gradient_sink, set_preclear

_FUSED = True


def set_fused_render(enable: bool) -> None:
    """``render`` runs as one fused native call per direction (default) or, when disabled, as the
    reference's two-op chain ``project`` -> ``rasterize`` (identical results either way)."""
    global _FUSED
    _FUSED = bool(enable)



def render(means3d, scales, quats, colors, opacities, *, viewmat, background, img_shape, f, c,
           glob_scale, clip_thresh, block_size=16, tile_rows=None):
    """Renders 3D Gaussians to a 2D image differentiably (reference jaxsplat/__init__.py:6-70).

    means3d (N,3), scales (N,3), quats (N,4) normalised (w,x,y,z), colors (N,3), opacities (N,1);
    viewmat (4,4) row-major world->camera, background (3,), img_shape (H,W), f (fx,fy),
    c (cx,cy).  Differentiable w.r.t. the five positional arrays.  Returns (H,W,3) float32.
    tile_rows=(begin, end) (extension, default None = reference behaviour) renders only that band
    of tile rows -- the unit of screen-tile-row sharding across GPUs (multiview.tile_row_bands).
    """
    if _FUSED and _ops._DEFAULT_TABLE is None:  # (another library's table: the plain op chain)
        return render_fused(means3d, scales, quats, colors, opacities, viewmat=viewmat,
                            background=background, img_shape=img_shape, f=f, c=c,
                            glob_scale=glob_scale, clip_thresh=clip_thresh, block_size=block_size,
                            tile_rows=tile_rows)
    (xys, depths, radii, conics, _num_tiles_hit, cum_tiles_hit) = project(
        means3d, scales, quats, viewmat, img_shape=img_shape, f=f, c=c, glob_scale=glob_scale,
        clip_thresh=clip_thresh, block_width=block_size)
    return rasterize(colors, opacities, background, xys, depths, radii, conics, cum_tiles_hit,
                     img_shape=img_shape, block_width=block_size, tile_rows=tile_rows)


__all__ = ["render", "render_views", "project", "rasterize", "set_fused_render",
           "set_gradient_sink", "set_preclear"]
