"""``rasterize`` -- torch twin of the reference's jaxsplat/_rasterize/__init__.py.

Same signature (``:40-68``) and custom-VJP structure (``:71-190``): gradients for (colors,
opacities, xys, conics), ``None`` for (background, depths, radii, cum_tiles_hit);
``img_alpha = 1 - final_Ts`` is computed on the host side like the reference (``:136``).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional, Tuple

import torch

from . import _ops


@dataclass(frozen=True)
class RasterizeDescriptor:
    """static arguments (reference _rasterize/__init__.py:10-37)"""
    num_points: int
    img_shape: Tuple[int, int]
    block_width: int
    band: Optional[Tuple[int, int]] = None  # extension: tile rows [begin, end) owned by this call

    def kwargs(self):
        return dict(num_points=self.num_points, img_shape=self.img_shape,
                    block_width=self.block_width, band=self.band)


class _Rasterize(torch.autograd.Function):
    @staticmethod
    def forward(ctx, desc, colors, opacities, background, xys, depths, radii, conics,
                cum_tiles_hit):
        final_Ts, final_idx, img = _ops.rasterize_fwd(colors, opacities, background, xys, depths,
                                                      radii, conics, cum_tiles_hit, **desc.kwargs())
        img_alpha = 1 - final_Ts
        ctx.desc = desc
        ctx.save_for_backward(colors, opacities, background, xys, depths, radii, conics,
                              cum_tiles_hit, final_Ts, final_idx)
        return img, img_alpha

    @staticmethod
    def backward(ctx, v_img, v_img_alpha):
        (colors, opacities, background, xys, depths, radii, conics, cum_tiles_hit, final_Ts,
         final_idx) = ctx.saved_tensors
        H, W = ctx.desc.img_shape
        dev = colors.device
        if v_img is None:
            v_img = torch.zeros((H, W, 3), dtype=torch.float32, device=dev)
        if v_img_alpha is None:
            v_img_alpha = torch.zeros((H, W, 1), dtype=torch.float32, device=dev)
        v_colors, v_opacity, v_xy, _v_xy_abs, v_conic = _ops.rasterize_bwd(
            colors, opacities, background, xys, depths, radii, conics, cum_tiles_hit, final_Ts,
            final_idx, v_img.contiguous(), v_img_alpha.contiguous(), **ctx.desc.kwargs())
        return None, v_colors, v_opacity, None, v_xy, None, None, v_conic, None


def rasterize(colors, opacities, background, xys, depths, radii, conics, cum_tiles_hit, *,
              img_shape, block_width, tile_rows=None):
    """Rasterizes projected Gaussians; returns the (H, W, 3) image.

    tile_rows=(begin, end) is an extension for screen-tile-row sharding: only those tile rows are
    rasterised (and differentiated); pixels outside the band are left unwritten."""
    desc = RasterizeDescriptor(num_points=colors.shape[0], img_shape=tuple(img_shape),
                               block_width=int(block_width),
                               band=None if tile_rows is None else (int(tile_rows[0]), int(tile_rows[1])))
    img, _img_alpha = _Rasterize.apply(desc, colors, opacities, background, xys, depths, radii,
                                       conics, cum_tiles_hit)
    return img
