"""``project`` -- torch twin of the reference's jaxsplat/_project/__init__.py.

Same public signature and return tuple (``:56-87``), same custom-VJP structure (``:90-208``):
gradients flow to (mean3ds, scales, quats), ``None`` for viewmat; integer outputs are
non-differentiable; the hidden 7th output ``compensation`` is dropped by ``project``.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Tuple

import torch

from . import _ops


@dataclass(frozen=True)
class ProjectDescriptor:
    """static arguments (reference _project/__init__.py:10-53)"""
    num_points: int
    img_shape: Tuple[int, int]
    f: Tuple[float, float]
    c: Tuple[float, float]
    glob_scale: float
    clip_thresh: float
    block_width: int

    def kwargs(self):
        return dict(num_points=self.num_points, img_shape=self.img_shape, f=self.f, c=self.c,
                    glob_scale=self.glob_scale, clip_thresh=self.clip_thresh,
                    block_width=self.block_width)


class _Project(torch.autograd.Function):
    @staticmethod
    def forward(ctx, desc: ProjectDescriptor, mean3ds, scales, quats, viewmat):
        (cov3ds, xys, depths, radii, conics, compensation, num_tiles_hit,
         cum_tiles_hit) = _ops.project_fwd(mean3ds, scales, quats, viewmat, **desc.kwargs())
        ctx.desc = desc
        ctx.save_for_backward(mean3ds, scales, quats, viewmat, cov3ds, xys, radii, conics,
                              compensation)
        ctx.mark_non_differentiable(radii, num_tiles_hit, cum_tiles_hit)
        return xys, depths, radii, conics, num_tiles_hit, cum_tiles_hit, compensation

    @staticmethod
    def backward(ctx, v_xy, v_depth, _v_radii, v_conic, _v_nth, _v_cth, v_compensation):
        (mean3ds, scales, quats, viewmat, cov3ds, xys, radii, conics,
         compensation) = ctx.saved_tensors
        n = ctx.desc.num_points
        dev = mean3ds.device

        def z(v, shape):
            return torch.zeros(shape, dtype=torch.float32, device=dev) if v is None else v.contiguous()

        v_mean3d, v_scale, v_quat, _v_cov2d, _v_cov3d = _ops.project_bwd(
            mean3ds, scales, quats, viewmat, cov3ds, xys, radii, conics, compensation,
            z(v_compensation, (n, 1)), z(v_xy, (n, 2)), z(v_depth, (n, 1)), z(v_conic, (n, 3)),
            **ctx.desc.kwargs())
        return None, v_mean3d, v_scale, v_quat, None


def project(mean3ds, scales, quats, viewmat, *, img_shape, f, c, glob_scale, clip_thresh,
            block_width):
    """Projects 3D Gaussians; returns (xys, depths, radii, conics, num_tiles_hit, cum_tiles_hit)."""
    desc = ProjectDescriptor(num_points=mean3ds.shape[0], img_shape=tuple(img_shape),
                             f=(float(f[0]), float(f[1])), c=(float(c[0]), float(c[1])),
                             glob_scale=float(glob_scale), clip_thresh=float(clip_thresh),
                             block_width=int(block_width))
    xys, depths, radii, conics, num_tiles_hit, cum_tiles_hit, _comp = _Project.apply(
        desc, mean3ds, scales, quats, viewmat)
    return xys, depths, radii, conics, num_tiles_hit, cum_tiles_hit
