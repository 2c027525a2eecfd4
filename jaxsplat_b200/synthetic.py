"""Seeded synthetic Gaussian scenes for the BASELINE.json configs.

Pure data generation on the CPU with torch.Generator (so every implementation -- ours, the
reference CUDA library, the CPU oracle -- sees bit-identical inputs).  No compute path here.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, Tuple

import torch


@dataclass
class Scene:
    name: str
    means3d: torch.Tensor
    scales: torch.Tensor
    quats: torch.Tensor
    colors: torch.Tensor
    opacities: torch.Tensor
    viewmat: torch.Tensor
    background: torch.Tensor
    img_shape: Tuple[int, int]
    f: Tuple[float, float]
    c: Tuple[float, float]
    glob_scale: float = 1.0
    clip_thresh: float = 0.01
    block_size: int = 16
    meta: Dict = field(default_factory=dict)

    def to(self, device):
        kw = {k: (v.to(device) if isinstance(v, torch.Tensor) else v) for k, v in self.__dict__.items()}
        return Scene(**kw)

    def params(self):
        return (self.means3d, self.scales, self.quats, self.colors, self.opacities)

    def render_kwargs(self):
        return dict(viewmat=self.viewmat, background=self.background, img_shape=self.img_shape,
                    f=self.f, c=self.c, glob_scale=self.glob_scale, clip_thresh=self.clip_thresh,
                    block_size=self.block_size)


def _viewmat_identity(tz: float) -> torch.Tensor:
    v = torch.eye(4, dtype=torch.float32)
    v[2, 3] = tz
    return v


def look_at(eye, target=(0.0, 0.0, 0.0), up=(0.0, 1.0, 0.0)) -> torch.Tensor:
    """row-major world->camera with +z forward (camera looks along +z), as the reference expects."""
    eye = torch.tensor(eye, dtype=torch.float64)
    target = torch.tensor(target, dtype=torch.float64)
    up = torch.tensor(up, dtype=torch.float64)
    fwd = target - eye
    fwd = fwd / fwd.norm()
    right = torch.linalg.cross(up, fwd)
    right = right / right.norm()
    true_up = torch.linalg.cross(fwd, right)
    R = torch.stack([right, true_up, fwd])  # rows
    v = torch.eye(4, dtype=torch.float64)
    v[:3, :3] = R
    v[:3, 3] = -R @ eye
    return v.to(torch.float32)


def benchmark_scene(num_points: int, img_shape: Tuple[int, int], scale_max: float = 0.5,
                    seed: int = 0, background: float = 1.0, name: str = "benchmark",
                    log_uniform_scales: Tuple[float, float] | None = None,
                    viewmat: torch.Tensor | None = None) -> Scene:
    """examples/benchmark.py:249-289 distribution: means U[-3,3]^3, quats N(0,1)^4 normalised,
    colors U[0,1), opacities U[0,1), scales U[0,scale_max), camera at z=+8 looking down -z...
    (identity rotation, translation (0,0,8)), f = (W/2, H/2), c = (W/2, H/2)."""
    g = torch.Generator().manual_seed(seed)
    H, W = img_shape
    means = (torch.rand((num_points, 3), generator=g) - 0.5) * 6.0
    if log_uniform_scales is None:
        scales = torch.rand((num_points, 3), generator=g) * scale_max
    else:
        lo, hi = log_uniform_scales
        u = torch.rand((num_points, 3), generator=g)
        scales = torch.exp(math.log(lo) + u * (math.log(hi) - math.log(lo)))
    quats = torch.randn((num_points, 4), generator=g)
    quats = quats / quats.norm(dim=1, keepdim=True)
    colors = torch.rand((num_points, 3), generator=g)
    opac = torch.rand((num_points, 1), generator=g)
    return Scene(name=name, means3d=means.float(), scales=scales.float(), quats=quats.float(),
                 colors=colors.float(), opacities=opac.float(),
                 viewmat=_viewmat_identity(8.0) if viewmat is None else viewmat,
                 background=torch.full((3,), background, dtype=torch.float32),
                 img_shape=(H, W), f=(W / 2.0, H / 2.0), c=(W / 2.0, H / 2.0),
                 meta=dict(seed=seed, scales=(f"U[0,{scale_max})" if log_uniform_scales is None else
                                              f"log-uniform[{log_uniform_scales[0]},{log_uniform_scales[1]}]")))


def single_image_scene(num_points: int = 10_000, img_shape=(256, 256), seed: int = 0) -> Scene:
    """examples/single_image.py:72-140 initialisation (C1): means U([-6,6]x[-6,6]x[-1,1]),
    scales U[0,0.5), uniform random rotations via the (u,v,w) formula, pre-sigmoid colors U[0,1)
    and opacities U[0.5,1); the activations (sigmoid, quat normalisation) are applied here so the
    Scene holds what `render` receives.  fov 90deg-ish: f = 0.5*W / tan(0.5*fov), camera z=8."""
    g = torch.Generator().manual_seed(seed)
    H, W = img_shape
    means = (torch.rand((num_points, 3), generator=g) - 0.5) * torch.tensor([12.0, 12.0, 2.0])
    scales = torch.rand((num_points, 3), generator=g) * 0.5
    u, v, w = (torch.rand((num_points, 1), generator=g) for _ in range(3))
    quats = torch.cat([torch.sqrt(1 - u) * torch.sin(2 * math.pi * v),
                       torch.sqrt(1 - u) * torch.cos(2 * math.pi * v),
                       torch.sqrt(u) * torch.sin(2 * math.pi * w),
                       torch.sqrt(u) * torch.cos(2 * math.pi * w)], dim=1)
    colors = torch.sigmoid(torch.rand((num_points, 3), generator=g))
    opac = torch.sigmoid(0.5 + 0.5 * torch.rand((num_points, 1), generator=g))
    fov = math.pi / 2.0
    focal = 0.5 * W / math.tan(0.5 * fov)
    return Scene(name="single_image", means3d=means.float(), scales=scales.float(),
                 quats=(quats / quats.norm(dim=1, keepdim=True)).float(), colors=colors.float(),
                 opacities=opac.float(), viewmat=_viewmat_identity(8.0),
                 background=torch.zeros(3, dtype=torch.float32), img_shape=(H, W),
                 f=(focal, focal), c=(W / 2.0, H / 2.0), meta=dict(seed=seed))


def ring_cameras(num_views: int, radius: float = 8.0, seed: int = 0):
    """C4: cameras on a ring of `radius` around the origin looking at it, seeded height jitter."""
    g = torch.Generator().manual_seed(1000 + seed)
    mats = []
    for k in range(num_views):
        ang = 2 * math.pi * k / num_views
        jitter = (torch.rand(1, generator=g).item() - 0.5) * 1.0
        mats.append(look_at((radius * math.sin(ang), jitter, -radius * math.cos(ang))))
    return mats


CONFIGS = {
    # name: (num_points, (H, W), scale_max)
    "C1": dict(kind="single_image", num_points=10_000, img_shape=(256, 256)),
    "C2": dict(kind="benchmark", num_points=100_000, img_shape=(512, 512), scale_max=0.5),
    "C3": dict(kind="benchmark", num_points=1_000_000, img_shape=(1080, 1920), scale_max=0.1),
    # C3's secondary distributions: small "trained-scene-like" splats (I ~ 2.8e6) and
    # examples/benchmark.py's literal scales < 0.5 (I ~ 2.6e8, list-building dominated)
    "C3s": dict(kind="benchmark", num_points=1_000_000, img_shape=(1080, 1920), scale_max=0.02),
    "C3b": dict(kind="benchmark", num_points=1_000_000, img_shape=(1080, 1920), scale_max=0.5),
    "C4": dict(kind="benchmark", num_points=3_000_000, img_shape=(1080, 1920),
               log_uniform_scales=(0.002, 0.06)),
    "C5": dict(kind="benchmark", num_points=6_000_000, img_shape=(2160, 3840), scale_max=0.1),
}


def make_config(name: str, seed: int = 0, num_points: int | None = None) -> Scene:
    cfg = dict(CONFIGS[name])
    kind = cfg.pop("kind")
    if num_points is not None:
        cfg["num_points"] = num_points
    if kind == "single_image":
        return single_image_scene(seed=seed, **cfg)
    s = benchmark_scene(seed=seed, name=name, **cfg)
    return s
