#!/usr/bin/env python
"""Fit Gaussians to one image with `jaxsplat_b200.render` -- the torch counterpart of the
reference's examples/single_image.py (same parameterisation and loss: quaternions normalised,
colours and opacities through a sigmoid, MSE against the ground truth, Adam with lr 0.01;
reference examples/single_image.py:13-66, :142-150).  Needs a B200; no image or video libraries:
the ground truth is a .npy array (H, W, 3) in [0, 1] or, by default, a seeded synthetic image, and
the result is written as .npy.

    python examples/single_image.py [--iters 100] [--num-points 10000] [--size 256 256]
                                    [--gt gt.npy] [--out fit.npy] [--native-adam]
"""
import argparse
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import jaxsplat_b200 as jsb  # noqa: E402
from jaxsplat_b200 import synthetic  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--iters", type=int, default=100)
    ap.add_argument("--num-points", type=int, default=10_000)
    ap.add_argument("--size", type=int, nargs=2, default=(256, 256), metavar=("H", "W"))
    ap.add_argument("--lr", type=float, default=0.01)
    ap.add_argument("--gt", default=None, help=".npy ground truth (H, W, 3) in [0, 1]")
    ap.add_argument("--out", default=None, help="write the fitted image here (.npy)")
    ap.add_argument("--native-adam", action="store_true", help="jaxsplat_b200.optim.Adam (one fused kernel per array)")
    args = ap.parse_args()
    dev = torch.device("cuda:0")

    if args.gt:
        gt = torch.from_numpy(np.load(args.gt).astype(np.float32))[..., :3]
        H, W = int(gt.shape[0]), int(gt.shape[1])
    else:  # a smooth seeded picture: coloured blobs on a gradient
        H, W = args.size
        yy, xx = torch.meshgrid(torch.linspace(0, 1, H), torch.linspace(0, 1, W), indexing="ij")
        gt = torch.stack([0.5 + 0.5 * torch.sin(6.0 * xx + 2.0 * yy), yy, 0.5 + 0.5 * torch.cos(9.0 * yy * xx)], -1)
    gt = gt.to(dev).contiguous()

    scene = synthetic.single_image_scene(num_points=args.num_points, img_shape=(H, W), seed=0).to(dev)
    kw = scene.render_kwargs()

    def logit(x):
        x = x.clamp(1e-6, 1 - 1e-6)
        return torch.log(x / (1 - x))

    params = {"means3d": scene.means3d.clone(), "scales": scene.scales.clone(), "quats": scene.quats.clone(),
              "colors": logit(scene.colors), "opacities": logit(scene.opacities)}
    for v in params.values():
        v.requires_grad_(True)
    if args.native_adam:
        from jaxsplat_b200.optim import Adam
        opt = Adam(list(params.values()), lr=args.lr)
    else:
        opt = torch.optim.Adam(list(params.values()), lr=args.lr)

    def render_fn():
        q = params["quats"] / params["quats"].norm(dim=1, keepdim=True)
        return jsb.render(params["means3d"], params["scales"], q, torch.sigmoid(params["colors"]),
                          torch.sigmoid(params["opacities"]), **kw)

    log_every, cum, cum_split = 50, 0.0, 0.0
    for i in range(args.iters):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        opt.zero_grad()
        loss = ((render_fn() - gt) ** 2).mean()
        loss.backward()
        opt.step()
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        cum += dt
        cum_split += dt
        if i % log_every == 0:
            print(f"iter {i} loss {float(loss.detach()):.4f}, {cum_split / max(1, min(i, log_every)) * 1000:.3f}ms avg per step")
            cum_split = 0.0
    print(f"done training in {cum:.3f}s ({cum / args.iters * 1000:.3f}ms avg per step), final loss {float(loss.detach()):.5f}")
    if args.out:
        with torch.no_grad():
            np.save(args.out, render_fn().cpu().numpy())


if __name__ == "__main__":
    main()
