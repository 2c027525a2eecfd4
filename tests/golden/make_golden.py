"""Generates the golden fixtures in tests/golden/ FROM THE REFERENCE'S OWN CUDA KERNELS.

Run on a GPU box where oracle/_ref/_jaxsplat*.so (the unmodified reference, built by
oracle/Makefile from /root/reference) is present:

    python tests/golden/make_golden.py --out gpurun_out/golden

Every array a fixture holds is an output of the reference library (through its XLA custom-call
targets and its internal ops::sort_and_bin), never of this repository's kernels or oracle.
The reference has no tests or vectors of its own, so these files are what
pins both the CPU oracle (tests/test_oracle_golden.py) and the CUDA path (tests/test_gpu_*.py).
"""
import argparse
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from jaxsplat_b200 import _ops, synthetic  # noqa: E402
from oracle import ref_cuda  # noqa: E402


def golden_scenes():
    """small scenes that exercise every cull branch, ragged image sizes and block sizes"""
    out = {}
    s = synthetic.benchmark_scene(3000, (80, 112), scale_max=0.3, seed=11, name="g_small16")
    out["g_small16"] = s
    # camera inside the cloud: near-plane culls, off-screen Gaussians, huge splats
    s = synthetic.benchmark_scene(2000, (72, 100), scale_max=0.6, seed=12, name="g_inside16")
    s.viewmat = synthetic._viewmat_identity(1.5)
    s.background = torch.tensor([0.1, 0.5, 0.9])
    out["g_inside16"] = s
    s = synthetic.benchmark_scene(1500, (60, 52), scale_max=0.25, seed=13, name="g_small8")
    s.block_size = 8
    out["g_small8"] = s
    s = synthetic.single_image_scene(2500, (64, 64), seed=14)
    s.name = "g_single16"
    out["g_single16"] = s
    # nothing visible: every Gaussian behind the near plane -> num_intersects == 0 (Q5)
    s = synthetic.benchmark_scene(300, (48, 48), scale_max=0.2, seed=15, name="g_empty16")
    s.viewmat = synthetic._viewmat_identity(-20.0)
    out["g_empty16"] = s
    return out


def run_reference(scene, table):
    dev = torch.device("cuda:0")
    sc = scene.to(dev)
    n = sc.means3d.shape[0]
    kw = dict(num_points=n, img_shape=sc.img_shape, f=sc.f, c=sc.c, glob_scale=sc.glob_scale,
              clip_thresh=sc.clip_thresh, block_width=sc.block_size)
    rkw = dict(num_points=n, img_shape=sc.img_shape, block_width=sc.block_size)
    o = {}
    (o["cov3ds"], o["xys"], o["depths"], o["radii"], o["conics"], o["compensation"],
     o["num_tiles_hit"], o["cum_tiles_hit"]) = _ops.project_fwd(
        sc.means3d, sc.scales, sc.quats, sc.viewmat, table=table, **kw)
    o["final_Ts"], o["final_idx"], o["img"] = _ops.rasterize_fwd(
        sc.colors, sc.opacities, sc.background, o["xys"], o["depths"], o["radii"], o["conics"],
        o["cum_tiles_hit"], table=table, **rkw)
    H, W = sc.img_shape
    g = torch.Generator().manual_seed(99)
    target = torch.rand((H, W, 3), generator=g).to(dev)
    o["target"] = target
    o["v_img"] = ((2.0 / (H * W * 3)) * (o["img"] - target)).contiguous()
    gg = torch.Generator().manual_seed(100)
    o["v_img_alpha"] = (torch.rand((H, W, 1), generator=gg) * 1e-3).to(dev)
    (o["v_colors"], o["v_opacity"], o["v_xy"], o["v_xy_abs"], o["v_conic"]) = _ops.rasterize_bwd(
        sc.colors, sc.opacities, sc.background, o["xys"], o["depths"], o["radii"], o["conics"],
        o["cum_tiles_hit"], o["final_Ts"], o["final_idx"], o["v_img"], o["v_img_alpha"],
        table=table, **rkw)
    gd = torch.Generator().manual_seed(101)
    o["v_depth"] = (torch.randn((n, 1), generator=gd) * 1e-4).to(dev)
    o["v_compensation"] = (torch.randn((n, 1), generator=gd) * 1e-4).to(dev)
    (o["v_mean3d"], o["v_scale"], o["v_quat"], o["v_cov2d"], o["v_cov3d"]) = _ops.project_bwd(
        sc.means3d, sc.scales, sc.quats, sc.viewmat, o["cov3ds"], o["xys"], o["radii"],
        o["conics"], o["compensation"], o["v_compensation"], o["v_xy"], o["v_depth"], o["v_conic"],
        table=table, **kw)
    torch.cuda.synchronize()
    I = int(o["cum_tiles_hit"][-1]) if n else 0
    desc = table.make_descriptor(n, sc.img_shape, (0, 0), (0, 0), 0.0, 0.0, sc.block_size)
    if I > 0:
        ids, bins = ref_cuda.sort_and_bin(desc, I, o["xys"], o["depths"], o["radii"],
                                          o["cum_tiles_hit"])
        o["gaussian_ids_sorted"], o["tile_bins"] = ids, bins
    full_desc = table.make_descriptor(n, sc.img_shape, sc.f, sc.c, sc.glob_scale, sc.clip_thresh,
                                      sc.block_size)
    res = {k: v.detach().cpu().numpy() for k, v in o.items()}
    res["descriptor"] = np.frombuffer(full_desc, dtype=np.uint8).copy()
    for k in ("means3d", "scales", "quats", "colors", "opacities", "viewmat", "background"):
        res["in_" + k] = getattr(scene, k).numpy()
    res["static"] = np.array([sc.img_shape[0], sc.img_shape[1], sc.f[0], sc.f[1], sc.c[0],
                              sc.c[1], sc.glob_scale, sc.clip_thresh, sc.block_size],
                             dtype=np.float64)
    return res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "golden"))
    args = ap.parse_args()
    os.makedirs(args.out, exist_ok=True)
    table = ref_cuda.table()
    for name, scene in golden_scenes().items():
        res = run_reference(scene, table)
        path = os.path.join(args.out, name + ".npz")
        np.savez_compressed(path, **res)
        I = int(res["cum_tiles_hit"][-1])
        print(f"{name}: N={scene.means3d.shape[0]} img={scene.img_shape} bw={scene.block_size} "
              f"I={I} visible={(res['radii'] > 0).sum()} -> {path} "
              f"({os.path.getsize(path) / 1024:.0f} KiB)")


if __name__ == "__main__":
    main()
