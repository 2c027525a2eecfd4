"""Runs ON the GPU box: batched cameras (jaxsplat_b200.render_views) against the
per-view loop over jaxsplat_b200.render on a C4-shaped step (3M Gaussians, 1080p, ring cameras).
usage: python tools/views_bench.py [num_views=8] [config=C4] [iters=10]  ->  one JSON line"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import jaxsplat_b200 as jsb  # noqa: E402
from jaxsplat_b200 import _lib, synthetic  # noqa: E402


def timed(fn, iters, warmup=3):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
    cfg = sys.argv[2] if len(sys.argv) > 2 else "C4"
    iters = int(sys.argv[3]) if len(sys.argv) > 3 else 10
    dev = torch.device("cuda:0")
    sc = synthetic.make_config(cfg).to(dev)
    H, W = sc.img_shape
    cams = torch.stack(synthetic.ring_cameras(B)).contiguous().to(dev)
    targets = torch.rand((B, H, W, 3), generator=torch.Generator().manual_seed(0)).to(dev)
    params = [p.clone().requires_grad_(True) for p in sc.params()]
    kw = sc.render_kwargs()
    kwv = {k: v for k, v in kw.items() if k != "viewmat"}

    def loop():
        for p in params:
            p.grad = None
        for b in range(B):
            kw["viewmat"] = cams[b]
            img = jsb.render(*params, **kw)
            (((img - targets[b]) ** 2).mean() / B).backward()

    def batched():
        for p in params:
            p.grad = None
        imgs = jsb.render_views(*params, viewmats=cams, **kwv)
        ((imgs - targets) ** 2).mean().backward()

    _lib.set_option("view_slots", max(B, 1))
    loop()
    g_loop = [p.grad.clone() for p in params]
    batched()
    g_b = [p.grad.clone() for p in params]
    rel = [float((a - b).abs().max() / b.abs().max().clamp_min(1e-30)) for a, b in zip(g_b, g_loop)]
    ms_loop = timed(loop, iters)
    ms_b = timed(batched, iters)
    st = _lib.stats()
    print(json.dumps({
        "workload": f"{cfg}: {sc.means3d.shape[0]} Gaussians, {W}x{H}, {B} ring cameras per step, "
                    "fwd+bwd of the mean MSE over the views",
        "loop_over_render": {"ms_per_step": ms_loop, "views_per_s": 1000.0 * B / ms_loop},
        "render_views": {"ms_per_step": ms_b, "views_per_s": 1000.0 * B / ms_b},
        "grad_rel_diff_batched_vs_loop": rel,
        "last_num_intersects": st["last_num_intersects"],
        "workspace_bytes_last_slot": st["workspace_bytes"],
    }))


if __name__ == "__main__":
    main()
