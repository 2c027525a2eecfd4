#!/usr/bin/env python
"""Runs ON the GPU box under torchrun (or plainly for one GPU): the two sharded configurations of
BASELINE.json that bench.py's headline line (C3, one view per rank and step) does not cover.

  views (config 4, C4): 3M Gaussians, 1080p, a batch of 64 ring cameras per step.  Rank r renders
        views r, r+G, ... with jaxsplat_b200.render_views (chunks of <= 8 views, one binning slot
        each), the five parameter gradients accumulate over its views, are packed into the flat
        14*N-float buffer and summed with ONE all-reduce per step.  Strong scaling of a fixed
        64-view batch; value = views/s of the whole job.
  bands (config 5, C5): 6M Gaussians, 3840x2160, ONE frame per step sharded by screen-tile rows
        (Gaussians replicated; multiview.band_step).  The bands are balanced by the per-row
        intersection histogram (multiview.tile_row_weights on the outputs of one `project` call,
        outside the timed region).  Strong scaling; value = frames/s; the summed loss is printed
        so lines at different G can be compared (same frame => same loss to float tolerance).

  python -m torch.distributed.run --nnodes=1 --nproc-per-node G --master-addr 127.0.0.1 \
      --master-port 29511 tools/multigpu_bench.py views|bands [--steps K] [--warmup W]
  --dry: CPU + gloo + a stand-in renderer and tiny sizes: checks this script's own control flow
         (sharding, chunking, reduction, timing reduction) where there is no GPU.
Timing: CUDA events around exactly K steps, barrier + synchronize on both sides, max over ranks.
Prints ONE JSON line on rank 0.
"""
import argparse
import json
import os
import sys
import time

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from jaxsplat_b200 import multiview, synthetic  # noqa: E402


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("mode", choices=["views", "bands"])
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--views", type=int, default=64)
    ap.add_argument("--chunk", type=int, default=8, help="views per render_views call")
    ap.add_argument("--config", default=None, help="default: C4 (views) / C5 (bands)")
    ap.add_argument("--num-points", type=int, default=None)
    ap.add_argument("--dry", action="store_true")
    return ap.parse_args()


def _dry_render_views(means, scales, quats, colors, opac, *, viewmats, img_shape, **_):
    H, W = img_shape
    cam = (viewmats[:, :3, :3].sum((1, 2)) + viewmats[:, :3, 3].sum(1)).view(-1, 1, 1, 1)
    base = (means.sum(0) * scales.mean(0) + quats[:, :3].cos().mean(0) + (colors * opac).mean(0))
    return (base.view(1, 1, 1, 3) * (1.0 + 0.01 * cam)).expand(viewmats.shape[0], H, W, 3)


def _dry_render(means, scales, quats, colors, opac, *, img_shape, tile_rows=None, block_size=16, **_):
    H, W = img_shape
    rows = torch.arange(H, dtype=torch.float32).view(H, 1, 1)
    base = means.sum(0) * scales.mean(0) + quats[:, :3].cos().mean(0) + (colors * opac).mean(0)
    img = (base.view(1, 1, 3) * (1.0 + 0.01 * rows)).expand(H, W, 3)
    if tile_rows is None:
        return img
    mask = torch.zeros((H, 1, 1))
    mask[tile_rows[0] * block_size:min(H, tile_rows[1] * block_size)] = 1.0
    return img * mask


class Timer:
    """CUDA events on the current stream (wall clock in the CPU dry run)"""

    def __init__(self, cuda):
        self.cuda = cuda

    def start(self):
        if self.cuda:
            self.a, self.b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            self.a.record()
        else:
            self.t0 = time.perf_counter()

    def stop_ms(self):
        if self.cuda:
            self.b.record()
            torch.cuda.synchronize()
            return self.a.elapsed_time(self.b)
        return 1e3 * (time.perf_counter() - self.t0)


def main():
    args = parse()
    real_stdout = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)  # NCCL banners go to stderr: stdout carries the one JSON line
    out = run(args)
    if out is not None:
        print(json.dumps(out), file=real_stdout, flush=True)


def run(args):
    """the measurement itself; returns the result dict on rank 0 (None elsewhere).  bench.py calls
    this for `--config C4` / `--config C5` with an argparse-like namespace."""
    torch.autograd.set_multithreading_enabled(False)  # (one backward chain: no worker-thread hand-off)
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    cuda = not args.dry
    if cuda:
        torch.cuda.set_device(local_rank)
        dev = torch.device("cuda", local_rank)
    else:
        dev = torch.device("cpu")
    if world > 1:
        if cuda:
            dist.init_process_group("nccl", device_id=dev)
        else:
            dist.init_process_group("gloo")

    def barrier():
        if world > 1:
            dist.barrier()
        if cuda:
            torch.cuda.synchronize()

    cfg = args.config or ("C4" if args.mode == "views" else "C5")
    if args.dry:
        scene = synthetic.benchmark_scene(args.num_points or 200, (48, 40), scale_max=0.3, seed=0)
    else:
        scene = synthetic.make_config(cfg, num_points=args.num_points)
    H, W = scene.img_shape
    N = scene.means3d.shape[0]
    sc = scene.to(dev)
    kw = sc.render_kwargs()
    if cuda:
        import jaxsplat_b200 as jsb
        from jaxsplat_b200 import _lib
        render_views, render = jsb.render_views, jsb.render
    else:
        render_views, render = _dry_render_views, _dry_render

    flat = multiview.FlatParams(sc.params())
    timer = Timer(cuda)
    extra = {}

    if args.mode == "views":
        B = args.views
        mine = multiview.shard_views(B, rank, world)
        cams = torch.stack(synthetic.ring_cameras(B)).contiguous()
        my_cams = cams[mine].contiguous().to(dev)
        # one seeded target per view (uniform noise: there are no image files), this rank's only
        targets = torch.stack([torch.rand((H, W, 3), generator=torch.Generator().manual_seed(77 + v))
                               for v in mine]).to(dev) if mine else torch.empty((0, H, W, 3), device=dev)
        chunk = max(1, min(args.chunk, max(len(mine), 1)))
        if cuda:
            _lib.set_option("view_slots", chunk)
        kwv = {k: v for k, v in kw.items() if k != "viewmat"}
        scale = 2.0 / float(B * H * W * 3)  # d(mean over all B views of the MSE)/d(img)

        def step():
            flat.zero_grad()
            total = torch.zeros((), dtype=torch.float32, device=dev)
            for a in range(0, len(mine), chunk):
                imgs = render_views(*flat.params, viewmats=my_cams[a:a + chunk], **kwv)
                diff = imgs.detach() - targets[a:a + chunk]
                total = total + torch.dot(diff.reshape(-1), diff.reshape(-1)) / float(B * H * W * 3)
                imgs.backward(gradient=diff * scale)  # accumulates into the flat buffer's views
            if world > 1:
                flat.reducer.all_reduce()
                dist.all_reduce(total)
            return total

        units, unit_name = B, "views/s"
        workload = (f"{cfg}: {N} Gaussians, {W}x{H}, {B}-view batch per step sharded over {world} "
                    f"rank(s) ({len(mine)} views on rank 0, render_views chunks of {chunk}), "
                    "fwd+bwd of the mean MSE over the views, one gradient all-reduce per step")
    else:
        bs = sc.block_size
        tiles_y = (H + bs - 1) // bs
        # planning pass (outside the timed region): intersections per tile row -> balanced bands
        if cuda:
            with torch.no_grad():
                xys, _d, radii, _c, _n, cum = jsb.project(
                    sc.means3d, sc.scales, sc.quats, sc.viewmat, img_shape=sc.img_shape, f=sc.f,
                    c=sc.c, glob_scale=sc.glob_scale, clip_thresh=sc.clip_thresh, block_width=bs)
            weights = multiview.tile_row_weights(xys, radii, sc.img_shape, bs)
            extra["num_intersects"] = int(cum[-1].item())
            del xys, radii, cum
        else:
            weights = torch.ones(tiles_y, dtype=torch.float64)
        bands = multiview.tile_row_bands(tiles_y, world, weights=weights.tolist())
        band = bands[rank]
        extra["bands"] = bands
        extra["band_intersections"] = [float(weights[a:b].sum()) for a, b in bands]
        target = torch.rand((H, W, 3), generator=torch.Generator().manual_seed(78)).to(dev)
        kwb = dict(kw)

        def render_band(*params, tile_rows=None):
            return render(*params, tile_rows=tile_rows, **kwb)

        def step():
            return multiview.band_step(flat, target, band, render_band, block_size=bs)

        units, unit_name = 1, "frames/s"
        workload = (f"{cfg}: {N} Gaussians, {W}x{H}, one frame per step sharded by screen-tile rows "
                    f"over {world} rank(s) (bands balanced by the per-row intersection histogram), "
                    "fwd+bwd of the MSE, one gradient all-reduce per step")

    loss = None
    for _ in range(max(args.warmup, 1)):
        loss = step()
    barrier()
    timer.start()
    for _ in range(args.steps):
        loss = step()
    ms = timer.stop_ms()
    barrier()
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    grad_norm = float(flat.grad.double().norm())
    if rank == 0:
        out = {"mode": args.mode, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 1),
               "ms_per_step": ms / args.steps, "value": units * args.steps / (ms / 1e3),
               "unit": unit_name, "scaling": "strong", "loss": float(loss),
               "grad_l2": grad_norm, "reduction": flat.reducer.mode, "workload": workload,
               "data": "synthetic", "dry_run_cpu": bool(args.dry)}
        out.update(extra)
        if cuda:
            out["workspace_bytes"] = _lib.stats()["workspace_bytes"]
    else:
        out = None
    if world > 1:
        dist.destroy_process_group()
    return out


if __name__ == "__main__":
    main()
