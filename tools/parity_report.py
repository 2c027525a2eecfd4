"""GPU-box report: ours vs the reference CUDA library on the BASELINE configs (parity + timing).
Usage: python tools/parity_report.py [C1 C2 C3 ...] [--iters K]"""
import argparse
import ctypes
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from jaxsplat_b200 import _ffi, _lib, _ops, synthetic  # noqa: E402
from oracle import ref_cuda  # noqa: E402


def run_pipeline(scene, table, target=None, feed=None):
    """feed: outputs of another implementation; when given, every op consumes THOSE as inputs
    (op-level parity) instead of its own upstream results (end-to-end parity)."""
    n = scene.means3d.shape[0]
    kw = dict(num_points=n, img_shape=scene.img_shape, f=scene.f, c=scene.c,
              glob_scale=scene.glob_scale, clip_thresh=scene.clip_thresh,
              block_width=scene.block_size)
    rkw = dict(num_points=n, img_shape=scene.img_shape, block_width=scene.block_size)
    out = {}
    (out["cov3ds"], out["xys"], out["depths"], out["radii"], out["conics"], out["compensation"],
     out["num_tiles_hit"], out["cum_tiles_hit"]) = _ops.project_fwd(
        scene.means3d, scene.scales, scene.quats, scene.viewmat, table=table, **kw)
    src = feed if feed is not None else out
    out["final_Ts"], out["final_idx"], out["img"] = _ops.rasterize_fwd(
        scene.colors, scene.opacities, scene.background, src["xys"], src["depths"], src["radii"],
        src["conics"], src["cum_tiles_hit"], table=table, **rkw)
    H, W = scene.img_shape
    if target is None:
        target = torch.zeros((H, W, 3), device=scene.means3d.device)
    v_img = (2.0 / (H * W * 3)) * (src["img"] - target)
    v_alpha = torch.zeros((H, W, 1), device=v_img.device)
    (out["v_colors"], out["v_opacity"], out["v_xy"], out["v_xy_abs"], out["v_conic"]) = \
        _ops.rasterize_bwd(scene.colors, scene.opacities, scene.background, src["xys"],
                           src["depths"], src["radii"], src["conics"], src["cum_tiles_hit"],
                           src["final_Ts"], src["final_idx"], v_img, v_alpha, table=table, **rkw)
    zeros = torch.zeros((n, 1), device=v_img.device)
    (out["v_mean3d"], out["v_scale"], out["v_quat"], out["v_cov2d"], out["v_cov3d"]) = \
        _ops.project_bwd(scene.means3d, scene.scales, scene.quats, scene.viewmat, src["cov3ds"],
                         src["xys"], src["radii"], src["conics"], src["compensation"], zeros,
                         src["v_xy"], zeros.clone(), src["v_conic"], table=table, **kw)
    return out


def time_pipeline(scene, table, target, iters, warmup=3):
    for _ in range(warmup):
        run_pipeline(scene, table, target)
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        run_pipeline(scene, table, target)
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2]


def compare(a, b, name, is_int):
    if is_int:
        bad = int((a != b).sum())
        return dict(name=name, mismatches=bad, total=a.numel())
    a64, b64 = a.double(), b.double()
    d = (a64 - b64).abs()
    bits = int((a.view(torch.int32) != b.view(torch.int32)).sum()) if a.dtype == torch.float32 else -1
    ref_inf = float(b64.abs().max())
    return dict(name=name, max_abs=float(d.max()), ref_inf=ref_inf,
                rel=float(d.max()) / max(ref_inf, 1e-30), bit_mismatch=bits, total=a.numel())


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("configs", nargs="*", default=["C1", "C2", "C3"])
    ap.add_argument("--iters", type=int, default=5)
    ap.add_argument("--no-ref", action="store_true")
    ap.add_argument("--out", default="gpurun_out/parity_report.json")
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    report = {}
    ours = _ffi.native_table()
    ref = None if args.no_ref or not ref_cuda.available() else ref_cuda.table()
    for cfg in args.configs:
        scene = synthetic.make_config(cfg).to(dev)
        H, W = scene.img_shape
        g = torch.Generator().manual_seed(123)
        target = torch.rand((H, W, 3), generator=g).to(dev)
        _ops.poison_outputs(True)
        mine = run_pipeline(scene, ours, target)
        torch.cuda.synchronize()
        _ops.poison_outputs(False)
        I = int(mine["cum_tiles_hit"][-1])
        entry = dict(N=scene.means3d.shape[0], img=[H, W], I=I, stats=_lib.stats())
        print(f"== {cfg}: N={entry['N']} img={H}x{W} I={I}", flush=True)
        if ref is not None:
            theirs = run_pipeline(scene, ref, target)
            torch.cuda.synchronize()
            rows = []
            ints = {"radii", "num_tiles_hit", "cum_tiles_hit", "final_idx"}
            for k in mine:
                rows.append(compare(mine[k], theirs[k], k, k in ints))
            oplevel = run_pipeline(scene, ours, target, feed=theirs)
            torch.cuda.synchronize()
            for k in oplevel:
                r = compare(oplevel[k], theirs[k], "op:" + k, k in ints)
                rows.append(r)
            # sorted ids + bins
            desc = ref.make_descriptor(entry["N"], scene.img_shape, (0, 0), (0, 0), 0.0, 0.0,
                                       scene.block_size)
            Iref = int(theirs["cum_tiles_hit"][-1])
            if Iref > 0:
                rid, rbins = ref_cuda.sort_and_bin(desc, Iref, theirs["xys"], theirs["depths"],
                                                   theirs["radii"], theirs["cum_tiles_hit"])
                # re-run our forward so the cached binning is the forward's
                _ops.rasterize_fwd(scene.colors, scene.opacities, scene.background, mine["xys"],
                                   mine["depths"], mine["radii"], mine["conics"],
                                   mine["cum_tiles_hit"], num_points=entry["N"],
                                   img_shape=scene.img_shape, block_width=scene.block_size)
                oid = torch.zeros(I, dtype=torch.int32, device=dev)
                obins = torch.zeros_like(rbins)
                got = _lib.lib().jsb200_debug_binning(
                    ctypes.c_void_p(torch.cuda.current_stream().cuda_stream), None, oid.data_ptr(),
                    obins.data_ptr(), None)
                torch.cuda.synchronize()
                rows.append(dict(name="num_intersects", mismatches=int(got != Iref), total=1))
                if got == Iref:
                    rows.append(compare(oid, rid, "gaussian_ids_sorted", True))
                    rows.append(compare(obins, rbins, "tile_bins", True))
            for r in rows:
                print("   ", json.dumps(r), flush=True)
            entry["parity"] = rows
            # reference self-noise on grads (atomic order)
            theirs2 = run_pipeline(scene, ref, target)
            entry["ref_self_noise"] = [compare(theirs2[k], theirs[k], k, False)
                                       for k in ("v_colors", "v_opacity", "v_xy", "v_conic",
                                                 "v_mean3d", "v_scale", "v_quat")]
            for r in entry["ref_self_noise"]:
                print("    ref-vs-ref", json.dumps(r), flush=True)
        entry["ms_ours"] = time_pipeline(scene, ours, target, args.iters)
        print(f"   ours fwd+bwd median ms: {entry['ms_ours']:.3f}", flush=True)
        if ref is not None:
            entry["ms_ref"] = time_pipeline(scene, ref, target, args.iters)
            print(f"   ref  fwd+bwd median ms: {entry['ms_ref']:.3f}  speedup {entry['ms_ref']/entry['ms_ours']:.2f}x", flush=True)
        report[cfg] = entry
        del scene, mine
        torch.cuda.empty_cache()
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    with open(args.out, "w") as f:
        json.dump(report, f, indent=1)


if __name__ == "__main__":
    main()
