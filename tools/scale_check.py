"""GPU: BASELINE configs 4 and 5 at full size on one GPU -- properties + timing vs the reference.
Usage: python tools/scale_check.py [C4 C5] [--no-ref]"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import gpu_util as G  # noqa: E402
from jaxsplat_b200 import _ffi, _lib, synthetic  # noqa: E402
from oracle import ref_cuda  # noqa: E402
from tools.parity_report import run_pipeline, time_pipeline  # noqa: E402


def main():
    cfgs = [a for a in sys.argv[1:] if not a.startswith("--")] or ["C4", "C5"]
    use_ref = "--no-ref" not in sys.argv and ref_cuda.available()
    ours = _ffi.native_table()
    out = {}
    for cfg in cfgs:
        scene = synthetic.make_config(cfg)
        if cfg == "C4":
            scene.viewmat = synthetic.ring_cameras(64, radius=8.0)[5]
        sc = scene.to("cuda:0")
        H, W = sc.img_shape
        target = torch.rand((H, W, 3), generator=torch.Generator().manual_seed(1)).to("cuda:0")
        mine = run_pipeline(sc, ours, target)
        torch.cuda.synchronize()
        I = int(mine["cum_tiles_hit"][-1])
        tiles = ((W + 15) // 16) * ((H + 15) // 16)
        got, ids, bins, keys = G.debug_binning(I, tiles, depths=mine["depths"])
        e = dict(N=sc.means3d.shape[0], img=[H, W], I=I, workspace_GB=_lib.stats()["workspace_bytes"] / 1e9)
        e["sorted"] = bool((keys[1:] >= keys[:-1]).all())
        cnt = (bins[:, 1] - bins[:, 0]).long()
        e["bins_partition"] = int(cnt.sum()) == I
        e["cumsum_ok"] = bool(torch.equal(torch.cumsum(mine["num_tiles_hit"].long().view(-1), 0),
                                          mine["cum_tiles_hit"].long().view(-1)))
        e["finite_grads"] = all(bool(torch.isfinite(mine[k]).all()) for k in
                                ("v_mean3d", "v_scale", "v_quat", "v_colors", "v_opacity"))
        del keys, ids
        torch.cuda.empty_cache()
        e["ms_ours"] = time_pipeline(sc, ours, target, iters=5)
        if use_ref:
            ref = ref_cuda.table()
            theirs = run_pipeline(sc, ref, target)
            torch.cuda.synchronize()
            e["final_idx_equal"] = bool(torch.equal(theirs["final_idx"], mine["final_idx"]))
            e["img_bits_equal"] = bool(torch.equal(theirs["img"].view(torch.int32), mine["img"].view(torch.int32)))
            e["radii_equal"] = bool(torch.equal(theirs["radii"], mine["radii"]))
            for k in ("v_mean3d", "v_scale", "v_quat", "v_colors", "v_opacity"):
                d = float((theirs[k].double() - mine[k].double()).abs().max())
                e["rel_" + k] = d / max(float(theirs[k].double().abs().max()), 1e-30)
            del theirs
            torch.cuda.empty_cache()
            e["ms_ref"] = time_pipeline(sc, ref, target, iters=3, warmup=1)
            e["speedup"] = e["ms_ref"] / e["ms_ours"]
        out[cfg] = e
        print(cfg, json.dumps(e), flush=True)
        del mine, sc
        torch.cuda.empty_cache()
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(out, open("gpurun_out/scale_check.json", "w"), indent=1)


if __name__ == "__main__":
    main()
