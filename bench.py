#!/usr/bin/env python
"""bench.py -- fwd+bwd render throughput of the splatting hot path on N B200s (one node).

Contract (driver): `python bench.py --gpus N --steps K --warmup W [--impl reference]`, launched
under torchrun for N > 1, prints ONE JSON line on rank 0.

Workload (BASELINE.json configs[2]): 1M synthetic Gaussians, 1920x1080, block_size 16; one step =
one view's render forward + backward of loss = mean((img - target)^2) w.r.t. the five parameter
arrays, through the public API (jaxsplat_b200.render + autograd) -> the fused C-ABI render calls.
At N > 1 the views shard across ranks (one view per rank per step, Gaussians replicated, weak
scaling); the render backward writes the 14*N-float parameter gradient in place into one flat
symmetric buffer (jaxsplat_b200.set_gradient_sink) that is summed with one all-reduce per step
(the in-switch multimem kernel of csrc/collective.cu where the node has NVLS, NCCL otherwise).

  value  : views/s, whole job, parameters/targets resident in HBM (CUDA events, max over ranks);
           the library's per-stage events are OFF during these K steps
  stages / roofline / pairs : per-stage device times from a separate pass of the same step with
           the stage events on (outside the timed region)
  e2e    : same step, but the step's host inputs (8-bit target image + camera, pinned) are copied H2D
           and the loss is read back D2H inside the timed region (wall clock around K steps);
           the copy of step k+1 runs on a copy stream under step k
  --impl reference : the UNMODIFIED reference CUDA library (oracle/_ref, built from
           /root/reference/lib by oracle/Makefile) driven through the same harness on rank 0.
"""
import argparse
import json
import os
import statistics
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "views/s, render fwd+bwd @1M Gaussians 1920x1080 (ms_per_step = fwd+bwd ms per view)"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="C3")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    return ap.parse_args()


class ClockSampler:
    """Samples SM clock / throttle reasons in-process through NVML every few milliseconds from a
    thread that runs from before the warm-up to after the timed region; `mark()` brackets the timed
    region so that the reported median is taken under load (nvidia-smi -lms 100 saw 0-2 samples of
    a 170 ms region)."""
    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown",
               0x4: "sw_power_cap"}

    def __init__(self, torch_device, period_s=0.004):
        self.samples, self.marks, self.stop_flag, self.thread = [], [], False, None
        self.period, self.err, self.h, self.nv = period_s, None, None, None
        try:
            import pynvml
            import torch
            pynvml.nvmlInit()
            self.nv = pynvml
            try:
                uuid = "GPU-" + str(torch.cuda.get_device_properties(torch_device).uuid)
                self.h = pynvml.nvmlDeviceGetHandleByUUID(uuid.encode())
            except Exception:
                vis = os.environ.get("CUDA_VISIBLE_DEVICES")
                idx = torch_device.index or 0
                if vis:
                    idx = int(vis.split(",")[idx])
                self.h = pynvml.nvmlDeviceGetHandleByIndex(idx)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
        except Exception as ex:  # no NVML: report it, never fail the bench
            self.err = f"{type(ex).__name__}: {ex}"

    def _run(self):
        nv = self.nv
        while not self.stop_flag:
            try:
                mhz = float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    reasons = int(nv.nvmlDeviceGetCurrentClocksEventReasons(self.h))
                except Exception:
                    reasons = int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
                self.samples.append((time.perf_counter(), mhz, reasons))
            except Exception as ex:
                self.err = f"{type(ex).__name__}: {ex}"
                return
            time.sleep(self.period)

    def start(self):
        if self.h is not None:
            self.thread = threading.Thread(target=self._run, daemon=True)
            self.thread.start()

    def mark(self):
        self.marks.append(time.perf_counter())

    def stop(self):
        self.stop_flag = True
        if self.thread:
            self.thread.join(timeout=1.0)
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [self.err or "no NVML samples"],
                    "samples": 0}
        lo, hi = (self.marks + [0.0, 1e30])[:2] if len(self.marks) >= 2 else (0.0, 1e30)
        inside = [s for s in self.samples if lo <= s[0] <= hi]
        where = "timed region"
        if len(inside) < 3:  # a very short timed region: fall back to warm-up + timed region
            inside, where = self.samples, "warm-up + timed region"
        reasons = set()
        for _, _, r in inside:
            for bit, name in self.REASONS.items():
                if r & bit:
                    reasons.add(name)
        return {"sm_mhz": statistics.median(s[1] for s in inside), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(reasons), "samples": len(inside), "sampled_over": where,
                "source": "NVML in-process, %.0f ms period" % (self.period * 1e3)}


def algorithmic_bytes(stage, N, I, P, T, rows=0, tiles_x=120, E=0, shift=0):
    """DESIGN.md 'algorithmic bytes per unit' (SURVEY.md 8d), per launch of the stage.
    rows = row entries (sum of tile-box heights) of the two-level list distribution; shift > 0:
    the lists are super-column lists of E 8-byte entries (2^shift tiles per super-column)."""
    if shift:
        # level 2 distributes 8-byte row entries into ceil(tiles_x / 2^shift) bins per row
        cols2 = ((tiles_x - 1) >> shift) + 1
        k2 = 256 if rows == 0 or E == 0 else max(32, min(256, (640 * 3 * rows // (4 * E)) & ~31))
        chunks = rows // k2 + (T // max(tiles_x, 1))
        lists = {"tile_offsets": rows * 8 + 3 * chunks * cols2 * 4 + (T >> shift) * (4 + 8),
                 "tile_scatter": rows * 8 + chunks * cols2 * 4 + E * 8,
                 "raster_fwd": P * 20 + N * 48 + E * 8,
                 "raster_bwd": P * 24 + N * 48 + E * 8 + N * 36}
        if stage in lists:
            return lists[stage]
    k2 = 256 if rows == 0 or I == 0 else max(32, min(256, (1536 * 3 * rows // (4 * I)) & ~31))
    chunks = rows // k2 + (T // max(tiles_x, 1))  # level-2 chunks (upper bound the grid uses)
    return {
        "project_fwd": N * 108,
        "bin_prep": N * (44 + 48 + 4 + 4 + 8),
        # resident sort, three passes over the live bits: keys read, (passes - 1) x (keys + ids
        # written and read back), ids written
        "depth_sort": N * (4 + 2 * 16 + 4),
        # level 1: order + box gathered twice (count, scatter), one 8-byte entry per covered row
        "emit": N * (4 + 8) * 2 + rows * 8,
        # level-2 counts: entries read once, count matrix written, scanned in place, tile bins
        "tile_offsets": rows * 8 + 3 * chunks * tiles_x * 4 + T * (4 + 8),
        # level-2 scatter: entries + run starts read, every id written once
        "tile_scatter": rows * 8 + chunks * tiles_x * 4 + I * 4,
        # rasterisers: compulsory HBM bytes only (pixels in/out, every record once, the id list
        # once, every gradient word once); the per-(tile, Gaussian) staging traffic is L2 hits
        "raster_fwd": P * 20 + N * 48 + I * 4,
        "raster_bwd": P * 24 + N * 48 + I * 4 + N * 36,
        # 188 B per Gaussian through the four-op API; the fused render neither writes v_cov2d /
        # v_cov3d nor reads the compensation or the 3D covariance (recomputed), and finishes the
        # opacity moment in place: 128 B
        "project_bwd": N * 128,
    }[stage]


def useful_pairs(sc, kw):
    """U_f = sum_p max(0, final_idx[p] + 1 - tile_bins[tile(p)].x), from the outputs of
    one forward (outside every timed region)."""
    import ctypes

    import torch

    from jaxsplat_b200 import _lib, _ops
    H, W = sc.img_shape
    bs = sc.block_size
    n = sc.means3d.shape[0]
    # (with the exact per-tile lists: final_idx of the super-column lists counts other tiles' entries)
    _lib.set_option("supercols", 0)
    try:
        outs, _tok = _ops.render_fwd(sc.means3d, sc.scales, sc.quats, sc.viewmat, sc.colors, sc.opacities,
                                     sc.background, num_points=n, img_shape=sc.img_shape, f=sc.f, c=sc.c,
                                     glob_scale=sc.glob_scale, clip_thresh=sc.clip_thresh, block_width=bs)
        final_idx = outs[7].view(H, W).long()
        tx, ty = (W + bs - 1) // bs, (H + bs - 1) // bs
        bins = torch.zeros((tx * ty, 2), dtype=torch.int32, device=final_idx.device)
        got = _lib.lib().jsb200_debug_binning(ctypes.c_void_p(torch.cuda.current_stream().cuda_stream),
                                              None, None, ctypes.c_void_p(bins.data_ptr()), None)
        torch.cuda.synchronize()
    finally:
        _lib.set_option("supercols", int(os.environ.get("JSB_SUPERCOLS", "2")))
    if got < 0:
        _lib.lib().jsb200_clear_error()
        return None
    first = bins[:, 0].long().view(ty, tx)
    count = (bins[:, 1] - bins[:, 0]).long().view(ty, tx)
    first_px = first.repeat_interleave(bs, 0).repeat_interleave(bs, 1)[:H, :W]
    count_px = count.repeat_interleave(bs, 0).repeat_interleave(bs, 1)[:H, :W]
    u = (final_idx + 1 - first_px).clamp(min=0)
    u = torch.where(count_px > 0, u, torch.zeros_like(u))
    return int(u.sum().item())


def sharded_config(args, real_stdout):
    """--config C4 / C5: the two sharded configurations of BASELINE.json (tools/multigpu_bench.py:
    a 64-view batch of 3M Gaussians split across the ranks, one all-reduce per batch; one
    3840x2160 frame of 6M Gaussians split into tile-row bands), reported in the bench contract."""
    import importlib.util
    import types
    spec = importlib.util.spec_from_file_location("multigpu_bench", os.path.join(ROOT, "tools", "multigpu_bench.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    ns = types.SimpleNamespace(mode="views" if args.config == "C4" else "bands", steps=args.steps,
                               warmup=max(args.warmup, 3), views=64, chunk=8, config=args.config,
                               num_points=None, dry=False)
    r = mod.run(ns)
    if r is None:
        return
    unit = r["unit"]
    out = {"metric": ("views/s, render fwd+bwd of a 64-view batch @3M Gaussians 1920x1080" if args.config == "C4"
                      else "frames/s, render fwd+bwd of one 3840x2160 frame @6M Gaussians, tile-row bands"),
           "value": r["value"], "unit": unit, "n_gpus": r["n_gpus"], "steps": r["steps"],
           "warmup": r["warmup"], "ms_per_step": r["ms_per_step"], "higher_is_better": True,
           "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
           "config": {"workload": r["workload"], "reduction": r["reduction"],
                      "timing": "CUDA events around exactly K steps, barrier+synchronize both sides, max over ranks"},
           "loss": r["loss"], "grad_l2": r["grad_l2"], "gpu_launches": None,
           "extra": {k: v for k, v in r.items() if k in ("bands", "band_intersections", "num_intersects", "workspace_bytes")}}
    print(json.dumps(out), file=real_stdout, flush=True)


def main():
    args = parse()
    # the contract is ONE JSON line on stdout: libraries that chat on fd 1 (NCCL prints its
    # version banner there) are pointed at stderr, the JSON goes to the saved descriptor
    real_stdout = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    if args.config in ("C4", "C5") and args.impl == "ours":
        return sharded_config(args, real_stdout)
    import torch
    import torch.distributed as dist

    # The backward of a step is one chain (loss -> render): run it in the calling thread instead of
    # handing it to autograd's worker thread and waiting for the wake-up (both arms; 40-50 us of
    # host time per step, visible where the GPU work is shorter than the host's: C1, C2)
    if os.environ.get("JSB_BENCH_AUTOGRAD_THREADS") != "1":
        torch.autograd.set_multithreading_enabled(False)

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference" and rank != 0:
        return  # the reference is single-GPU: rank 0 alone runs it
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    use_dist = world > 1 and args.impl == "ours"
    if use_dist:
        dist.init_process_group("nccl", device_id=dev)

    import jaxsplat_b200 as jsb
    from jaxsplat_b200 import _lib, _ops, synthetic

    if args.impl == "reference":
        from oracle import ref_cuda
        if not ref_cuda.available():
            print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/_jaxsplat*.so not built"}),
                  file=real_stdout, flush=True)
            return
        _ops.set_default_table(ref_cuda.table())

    scene = synthetic.make_config(args.config)
    H, W = scene.img_shape
    N = scene.means3d.shape[0]
    if world > 1 and args.impl == "ours":
        scene.viewmat = synthetic.ring_cameras(world, radius=8.0)[rank]  # one view per rank
    sc = scene.to(dev)
    gen = torch.Generator().manual_seed(1234 + rank)
    # the target image is what a data loader hands over: 8 bits per channel (uniform noise, there
    # are no image files here), decoded to float32 in [0, 1] on the device
    target_host = (torch.rand((H, W, 3), generator=gen) * 255.0).round().to(torch.uint8).pin_memory()
    viewmat_host = scene.viewmat.clone().pin_memory()
    target = target_host.to(dev).to(torch.float32).mul_(1.0 / 255.0)
    kw = sc.render_kwargs()

    # A/B switches for profiling runs (defaults are the product configuration)
    if args.impl == "ours":
        if os.environ.get("JSB_ROWCOL") == "0":
            _lib.set_option("rowcol", 0)
        if os.environ.get("JSB_FWDWARP") == "0":  # per-CTA batch forward instead of the barrier-free one
            _lib.set_option("fwdwarp", 0)
        if os.environ.get("JSB_BWDMOM") == "0":  # hit-driven backward emits finished gradients
            _lib.set_option("bwdmom", 0)
        if os.environ.get("JSB_BWDHITS") == "0":  # per-tile batch backward instead of the hit-driven one
            _lib.set_option("bwdhits", 0)
        if os.environ.get("JSB_SUPERCOLS"):  # tiles per super-column of the fused render's lists (log2; 0 = exact lists)
            _lib.set_option("supercols", int(os.environ["JSB_SUPERCOLS"]))
        if os.environ.get("JSB_PRECLEAR") == "0":  # the backward's accumulators cleared by its own memset
            jsb.set_preclear(False)
        if os.environ.get("JSB_PDL"):  # programmatic dependent launch between the step's kernels
            _lib.set_option("pdl", int(os.environ["JSB_PDL"]))
        if os.environ.get("JSB_SORT1K") == "0":  # four onesweep passes instead of the resident sort
            _lib.set_option("sort1k", 0)
        if os.environ.get("JSB_FUSED") == "0":
            jsb.set_fused_render(False)
    # N > 1: the five gradients are packed into one flat symmetric buffer so the all-reduce is a
    # single call: the in-switch multimem kernel (csrc/collective.cu) where the node has NVLS
    from jaxsplat_b200.multiview import SwitchReducedBuffer
    flat_n = sum(p.numel() for p in sc.params())
    if use_dist and os.environ.get("JSB_BENCH_NCCL") == "1":
        reducer = None
        flat_raw = torch.zeros(flat_n + 5 * N, dtype=torch.float32, device=dev)
        flat_g, flat_scratch = flat_raw[:flat_n], flat_raw[flat_n:]
    else:
        reducer = SwitchReducedBuffer(flat_n, dev, scratch_numel=5 * N) if use_dist else None
        flat_g = reducer.tensor if reducer is not None else None
        flat_scratch = reducer.scratch if reducer is not None else None
    params = [p.clone().requires_grad_(True) for p in sc.params()]
    # the render backward writes the five parameter gradients straight into the flat buffer the
    # all-reduce sums in place (no packing pass); checked after the first step, torch.cat otherwise
    sink_views = None
    if use_dist and args.impl == "ours" and os.environ.get("JSB_BENCH_NOSINK") != "1":
        sink_views, off = [], 0
        for q in params:
            sink_views.append(flat_g[off:off + q.numel()].view(q.shape))
            off += q.numel()
        jsb.set_gradient_sink(sink_views + [flat_scratch])
    sink_ok = [False]
    # the MSE and its cotangent come from one kernel of the library (csrc/loss.cu: 0.841 vs 0.859 ms
    # per step, profiles/r02a_bench_nativeloss.json); JSB_BENCH_NATIVE_LOSS=0 = three torch kernels
    # (the reference arm always uses those: its library has no loss kernel)
    native_loss = args.impl == "ours" and os.environ.get("JSB_BENCH_NATIVE_LOSS") != "0"
    if native_loss:
        from jaxsplat_b200 import losses

    def step(view, tgt):
        for q in params:
            q.grad = None  # optimizer.zero_grad(set_to_none=True): no memset, no accumulate pass
        k = dict(kw)
        k["viewmat"] = view
        img = jsb.render(*params, **k)
        # MSE and its gradient written out by hand (sub, dot, scale instead of the 9 kernels that
        # autograd launches for ((img - tgt) ** 2).mean()); the render backward is autograd's
        if native_loss:  # loss + cotangent in one native kernel
            loss, v_img = losses.mse_loss_and_grad(img.detach(), tgt)
            img.backward(gradient=v_img)
        else:
            diff = img.detach() - tgt
            flat = diff.view(-1)
            loss = torch.dot(flat, flat) / flat.numel()
            img.backward(gradient=diff * (2.0 / diff.numel()))
        if use_dist:
            # the five parameter gradients in the one flat (symmetric) buffer, summed over the
            # ranks in a single call
            if not sink_ok[0]:
                sink_ok[0] = sink_views is not None and all(
                    q.grad.data_ptr() == v.data_ptr() for q, v in zip(params, sink_views))
            if not sink_ok[0]:
                torch.cat([q.grad.reshape(-1) for q in params], out=flat_g)
            if reducer is not None:
                reducer.all_reduce()
            else:
                dist.all_reduce(flat_g)
        return loss

    def barrier():
        if use_dist:
            dist.barrier()
        torch.cuda.synchronize()

    # ---------------- device-resident timing (value) ----------------
    sampler = ClockSampler(dev)
    sampler.start()
    for _ in range(max(args.warmup, 3)):
        step(sc.viewmat, target)
    barrier()
    launches0 = _lib.stats()["kernel_launches"] if args.impl == "ours" else 0
    # the reference's stock call path spends its step in cudaMalloc/cudaFree/syncs and its time
    # swings from run to run: time the K steps `repeats` times and report min / median too
    repeats = 5 if args.impl == "reference" else 1
    rep_ms = []
    for rep in range(repeats):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        if rep == 0:
            sampler.mark()
        e0.record()
        for _ in range(args.steps):
            step(sc.viewmat, target)
        e1.record()
        barrier()
        if rep == 0:
            sampler.mark()
        rep_ms.append(e0.elapsed_time(e1))
    clocks = sampler.stop()
    ms_total = statistics.median(rep_ms)
    stages = None
    launches = 0
    if args.impl == "ours":
        launches = _lib.stats()["kernel_launches"] - launches0
        # per-stage device times: a separate pass of the same step with the library's stage events
        # switched on, OUTSIDE the timed region (two event records per stage would otherwise sit
        # between the kernels of the steps being timed)
        stage_steps = min(args.steps, 50)
        _lib.set_option("timing", 1)
        _lib.stage_times(torch.cuda.current_stream().cuda_stream, reset=True)
        for _ in range(stage_steps):
            step(sc.viewmat, target)
        barrier()
        stages = _lib.stage_times(torch.cuda.current_stream().cuda_stream, reset=True)
        _lib.set_option("timing", 0)
    t = torch.tensor([ms_total], device=dev, dtype=torch.float64)
    if use_dist:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total = float(t.item())
    ms_per_step = ms_total / args.steps
    n_views = world if args.impl == "ours" else 1
    value = n_views * args.steps / (ms_total / 1e3)

    # ---------------- end to end: host inputs in, loss out, every step ----------------
    # Every step's target image + camera are copied from pinned host memory and every step's loss
    # is read back, all inside the timed region.  Like any input pipeline the copy of step k+1
    # runs on a copy stream under step k (two device buffers), and the host reads the loss of step
    # k-1 while step k is in flight, so neither PCIe nor the read-back stalls the GPU.
    # With the native loss the 8-bit image is the loss kernel's target as it is (decoded in
    # registers); the eager loss needs it decoded to floats by two more kernels on the copy stream.
    copy_stream = torch.cuda.Stream(device=dev)
    tgt_u8 = [torch.empty((H, W, 3), dtype=torch.uint8, device=dev) for _ in range(2)]
    tgt_buf = tgt_u8 if native_loss else [torch.empty_like(target) for _ in range(2)]
    view_buf = [torch.empty_like(sc.viewmat) for _ in range(2)]
    loss_pinned = torch.zeros(2, dtype=torch.float32).pin_memory()
    ev_copied = [torch.cuda.Event() for _ in range(2)]
    ev_free = [torch.cuda.Event() for _ in range(2)]
    ev_loss = [torch.cuda.Event() for _ in range(2)]

    def issue_copy(k):
        b = k % 2
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(ev_free[b])  # the step that last read this buffer is done
            tgt_u8[b].copy_(target_host, non_blocking=True)       # H2D: 3 bytes per pixel
            if not native_loss:
                tgt_buf[b].copy_(tgt_u8[b]).mul_(1.0 / 255.0)     # decode on the device
            view_buf[b].copy_(viewmat_host, non_blocking=True)
            ev_copied[b].record(copy_stream)

    def e2e_loop(steps):
        main = torch.cuda.current_stream()
        for b in range(2):
            ev_free[b].record(main)
        issue_copy(0)
        last = 0.0
        for k in range(steps):
            b = k % 2
            main.wait_event(ev_copied[b])
            loss = step(view_buf[b], tgt_buf[b])
            ev_free[b].record(main)
            loss_pinned[b:b + 1].copy_(loss.detach().reshape(1), non_blocking=True)  # D2H of the result
            ev_loss[b].record(main)
            if k + 1 < steps:
                issue_copy(k + 1)  # after step k is enqueued: the host never delays the GPU's next kernel
            if k >= 1:
                ev_loss[1 - b].synchronize()
                last = float(loss_pinned[1 - b])
        ev_loss[(steps - 1) % 2].synchronize()
        return float(loss_pinned[(steps - 1) % 2])

    e2e_loop(3)
    barrier()
    t0 = time.perf_counter()
    loss_host = e2e_loop(args.steps)
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    t = torch.tensor([e2e_s], device=dev, dtype=torch.float64)
    if use_dist:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_s = float(t.item())
    e2e = {"value": n_views * args.steps / e2e_s, "unit": "views/s",
           "ms_per_step": 1e3 * e2e_s / args.steps,
           "h2d_bytes_per_step": target_host.numel() * target_host.element_size() + viewmat_host.numel() * 4,
           "d2h_bytes_per_step": 4}

    if rank != 0:
        if use_dist:
            dist.destroy_process_group()
        return

    I = _lib.stats()["last_num_intersects"] if args.impl == "ours" else None
    rows = _lib.stats()["last_row_entries"] if args.impl == "ours" else 0
    lst = _lib.list_stats() if args.impl == "ours" else {"super_entries": 0, "super_shift": 0}
    T = ((W + 15) // 16) * ((H + 15) // 16)
    out = {
        "metric": METRIC, "value": value, "unit": "views/s", "n_gpus": world if args.impl == "ours" else 1,
        "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": {"workload": f"{args.config}: {N} Gaussians (benchmark.py distribution, "
                               f"scales {scene.meta.get('scales', '?')}, seed 0), "
                               f"{W}x{H}, block_size 16, fwd+bwd of MSE loss, 1 view/GPU/step",
                   "num_intersects": I, "parallelism": (f"view-parallel dp{world}, gradient all-reduce: "
                                   f"{reducer.mode if reducer is not None else 'nccl'}, gradients "
                                   f"{'written in place into' if sink_ok[0] else 'packed (torch.cat) into'} "
                                   "the reduced buffer") if world > 1 else "single GPU",
                   "lists": (f"super-column lists, {1 << lst['super_shift']} tiles per super-column, "
                             f"{lst['super_entries']} 8-byte entries" if lst["super_shift"] else "exact per-tile id lists"),
                   "l2": "per-step working set ~0.5 GB (keys/ids ping-pong 0.2 GB, records, images, "
                         "parameters, gradients) exceeds the 126 MB L2; no explicit flush",
                   "host": "torch autograd in the calling thread (set_multithreading_enabled(False))"
                           if os.environ.get("JSB_BENCH_AUTOGRAD_THREADS") != "1" else "torch autograd worker thread",
                   "timing": "CUDA events around exactly K steps, barrier+synchronize both sides, max over ranks"},
        "clocks": clocks, "e2e": e2e, "gpu_launches": launches,
        "loss": loss_host,
    }
    if args.impl == "reference":
        out["impl"] = "reference"
        out["config"]["reference"] = "unmodified yklcs/jaxsplat lib/*.cu built for sm_100a (oracle/_ref), same harness"
        out["gpu_launches"] = None
        out["repeats"] = {"ms_per_step": [m / args.steps for m in rep_ms],
                          "min_ms_per_step": min(rep_ms) / args.steps,
                          "median_ms_per_step": statistics.median(rep_ms) / args.steps,
                          "note": "K timed steps repeated 5 times; value / ms_per_step are the median "
                                  "repeat (the stock call path's cudaMalloc/cudaFree/syncs make it swing)"}
        if args.config == "C3":
            # sum of the reference's own kernel durations for one C3 fwd+bwd step, from the committed
            # ncu launch list of this same arm (cold-cache, serialised launches: an upper bound of
            # its kernel time inside a step; no allocator or sync time in it)
            out["reference_kernel_ms"] = {"value": 4.61,
                                          "source": "profiles/r02a_reference_launches_C3.csv (38 launches per step)"}
        # the reference has no CPU implementation of this path (its kernels are CUDA-only): this arm
        # is its own CUDA library on the same B200; the host-core timing is the `cpu_baseline` of
        # the default arm (oracle/oracle.c)
        out["cpu_baseline"] = {"value": value, "unit": "views/s", "cores": 0, "kind": "reference",
                               "sample": f"{args.steps} full fwd+bwd views of {args.config} through the "
                                         "unmodified reference CUDA kernels on this GPU (no host-core "
                                         "compute: the reference has no CPU path)"}
    if stages is not None:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        which_peak = "measured (MEASURED_PEAKS.json hbm_gbs)" if peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
        P = H * W
        per_stage = {}
        for k, v in stages.items():
            if v["calls"] == 0:
                continue
            ms = v["ms"] / v["calls"]
            b = algorithmic_bytes(k, N, I, P, T, rows, (W + 15) // 16, lst["super_entries"], lst["super_shift"])
            per_stage[k] = {"ms_per_launch": ms, "launches_per_step": v["calls"] / stage_steps,
                            "algorithmic_bytes": b, "gbs": b / (ms * 1e-3) / 1e9,
                            "hbm_frac": b / (ms * 1e-3) / 1e9 / peak}
            if k == "project_fwd":
                # SURVEY 8d's 108 B per Gaussian is the projection alone; the fused kernel also reads
                # colours + opacities and writes the rasteriser records, depth keys, ids and tile
                # boxes (what bin_prep would), and no longer writes cov3d / compensation (lean
                # residuals): 152 B per Gaussian actually moved
                per_stage[k]["moved_bytes"] = N * 152
                per_stage[k]["moved_hbm_frac"] = N * 152 / (ms * 1e-3) / 1e9 / peak
        dom = max(per_stage, key=lambda k: per_stage[k]["ms_per_launch"] * per_stage[k]["launches_per_step"])
        d = per_stage[dom]
        prof = {}
        try:
            prof = json.load(open(os.path.join(ROOT, "profiles", "dominant_kernel_traffic.json")))
        except Exception:
            pass
        traffic = prof.get(dom) if args.config == "C3" else None
        f_clk = (clocks.get("sm_mhz") or 1965.0) * 1e6
        # SURVEY.md 8d unit of the rasterisers: useful pixel-Gaussian pairs, from the outputs
        U_f = useful_pairs(sc, kw)
        pair_roof = 148 * 128 * f_clk / 15.0   # every pair evaluated at >= 15 FP32-pipe slots
        out["pairs"] = {"U_f": U_f, "roof_pairs_per_s": pair_roof,
                        "note": "U_f = sum over pixels of (final_idx + 1 - first list entry of the "
                                "tile): the pairs a tile-walking rasteriser without culling evaluates. "
                                "This library's conservative 8x8-quadrant cull skips most of them "
                                "unevaluated, so pairs/s can exceed the evaluate-everything roof; the "
                                "roofline object is in issued warp instructions instead"}
        for k in ("raster_fwd", "raster_bwd"):
            if k in per_stage and U_f:
                out["pairs"][k + "_pairs_per_s"] = U_f / (per_stage[k]["ms_per_launch"] * 1e-3)
        if dom in ("raster_fwd", "raster_bwd"):
            # issue-slot roofline of the dominant rasteriser in SURVEY 8d's unit: useful pixel-
            # Gaussian pairs per second against 148 SMs x 128 FP32 lanes x f_clk / 15 slots per pair
            # (U_f counts every pair of the walked list prefixes; the kernel evaluates the subset its
            # 8x8-quadrant pre-filter lets through, so where most pairs are skipped -- C3b's large
            # splats -- pairs/s can pass the evaluate-everything roof: 0.47 at C3, 1.06 at C3b; the
            # issue-slot view beside it cannot).  The raw issue-slot view: warp
            # instructions of one launch (ncu smsp__inst_executed.sum of the same kernel on the same
            # workload, committed under profiles/) / its device time, against 4 schedulers x 148 SMs.
            inst = (prof.get("_inst_executed") or {}).get(dom) if args.config == "C3" else None
            peak_issue = 148 * 4 * f_clk
            ach_i = (inst / (d["ms_per_launch"] * 1e-3)) if inst else None
            ach_p = (U_f / (d["ms_per_launch"] * 1e-3)) if U_f else None
            out["roofline"] = {"bound": "issue", "kernel": dom, "achieved": ach_p, "peak": pair_roof,
                               "unit": "pairs/s", "frac": (ach_p / pair_roof) if ach_p else None,
                               "traffic": traffic,
                               "issue_slots": {"inst_per_launch": inst, "achieved": ach_i, "peak": peak_issue,
                                               "unit": "warp-inst/s", "frac": (ach_i / peak_issue) if ach_i else None,
                                               "inst_source": prof.get("_inst_source")},
                               "hbm_view": {"achieved_gbs": d["gbs"], "peak_gbs": peak,
                                            "frac": d["hbm_frac"], "peak_source": which_peak},
                               "note": "the rasterisers are SM-issue bound (no dense contraction, no "
                                       "tensor cores; DRAM throughput 2-4 %); the HBM stages' fractions "
                                       "are in 'stages'"}
            if ach_p and ach_p > pair_roof:
                out["roofline"]["note"] += ("; frac > 1: U_f counts pairs the quadrant pre-filter "
                                            "skips unevaluated (see 'pairs')")
        else:
            out["roofline"] = {"bound": "hbm", "kernel": dom, "achieved": d["gbs"], "peak": peak,
                               "unit": "GB/s", "frac": d["hbm_frac"], "traffic": traffic,
                               "peak_source": which_peak}
        out["stages"] = per_stage

    # (the CPU oracle is timed on the configurations it finishes in seconds; C3b's 2.6e8
    # intersections would take minutes per view)
    if args.impl == "ours" and world == 1 and not args.no_cpu_baseline and args.config in ("C1", "C2", "C3", "C3s"):
        from oracle import cpu
        s_np = cpu.scene_to_numpy(scene)
        tgt_np = (target_host.to(torch.float32) / 255.0).numpy()
        cpu.render_fwd_bwd(s_np, tgt_np)  # warm
        reps, t0 = 0, time.perf_counter()
        while reps < 3 or (time.perf_counter() - t0 < 10.0 and reps < 20):
            cpu.render_fwd_bwd(s_np, tgt_np)
            reps += 1
        dt = time.perf_counter() - t0
        out["cpu_baseline"] = {"value": reps / dt, "unit": "views/s", "cores": cpu.num_threads(),
                               "kind": "port",
                               "sample": f"{reps} full fwd+bwd views of the same {args.config} workload "
                                         f"(oracle/oracle.c, OpenMP, {dt / reps:.2f} s/view); the reference "
                                         "itself has no CPU path, and BASELINE.json's pure-JAX jit oracle is "
                                         "replaced by this C restatement because JAX cannot be installed here"}
    print(json.dumps(out), file=real_stdout, flush=True)
    if use_dist:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
