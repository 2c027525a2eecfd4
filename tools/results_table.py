"""Markdown rows of BASELINE.md's results table from the bench lines of one measurement set.
usage: python tools/results_table.py [dir=profiles] [tag=r03]"""
import json
import os
import sys

d = sys.argv[1] if len(sys.argv) > 1 else "profiles"
tag = sys.argv[2] if len(sys.argv) > 2 else "r03"


def line(name):
    p = os.path.join(d, name)
    if not os.path.exists(p):
        return None
    txt = open(p).read().strip().splitlines()
    for t in reversed(txt):
        try:
            return json.loads(t)
        except Exception:
            continue
    return None


for cfg, ours, ref in (("C1", f"{tag}_bench_C1.json", f"{tag}_bench_reference_C1.json"),
                       ("C2", f"{tag}_bench_C2.json", f"{tag}_bench_reference_C2.json"),
                       ("C3", f"{tag}_bench.json", f"{tag}_bench_reference.json"),
                       ("C3s", f"{tag}_bench_C3s.json", f"{tag}_bench_reference_C3s.json"),
                       ("C3b", f"{tag}_bench_C3b.json", f"{tag}_bench_reference_C3b.json"),
                       ("C4", f"{tag}_bench_C4_1gpu.json", None), ("C5", f"{tag}_bench_C5_1gpu.json", None)):
    o, r = line(ours), line(ref) if ref else None
    if not o:
        print(cfg, "missing")
        continue
    e = (o.get("e2e") or {}).get("ms_per_step")
    cpu = (o.get("cpu_baseline") or {})
    row = f"| {cfg} | I={o.get('config', {}).get('num_intersects')} | ours {o['ms_per_step']:.3f} ms"
    if e:
        row += f" (e2e {e:.3f})"
    if cpu.get("value"):
        row += f" | cpu {cpu['value']:.3f} {cpu.get('unit')} on {cpu.get('cores')} cores"
    if r:
        re = (r.get("e2e") or {}).get("ms_per_step")
        row += f" | ref {r['ms_per_step']:.2f} ms"
        if r.get("ms_per_step_min"):
            row += f" [min {r['ms_per_step_min']:.2f}]"
        if re:
            row += f" (e2e {re:.2f})"
        row += f" | x{r['ms_per_step'] / o['ms_per_step']:.1f}"
        if e and re:
            row += f" (e2e x{re / e:.1f})"
        rk = r.get("reference_kernel_ms")
        if isinstance(rk, dict):
            rk = rk.get("value") or rk.get("ms")
        if isinstance(rk, (int, float)):
            row += f" | ref kernels {rk:.2f} ms"
    if o.get("stages"):
        row += " | " + ", ".join(f"{k} {v['ms_per_launch'] * 1000:.1f}" for k, v in o["stages"].items())
    rf = o.get("roofline") or {}
    if rf:
        row += f" | roofline {rf.get('bound')} {rf.get('frac')}"
    print(row)
