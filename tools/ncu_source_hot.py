"""Per-instruction warp-stall samples of an `ncu --page source --csv --print-source sass` export:
for every kernel the hottest SASS instructions and the contiguous hot loop (rows executed at least
30 % as often as the most executed one).  usage: python tools/ncu_source_hot.py source.csv [raw.csv]"""
import csv
import sys


def sections(rows):
    starts = [i for i, r in enumerate(rows) if r and r[0] == "Address"]
    for n, a in enumerate(starts):
        b = starts[n + 1] - 1 if n + 1 < len(starts) else len(rows)
        name = rows[a - 1][1] if a >= 1 and len(rows[a - 1]) > 1 else "?"
        yield name, rows[a], [r for r in rows[a + 1:b] if len(r) == len(rows[a])]


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    for name, hdr, body in sections(rows):
        isrc, iall, iex = (hdr.index("Source"), hdr.index("Warp Stall Sampling (All Samples)"),
                           hdr.index("Instructions Executed"))
        out = [(k, r[isrc].strip(), int(r[iall] or 0), int(r[iex] or 0)) for k, r in enumerate(body)]
        total = sum(s for _, _, s, _ in out)
        mx = max(e for _, _, _, e in out)
        hot = [k for k, _, _, e in out if e > 0.3 * mx]
        lo, hi = hot[0], hot[-1]
        inloop = sum(s for k, _, s, _ in out if lo <= k <= hi)
        print(f"===== {name[:100]}")
        print(f"warp-stall samples {total}, instructions executed {sum(e for _, _, _, e in out)}; "
              f"hot loop = rows {lo}..{hi}: {inloop} samples ({100.0 * inloop / max(total, 1):.1f} %)")
        print("-- 15 hottest instructions outside the loop (row, samples, executed, SASS)")
        for k, src, s, e in sorted((x for x in out if not lo <= x[0] <= hi), key=lambda x: -x[2])[:15]:
            print(f"{k:5d} {s:6d} {e:9d}  {src[:100]}")
        print("-- the hot loop (first 160 rows)")
        for k, src, s, e in out[lo:min(hi + 1, lo + 160)]:
            print(f"{k:5d} {s:6d} {e:9d}  {src[:100]}")
    if len(sys.argv) > 2:
        raw = list(csv.reader(open(sys.argv[2])))
        head, units = raw[0], raw[1]
        ki = head.index("Kernel Name")
        for r in raw[2:]:
            print(f"===== stall reasons per issued instruction: {r[ki][:80]}")
            for i, h in enumerate(head):
                if ("smsp__average_warps_issue_stalled" in h and "not_issued" not in h) or h in (
                        "smsp__issue_active.avg.pct_of_peak_sustained_active",
                        "smsp__warps_active.avg.per_cycle_active",
                        "smsp__warps_eligible.avg.per_cycle_active",
                        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
                        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
                        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
                        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
                        "gpu__time_duration.sum", "smsp__inst_executed.sum"):
                    print(f"    {h:86s} {r[i]} {units[i]}")


if __name__ == "__main__":
    main()
