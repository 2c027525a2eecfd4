#!/bin/bash
# Runs ON the GPU box (through gpurun, ONE GPU): everything round 1 left unmeasured, plus the
# regression set, in one call.  usage: tools/gpu_round2_open.sh [TAG]
#   GPU tests (incl. the packed-vs-scalar rasteriser tests added after the last round-1 GPU run),
#   bench lines for C3 and its two secondary scale distributions (C3s, C3b), the
#   reference arm, and the single-GPU points of the two sharded configurations (C4 views, C5 bands).
TAG=${1:-r02a}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm,power.limit --format=csv > $OUT/${TAG}_gpu.txt 2>&1
nproc >> $OUT/${TAG}_gpu.txt
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/${TAG}_pytest.log 2>&1
tail -3 $OUT/${TAG}_pytest.log
timeout 600 python bench.py > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err
timeout 300 python bench.py --impl reference --steps 30 --warmup 5 > $OUT/${TAG}_bench_reference.json 2> $OUT/${TAG}_bench_reference.err
# the experimental options (written without a GPU): their parity tests, then the A/B lines
JSB_EXPERIMENTAL=1 timeout 300 python -m pytest tests -m gpu -q -k "live_bit or one_barrier or raster_bwd_event or native_mse" > $OUT/${TAG}_pytest_sortlive.log 2>&1
tail -3 $OUT/${TAG}_pytest_sortlive.log
JSB_SORTLIVE=1 timeout 300 python bench.py --steps 100 --warmup 10 --no-cpu-baseline > $OUT/${TAG}_bench_sortlive.json 2> $OUT/${TAG}_bench_sortlive.err
JSB_FWD1BAR=1 timeout 300 python bench.py --steps 100 --warmup 10 --no-cpu-baseline > $OUT/${TAG}_bench_fwd1bar.json 2> $OUT/${TAG}_bench_fwd1bar.err
JSB_BENCH_NATIVE_LOSS=1 timeout 300 python bench.py --steps 100 --warmup 10 --no-cpu-baseline > $OUT/${TAG}_bench_nativeloss.json 2> $OUT/${TAG}_bench_nativeloss.err
JSB_BWD1BAR=1 timeout 300 python bench.py --steps 100 --warmup 10 --no-cpu-baseline > $OUT/${TAG}_bench_bwd1bar.json 2> $OUT/${TAG}_bench_bwd1bar.err
for C in C3s C3b; do
  timeout 400 python bench.py --config $C --steps 50 --warmup 5 --no-cpu-baseline > $OUT/${TAG}_bench_$C.json 2> $OUT/${TAG}_bench_$C.err
  timeout 400 python bench.py --config $C --impl reference --steps 10 --warmup 3 > $OUT/${TAG}_bench_reference_$C.json 2> $OUT/${TAG}_bench_reference_$C.err
done
# reference kernel-time baseline: launch list of the reference arm's C3 step
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file $OUT/${TAG}_reference_launches_C3.csv python bench.py --impl reference --steps 2 --warmup 1 > $OUT/${TAG}_reference_ncu.log 2>&1
# depth sort vs CUB on the same keys
timeout 120 tools/microbench/depth_sort_vs_cub 1000000 200 > $OUT/${TAG}_depth_sort_vs_cub.json 2> $OUT/${TAG}_depth_sort_vs_cub.err
timeout 120 tools/microbench/depth_sort_vs_cub 3000000 100 >> $OUT/${TAG}_depth_sort_vs_cub.json 2>> $OUT/${TAG}_depth_sort_vs_cub.err
timeout 400 python tools/multigpu_bench.py views --steps 3 --warmup 2 > $OUT/${TAG}_views_C4_1gpu.json 2> $OUT/${TAG}_views_C4_1gpu.err
timeout 400 python tools/multigpu_bench.py bands --steps 5 --warmup 2 > $OUT/${TAG}_bands_C5_1gpu.json 2> $OUT/${TAG}_bands_C5_1gpu.err
python - <<EOF
import glob, json
for f in sorted(glob.glob("${OUT}/${TAG}_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        e = d.get('e2e') or {}
        print(f, 'ms/step', round(d['ms_per_step'], 4), d.get('unit'), round(d.get('value', 0), 2),
              'e2e', e.get('ms_per_step'))
    except Exception as ex:
        print(f, 'unreadable', ex)
EOF
