#!/bin/bash
# Runs ON the GPU box: the round's single-GPU measurement set (tests, every bench configuration with
# the reference arm beside it, launch list, ncu --set full of the top kernels, memcheck).
TAG=${1:-r03}
OUT=gpurun_out
mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -q -x > $OUT/${TAG}_pytest.log 2>&1
tail -3 $OUT/${TAG}_pytest.log
timeout 600 python bench.py > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err
timeout 300 python bench.py --impl reference --steps 30 --warmup 5 > $OUT/${TAG}_bench_reference.json 2> $OUT/${TAG}_bench_reference.err
for C in C1 C2 C3s C3b; do
  timeout 400 python bench.py --config $C --steps 50 --warmup 5 > $OUT/${TAG}_bench_$C.json 2> $OUT/${TAG}_bench_$C.err
  timeout 400 python bench.py --config $C --impl reference --steps 10 --warmup 3 > $OUT/${TAG}_bench_reference_$C.json 2> $OUT/${TAG}_bench_reference_$C.err
done
timeout 600 python bench.py --config C4 --steps 3 --warmup 3 > $OUT/${TAG}_bench_C4_1gpu.json 2> $OUT/${TAG}_bench_C4_1gpu.err
timeout 600 python bench.py --config C5 --steps 5 --warmup 3 > $OUT/${TAG}_bench_C5_1gpu.json 2> $OUT/${TAG}_bench_C5_1gpu.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv \
    --log-file $OUT/${TAG}_launches_C3.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline \
    > $OUT/${TAG}_launches_bench.log 2>&1
# (the .ncu-rep of a whole step with sources is larger than what gpurun copies back: exported to
# CSV pages here, the report itself stays on the box)
timeout 900 ncu --set full --clock-control none --import-source on -s 28 -c 14 -o /tmp/${TAG}_ncu_full -f \
    python tools/ncu_step.py C3 4 > $OUT/${TAG}_ncu_full.log 2>&1
ncu -i /tmp/${TAG}_ncu_full.ncu-rep --page raw --csv > $OUT/${TAG}_ncu_full_raw.csv 2>/dev/null
ncu -i /tmp/${TAG}_ncu_full.ncu-rep --page source --csv --print-source sass -k regex:"raster_bwd_hits|raster_fwd_warp" > $OUT/${TAG}_ncu_source_rasterisers.csv 2>/dev/null
ls -la /tmp/${TAG}_ncu_full.ncu-rep $OUT/${TAG}_ncu_*.csv
SEL="test_super_column or test_block_size_8 or test_lean_residuals or test_render_views_equals or test_fused_render_equals or test_hit_driven_backward_through"
timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 7 --print-limit 20 \
    python -m pytest tests -m gpu -q -x -k "$SEL" > $OUT/${TAG}_sanitize_memcheck.log 2>&1
echo "memcheck rc=$?"; grep -E "ERROR SUMMARY|passed|failed|Invalid|out of bounds" $OUT/${TAG}_sanitize_memcheck.log | tail -6
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 7 --print-limit 10 \
    python -m pytest tests -m gpu -q -x -k "test_super_column_lists_render_the_same_image_and_gradients and (ragged or tiny) or test_block_size_8_tuned_kernels_equal_the_generic_ones and tiny or test_hit_driven_backward_equals_the_batch_kernel and not C2" > $OUT/${TAG}_sanitize_racecheck.log 2>&1
echo "racecheck rc=$?"; grep -E "RACECHECK SUMMARY|passed|failed" $OUT/${TAG}_sanitize_racecheck.log | tail -4
python - <<PY
import json, glob
for f in sorted(glob.glob("${OUT}/${TAG}_bench*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        e = d.get('e2e') or {}
        print(f, 'ms/step', round(d['ms_per_step'], 4), d.get('unit'), round(d.get('value', 0), 2), 'e2e', e.get('ms_per_step'), 'cpu', (d.get('cpu_baseline') or {}).get('value'))
    except Exception as ex:
        print(f, 'unreadable', ex)
PY
