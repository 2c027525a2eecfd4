#!/bin/bash
# Runs ON the GPU box: GPU tests, bench A/B (default / old sort / batch backward), launch list
TAG=${1:-r02c}
OUT=gpurun_out
mkdir -p $OUT
timeout 1200 python -m pytest tests -m gpu -q -x > $OUT/${TAG}_pytest.log 2>&1
tail -15 $OUT/${TAG}_pytest.log
timeout 300 python bench.py --steps 100 --warmup 10 --no-cpu-baseline > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err
JSB_SORT1K=0 timeout 300 python bench.py --steps 100 --warmup 10 --no-cpu-baseline > $OUT/${TAG}_bench_oldsort.json 2> $OUT/${TAG}_bench_oldsort.err
JSB_BWDHITS=0 timeout 300 python bench.py --steps 100 --warmup 10 --no-cpu-baseline > $OUT/${TAG}_bench_batchbwd.json 2> $OUT/${TAG}_bench_batchbwd.err
timeout 300 python bench.py --config C3s --steps 50 --warmup 5 --no-cpu-baseline > $OUT/${TAG}_bench_C3s.json 2> $OUT/${TAG}_bench_C3s.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv \
    --log-file $OUT/${TAG}_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline \
    > $OUT/${TAG}_launches_bench.log 2>&1
timeout 900 compute-sanitizer --tool memcheck python -m pytest tests -m gpu -q -x -k "hit_driven or (resident_depth_sort and (C1 or ragged or all_culled))" > $OUT/${TAG}_sanitizer.log 2>&1
tail -4 $OUT/${TAG}_sanitizer.log
python - <<PY
import json
for name in ("bench", "bench_oldsort", "bench_batchbwd", "bench_C3s"):
    f = "${OUT}/${TAG}_%s.json" % name
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, 'ms/step', round(d['ms_per_step'], 4), 'e2e', round(d['e2e']['ms_per_step'], 4), 'loss', d['loss'])
        print(' | '.join(f"{k} {v['ms_per_launch']*1000:.1f}x{v['launches_per_step']:.0f}" for k, v in d['stages'].items()))
    except Exception as e:
        print(f, 'unreadable', e)
PY
