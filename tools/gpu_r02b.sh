#!/bin/bash
# Runs ON the GPU box: GPU tests, bench (default / old sort), launch list, sanitizer on the new kernels
TAG=${1:-r02b}
OUT=gpurun_out
mkdir -p $OUT
timeout 1200 python -m pytest tests -m gpu -q -x > $OUT/${TAG}_pytest.log 2>&1
tail -15 $OUT/${TAG}_pytest.log
timeout 300 python bench.py --steps 100 --warmup 10 --no-cpu-baseline > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err
JSB_SORT1K=0 timeout 300 python bench.py --steps 100 --warmup 10 --no-cpu-baseline > $OUT/${TAG}_bench_oldsort.json 2> $OUT/${TAG}_bench_oldsort.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv \
    --log-file $OUT/${TAG}_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline \
    > $OUT/${TAG}_launches_bench.log 2>&1
timeout 600 compute-sanitizer --tool memcheck python -m pytest tests -m gpu -q -x -k "resident_depth_sort and (C1 or ragged or all_culled or one_depth) or test_project_fwd or test_project_bwd" > $OUT/${TAG}_sanitizer.log 2>&1
tail -5 $OUT/${TAG}_sanitizer.log
python - <<PY
import json
for name in ("bench", "bench_oldsort"):
    f = "${OUT}/${TAG}_%s.json" % name
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, 'ms/step', round(d['ms_per_step'], 4), 'e2e', round(d['e2e']['ms_per_step'], 4), d['clocks'])
        print(' | '.join(f"{k} {v['ms_per_launch']*1000:.1f}x{v['launches_per_step']:.0f}" for k, v in d['stages'].items()))
        print(d.get('roofline'), d.get('pairs'))
    except Exception as e:
        print(f, 'unreadable', e)
PY
