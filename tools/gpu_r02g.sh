#!/bin/bash
TAG=${1:-r02g}
OUT=gpurun_out
mkdir -p $OUT
timeout 1200 python -m pytest tests -m gpu -q -x > $OUT/${TAG}_pytest.log 2>&1
tail -6 $OUT/${TAG}_pytest.log
timeout 300 python bench.py --steps 100 --warmup 10 --no-cpu-baseline > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err
JSB_SPEC=0 timeout 300 python bench.py --steps 100 --warmup 10 --no-cpu-baseline > $OUT/${TAG}_bench_nospec.json 2> $OUT/${TAG}_bench_nospec.err
for v in bwdh24 bwdh28 bwdh32 bwdh32s16; do
  JSB200_LIB=$PWD/build/alt/$v.so timeout 300 python bench.py --steps 100 --warmup 10 --no-cpu-baseline > $OUT/${TAG}_bench_$v.json 2> $OUT/${TAG}_bench_$v.err
done
timeout 300 python bench.py --config C2 --steps 200 --warmup 10 --no-cpu-baseline > $OUT/${TAG}_bench_C2.json 2> $OUT/${TAG}_bench_C2.err
JSB_SPEC=0 timeout 300 python bench.py --config C2 --steps 200 --warmup 10 --no-cpu-baseline > $OUT/${TAG}_bench_C2_nospec.json 2> $OUT/${TAG}_bench_C2_nospec.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv \
    --log-file $OUT/${TAG}_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline \
    > $OUT/${TAG}_launches_bench.log 2>&1
python - <<PY
import json, glob
for f in sorted(glob.glob("${OUT}/${TAG}_bench*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, 'ms/step', round(d['ms_per_step'], 4), 'e2e', round(d['e2e']['ms_per_step'], 4), 'loss', d['loss'])
        print(' | '.join(f"{k} {v['ms_per_launch']*1000:.1f}x{v['launches_per_step']:.0f}" for k, v in d['stages'].items()))
    except Exception as e:
        print(f, 'unreadable', e)
PY
