#!/bin/bash
TAG=${1:-r02j}
OUT=gpurun_out
mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -q -x --durations=8 > $OUT/${TAG}_pytest.log 2>&1
tail -16 $OUT/${TAG}_pytest.log
timeout 300 python bench.py --steps 100 --warmup 10 --no-cpu-baseline > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err
for v in rc44 rc84 rc48; do
  JSB200_LIB=$PWD/build/alt/$v.so timeout 300 python bench.py --steps 100 --warmup 10 --no-cpu-baseline > $OUT/${TAG}_bench_$v.json 2> $OUT/${TAG}_bench_$v.err
done
timeout 300 python tools/views_bench.py 8 C4 5 > $OUT/${TAG}_views_C4.json 2> $OUT/${TAG}_views_C4.err
python - <<PY
import json, glob
for f in sorted(glob.glob("${OUT}/${TAG}_bench*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, 'ms/step', round(d['ms_per_step'], 4), 'e2e', round(d['e2e']['ms_per_step'], 4), 'loss', d['loss'])
        print(' | '.join(f"{k} {v['ms_per_launch']*1000:.1f}x{v['launches_per_step']:.0f}" for k, v in d['stages'].items()))
    except Exception as e:
        print(f, 'unreadable', e)
print(open("${OUT}/${TAG}_views_C4.json").read()[:1200])
PY
