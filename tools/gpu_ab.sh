#!/bin/bash
# Runs ON the GPU box: GPU tests, then bench.py with the default options and with each VAR=VALUE given
# usage: tools/gpu_ab.sh TAG [ENV=VAL ...] ; extra configs: CONFIGS="C3s C2" tools/gpu_ab.sh ...
TAG=$1; shift
OUT=gpurun_out
mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -q -x > $OUT/${TAG}_pytest.log 2>&1
tail -6 $OUT/${TAG}_pytest.log
timeout 300 python bench.py --steps 100 --warmup 10 --no-cpu-baseline > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err
for kv in "$@"; do
  name=$(echo $kv | tr '=' '_')
  env $kv timeout 300 python bench.py --steps 100 --warmup 10 --no-cpu-baseline > $OUT/${TAG}_bench_${name}.json 2> $OUT/${TAG}_bench_${name}.err
done
for C in $CONFIGS; do
  timeout 300 python bench.py --config $C --steps 50 --warmup 5 --no-cpu-baseline > $OUT/${TAG}_bench_$C.json 2> $OUT/${TAG}_bench_$C.err
done
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
