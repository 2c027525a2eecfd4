#!/bin/bash
# Runs ON the GPU box: the GPU suite with programmatic dependent launch on, then bench.py A/B
# (JSB_PDL=0 / 1, interleaved twice) at C3 and once at the launch-bound configurations C1 / C2.
TAG=${1:-r04a}
OUT=gpurun_out
mkdir -p $OUT
JSB_PDL=1 timeout 1500 python -m pytest tests -m gpu -q -x > $OUT/${TAG}_pytest_pdl1.log 2>&1
tail -4 $OUT/${TAG}_pytest_pdl1.log
for rep in a b; do
  for v in 0 1; do
    JSB_PDL=$v timeout 300 python bench.py --steps 100 --warmup 10 --no-cpu-baseline > $OUT/${TAG}_bench_C3_pdl${v}${rep}.json 2> $OUT/${TAG}_bench_C3_pdl${v}${rep}.err
  done
done
for C in C1 C2; do
  for v in 0 1; do
    JSB_PDL=$v timeout 300 python bench.py --config $C --steps 100 --warmup 10 --no-cpu-baseline > $OUT/${TAG}_bench_${C}_pdl${v}.json 2> $OUT/${TAG}_bench_${C}_pdl${v}.err
  done
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
