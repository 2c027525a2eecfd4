#!/bin/bash
# Runs ON the GPU box: GPU suite, then bench.py with the accumulators cleared on the side stream
# (default) against the backward's own memset (JSB_PRECLEAR=0), interleaved twice.
OUT=gpurun_out
mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -q -x > $OUT/r04c_pytest.log 2>&1
tail -4 $OUT/r04c_pytest.log
for rep in a b; do
  for v in 0 1; do
    JSB_PRECLEAR=$v timeout 300 python bench.py --steps 100 --warmup 10 --no-cpu-baseline > $OUT/r04c_bench_C3_preclear${v}${rep}.json 2> $OUT/r04c_bench_C3_preclear${v}${rep}.err
  done
done
python - <<PY
import json, glob
for f in sorted(glob.glob("${OUT}/r04c_bench*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, 'ms/step', round(d['ms_per_step'], 4), 'e2e', round(d['e2e']['ms_per_step'], 4), 'loss', d['loss'])
    except Exception as e:
        print(f, 'unreadable', e)
PY
