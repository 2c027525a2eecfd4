#!/bin/bash
OUT=gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x > $OUT/r02r_pytest.log 2>&1
tail -5 $OUT/r02r_pytest.log
bash tools/gpu_ab_builds.sh r02r default build/alt/fsk2.so
for s in 2 4 5; do
  JSB_SUPERCOLS=$s timeout 300 python bench.py --steps 100 --warmup 10 --no-cpu-baseline > $OUT/r02r_sc$s.json 2> $OUT/r02r_sc$s.err
  python - <<PY
import json
try:
    d = json.loads(open("$OUT/r02r_sc$s.json").read().strip().splitlines()[-1])
    print("supercols $s", 'ms/step', round(d['ms_per_step'], 4), 'loss', d['loss'], ' | '.join(f"{k} {v['ms_per_launch']*1000:.1f}" for k, v in d['stages'].items()))
except Exception as e:
    print("supercols $s", 'unreadable', e)
PY
done
