#!/bin/bash
OUT=gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x -k "mse or fit" > $OUT/r02v_pytest.log 2>&1; tail -3 $OUT/r02v_pytest.log
timeout 300 python bench.py --steps 100 --warmup 10 --no-cpu-baseline > $OUT/r02v_bench.json 2> $OUT/r02v_bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/r02v_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > $OUT/r02v_launches_bench.log 2>&1
python - <<PY
import json
d = json.loads(open("gpurun_out/r02v_bench.json").read().strip().splitlines()[-1])
print('ms/step', round(d['ms_per_step'], 4), 'e2e', round(d['e2e']['ms_per_step'], 4), 'loss', d['loss'])
print(' | '.join(f"{k} {v['ms_per_launch']*1000:.1f}" for k, v in d['stages'].items()))
PY
