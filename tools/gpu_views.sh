#!/bin/bash
# Runs ON the GPU box: the batched-camera / drop-in-module tests, a regression subset, the C4
# views benchmark and one short bench run.
TAG=${1:-r01v}
OUT=gpurun_out
mkdir -p $OUT
timeout 400 python -m pytest tests/test_gpu_views_and_dropin.py tests/test_gpu_golden.py -q -m gpu > $OUT/${TAG}_pytest_new.log 2>&1
tail -15 $OUT/${TAG}_pytest_new.log
timeout 300 python -m pytest tests/test_gpu_api_and_scale.py -q -m gpu -k "render or fused or band or token" > $OUT/${TAG}_pytest_api.log 2>&1
tail -5 $OUT/${TAG}_pytest_api.log
timeout 200 python tools/views_bench.py 8 C4 10 > $OUT/${TAG}_views_C4.json 2> $OUT/${TAG}_views_C4.err
cat $OUT/${TAG}_views_C4.json; tail -3 $OUT/${TAG}_views_C4.err
timeout 200 python bench.py --steps 100 --warmup 10 --no-cpu-baseline > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err
python - <<PY
import json
d = json.loads(open("$OUT/${TAG}_bench.json").read().strip().splitlines()[-1])
print('ms/step', round(d['ms_per_step'], 4), 'e2e', round(d['e2e']['ms_per_step'], 4))
print(' | '.join(f"{k} {v['ms_per_launch']*1000:.1f}" for k, v in d['stages'].items()))
PY
