#!/bin/bash
# Runs ON the GPU box: the whole GPU test-suite, smoke(), and one bench run.
TAG=${1:-r01f}
OUT=gpurun_out
mkdir -p $OUT
( time timeout 1200 python -m pytest tests -m gpu -q --durations=8 ) > $OUT/${TAG}_pytest.log 2>&1
tail -16 $OUT/${TAG}_pytest.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 300 python bench.py --steps 100 --warmup 10 --no-cpu-baseline > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err
python - <<PY
import json
d = json.loads(open("$OUT/${TAG}_bench.json").read().strip().splitlines()[-1])
print('ms/step', round(d['ms_per_step'], 4), 'e2e', round(d['e2e']['ms_per_step'], 4), 'launches', d['gpu_launches'])
print(' | '.join(f"{k} {v['ms_per_launch']*1000:.1f}" for k, v in d['stages'].items()))
PY
