#!/bin/bash
# Runs ON the GPU box: parity tests of the rasterisers, then the bench with the packed FP32x2
# kernels (default) and with the scalar ones (A/B).
TAG=${1:-r01x}
OUT=gpurun_out
mkdir -p $OUT
timeout 500 python -m pytest tests/test_gpu_golden.py tests/test_gpu_oracle_and_edges.py tests/test_gpu_views_and_dropin.py -q -m gpu -x > $OUT/${TAG}_pytest.log 2>&1
tail -6 $OUT/${TAG}_pytest.log
timeout 200 python bench.py --steps 100 --warmup 10 --no-cpu-baseline > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err
JSB_BWDPK=0 JSB_FWDPK=0 timeout 200 python bench.py --steps 100 --warmup 10 --no-cpu-baseline > $OUT/${TAG}_bench_scalar.json 2> $OUT/${TAG}_bench_scalar.err
python - <<PY
import json
for sfx in ("", "_scalar"):
    f = "$OUT/${TAG}_bench%s.json" % sfx
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(sfx or "packed", 'ms/step', round(d['ms_per_step'], 4), 'e2e', round(d['e2e']['ms_per_step'], 4), 'loss', d.get('loss'))
        print(' | '.join(f"{k} {v['ms_per_launch']*1000:.1f}" for k, v in d['stages'].items()))
    except Exception as e:
        print(f, 'unreadable', e)
PY
