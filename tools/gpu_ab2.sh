#!/bin/bash
# Runs ON the GPU box: GPU tests, bench in the product configuration, bench with named env
# switches (A/B), ncu launch list.   usage: tools/gpu_ab2.sh TAG "ENV1=0" "ENV2=0" ...
TAG=$1; shift
OUT=gpurun_out
mkdir -p $OUT
timeout 1200 python -m pytest tests -m gpu -q > $OUT/${TAG}_pytest.log 2>&1
tail -12 $OUT/${TAG}_pytest.log
timeout 300 python bench.py --steps 100 --warmup 10 --no-cpu-baseline > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err
for sw in "$@"; do
  env $sw timeout 300 python bench.py --steps 100 --warmup 10 --no-cpu-baseline > $OUT/${TAG}_bench_${sw%%=*}.json 2> $OUT/${TAG}_bench_${sw%%=*}.err
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv \
    --log-file $OUT/${TAG}_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline \
    > $OUT/${TAG}_launches_bench.log 2>&1
python - <<EOF
import json, glob
for f in sorted(glob.glob("${OUT}/${TAG}_bench*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, 'ms/step', round(d['ms_per_step'], 4), 'e2e', round(d['e2e']['ms_per_step'], 4))
        print(' | '.join(f"{k} {v['ms_per_launch']*1000:.1f}x{v['launches_per_step']:.0f}" for k, v in d['stages'].items()))
    except Exception as e:
        print(f, 'unreadable', e)
EOF
