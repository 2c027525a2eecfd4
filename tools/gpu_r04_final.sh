#!/bin/bash
# Runs ON the GPU box: GPU suite, bench.py as the driver runs it, the plain-launch A/B (JSB_PDL=0),
# every other configuration, and the ncu launch list of the default command.
TAG=${1:-r04}
OUT=gpurun_out
mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -q -x > $OUT/${TAG}_pytest.log 2>&1
tail -4 $OUT/${TAG}_pytest.log
timeout 600 python bench.py > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err
JSB_PDL=0 timeout 300 python bench.py --steps 100 --warmup 10 --no-cpu-baseline > $OUT/${TAG}_bench_C3_pdl0.json 2> $OUT/${TAG}_bench_C3_pdl0.err
timeout 300 python bench.py --steps 100 --warmup 10 --no-cpu-baseline > $OUT/${TAG}_bench_C3_pdl1.json 2> $OUT/${TAG}_bench_C3_pdl1.err
for C in C1 C2 C3s C3b; do
  timeout 300 python bench.py --config $C --steps 50 --warmup 5 --no-cpu-baseline > $OUT/${TAG}_bench_$C.json 2> $OUT/${TAG}_bench_$C.err
done
timeout 400 python bench.py --config C4 --steps 3 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_bench_C4_1gpu.json 2> $OUT/${TAG}_bench_C4_1gpu.err
timeout 400 python bench.py --config C5 --steps 5 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_bench_C5_1gpu.json 2> $OUT/${TAG}_bench_C5_1gpu.err
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
    --log-file $OUT/${TAG}_launches_C3.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline \
    > $OUT/${TAG}_launches_bench.log 2>&1
python - <<PY
import json, glob
for f in sorted(glob.glob("${OUT}/${TAG}_bench*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        e = d.get('e2e') or {}
        print(f, 'ms/step', round(d['ms_per_step'], 4), d.get('unit'), round(d.get('value', 0), 2), 'e2e', e.get('ms_per_step'))
        if 'stages' in d:
            print(' | '.join(f"{k} {v['ms_per_launch']*1000:.1f}x{v['launches_per_step']:.0f}" for k, v in d['stages'].items()))
    except Exception as e:
        print(f, 'unreadable', e)
PY
