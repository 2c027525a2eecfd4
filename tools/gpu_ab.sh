#!/bin/bash
# Runs ON the GPU box: GPU tests, then the bench in the product configuration and with single
# features switched off (A/B), then the ncu launch list of the product configuration.
TAG=${1:-r01k}
OUT=gpurun_out
mkdir -p $OUT
timeout 1200 python -m pytest tests -m gpu -x -q > $OUT/${TAG}_pytest.log 2>&1
tail -15 $OUT/${TAG}_pytest.log
timeout 300 python bench.py --steps 100 --warmup 10 --no-cpu-baseline > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err
JSB_ROWCOL=0 timeout 300 python bench.py --steps 100 --warmup 10 --no-cpu-baseline > $OUT/${TAG}_bench_norowcol.json 2> $OUT/${TAG}_bench_norowcol.err
JSB_FUSED=0 timeout 300 python bench.py --steps 100 --warmup 10 --no-cpu-baseline > $OUT/${TAG}_bench_nofused.json 2> $OUT/${TAG}_bench_nofused.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv \
    --log-file $OUT/${TAG}_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline \
    > $OUT/${TAG}_launches_bench.log 2>&1
python - <<EOF
import json
for sfx in ("", "_norowcol", "_nofused"):
    f = "${OUT}/${TAG}_bench%s.json" % sfx
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, 'ms/step', round(d['ms_per_step'], 4), 'e2e', round(d['e2e']['ms_per_step'], 4))
        if 'stages' in d:
            print(' | '.join(f"{k} {v['ms_per_launch']*1000:.1f}x{v['launches_per_step']:.0f}" for k, v in d['stages'].items()))
    except Exception as e:
        print(f, 'unreadable', e)
        import subprocess
        print(subprocess.run(['tail', '-5', f.replace('.json', '.err')], capture_output=True, text=True).stdout)
EOF
