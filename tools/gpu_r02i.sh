#!/bin/bash
TAG=${1:-r02i}
OUT=gpurun_out
mkdir -p $OUT
timeout 1200 python -m pytest tests -m gpu -q -x > $OUT/${TAG}_pytest.log 2>&1
tail -6 $OUT/${TAG}_pytest.log
timeout 300 python bench.py --steps 100 --warmup 10 --no-cpu-baseline > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err
JSB_BWDMOM=0 timeout 300 python bench.py --steps 100 --warmup 10 --no-cpu-baseline > $OUT/${TAG}_bench_nomoments.json 2> $OUT/${TAG}_bench_nomoments.err
timeout 300 python bench.py --config C3s --steps 50 --warmup 5 --no-cpu-baseline > $OUT/${TAG}_bench_C3s.json 2> $OUT/${TAG}_bench_C3s.err
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"resident_depth_sort|raster_bwd_hits|rc_col_scatter|rc_row_scatter" -s 12 -c 4 -o $OUT/${TAG}_prof -f python tools/ncu_step.py C3 5 > $OUT/${TAG}_ncu.log 2>&1
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
