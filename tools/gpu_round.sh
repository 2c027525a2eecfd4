#!/bin/bash
# Runs ON the GPU box (through gpurun): GPU tests, both bench arms, the ncu launch list of the
# bench command and one `ncu --set full` capture of our kernels over one fwd+bwd step.
# usage: tools/gpu_round.sh TAG [full]
TAG=${1:-r01x}
FULL=${2:-}
OUT=gpurun_out
mkdir -p $OUT
KREGEX='regex:^(rc_|project_|bin_prep|fingerprint|radix_|onesweep|upsweep|scan_|box_gather|emit_|chunk_|group_|hist_apply|raster_|inclusive_scan|add_block|fill_background)'
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm,power.limit --format=csv > $OUT/${TAG}_gpu.txt 2>&1
nproc >> $OUT/${TAG}_gpu.txt
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/${TAG}_pytest.log 2>&1
tail -3 $OUT/${TAG}_pytest.log
timeout 600 python bench.py > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err
timeout 300 python bench.py --impl reference --steps 30 --warmup 5 > $OUT/${TAG}_bench_reference.json 2> $OUT/${TAG}_bench_reference.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv \
    --log-file $OUT/${TAG}_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline \
    > $OUT/${TAG}_launches_bench.log 2>&1
if [ -n "$FULL" ]; then
  # step 1 warms (arena growth), step 2 is captured: skip the first step's launches of our kernels
  timeout 900 ncu --set full --clock-control none --import-source on -k "$KREGEX" \
      --launch-skip 22 -c 22 -f -o $OUT/${TAG}_full python tools/ncu_step.py C3 3 \
      > $OUT/${TAG}_full.log 2>&1
  ncu -i $OUT/${TAG}_full.ncu-rep --page raw --csv > $OUT/${TAG}_full_raw.csv 2>/dev/null
fi
python - <<EOF
import json
for f in ("${OUT}/${TAG}_bench.json", "${OUT}/${TAG}_bench_reference.json"):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, 'ms/step', round(d['ms_per_step'], 4), 'e2e', round(d['e2e']['ms_per_step'], 4))
        if 'stages' in d:
            print(' | '.join(f"{k} {v['ms_per_launch']*1000:.1f}x{v['launches_per_step']:.0f}" for k, v in d['stages'].items()))
    except Exception as e:
        print(f, 'unreadable', e)
EOF
