#!/bin/bash
# Runs ON the GPU box: one `ncu --set full` capture of every kernel of the library over one
# steady-state fwd+bwd step (the ncu part of tools/gpu_round.sh on its own).  usage: TAG
TAG=${1:-r01x}
OUT=gpurun_out
mkdir -p $OUT
KREGEX='regex:^(rc_|project_|bin_prep|fingerprint|radix_|onesweep|upsweep|scan_|box_gather|emit_|chunk_|group_|hist_apply|raster_|inclusive_scan|add_block|fill_background)'
timeout 900 ncu --set full --clock-control none --import-source on -k "$KREGEX" \
    --launch-skip 22 -c 22 -f -o $OUT/${TAG}_full python tools/ncu_step.py C3 3 > $OUT/${TAG}_full.log 2>&1
ncu -i $OUT/${TAG}_full.ncu-rep --page raw --csv > $OUT/${TAG}_full_raw.csv 2>/dev/null
tail -3 $OUT/${TAG}_full.log
