#!/bin/bash
# Runs ON the GPU box: one `ncu --set full` capture of the kernels matching REGEX (steady state).
# usage: tools/gpu_ncu_kernels.sh TAG REGEX [COUNT] [SKIP]
TAG=$1; REGEX=$2; COUNT=${3:-4}; SKIP=${4:-4}
OUT=gpurun_out
mkdir -p $OUT
timeout 900 ncu --set full --clock-control none --import-source on -k "regex:$REGEX" \
    --launch-skip $SKIP -c $COUNT -f -o $OUT/${TAG} python tools/ncu_step.py C3 4 > $OUT/${TAG}.log 2>&1
ncu -i $OUT/${TAG}.ncu-rep --page raw --csv > $OUT/${TAG}_raw.csv 2>/dev/null
ncu -i $OUT/${TAG}.ncu-rep --page source --csv --print-source sass,cuda > $OUT/${TAG}_src.csv 2>/dev/null
tail -3 $OUT/${TAG}.log
