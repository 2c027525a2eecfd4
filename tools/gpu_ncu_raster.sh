#!/bin/bash
# Runs ON the GPU box: one `ncu --set full` capture of the two tile rasterisers (warm step) and the
# raw / source pages as CSV.
TAG=${1:-r01n}
OUT=gpurun_out
mkdir -p $OUT
timeout 600 ncu --set full --clock-control none --import-source on -k "regex:^raster_" \
    --launch-skip 2 -c 2 -f -o $OUT/${TAG}_raster python tools/ncu_step.py C3 3 > $OUT/${TAG}_raster.log 2>&1
ncu -i $OUT/${TAG}_raster.ncu-rep --page raw --csv > $OUT/${TAG}_raster_raw.csv 2>/dev/null
ncu -i $OUT/${TAG}_raster.ncu-rep --page source --csv --print-source sass > $OUT/${TAG}_raster_source.csv 2>/dev/null
tail -3 $OUT/${TAG}_raster.log
ls -la $OUT | tail -5
