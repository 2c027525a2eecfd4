#!/bin/bash
TAG=${1:-r02d}
OUT=gpurun_out
mkdir -p $OUT
timeout 300 python tools/hits_debug.py > $OUT/${TAG}_hits_debug.log 2>&1
tail -20 $OUT/${TAG}_hits_debug.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"resident_depth_sort|raster_bwd_hits|raster_fwd16x2" -s 9 -c 3 -o $OUT/${TAG}_prof -f python tools/ncu_step.py C3 5 > $OUT/${TAG}_ncu.log 2>&1
tail -3 $OUT/${TAG}_ncu.log
ls -la $OUT/${TAG}_prof.ncu-rep
