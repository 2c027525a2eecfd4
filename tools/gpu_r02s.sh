#!/bin/bash
OUT=gpurun_out
for m in ring plain; do
  if [ $m = plain ]; then export JSB_OPTS=supercols=3,fwdsuper=0; else export JSB_OPTS=supercols=3; fi
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:"raster_fwd" -s 2 -c 1 -o $OUT/r02s_fwd_$m -f python tools/ncu_step.py C3 4 > $OUT/r02s_ncu_$m.log 2>&1
done
export JSB_OPTS=supercols=0
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"raster_fwd" -s 2 -c 1 -o $OUT/r02s_fwd_exact -f python tools/ncu_step.py C3 4 > $OUT/r02s_ncu_exact.log 2>&1
ls -la $OUT/r02s*
