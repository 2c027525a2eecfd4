#!/bin/bash
# Runs ON the GPU box: the shortest possible A/B of builds (bench only, 60 steps each).
# usage: tools/gpu_ab_min.sh TAG lib1.so ...   ("default" = the in-tree build)
TAG=$1; shift
mkdir -p gpurun_out
for lib in "$@"; do
  name=$(basename $lib .so)
  if [ "$lib" = "default" ]; then unset JSB200_LIB; else export JSB200_LIB=$PWD/$lib; fi
  timeout 60 python bench.py --steps 60 --warmup 8 --no-cpu-baseline > gpurun_out/${TAG}_${name}.json 2> gpurun_out/${TAG}_${name}.err
  python -c "
import json
d = json.loads(open('gpurun_out/${TAG}_${name}.json').read().strip().splitlines()[-1])
print('$name', round(d['ms_per_step'], 4), d['loss'], round(d['stages']['raster_fwd']['ms_per_launch']*1e3, 1), round(d['stages']['raster_bwd']['ms_per_launch']*1e3, 1))
"
done
