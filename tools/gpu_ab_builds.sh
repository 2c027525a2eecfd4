#!/bin/bash
# Runs ON the GPU box: bench.py against several builds of the library (JSB200_LIB).
# usage: tools/gpu_ab_builds.sh TAG lib1.so lib2.so ...   ("default" = the in-tree build)
TAG=$1; shift
OUT=gpurun_out
mkdir -p $OUT
for lib in "$@"; do
  name=$(basename $lib .so)
  if [ "$lib" = "default" ]; then unset JSB200_LIB; else export JSB200_LIB=$PWD/$lib; fi
  timeout 300 python bench.py --steps 100 --warmup 10 --no-cpu-baseline > $OUT/${TAG}_${name}.json 2> $OUT/${TAG}_${name}.err
  python - <<EOF
import json
try:
    d = json.loads(open("$OUT/${TAG}_${name}.json").read().strip().splitlines()[-1])
    print("$name", 'ms/step', round(d['ms_per_step'], 4), 'loss', d['loss'], ' | '.join(f"{k} {v['ms_per_launch']*1000:.1f}" for k, v in d['stages'].items()))
except Exception as e:
    print("$name", 'unreadable', e)
EOF
done
