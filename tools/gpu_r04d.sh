#!/bin/bash
# Runs ON the GPU box: GPU suite, then the in-tree build against the previous one (build/alt/r04c.so)
OUT=gpurun_out
mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -q -x > $OUT/r04d_pytest.log 2>&1
tail -4 $OUT/r04d_pytest.log
bash tools/gpu_ab_builds.sh r04d build/alt/r04c.so default
bash tools/gpu_ab_builds.sh r04d2 build/alt/r04c.so default
