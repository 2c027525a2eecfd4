#!/bin/bash
# Runs ON the GPU box: GPU suite, then the in-tree build against the previous one (build/alt/r04.so)
OUT=gpurun_out
mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -q -x > $OUT/r04b_pytest.log 2>&1
tail -4 $OUT/r04b_pytest.log
bash tools/gpu_ab_builds.sh r04b build/alt/r04.so default
bash tools/gpu_ab_builds.sh r04b2 build/alt/r04.so default
