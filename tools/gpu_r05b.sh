#!/bin/bash
# Runs ON the GPU box: GPU suite, then the in-tree build against the previous one (build/alt/r05.so)
OUT=gpurun_out
mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -q -x > $OUT/r05b_pytest.log 2>&1
tail -4 $OUT/r05b_pytest.log
bash tools/gpu_ab_builds.sh r05b build/alt/r05.so default
bash tools/gpu_ab_builds.sh r05b2 build/alt/r05.so default
