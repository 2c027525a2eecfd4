#!/bin/bash
# Runs ON the GPU box through `gpurun --gpus G`: the three multi-GPU lines at G ranks.
# usage: tools/gpu_round2_multi.sh G [TAG]
#   bench.py (C3, one view per rank and step: weak scaling, gradient sink + in-switch all-reduce),
#   C4 64-view batch (strong scaling, one all-reduce per batch), C5 frame by tile-row bands.
G=${1:-2}
TAG=${2:-r02m}
OUT=gpurun_out
mkdir -p $OUT
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1"
timeout 600 $RUN --master-port 29501 bench.py --gpus $G --steps 200 --warmup 20 > $OUT/${TAG}_bench_C3_${G}gpu.json 2> $OUT/${TAG}_bench_C3_${G}gpu.err
# experimental: appearance gradients reduced on a side stream under project_bwd (bench.py docstring)
JSB_BENCH_OVERLAP=1 timeout 600 $RUN --master-port 29504 bench.py --gpus $G --steps 200 --warmup 20 > $OUT/${TAG}_bench_C3_overlap_${G}gpu.json 2> $OUT/${TAG}_bench_C3_overlap_${G}gpu.err
timeout 600 $RUN --master-port 29502 tools/multigpu_bench.py views --steps 3 --warmup 2 > $OUT/${TAG}_views_C4_${G}gpu.json 2> $OUT/${TAG}_views_C4_${G}gpu.err
timeout 600 $RUN --master-port 29503 tools/multigpu_bench.py bands --steps 5 --warmup 2 > $OUT/${TAG}_bands_C5_${G}gpu.json 2> $OUT/${TAG}_bands_C5_${G}gpu.err
for f in $OUT/${TAG}_*_${G}gpu.json; do echo $f; head -c 600 $f; echo; done
