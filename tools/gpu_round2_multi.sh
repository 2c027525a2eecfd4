#!/bin/bash
# Runs ON the GPU box through `gpurun --gpus G`: the multi-GPU lines at G ranks.
# usage: tools/gpu_round2_multi.sh G [TAG]
#   bench.py (C3, one view per rank and step: weak scaling, gradient sink + in-switch all-reduce),
#   the same with the appearance gradients reduced under project_bwd, the in-switch all-reduce
#   check, C4 64-view batch (strong scaling, one all-reduce per batch), C5 frame by tile-row bands.
G=${1:-2}
TAG=${2:-r03m}
OUT=gpurun_out
mkdir -p $OUT
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1"
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv,noheader | head -2 > $OUT/${TAG}_${G}gpu_info.txt
timeout 600 $RUN --master-port 29501 bench.py --gpus $G --steps 100 --warmup 10 > $OUT/${TAG}_bench_C3_${G}gpu.json 2> $OUT/${TAG}_bench_C3_${G}gpu.err

if [ "$G" = "8" ]; then
timeout 300 $RUN --master-port 29505 tools/allreduce_check.py 1000000 > $OUT/${TAG}_allreduce_${G}gpu.json 2> $OUT/${TAG}_allreduce_${G}gpu.err
fi
timeout 600 $RUN --master-port 29502 bench.py --gpus $G --config C4 --steps 3 --warmup 3 > $OUT/${TAG}_bench_C4_${G}gpu.json 2> $OUT/${TAG}_bench_C4_${G}gpu.err
timeout 600 $RUN --master-port 29503 bench.py --gpus $G --config C5 --steps 5 --warmup 3 > $OUT/${TAG}_bench_C5_${G}gpu.json 2> $OUT/${TAG}_bench_C5_${G}gpu.err
for f in $OUT/${TAG}_*_${G}gpu.json; do echo $f; head -c 700 $f; echo; done
