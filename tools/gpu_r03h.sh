#!/bin/bash
OUT=gpurun_out
bash tools/gpu_ab_builds.sh r03h default build/alt/bwd24.so build/alt/bwd28.so build/alt/fwd9.so build/alt/fwd7.so
JSB_SUPERCOLS=1 bash tools/gpu_ab_builds.sh r03h_sc1 default
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 1200 compute-sanitizer --tool racecheck --error-exitcode 7 --print-limit 10 \
    python -m pytest tests -m gpu -q -x -k "test_super_column_lists_render_the_same_image_and_gradients and (ragged or tiny or band) or test_block_size_8_tuned_kernels_equal_the_generic_ones and tiny or test_lean_residuals" > $OUT/r03h_racecheck.log 2>&1
echo "racecheck rc=$?"; grep -E "RACECHECK SUMMARY|passed|failed|hazard" $OUT/r03h_racecheck.log | tail -8
