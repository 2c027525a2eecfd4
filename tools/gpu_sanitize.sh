#!/bin/bash
# Runs ON the GPU box: compute-sanitizer memcheck + racecheck over the small-scene GPU tests.
OUT=gpurun_out
mkdir -p $OUT
SEL="test_project_fwd or test_rasterize_fwd_and_binning or test_rasterize_bwd or test_project_bwd or test_row_column_lists or test_fused_render_equals or test_empty_frame or test_ragged_counts or test_block_sizes or test_two_pixels or test_fused_backward_reuses or test_native_adam or test_render_views"
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 7 --print-limit 20 \
    python -m pytest tests -m gpu -q -x -k "$SEL" > $OUT/sanitize_memcheck.log 2>&1
echo "memcheck rc=$?"; grep -E "ERROR SUMMARY|passed|failed|Invalid|out of bounds" $OUT/sanitize_memcheck.log | tail -8
timeout 1500 compute-sanitizer --tool racecheck --error-exitcode 7 --print-limit 20 \
    python -m pytest tests -m gpu -q -x -k "test_row_column_lists or test_fused_render_equals or test_two_pixels or test_render_views_equals" > $OUT/sanitize_racecheck.log 2>&1
echo "racecheck rc=$?"; grep -E "RACECHECK SUMMARY|passed|failed|hazard" $OUT/sanitize_racecheck.log | tail -12
