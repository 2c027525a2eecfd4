#!/bin/bash
OUT=gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"rc_row_scatter|rc_col_scatter|rc_row_count|rc_col_count|rc_col_scan|rc_row_scan" -s 12 -c 6 -o $OUT/r02y_lists -f python tools/ncu_step.py C3 4 > $OUT/r02y_ncu.log 2>&1
ls -la $OUT/r02y*
