#!/bin/bash
# round 2, call O: ncu of the bulk-copy-fed match_cost_grad kernel
OUT=gpurun_out/r2_o
mkdir -p $OUT
timeout 300 ncu --set full --import-source on --clock-control none -k regex:matchcostgrad_tma -s 12 -c 1 -o $OUT/mg_tma -f python tools/am_grad_quick.py > $OUT/ncu.log 2>&1
tail -3 $OUT/ncu.log
ls -la $OUT
