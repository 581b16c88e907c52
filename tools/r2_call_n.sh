#!/bin/bash
# round 2, call N: match_cost_grad fed by bulk copies (TMA 1-D) instead of the register ring
OUT=gpurun_out/r2_n
mkdir -p $OUT
timeout 600 python -m pytest tests/test_tfops_gpu.py -q -x -p no:cacheprovider -k "emd" > $OUT/tests.log 2>&1; tail -3 $OUT/tests.log
MPB_MG_CG=1 timeout 600 python -m pytest tests/test_tfops_gpu.py -q -x -p no:cacheprovider -k "emd" > $OUT/tests_cg1.log 2>&1; tail -1 $OUT/tests_cg1.log
TAG=cg2 timeout 120 python tools/am_grad_quick.py 2>&1 | tail -2
TAG=cg1 MPB_MG_CG=1 timeout 120 python tools/am_grad_quick.py 2>&1 | tail -2
TAG=ring MPB_MG_CG=0 timeout 120 python tools/am_grad_quick.py 2>&1 | tail -2
for R in 48 56 72 80 96 160 240 400; do TAG=cg2_rows$R MPB_MS_ROWS=$R timeout 120 python tools/am_grad_quick.py 2>&1 | tail -2; done
