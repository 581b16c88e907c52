#!/bin/bash
# round 2, call P: match_cost on the bulk-copy feed
OUT=gpurun_out/r2_p
mkdir -p $OUT
timeout 600 python -m pytest tests/test_tfops_gpu.py tests/test_pointset_loss_gpu.py -q -x -p no:cacheprovider > $OUT/tests.log 2>&1; tail -3 $OUT/tests.log
MPB_MC_CG=1 timeout 600 python -m pytest tests/test_tfops_gpu.py -q -x -p no:cacheprovider -k emd > $OUT/tests_cg1.log 2>&1; tail -1 $OUT/tests_cg1.log
TAG=cost_cg2 timeout 120 python tools/am_grad_quick.py 2>&1 | tail -2
TAG=cost_cg1 MPB_MC_CG=1 timeout 120 python tools/am_grad_quick.py 2>&1 | tail -2
TAG=cost_ring MPB_MC_CG=0 timeout 120 python tools/am_grad_quick.py 2>&1 | tail -2
for R in 40 64; do TAG=rows$R MPB_MS_ROWS=$R timeout 120 python tools/am_grad_quick.py 2>&1 | tail -2; done
