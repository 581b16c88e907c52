#!/bin/bash
# round 2, call R: approxmatch skips the exact-zero terms of the fine levels
OUT=gpurun_out/r2_r
mkdir -p $OUT
timeout 600 python -m pytest tests/test_tfops_gpu.py tests/test_pointset_loss_gpu.py -q -x -p no:cacheprovider > $OUT/tests.log 2>&1; tail -2 $OUT/tests.log
timeout 200 python tools/am_quick.py 2>&1 | tail -4
