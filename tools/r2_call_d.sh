#!/bin/bash
# round 2, call D: h3 with unrounded operands, point-set training losses, whole suite
OUT=gpurun_out/r2_d
mkdir -p $OUT
echo "== forward parity h3 (+ gradients)"
MPB_PRECISION=h3 timeout 300 python tools/check_network.py --bwd > $OUT/check_network_h3.txt 2>&1
grep -h "feat\|map_features\|inst_xyz_map_local\|centroids \|alpha_bins\|cen_z_offs\|total\|median" $OUT/check_network_h3.txt
echo "== new tests"
timeout 1200 python -m pytest tests/test_h3_gpu.py tests/test_pointset_loss_gpu.py -q -p no:cacheprovider > $OUT/new_tests.log 2>&1
tail -60 $OUT/new_tests.log
echo "== whole suite"
timeout 1500 python -m pytest tests -m gpu -q -rs -p no:cacheprovider > $OUT/pytest_gpu.log 2>&1
tail -40 $OUT/pytest_gpu.log
echo "== bench h3"
MPB_PRECISION=h3 timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-ops > $OUT/bench_h3.json 2> $OUT/bench_h3.err
tail -c 1500 $OUT/bench_h3.json; tail -5 $OUT/bench_h3.err
