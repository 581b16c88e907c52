#!/bin/bash
# round 2, call B: whole GPU suite in the new default precision, streaming EMD kernels (row-tile sweep), memcheck of the EMD ops
OUT=gpurun_out/r2_b
mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -q -x -p no:cacheprovider > $OUT/pytest_gpu.log 2>&1
tail -15 $OUT/pytest_gpu.log
for R in 8 16 32 64; do
  echo "== MPB_MS_ROWS=$R"
  MPB_MS_ROWS=$R timeout 300 python tools/bench_tfops.py > $OUT/tfops_rows$R.json 2>&1
  grep -A3 '"am_cost_b32_n1024"\|"am_grad_b32_n1024"\|"am_cost_b32_n2304"\|"am_grad_b32_n2304"' $OUT/tfops_rows$R.json | grep "med_us\|am_"
done
timeout 300 python tools/bench_tfops.py > $OUT/tfops_auto.json 2>&1
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_tfops_gpu.py -k "emd_vs_oracle or nn_grad or nn_bit_exact" -q -x -p no:cacheprovider > $OUT/memcheck_tfops.log 2>&1
echo "memcheck rc=$?"; tail -5 $OUT/memcheck_tfops.log
