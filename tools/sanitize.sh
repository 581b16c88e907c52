#!/bin/bash
# compute-sanitizer target: memcheck + racecheck on the GEMM unit cases (tf32 / x3 / h3) and on both
# point-set ops (incl. the bulk-copy / mbarrier EMD kernels), memcheck on the network's SIMT kernels.  Run on a GPU box:  bash tools/sanitize.sh [outdir]   -> <outdir>/sanitize_*.log, one summary line each.
# (--report-api-errors no: the lazy module load behind the first cudaLaunchKernel of a library makes the runtime probe
#  cuKernelGetFunction, which memcheck would otherwise list as an "invalid resource handle" API error.)
OUT=${1:-gpurun_out/sanitize}
mkdir -p $OUT
run() {   # name tool pytest-args...
  local name=$1 tool=$2; shift 2
  timeout 1200 compute-sanitizer --tool $tool --report-api-errors no --error-exitcode 9 \
      python -m pytest "$@" -q -x -p no:cacheprovider > $OUT/sanitize_${name}_${tool}.log 2>&1
  echo "$name $tool rc=$? : $(grep -h 'ERROR SUMMARY\|RACECHECK SUMMARY' $OUT/sanitize_${name}_${tool}.log | tail -1) : $(tail -1 $OUT/sanitize_${name}_${tool}.log)"
}
run tfops memcheck tests/test_tfops_gpu.py -k "emd_vs_oracle or nn_grad or nn_bit_exact or kats"
run gemm memcheck tests/test_h3_gpu.py tests/test_x3_gpu.py tests/test_tc_gemm_gpu.py -k "gemm_matches_fp64 or (test_fwd and not splitk) or split16 or weight_split"
run tfops racecheck tests/test_tfops_gpu.py -k "emd_vs_oracle and not 1024 and not 2048 and not 200-200 or nn_grad and not 2304 or kats"
run gemm racecheck tests/test_h3_gpu.py tests/test_x3_gpu.py -k "gemm_matches_fp64 and (1-8-16 or 2-12-12-1 or 3-12-12)"
# single-launch batch norm (grid barrier), crop/pool, stem, small-FC and xyz-head kernels
run netk memcheck tests/test_net_kernels_gpu.py
