#!/bin/bash
# round 2, final evidence 2/2 (same build as 1/2): ncu launch list of the captured step, ncu --set full of the six GEMM
# launches and of the point-set ops, compute-sanitizer
OUT=gpurun_out/r2_q
mkdir -p $OUT
echo "== ncu launch list of the captured step"
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active --clock-control none --csv --log-file $OUT/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-ops > $OUT/ncu_bench.log 2>&1
python tools/ncu_launch_shares.py $OUT/launches.csv $OUT/r2 | head -40
echo "== ncu --set full, six GEMM launches (h3 forward)"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:tc_gemm -o $OUT/r2_ncu_gemm -f python tools/ncu_gemm.py --h3 > $OUT/ncu_gemm.log 2>&1
tail -3 $OUT/ncu_gemm.log
echo "== ncu point-set ops"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'nn_distance_kernel|nn_distance_grad|approxmatch_cluster_kernel|matchcost_tma|matchcostgrad_tma|ms_sum_planes|matchcost_final' -o $OUT/r2_ncu_tfops -f python tools/ncu_tfops.py > $OUT/ncu_tfops.log 2>&1
tail -3 $OUT/ncu_tfops.log
ls -la $OUT/*.ncu-rep
echo "== sanitizers"
bash tools/sanitize.sh $OUT
