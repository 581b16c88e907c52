#!/bin/bash
# round 2, call G: epilogue column prefetch + trimmed EMD gradient kernel; profiles refresh (launch shares, six GEMM launches, timeline)
OUT=gpurun_out/r2_g
mkdir -p $OUT
timeout 900 python -m pytest tests/test_tfops_gpu.py tests/test_h3_gpu.py tests/test_x3_gpu.py tests/test_tc_gemm_gpu.py tests/test_network_gpu.py tests/test_zz_late_additions_gpu.py tests/test_pointset_loss_gpu.py -q -x -p no:cacheprovider 2>&1 | tail -4
timeout 300 python tools/bench_tfops.py > $OUT/tfops_minb2.json 2>&1
echo "grad minb=2: $(grep -A1 '"am_cost_b32_n1024"\|"am_grad_b32_n1024"\|"am_grad_b32_n2304"' $OUT/tfops_minb2.json | grep med_us | tr '\n' ' ')"
for P in h3 tf32; do
MPB_PRECISION=$P timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-ops > $OUT/bench_$P.json 2> $OUT/bench_$P.err
python -c "
import json; d=json.load(open('$OUT/bench_$P.json')); print('$P %.3f ms/step e2e %.0f alone %.2f' % (d['ms_per_step'], d['e2e']['value'], d['roofline']['alone']['ms']))"
done
echo "== step timeline"
timeout 300 python tools/step_timeline.py > $OUT/step_timeline.txt 2>&1; tail -25 $OUT/step_timeline.txt
cp gpurun_out/step_kernels.csv $OUT/ 2>/dev/null
echo "== ncu launch list of the captured step"
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active --clock-control none --csv --log-file $OUT/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-ops > $OUT/ncu_bench.log 2>&1
python tools/ncu_launch_shares.py $OUT/launches.csv $OUT/r2 | head -30
echo "== ncu --set full, six GEMM launches (h3 forward)"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:tc_gemm -o $OUT/r2_ncu_gemm python tools/ncu_gemm.py --h3 > $OUT/ncu_gemm.log 2>&1
tail -8 $OUT/ncu_gemm.log
echo "== grad kernel at 3 CTAs/SM (rebuild)"
MPB_NVCC_EXTRA="-DMPB_MS_GRAD_MINB=3" python -c "
from monopsr_b200 import build; build.build()" > /dev/null 2>&1
touch monopsr_b200/csrc/approxmatch.cu
MPB_NVCC_EXTRA="-DMPB_MS_GRAD_MINB=3" python -c "
from monopsr_b200 import build; build.build()" > $OUT/rebuild.log 2>&1
timeout 300 python tools/bench_tfops.py > $OUT/tfops_minb3.json 2>&1
echo "grad minb=3: $(grep -A1 '"am_grad_b32_n1024"\|"am_grad_b32_n2304"' $OUT/tfops_minb3.json | grep med_us | tr '\n' ' ')"
