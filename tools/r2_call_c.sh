#!/bin/bash
# round 2, call C: fp16-split forward (h3) first run, streaming EMD kernels after the load fix
OUT=gpurun_out/r2_c
mkdir -p $OUT
echo "== h3 tests"
timeout 900 python -m pytest tests/test_h3_gpu.py -q -p no:cacheprovider > $OUT/h3_tests.log 2>&1
tail -40 $OUT/h3_tests.log
echo "== tf32 / x3 / h3 per shape"
timeout 600 python tools/h3_sweep.py > $OUT/h3_sweep.txt 2>&1
cat $OUT/h3_sweep.txt
echo "== forward parity h3 (+ gradients)"
MPB_PRECISION=h3 timeout 300 python tools/check_network.py --bwd > $OUT/check_network_h3.txt 2>&1
grep -h "feat\|map_features\|inst_xyz_map_local\|centroids \|alpha_bins\|cen_z_offs\|total\|median" $OUT/check_network_h3.txt
MPB_PRECISION=x3 timeout 300 python tools/check_network.py --bwd > $OUT/check_network_x3.txt 2>&1
grep -h "median" $OUT/check_network_x3.txt
echo "== bench h3 / x3 / tf32"
for P in h3 x3 tf32; do
MPB_PRECISION=$P timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-ops > $OUT/bench_$P.json 2> $OUT/bench_$P.err
python - <<PY
import json
try:
    d = json.load(open("$OUT/bench_$P.json"))
    print("$P", "%.3f ms/step" % d["ms_per_step"], "%.0f crops/s" % d["value"], "e2e %.0f" % d["e2e"]["value"], "gemm alone", d["roofline"]["kernel"])
except Exception as e:
    print("$P bench failed:", e); print(open("$OUT/bench_$P.err").read()[-1500:])
PY
done
echo "== EMD streaming kernels"
for R in 16 32 64; do
  echo "rows=$R"
  MPB_MS_ROWS=$R timeout 300 python tools/bench_tfops.py > $OUT/tfops_rows$R.json 2>&1
  grep -A1 '"am_cost_b32_n1024"\|"am_grad_b32_n1024"\|"am_cost_b32_n2304"\|"am_grad_b32_n2304"' $OUT/tfops_rows$R.json | grep "med_us" | tr '\n' ' '; echo
done
timeout 300 python -m pytest tests/test_tfops_gpu.py -q -x -p no:cacheprovider -k emd 2>&1 | tail -3
