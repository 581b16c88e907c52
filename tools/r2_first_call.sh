#!/bin/bash
# First GPU call of the next round: everything that was written after the round-1 GPU budget ran out, in one go.
#   /usr/local/graft/bin/gpurun --timeout 1500 -- 'bash tools/r2_first_call.sh'
# Outputs land in gpurun_out/r2_first/ (merged back by gpurun).  Each step runs under its own timeout so that a hang
# in the (never run) 3xTF32 kernel cannot eat the call; its mbarrier waits are bounded and trap instead of spinning.
OUT=gpurun_out/r2_first
mkdir -p $OUT
echo "== x3 cases directly" | tee $OUT/summary.txt
timeout 600 python -m pytest tests/x3_gpu_cases.py -q -p no:cacheprovider > $OUT/x3_cases.log 2>&1
tail -60 $OUT/x3_cases.log | tee -a $OUT/summary.txt
echo "== late GPU tests (xfail-marked: look for XPASS)" | tee -a $OUT/summary.txt
timeout 1200 python -m pytest tests/test_zz_late_additions_gpu.py -q -rxX -s --runxfail -p no:cacheprovider > $OUT/late_tests.log 2>&1
tail -120 $OUT/late_tests.log | tee -a $OUT/summary.txt
echo "== 3xTF32 vs single-pass per shape (us, error vs fp64)" | tee -a $OUT/summary.txt
timeout 300 python tools/x3_sweep.py > $OUT/x3_sweep.txt 2>&1
cat $OUT/x3_sweep.txt | tee -a $OUT/summary.txt
echo "== forward parity of the engine, default and x3" | tee -a $OUT/summary.txt
timeout 300 python tools/check_network.py > $OUT/check_network_tf32.txt 2>&1
MPB_PRECISION=x3 timeout 300 python tools/check_network.py > $OUT/check_network_x3.txt 2>&1
grep -h "inst_xyz_map_local\|map_features\|centroids \|alpha_bins\|total" $OUT/check_network_tf32.txt $OUT/check_network_x3.txt | tee -a $OUT/summary.txt
echo "== bench, default and x3" | tee -a $OUT/summary.txt
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-ops > $OUT/bench_tf32.json 2> $OUT/bench_tf32.err
MPB_PRECISION=x3 timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-ops > $OUT/bench_x3.json 2> $OUT/bench_x3.err
python - <<'PY' | tee -a $OUT/summary.txt
import json
for tag in ("tf32", "x3"):
    try:
        d = json.load(open("gpurun_out/r2_first/bench_%s.json" % tag))
        print(tag, "%.3f ms/step" % d["ms_per_step"], "%.0f crops/s" % d["value"], "e2e %.0f" % d["e2e"]["value"],
              "roofline.frac %.3f" % d["roofline"]["frac"], d["roofline"]["kernel"])
    except Exception as e:
        print(tag, "bench failed:", e)
PY
