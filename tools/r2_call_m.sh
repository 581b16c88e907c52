#!/bin/bash
# round 2, call M: decoder batch norm as one launch per direction (grid barrier)
OUT=gpurun_out/r2_m
mkdir -p $OUT
timeout 600 python -m pytest tests/test_net_kernels_gpu.py tests/test_network_gpu.py tests/test_zz_late_additions_gpu.py -q -x -p no:cacheprovider 2>&1 | tail -3
run() { # tag, env...
  local tag=$1; shift
  env "$@" timeout 200 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-ops > $OUT/bench_$tag.json 2> $OUT/bench_$tag.err
  python - <<PY
import json
try:
    d = json.load(open("$OUT/bench_$tag.json"))
    print("$tag", "%.3f ms/step" % d["ms_per_step"], "e2e %.0f" % d["e2e"]["value"], "frac %.3f" % d["roofline"]["frac"])
except Exception as e:
    print("$tag failed", e); print(open("$OUT/bench_$tag.err").read()[-600:])
PY
}
run fused A=1
run unfused MPB_BN_FUSED=0
run fused2 A=1
run unfused2 MPB_BN_FUSED=0
timeout 200 python tools/step_timeline.py > $OUT/step_timeline.txt 2>&1; grep "step span\|concurrency\|bn_" $OUT/step_timeline.txt
cp gpurun_out/step_kernels.csv $OUT/ 2>/dev/null
