#!/bin/bash
# round 2, call I: critical-path SIMT kernels (crop+pool, xyz head, small FC backward) rewritten; FC-stream priority
OUT=gpurun_out/r2_j
mkdir -p $OUT
timeout 900 python -m pytest tests/test_net_kernels_gpu.py tests/test_network_gpu.py tests/test_h3_gpu.py tests/test_zz_late_additions_gpu.py tests/test_pointset_loss_gpu.py -q -x -p no:cacheprovider 2>&1 | tail -4
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
run base A=1

run base2 A=1

run tf32 MPB_PRECISION=tf32
timeout 200 python tools/step_timeline.py > $OUT/step_timeline.txt 2>&1; grep "step span\|concurrency" $OUT/step_timeline.txt
cp gpurun_out/step_kernels.csv $OUT/ 2>/dev/null
