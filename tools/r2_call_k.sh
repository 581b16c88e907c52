#!/bin/bash
# round 2, call K: head train-op under the tail of the towers' backward pass (start unit sweep)
OUT=gpurun_out/r2_k
mkdir -p $OUT
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
run early_now MPB_EARLY_OPT=1
run early_u16 MPB_EARLY_OPT=1 MPB_EARLY_OPT_UNIT=16
run early_u10 MPB_EARLY_OPT=1 MPB_EARLY_OPT_UNIT=10
run early_u6 MPB_EARLY_OPT=1 MPB_EARLY_OPT_UNIT=6
run early_u3 MPB_EARLY_OPT=1 MPB_EARLY_OPT_UNIT=3
run early_u0 MPB_EARLY_OPT=1 MPB_EARLY_OPT_UNIT=0
run base2 A=1
MPB_EARLY_OPT=1 MPB_EARLY_OPT_UNIT=6 timeout 300 python -m pytest tests/test_network_gpu.py -q -x -p no:cacheprovider -k "graph_replay" 2>&1 | tail -2
