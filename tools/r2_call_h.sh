#!/bin/bash
# round 2, call H: fold fused into the split-copy pass; backward knobs (A-tile multicast clusters, wgrad tiling)
OUT=gpurun_out/r2_h
mkdir -p $OUT
timeout 900 python -m pytest tests/test_h3_gpu.py tests/test_x3_gpu.py tests/test_network_gpu.py tests/test_zz_late_additions_gpu.py tests/test_pointset_loss_gpu.py tests/test_tfops_gpu.py -q -x -p no:cacheprovider 2>&1 | tail -3
run() { # tag, env...
  local tag=$1; shift
  env "$@" timeout 200 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-ops > $OUT/bench_$tag.json 2> $OUT/bench_$tag.err
  python - <<PY
import json
try:
    d = json.load(open("$OUT/bench_$tag.json"))
    print("$tag", "%.3f ms/step" % d["ms_per_step"], "e2e %.0f" % d["e2e"]["value"], "alone %.2f ms" % d["roofline"]["alone"]["ms"])
except Exception as e:
    print("$tag failed", e); print(open("$OUT/bench_$tag.err").read()[-600:])
PY
}
run base A=1
run cluster2 MPB_TC_CLUSTER=2
run cluster4 MPB_TC_CLUSTER=4
run wgbn256 MPB_WGRAD_BN=256
run wgfill075 MPB_WGRAD_FILL=0.75
run wgfill1 MPB_WGRAD_FILL=1.0
run wgfill035 MPB_WGRAD_FILL=0.35
run fill06_sk128 MPB_TILE_FILL=0.6 MPB_SHORTK_BN=128
run base2 A=1
