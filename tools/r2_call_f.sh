#!/bin/bash
# round 2, call F: EMD streaming kernels (single-wave tiling, primed loads), fused split16 producers, tile knobs for h3
OUT=gpurun_out/r2_f
mkdir -p $OUT
timeout 600 python -m pytest tests/test_tfops_gpu.py tests/test_h3_gpu.py tests/test_network_gpu.py -q -x -p no:cacheprovider 2>&1 | tail -3
for R in auto 32 64 80 128; do
  if [ $R = auto ]; then unset MPB_MS_ROWS; else export MPB_MS_ROWS=$R; fi
  timeout 300 python tools/bench_tfops.py > $OUT/tfops_rows$R.json 2>&1
  echo "rows=$R $(grep -A1 '"am_cost_b32_n1024"\|"am_grad_b32_n1024"\|"am_cost_b32_n2304"\|"am_grad_b32_n2304"' $OUT/tfops_rows$R.json | grep med_us | tr '\n' ' ')"
done
unset MPB_MS_ROWS
run() { # tag, env...
  local tag=$1; shift
  env "$@" timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-ops > $OUT/bench_$tag.json 2> $OUT/bench_$tag.err
  python - <<PY
import json
try:
    d = json.load(open("$OUT/bench_$tag.json"))
    print("$tag", "%.3f ms/step" % d["ms_per_step"], "e2e %.0f" % d["e2e"]["value"], "alone %.2f ms" % d["roofline"]["alone"]["ms"], "frac %.3f" % d["roofline"]["frac"])
except Exception as e:
    print("$tag failed", e); print(open("$OUT/bench_$tag.err").read()[-800:])
PY
}
run base A=1
run shortk128 MPB_SHORTK_BN=128
run fill07 MPB_TILE_FILL=0.7
run fill06_sk128 MPB_TILE_FILL=0.6 MPB_SHORTK_BN=128
run csk MPB_CSK=1
run base2 A=1
run tf32 MPB_PRECISION=tf32
run x3 MPB_PRECISION=x3
