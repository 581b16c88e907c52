#!/bin/bash
# round 2, last call: suite + default bench + approxmatch ncu on the final tree (approxmatch sweep unroll 8)
OUT=gpurun_out/r2_s
mkdir -p $OUT
timeout 1200 python -m pytest tests -m gpu -q -p no:cacheprovider > $OUT/pytest_gpu.log 2>&1
tail -2 $OUT/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -1
timeout 900 python bench.py > $OUT/bench.json 2> $OUT/bench.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/r2_s/bench.json"))
print({k: d[k] for k in ("value", "ms_per_step", "dtype", "gpu_launches")}, d["e2e"]["value"], d["roofline"]["frac"], d["clocks"])
o = d["ops"]["cfg4_approxmatch_b32_n1024"]
print("cfg4", o["match_us"]["med"], o["cost_us"]["med"], o["grad_us"]["med"], o["roofline_match"]["frac"], o["ref_gpu_us"]["match"]["med"])
o = d["ops"]["approxmatch_b32_n2304_inmodel"]
print("2304", o["match_us"]["med"], o["cost_us"]["med"], o["grad_us"]["med"], o["roofline_match"]["frac"])
PY
timeout 300 ncu --set full --clock-control none --import-source on -k regex:approxmatch_cluster_kernel -c 2 -o $OUT/r2_ncu_approxmatch -f python tools/ncu_tfops.py > $OUT/ncu_am.log 2>&1
tail -2 $OUT/ncu_am.log
