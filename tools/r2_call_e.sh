#!/bin/bash
# round 2, call E: suite in the h3 default, default bench (ops block), ncu of the point-set ops, sanitizers
OUT=gpurun_out/r2_e
mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider > $OUT/pytest_gpu.log 2>&1
tail -8 $OUT/pytest_gpu.log
echo "== default bench"
( time timeout 900 python bench.py > $OUT/bench.json 2> $OUT/bench.err ) 2>&1 | grep real
tail -c 600 $OUT/bench.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/r2_e/bench.json"))
print({k: d[k] for k in ("value", "ms_per_step", "dtype")}, d["e2e"], d["roofline"]["frac"], d["roofline"]["alone"], d["roofline"]["tensor_busy_frac"])
for k, v in d.get("ops", {}).items():
    print(k, json.dumps(v)[:900])
print(d.get("cpu_baseline"))
PY
echo "== ncu point-set ops"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'nn_distance|approxmatch_cluster_kernel|matchcost_stream|matchcostgrad_stream' -o $OUT/r2_ncu_tfops python tools/ncu_tfops.py > $OUT/ncu_tfops.log 2>&1
tail -3 $OUT/ncu_tfops.log
ls -la $OUT/*.ncu-rep
echo "== sanitizers"
bash tools/sanitize.sh $OUT
