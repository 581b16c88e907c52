#!/bin/bash
# round 2, final evidence 1/2 (one build): GPU suite, smoke, default bench + reference arm, step timeline
OUT=gpurun_out/r2_q
mkdir -p $OUT
timeout 1200 python -m pytest tests -m gpu -q -p no:cacheprovider > $OUT/pytest_gpu.log 2>&1
tail -4 $OUT/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
echo "== default bench"
( time timeout 900 python bench.py > $OUT/bench.json 2> $OUT/bench.err ) 2>&1 | grep real
tail -c 400 $OUT/bench.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/r2_q/bench.json"))
print({k: d[k] for k in ("value", "ms_per_step", "dtype", "gpu_launches")}, d["e2e"], d["roofline"]["frac"], d["clocks"])
for k, v in d.get("ops", {}).items():
    print(k, json.dumps(v)[:1200])
print(d.get("cpu_baseline"))
PY
echo "== reference arm"
( time timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $OUT/bench_ref.json 2> $OUT/bench_ref.err ) 2>&1 | grep real
tail -c 600 $OUT/bench_ref.json
echo "== tf32 fast mode, same build"
MPB_PRECISION=tf32 timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-ops > $OUT/bench_tf32.json 2> $OUT/bench_tf32.err
python -c "
import json; d=json.load(open('$OUT/bench_tf32.json')); print('tf32 %.3f ms/step e2e %.0f frac %.3f' % (d['ms_per_step'], d['e2e']['value'], d['roofline']['frac']))"
echo "== step timeline"
timeout 300 python tools/step_timeline.py > $OUT/step_timeline.txt 2>&1; tail -32 $OUT/step_timeline.txt | head -12
cp gpurun_out/step_kernels.csv $OUT/ 2>/dev/null
echo "== parity table"
timeout 300 python tools/check_network.py > $OUT/check_network_h3.txt 2>&1; tail -25 $OUT/check_network_h3.txt
