#!/bin/bash
# round 2: tower gradients in two buckets (second one under the first tower's train-op) vs one bucket; N GPUs
N=${1:-2}
OUT=gpurun_out/r2_dps$N
mkdir -p $OUT
export MASTER_ADDR=127.0.0.1
timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tools/dp_check.py > $OUT/dp_check.log 2>&1
grep "step \|ms/step\|DP CHECK" $OUT/dp_check.log
run() {
  tag=$1; shift
  env "$@" timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 30 --warmup 5 --no-ops --no-cpu-baseline > $OUT/bench_$tag.json 2> $OUT/bench_$tag.err
  python - <<PY
import json
try:
    d = json.loads([l for l in open("$OUT/bench_$tag.json") if l.startswith("{")][-1])
    print("$tag N=%d %.3f ms/step %.0f crops/s e2e %.0f" % (d["n_gpus"], d["ms_per_step"], d["value"], d["e2e"]["value"]))
except Exception as e:
    print("bench failed", e); print(open("$OUT/bench_$tag.err").read()[-1500:])
PY
}
run split2 MPB_DP_TOWER_SPLIT=1
run split1 MPB_DP_TOWER_SPLIT=0
run split2b MPB_DP_TOWER_SPLIT=1
timeout 120 python bench.py --gpus 1 --steps 30 --warmup 5 --no-ops --no-cpu-baseline > $OUT/bench_n1.json 2>/dev/null
python -c "
import json; d=json.load(open('$OUT/bench_n1.json')); print('N=1 %.3f ms/step %.0f crops/s' % (d['ms_per_step'], d['value']))"
