#!/bin/bash
# round 2, multi-GPU call: data-parallel step as one CUDA graph with bucketed NCCL all-reduces captured inside
N=${1:-2}
OUT=gpurun_out/r2_dp$N
mkdir -p $OUT
export MASTER_ADDR=127.0.0.1
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tools/dp_check.py > $OUT/dp_check.log 2>&1
tail -12 $OUT/dp_check.log
for B in 4 8; do
MPB_DP_BUCKETS=$B timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 20 --warmup 5 > $OUT/bench_b$B.json 2> $OUT/bench_b$B.err
python - <<PY
import json
try:
    d = json.load(open("$OUT/bench_b$B.json"))
    print("buckets $B: N=%d %.3f ms/step %.0f crops/s e2e %.0f" % (d["n_gpus"], d["ms_per_step"], d["value"], d["e2e"]["value"]))
except Exception as e:
    print("bench failed", e); print(open("$OUT/bench_b$B.err").read()[-1500:])
PY
done
MPB_DP_GRAPH=0 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --steps 20 --warmup 5 > $OUT/bench_split.json 2> $OUT/bench_split.err
python -c "
import json; d=json.load(open('$OUT/bench_split.json')); print('split graphs: N=%d %.3f ms/step %.0f crops/s' % (d['n_gpus'], d['ms_per_step'], d['value']))"
timeout 300 python bench.py --gpus 1 --steps 20 --warmup 5 --no-ops --no-cpu-baseline > $OUT/bench_n1.json 2>/dev/null
python -c "
import json; d=json.load(open('$OUT/bench_n1.json')); print('N=1 %.3f ms/step %.0f crops/s' % (d['ms_per_step'], d['value']))"
