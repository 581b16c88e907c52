#!/bin/bash
# round 2, call L: split-copy weight pass by row-length class; match_cost at 4 CTAs/SM
OUT=gpurun_out/r2_l
mkdir -p $OUT
timeout 600 python -m pytest tests/test_h3_gpu.py tests/test_tfops_gpu.py tests/test_network_gpu.py -q -x -p no:cacheprovider 2>&1 | tail -2
timeout 300 python tools/bench_tfops.py > $OUT/tfops.json 2>&1
echo "cost/grad 1024, cost/grad 2304: $(grep -A1 '"am_cost_b32_n1024"\|"am_grad_b32_n1024"\|"am_cost_b32_n2304"\|"am_grad_b32_n2304"' $OUT/tfops.json | grep med_us | tr '\n' ' ')"
for t in a b; do
timeout 200 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-ops > $OUT/bench_$t.json 2> $OUT/bench_$t.err
python -c "
import json; d=json.load(open('$OUT/bench_$t.json')); print('h3 %.3f ms/step e2e %.0f frac %.3f' % (d['ms_per_step'], d['e2e']['value'], d['roofline']['frac']))"
done
timeout 200 python tools/step_timeline.py > $OUT/step_timeline.txt 2>&1; grep "step span\|concurrency\|split16_weights\|opt_adam" $OUT/step_timeline.txt
cp gpurun_out/step_kernels.csv $OUT/ 2>/dev/null
