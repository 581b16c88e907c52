#!/bin/bash
# bench.py under different tile-planning knobs (ms/step); usage: tools/knob_sweep.sh "CSK WGRAD_BN WGRAD_FILL TILE_FILL" ...
for cfg in "$@"; do
  set -- $cfg
  r=$(MPB_CSK=$1 MPB_WGRAD_BN=$2 MPB_WGRAD_FILL=$3 MPB_TILE_FILL=$4 MPB_CSK_BN=${5:-256} python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-ops 2>/dev/null | python -c "import json,sys; d=json.load(sys.stdin); print('%.3f ms/step  gemm-only %s' % (d['ms_per_step'], d['roofline']['kernel'].split(',')[-1]))")
  echo "CSK=$1 WGRAD_BN=$2 WGRAD_FILL=$3 TILE_FILL=$4 CSK_BN=${5:-256}: $r"
done
