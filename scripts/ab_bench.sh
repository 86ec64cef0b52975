#!/bin/bash
# A/B two builds of the library ON THE SAME BOX, alternating (box-to-box variance is +-2 %, more than most
# single changes).  Usage (from the repo root, two libraries prepared under ab_tmp/):
#   gpurun -- 'bash scripts/ab_bench.sh 3'
# prints one line per run: variant, images/s, ms/step, in-step GEMM TFLOP/s, median SM clock.
n=${1:-3}
for i in $(seq 1 "$n"); do
  for v in old new; do
    cp "ab_tmp/lib_$v.so" ucf_vit_b200/lib/libucfvit_b200.so
    python bench.py --no-cpu-baseline 2>&1 | tail -n 1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$v', round(d['value']), round(d['ms_per_step'],2), round(d['roofline']['achieved']), d['clocks']['sm_mhz'])"
  done
done
