#!/bin/bash
# every BASELINE config at N = $1 GPUs (default 1); one JSON line each under gpurun_out/
N=${1:-1}
TAG=${2:-r02}
STEPS=${3:-10}
mkdir -p gpurun_out
for c in ${CONFIGS:-vit_b16 vit_tiny mae_vitl_fsdp diffusion_fsdp unetr_128 sap_4096_L1024 sap_4096_L4096}; do
  if [ "$N" = "1" ]; then
    timeout 600 python bench.py --config $c --steps $STEPS --warmup 3 > gpurun_out/${TAG}_bench_${c}_n${N}.json 2> gpurun_out/${TAG}_bench_${c}_n${N}.err
  else
    timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --config $c --gpus $N --steps $STEPS --warmup 3 > gpurun_out/${TAG}_bench_${c}_n${N}.json 2> gpurun_out/${TAG}_bench_${c}_n${N}.err
  fi
  echo "== $c rc=$?"; tail -c 1500 gpurun_out/${TAG}_bench_${c}_n${N}.json; tail -5 gpurun_out/${TAG}_bench_${c}_n${N}.err
done
