#!/bin/bash
# ncu evidence of the current build (1 GPU): launch list of one bench step + --set full captures of the hot kernels.
# Run only after the plain commands have exited 0 without ncu.
TAG=${1:-r02}
O=gpurun_out
M="gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"
[ -n "$SKIP_LIST" ] || python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $O/${TAG}_plain_bench.json 2>$O/${TAG}_plain_bench.err || exit 1
[ -n "$SKIP_LIST" ] || timeout 900 ncu --metrics $M --clock-control none -s 1200 -c 700 --csv --log-file $O/${TAG}_launches_metrics.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $O/${TAG}_ncu_bench.log 2>&1
[ -n "$SKIP_LIST" ] || python scripts/ncu_launches_summary.py $O/${TAG}_launches_metrics.csv > $O/${TAG}_step_kernel_table.txt 2>&1
cap() {  # name, kernel regex, script args...
  local name=$1 regex=$2; shift 2
  python "$@" > /dev/null 2>&1 || { echo "plain run of $name failed"; return; }
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:$regex -s 2 -c 1 -f -o $O/${TAG}_$name python "$@" > $O/${TAG}_ncu_$name.log 2>&1
  ncu -i $O/${TAG}_$name.ncu-rep --page details --csv > $O/${TAG}_$name.details.csv 2>/dev/null
}
cap attn_fwd_n197 attn_fwd_kernel scripts/gpu_one_kernel.py attn_fwd 256 197 12 64
cap attn_bwd_n197 attn_bwd2_kernel scripts/gpu_one_kernel.py attn_bwd 256 197 12 64
cap attn_fwd_n4096 attn_fwd_kernel scripts/gpu_one_kernel.py attn_fwd 4 4096 12 64
cap attn_bwd_n4096 attn_bwd2_kernel scripts/gpu_one_kernel.py attn_bwd 4 4096 12 64
cap fc1_gelu gemm2_bf16_kernel scripts/gpu_one_kernel.py fc1_gelu
cap sap_gather sap_gather scripts/gpu_one_kernel.py sap_gather
cap var_attn_fwd var_attn_fwd scripts/gpu_one_kernel.py var_attn
ls -la $O | grep ${TAG}_ | head -40
cat $O/${TAG}_step_kernel_table.txt
