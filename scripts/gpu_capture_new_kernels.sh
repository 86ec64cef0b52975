#!/bin/bash
# ncu evidence for the UNETR-decoder InstanceNorm kernels and the SAP front-end kernels (1 GPU).
TAG=${1:-r02}
O=gpurun_out
M="gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum"
for w in inorm canny; do
  python scripts/gpu_one_kernel.py $w > /dev/null 2>&1 || { echo "plain run of $w failed"; continue; }
  timeout 600 ncu --metrics $M --clock-control none -k regex:"inorm_|canny_|blur_u8" --csv --log-file $O/${TAG}_${w}_launches.csv python scripts/gpu_one_kernel.py $w > $O/${TAG}_ncu_${w}.log 2>&1
done
cap() {  # name, kernel regex, skip, script args...
  local name=$1 regex=$2 skip=$3; shift 3
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:$regex -s $skip -c 1 -f -o $O/${TAG}_$name python "$@" > $O/${TAG}_ncu_$name.log 2>&1
  ncu -i $O/${TAG}_$name.ncu-rep --page details --csv > $O/${TAG}_$name.details.csv 2>/dev/null
}
cap inorm_bwd_apply inorm_bwd_apply_kernel 3 scripts/gpu_one_kernel.py inorm
cap inorm_stats inorm_stats_kernel 3 scripts/gpu_one_kernel.py inorm
cap canny_nms canny_nms_kernel 2 scripts/gpu_one_kernel.py canny
cap blur5 blur_u8_kernel 4 scripts/gpu_one_kernel.py canny
ls -la $O | grep ${TAG}_ | grep -E "inorm|canny|blur"
