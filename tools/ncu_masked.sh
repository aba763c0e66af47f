#!/bin/bash
# One `ncu --set full` capture of the main kernel of each mask-heavy operator (BASELINE config 5):
#   gpurun --timeout 1500 -- 'bash tools/ncu_masked.sh <tag>'
# (every operator is run once WITHOUT ncu first; all ncu runs of a call count as one)
TAG=${1:-r02}
K='regex:ew_kernel|ensemble_kernel|stencil_tile_kernel|shapiro2_kernel|tfp_tile_kernel'
OPS="relvort_masked30 jacobian_masked30 gradient_c3_masked30 shapiro2_filter_masked30 alevelhum_c5_masked30 stddevValue_masked5 thermalFrontParameter_masked30 vesselIcingMertins_masked30 $2"
python tools/opbench.py --ops "$(echo $OPS | tr ' ' ',')" --seconds 0.05 > gpurun_out/plain_masked_$TAG.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/plain_masked_$TAG.log; exit 1; }
cat gpurun_out/plain_masked_$TAG.log
for op in $OPS; do
  ncu --set full --clock-control none --import-source on -k "$K" -s 1 -c 1 -f -o gpurun_out/prof_${TAG}_$op python tools/opbench.py --ops $op --seconds 0.01 > gpurun_out/ncu_${TAG}_$op.log 2>&1
  echo "$op ncu rc=$?"
  # (gpurun brings back at most 64 MiB: keep the raw page and the per-instruction page as compressed csv, drop the report)
  ncu -i gpurun_out/prof_${TAG}_$op.ncu-rep --page raw --csv > gpurun_out/prof_${TAG}_${op}_raw.csv 2>/dev/null
  ncu -i gpurun_out/prof_${TAG}_$op.ncu-rep --page source --csv 2>/dev/null | gzip > gpurun_out/prof_${TAG}_${op}_source.csv.gz
  rm -f gpurun_out/prof_${TAG}_$op.ncu-rep
done
