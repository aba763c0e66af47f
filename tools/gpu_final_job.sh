#!/bin/bash
# Shorter measurement pass (the fused chain and the TFP tile kernel are unchanged since their last ncu captures):
#   gpurun --timeout 900 -- 'bash tools/gpu_final_job.sh <tag>'
# GPU parity suite, bench (product and reference arm), every operator (plain and 30 % masked) with clocks, and one
# `ncu --set full` capture each of the shapiro2 kernel and of the masked stddevValue kernel.
TAG=${1:-r02}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu_$TAG.log
python bench.py --steps 20 --warmup 3 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench rc=$?"; head -c 400 gpurun_out/bench_$TAG.json; echo; tail -3 gpurun_out/bench_$TAG.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_$TAG.json 2> gpurun_out/bench_ref_$TAG.err; echo "ref rc=$?"; head -c 200 gpurun_out/bench_ref_$TAG.json; echo
python tools/opbench.py --all --json gpurun_out/opbench_$TAG.json > gpurun_out/opbench_$TAG.log 2>&1; echo "opbench rc=$?"; cat gpurun_out/opbench_$TAG.log
python tools/opbench.py --all --mask 0.3 --json gpurun_out/opbench_${TAG}_masked.json > gpurun_out/opbench_${TAG}_masked.log 2>&1; echo "opbench masked rc=$?"; cat gpurun_out/opbench_${TAG}_masked.log
S="python tools/opbench.py --ops shapiro2_filter --seconds 0.05"
$S > gpurun_out/plain_$TAG.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:shapiro2 -s 2 -c 1 -f -o gpurun_out/prof_${TAG}_shapiro2 $S > gpurun_out/ncu_shapiro2_$TAG.log 2>&1; echo "shapiro2 capture rc=$?"
E="python tools/opbench.py --ops stddevValue_masked5 --seconds 0.05"
$E > gpurun_out/plain_$TAG.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:ensemble_kernel -s 2 -c 1 -f -o gpurun_out/prof_${TAG}_stddev_masked $E > gpurun_out/ncu_stddev_masked_$TAG.log 2>&1; echo "stddev capture rc=$?"
du -sh gpurun_out
