#!/bin/bash
# One measurement pass on one B200:  gpurun --timeout 2400 -- 'bash tools/gpu_round_job.sh <tag>'
# GPU parity suite, bench (product and reference arm), every operator (plain and 30 % masked) with clocks, the ncu launch
# list of the bench command and one `ncu --set full` capture each of the fused chain kernel and of the TFP tile kernel.
# Outputs under gpurun_out/ with the tag in the name; copy what should be judged to profiles/
# (tools/ncu_summary.py <rep> --traffic profiles/chain_traffic.json feeds bench.py's roofline.traffic).
TAG=${1:-r02}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_gpu_$TAG.log
python bench.py --steps 20 --warmup 3 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench rc=$?"; head -c 600 gpurun_out/bench_$TAG.json; echo; tail -5 gpurun_out/bench_$TAG.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_$TAG.json 2> gpurun_out/bench_ref_$TAG.err; echo "ref rc=$?"; head -c 300 gpurun_out/bench_ref_$TAG.json; echo
python tools/opbench.py --all --json gpurun_out/opbench_$TAG.json > gpurun_out/opbench_$TAG.log 2>&1; echo "opbench rc=$?"; cat gpurun_out/opbench_$TAG.log
python tools/opbench.py --all --mask 0.3 --json gpurun_out/opbench_${TAG}_masked.json > gpurun_out/opbench_${TAG}_masked.log 2>&1; echo "opbench masked rc=$?"; cat gpurun_out/opbench_${TAG}_masked.log
# ---- ncu (the same commands have just exited 0 without it)
B="python bench.py --steps 2 --warmup 3 --no-ops --no-cpu"
$B > gpurun_out/plain_$TAG.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$TAG.csv $B > gpurun_out/ncu_launches_$TAG.log 2>&1; echo "launch list rc=$?"
$B > gpurun_out/plain_$TAG.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:ew_kernel -s 4 -c 1 -f -o gpurun_out/prof_${TAG}_chain $B > gpurun_out/ncu_chain_$TAG.log 2>&1; echo "chain capture rc=$?"
T="python tools/opbench.py --ops thermalFrontParameter --seconds 0.05"
$T > gpurun_out/plain_$TAG.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:tfp_tile -s 2 -c 1 -f -o gpurun_out/prof_${TAG}_tfp $T > gpurun_out/ncu_tfp_$TAG.log 2>&1; echo "tfp capture rc=$?"
du -sh gpurun_out
