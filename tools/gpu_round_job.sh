#!/bin/bash
# One measurement pass on one B200:  gpurun --timeout 2400 -- 'bash tools/gpu_round_job.sh <tag>'
# GPU parity suite, bench (product and reference arm), every operator (plain and 30 % masked) with clocks, the ncu launch
# list of the bench command.  Outputs under gpurun_out/ with the tag in the name; copy what should be judged to profiles/.
TAG=${1:-r02}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_gpu_$TAG.log
python bench.py --steps 20 --warmup 3 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench rc=$?"; head -c 600 gpurun_out/bench_$TAG.json; echo; tail -5 gpurun_out/bench_$TAG.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_$TAG.json 2> gpurun_out/bench_ref_$TAG.err; echo "ref rc=$?"; head -c 300 gpurun_out/bench_ref_$TAG.json; echo
python tools/opbench.py --all --json gpurun_out/opbench_$TAG.json > gpurun_out/opbench_$TAG.log 2>&1; echo "opbench rc=$?"; cat gpurun_out/opbench_$TAG.log
python tools/opbench.py --all --mask 0.3 --json gpurun_out/opbench_${TAG}_masked.json > gpurun_out/opbench_${TAG}_masked.log 2>&1; echo "opbench masked rc=$?"; cat gpurun_out/opbench_${TAG}_masked.log
