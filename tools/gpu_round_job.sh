#!/bin/bash
# The measurement pass behind profiles/r01_*_f*: bench (product and reference arm), every operator (plain and 30 % masked),
# the ncu launch list of the bench command.  Run with gpurun on one B200:  gpurun --timeout 1500 -- 'bash tools/gpu_round_job.sh'
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
python bench.py --steps 20 --warmup 3 > gpurun_out/bench_r01_f.json 2> gpurun_out/bench_r01_f.err; echo "bench rc=$?"; head -c 400 gpurun_out/bench_r01_f.json; echo
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_r01_f.json 2> gpurun_out/bench_ref_r01_f.err; echo "ref rc=$?"
python tools/opbench.py --reps 10 --json gpurun_out/opbench_r01_f.json > gpurun_out/opbench_r01_f.log 2>&1; echo "opbench rc=$?"; cat gpurun_out/opbench_r01_f.log
python tools/opbench.py --reps 10 --mask 0.3 --json gpurun_out/opbench_r01_f_masked.json > gpurun_out/opbench_r01_f_masked.log 2>&1; echo "opbench masked rc=$?"; tail -70 gpurun_out/opbench_r01_f_masked.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r01_f.csv python bench.py --steps 2 --warmup 1 --no-cpu > gpurun_out/ncu_launches.log 2>&1; echo "launchlist rc=$?"
