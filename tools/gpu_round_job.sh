mkdir -p gpurun_out
python bench.py --steps 20 --warmup 3 > gpurun_out/bench_r01_d.json 2> gpurun_out/bench_r01_d.err; echo "bench rc=$?"; cat gpurun_out/bench_r01_d.json | head -c 3000
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_r01_d.json 2> gpurun_out/bench_ref_r01_d.err; echo "ref rc=$?"; cat gpurun_out/bench_ref_r01_d.json | head -c 1500
python tools/opbench.py --reps 10 --json gpurun_out/opbench_r01_d.json > gpurun_out/opbench_r01_d.log 2>&1; echo "opbench rc=$?"; cat gpurun_out/opbench_r01_d.log
python tools/opbench.py --reps 10 --mask 0.3 --json gpurun_out/opbench_r01_d_masked.json > gpurun_out/opbench_r01_d_masked.log 2>&1; echo "opbench masked rc=$?"; tail -45 gpurun_out/opbench_r01_d_masked.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r01_d.csv python bench.py --steps 2 --warmup 1 --no-cpu > gpurun_out/ncu_launches.log 2>&1; echo "launchlist rc=$?"
ncu --set full --clock-control none --import-source on -k regex:ew_kernel -s 2 -c 1 -o gpurun_out/prof_r01_chain_final -f python tools/opbench.py --ops alevel_chain --reps 3 > gpurun_out/ncu_chain_final.log 2>&1; echo "ncu rc=$?"
