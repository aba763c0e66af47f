#!/bin/bash
# TEMPORARY tuning driver: parity first, then the chain kernel shapes side by side
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/tune_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/tune_pytest.log
tail -3 gpurun_out/tune_pytest.log
for v in 0 1 2 3 4; do
    echo "== variant $v"
    FCB200_CHAIN_VARIANT=$v python tools/opbench.py --ops alevel_chain --reps 20 2>&1 | tail -1
done > gpurun_out/tune_chain.log 2>&1
cat gpurun_out/tune_chain.log
python tools/opbench.py --ops aleveltemp_c3,alevelhum_c1,alevelhum_c5,alevelthe_c1,alevelducting_c1,windCooling,pleveltemp_c4,plevelhum_c1,plevelhum_c7,fieldOPERfield_add,vesselIcingOverland --reps 20 2>&1 | tee gpurun_out/tune_pow_ops.log | tail -12
FCB200_CHAIN_VARIANT=1 python tools/opbench.py --ops alevel_chain --mask 0.3 --reps 20 2>&1 | tail -1
