#!/bin/bash
# round 2, run 1: diffusers probe, new parity tests, batch-32 evaluation timing + launch list
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
python -c "import diffusers; print('diffusers', diffusers.__version__)" > gpurun_out/diffusers_probe.txt 2>&1
pip download diffusers==0.9.0 --no-deps -d /tmp/x >> gpurun_out/diffusers_probe.txt 2>&1 | tail -1
(timeout 1500 python -m pytest tests/test_parity_bench_gpu.py -m gpu -q -s 2>&1 | tail -60) > gpurun_out/pytest_parity.txt 2>&1
(timeout 900 python -m pytest tests -m gpu -x -q --deselect tests/test_parity_bench_gpu.py 2>&1 | tail -8) > gpurun_out/pytest_gpu.txt 2>&1
timeout 300 python scratch/prof_eval.py 32 > gpurun_out/eval32.txt 2>&1
timeout 300 python scratch/prof_b.py 32 > gpurun_out/tc_b32.txt 2>&1
cp gpurun_out/tc_launches.csv gpurun_out/tc_launches_b32.csv
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_eval32.csv python scratch/prof_eval.py 32 ncu > gpurun_out/ncu_eval32.log 2>&1
cat gpurun_out/diffusers_probe.txt | tail -3; cat gpurun_out/pytest_parity.txt; cat gpurun_out/pytest_gpu.txt; cat gpurun_out/eval32.txt; head -30 gpurun_out/tc_b32.txt
