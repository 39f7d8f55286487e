#!/bin/bash
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout -k 5 120 scratch/mma_bench.bin > gpurun_out/mma_bench.txt 2>&1
cat gpurun_out/mma_bench.txt
timeout -k 10 900 python -m pytest tests/test_tc_gpu.py tests/test_unet_gpu.py -q -x > gpurun_out/pytest_tc.txt 2>&1
echo "rc=$?" >> gpurun_out/pytest_tc.txt
tail -12 gpurun_out/pytest_tc.txt
timeout -k 10 300 python scratch/prof_eval.py 32 > gpurun_out/eval32_r3.txt 2>&1
cat gpurun_out/eval32_r3.txt
timeout -k 10 1500 python -m pytest tests/test_parity_bench_gpu.py -m gpu -q -s > gpurun_out/pytest_parity_full.txt 2>&1
grep -E "kernels vs|floor|riffuse|uint8|waveform|FAILED|passed|failed|Error" gpurun_out/pytest_parity_full.txt | head -60
timeout -k 10 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_eval32_r3.csv python scratch/prof_eval.py 32 ncu > gpurun_out/ncu_eval32.log 2>&1
