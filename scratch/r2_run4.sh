#!/bin/bash
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout -k 10 900 python -m pytest tests/test_tc_gpu.py tests/test_unet_gpu.py -q -x > gpurun_out/pytest_tc.txt 2>&1
echo "rc=$?" >> gpurun_out/pytest_tc.txt
tail -12 gpurun_out/pytest_tc.txt
timeout -k 10 300 python scratch/prof_eval.py 32 > gpurun_out/eval32_r4.txt 2>&1
RF_GEMM_320_MIN_KB=5 timeout -k 10 300 python scratch/prof_eval.py 32 > gpurun_out/eval32_r4_320all.txt 2>&1
RF_GEMM_320_MIN_KB=9999 timeout -k 10 300 python scratch/prof_eval.py 32 > gpurun_out/eval32_r4_no320.txt 2>&1
cat gpurun_out/eval32_r4.txt gpurun_out/eval32_r4_320all.txt gpurun_out/eval32_r4_no320.txt
timeout -k 10 300 python scratch/prof_b.py 32 > gpurun_out/tc_b32_r4.txt 2>&1
RF_GEMM_320_MIN_KB=5 timeout -k 10 300 python scratch/prof_b.py 32 > gpurun_out/tc_b32_r4_320all.txt 2>&1
head -45 gpurun_out/tc_b32_r4.txt
grep -E "N=  320 K=   320|N=  640 K=   640" gpurun_out/tc_b32_r4_320all.txt
timeout -k 10 1500 python -m pytest tests/test_parity_bench_gpu.py -m gpu -q -s > gpurun_out/pytest_parity_full.txt 2>&1
grep -E "kernels vs|floor|apart|riffuse|uint8|waveform|FAILED|passed|failed|Error" gpurun_out/pytest_parity_full.txt | head -60
