#!/bin/bash
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout -k 10 900 python -m pytest tests/test_tc_gpu.py tests/test_unet_gpu.py -q -x > gpurun_out/pytest_tc.txt 2>&1
echo "rc=$?" >> gpurun_out/pytest_tc.txt
tail -8 gpurun_out/pytest_tc.txt
timeout -k 10 300 python scratch/prof_eval.py 32 > gpurun_out/eval32_bres.txt 2>&1
RF_GEMM_BRES=0 timeout -k 10 300 python scratch/prof_eval.py 32 > gpurun_out/eval32_nobres.txt 2>&1
cat gpurun_out/eval32_bres.txt gpurun_out/eval32_nobres.txt
timeout -k 10 300 python scratch/prof_b.py 32 > gpurun_out/tc_b32_bres.txt 2>&1
head -12 gpurun_out/tc_b32_bres.txt
