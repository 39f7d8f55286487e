#!/bin/bash
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout -k 10 900 python -m pytest tests/test_tc_gpu.py tests/test_unet_gpu.py -q -x > gpurun_out/pytest_tc.txt 2>&1
echo "rc=$?" >> gpurun_out/pytest_tc.txt
tail -6 gpurun_out/pytest_tc.txt
for sl in 2 1 3; do
  RF_GEMM_SLABS=$sl timeout -k 10 300 python scratch/prof_eval.py 32 > gpurun_out/eval32_slabs$sl.txt 2>&1
  echo "slabs=$sl: $(cat gpurun_out/eval32_slabs$sl.txt)"
done
timeout -k 10 300 python scratch/prof_b.py 32 > gpurun_out/tc_b32_slabs2.txt 2>&1
RF_GEMM_SLABS=3 timeout -k 10 300 python scratch/prof_b.py 32 > gpurun_out/tc_b32_slabs3.txt 2>&1
head -30 gpurun_out/tc_b32_slabs2.txt
grep "bn= 160" gpurun_out/tc_b32_slabs3.txt | head -8
