#!/bin/bash
# round 2, run 2: CTA-pair GEMM kernel correctness + A/B timing, parity numbers with the accurate SiLU
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout -k 10 900 python -m pytest tests/test_tc_gpu.py -q -x -k "pair or gemm or conv" > gpurun_out/pytest_tc.txt 2>&1
echo "tc rc=$?" >> gpurun_out/pytest_tc.txt
tail -15 gpurun_out/pytest_tc.txt
timeout -k 10 300 python scratch/prof_eval.py 32 > gpurun_out/eval32_pair.txt 2>&1
RF_GEMM_PAIR=0 timeout -k 10 300 python scratch/prof_eval.py 32 > gpurun_out/eval32_nopair.txt 2>&1
RF_SILU_TANH=1 timeout -k 10 300 python scratch/prof_eval.py 32 > gpurun_out/eval32_tanh.txt 2>&1
cat gpurun_out/eval32_pair.txt gpurun_out/eval32_nopair.txt gpurun_out/eval32_tanh.txt
timeout -k 10 300 python scratch/prof_b.py 32 > gpurun_out/tc_b32_pair.txt 2>&1
RF_GEMM_BN=256 timeout -k 10 300 python scratch/prof_b.py 32 > gpurun_out/tc_b32_pair256.txt 2>&1
RF_GEMM_BN=160 timeout -k 10 300 python scratch/prof_b.py 32 > gpurun_out/tc_b32_pair160.txt 2>&1
RF_GEMM_BN=128 timeout -k 10 300 python scratch/prof_b.py 32 > gpurun_out/tc_b32_pair128.txt 2>&1
head -32 gpurun_out/tc_b32_pair.txt
timeout -k 10 1500 python -m pytest tests/test_parity_bench_gpu.py -m gpu -q -s > gpurun_out/pytest_parity_full.txt 2>&1
grep -E "kernels vs|floor|riffuse|uint8|waveform|FAILED|passed|failed|Error" gpurun_out/pytest_parity_full.txt | head -60
RF_SILU_TANH=1 timeout -k 10 600 python -m pytest tests/test_parity_bench_gpu.py -m gpu -q -s -k "fp16_storage_floor" > gpurun_out/pytest_parity_tanh.txt 2>&1
grep -E "kernels vs" gpurun_out/pytest_parity_tanh.txt
