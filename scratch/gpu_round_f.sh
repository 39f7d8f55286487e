#!/bin/bash
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
python scratch/attn_one.py 16 4096 40 2>&1 | tail -1
python scratch/gemm_one.py 65536 2560 320 geglu 2>&1 | tail -1
python scratch/gemm_one.py 65536 320 320 res 2>&1 | tail -1
python scratch/gemm_one.py 65536 320 320 plain 2>&1 | tail -1
python scratch/gemm_one.py 16384 640 640 plain 2>&1 | tail -1
python scratch/gemm_one.py 65536 1280 1280 plain 2>&1 | tail -1
timeout 600 ncu --set full --import-source on --clock-control none -k regex:k_flash_attn1 -s 3 -c 1 -o gpurun_out/attn1 -f python scratch/attn_one.py 16 4096 40 > gpurun_out/ncu_attn.log 2>&1
timeout 600 ncu --set full --import-source on --clock-control none -k regex:k_tc_gemm -s 4 -c 1 -o gpurun_out/gemm_geglu -f python scratch/gemm_one.py 65536 2560 320 geglu > gpurun_out/ncu_gemm1.log 2>&1
timeout 600 ncu --set full --import-source on --clock-control none -k regex:k_tc_gemm -s 4 -c 1 -o gpurun_out/gemm_k320 -f python scratch/gemm_one.py 65536 320 320 plain > gpurun_out/ncu_gemm2.log 2>&1
timeout 600 ncu --set full --import-source on --clock-control none -k regex:'k_istft_dec|k_stft_dec' -s 4 -c 2 -o gpurun_out/gl_dec -f python scratch/prof_gl.py 64 4 > gpurun_out/ncu_gl.log 2>&1
ls -la gpurun_out/*.ncu-rep
