#!/bin/bash
# self-attention 64x64 (16 images x 8 heads x 4096^2, d = 40): time + ncu --set full with source view
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout -k 10 120 python scratch/attn_one.py 16 4096 40 2>&1 | tail -1
timeout -k 10 600 ncu --set full --clock-control none --import-source on -k regex:k_flash_attn1 -s 3 -c 1 -o gpurun_out/r02d_attn1 -f python scratch/attn_one.py 16 4096 40 > gpurun_out/ncu_attn.log 2>&1
tail -2 gpurun_out/ncu_attn.log
