#!/bin/bash
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_tc_gpu.py tests/test_unet_gpu.py -x -q 2>&1 | tail -5
python scratch/gemm_one.py 65536 2560 320 geglu 2>&1 | tail -1
python scratch/gemm_one.py 65536 320 320 res 2>&1 | tail -1
python scratch/gemm_one.py 65536 320 320 plain 2>&1 | tail -1
python scratch/gemm_one.py 16384 640 640 plain 2>&1 | tail -1
timeout 300 python scratch/bench_b.py 2>&1 | tail -5
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/unet_launches_b8.csv python scratch/unet_one.py 8 > gpurun_out/unet_one.log 2>&1
tail -1 gpurun_out/unet_one.log
