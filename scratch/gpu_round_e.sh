#!/bin/bash
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_tc_gpu.py tests/test_unet_gpu.py -x -q 2>&1 | tail -5
timeout 300 python scratch/bench_b.py 2>&1 | tail -6
timeout 300 python scratch/prof_b.py 8 2>&1 | tail -28
RF_GEMM_SPLITK=0 timeout 300 python scratch/bench_b.py 2>&1 | tail -4
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/unet_launches_b8.csv python scratch/unet_one.py 8 > gpurun_out/unet_one.log 2>&1
tail -2 gpurun_out/unet_one.log
