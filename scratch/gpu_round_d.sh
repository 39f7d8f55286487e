#!/bin/bash
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q -s 2>&1 | grep -v "^$" | tail -25 > gpurun_out/pytest_gpu.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/unet_launches_b8.csv python scratch/unet_one.py 8 > gpurun_out/unet_one.log 2>&1
tail -12 gpurun_out/pytest_gpu.txt; tail -2 gpurun_out/unet_one.log
