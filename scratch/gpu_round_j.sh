#!/bin/bash
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
(timeout 250 python -m pytest tests/test_tc_gpu.py tests/test_unet_gpu.py -x -q 2>&1 | tail -3)
(timeout 200 python scratch/bench_b.py 2>&1 | tail -4)
(timeout 200 python scratch/prof_b.py 8 2>&1 | tail -26)
