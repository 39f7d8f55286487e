#!/bin/bash
cd "$GRAFT_REPO_ROOT"
python scratch/gemm_one.py 65536 2560 320 geglu 2>&1 | tail -1
python scratch/gemm_one.py 65536 320 320 res 2>&1 | tail -1
python scratch/gemm_one.py 65536 320 320 plain 2>&1 | tail -1
python scratch/gemm_one.py 16384 640 640 res 2>&1 | tail -1
python scratch/gemm_one.py 65536 1280 1280 plain 2>&1 | tail -1
(timeout 250 python -m pytest tests/test_tc_gpu.py tests/test_unet_gpu.py -x -q 2>&1 | tail -3)
(timeout 200 python scratch/bench_b.py 2>&1 | tail -4)
