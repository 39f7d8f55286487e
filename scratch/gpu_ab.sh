#!/bin/bash
cd "$GRAFT_REPO_ROOT"
for v in cur va vb; do
  if [ $v = cur ]; then unset RF_B200_LIB; else export RF_B200_LIB=$PWD/scratch/variants/$v.so; fi
  echo "== $v"
  python scratch/gemm_one.py 65536 2560 320 geglu 2>&1 | tail -1
  python scratch/gemm_one.py 65536 320 320 res 2>&1 | tail -1
  python scratch/gemm_one.py 65536 320 320 plain 2>&1 | tail -1
  python scratch/gemm_one.py 16384 640 640 plain 2>&1 | tail -1
  python scratch/gemm_one.py 65536 2560 320 plain 2>&1 | tail -1
done
unset RF_B200_LIB
timeout 600 python -m pytest tests/test_tc_gpu.py tests/test_unet_gpu.py -x -q 2>&1 | tail -2
timeout 300 python scratch/bench_b.py 2>&1 | tail -4
