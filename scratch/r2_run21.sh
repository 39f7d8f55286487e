#!/bin/bash
# deferred row maximum in the single-pass attention kernel: parity tests and timings with RF_ATTN_DEFER_MAX = 1 / 0
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
for dm in 1 0; do
  echo "== RF_ATTN_DEFER_MAX=$dm"
  RF_ATTN_DEFER_MAX=$dm timeout -k 10 600 python -m pytest tests/test_tc_gpu.py -q -x -k "attention" 2>&1 | tail -2
  RF_ATTN_DEFER_MAX=$dm timeout -k 10 120 python scratch/attn_one.py 16 4096 40 2>&1 | tail -1
  RF_ATTN_DEFER_MAX=$dm timeout -k 10 120 python scratch/attn_one.py 16 1024 80 2>&1 | tail -1
  RF_ATTN_DEFER_MAX=$dm timeout -k 10 300 python scratch/prof_eval.py 32 2>&1 | tail -1
done 2>&1 | tee gpurun_out/attn_defer_ab.txt
