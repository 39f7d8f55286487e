#!/bin/bash
# PDL A/B: evaluation time at batch 32 and 1 with RF_PDL = 1 (small launches only, the default), 0 (never), 2 (every
# instrumented launch), then the tensor-core / UNet parity tests
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
for b in 32 1; do
  for pdl in 1 0 2; do
    echo "== B=$b RF_PDL=$pdl"; RF_PDL=$pdl timeout -k 10 300 python scratch/prof_eval.py $b 2>&1 | tail -1
  done
done 2>&1 | tee gpurun_out/pdl_ab.txt
timeout -k 10 1200 python -m pytest tests/test_tc_gpu.py tests/test_unet_gpu.py tests/test_clip_gpu.py -q -x > gpurun_out/pytest_tc.txt 2>&1
echo "rc=$?" >> gpurun_out/pytest_tc.txt
tail -5 gpurun_out/pytest_tc.txt
