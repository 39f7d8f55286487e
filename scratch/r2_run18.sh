#!/bin/bash
# single-request regime (2 images): per-launch table of the tensor-core kernels and the launch list of one evaluation
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout -k 10 300 python scratch/prof_b.py 1 > gpurun_out/tc_b1.txt 2>&1
head -60 gpurun_out/tc_b1.txt
timeout -k 10 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/eval1_launches.csv python scratch/prof_eval.py 1 ncu > gpurun_out/eval1_ncu.log 2>&1
python scratch/agg_launches.py gpurun_out/eval1_launches.csv | head -30
