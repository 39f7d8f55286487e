#!/bin/bash
# compute-sanitizer over the Griffin-Lim kernels of the final build (small batch): memcheck, then racecheck (shared memory)
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout -k 10 500 compute-sanitizer --tool memcheck --print-limit 20 python scratch/prof_gl.py 2 3 > gpurun_out/sanitizer_memcheck.txt 2>&1
tail -6 gpurun_out/sanitizer_memcheck.txt
timeout -k 10 500 compute-sanitizer --tool racecheck --racecheck-report analysis --print-limit 20 python scratch/prof_gl.py 1 2 > gpurun_out/sanitizer_racecheck.txt 2>&1
tail -12 gpurun_out/sanitizer_racecheck.txt
