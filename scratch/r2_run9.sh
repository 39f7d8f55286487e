#!/bin/bash
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout -k 5 120 scratch/mma_bench.bin > gpurun_out/mma_bench.txt 2>&1
cat gpurun_out/mma_bench.txt
timeout -k 10 600 ncu --set full --clock-control none --import-source on -k regex:'k_istft_half|k_stft_half' -s 2 -c 2 -o gpurun_out/r02_gl -f python scratch/prof_gl.py 64 4 > gpurun_out/ncu_gl.log 2>&1
tail -2 gpurun_out/ncu_gl.log
timeout -k 10 300 python bench.py --workload gl --steps 10 --warmup 3 > gpurun_out/bench_gl.json 2> gpurun_out/bench_gl.err
cut -c1-300 gpurun_out/bench_gl.json
