#!/bin/bash
# ncu launch lists of the final build (cold-cache, serialised: compare shares, not absolute times)
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout -k 10 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/r02d_clip_launches.csv python bench.py --steps 1 --warmup 3 --evals 1 --clips 32 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
python scratch/agg_launches.py gpurun_out/r02d_clip_launches.csv 0.75 | head -24
timeout -k 10 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 2000 --csv --log-file gpurun_out/r02d_gl_launches.csv python scratch/prof_gl.py 64 32 > gpurun_out/gl_under_ncu.log 2>&1
python scratch/agg_launches.py gpurun_out/r02d_gl_launches.csv | head -16
