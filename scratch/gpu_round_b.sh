#!/bin/bash
set -x
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 > gpurun_out/pytest_gpu.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.txt 2>&1
timeout 900 python bench.py --steps 2 --warmup 3 > gpurun_out/bench_clip.json 2> gpurun_out/bench_clip.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file gpurun_out/launches_clip.csv python bench.py --steps 1 --warmup 3 --evals 1 --clips 8 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_tc_gemm|k_flash_attn' -s 60 -c 12 -o gpurun_out/prof_unet -f python scratch/prof_b.py 4 > gpurun_out/prof_unet.log 2>&1
ls -la gpurun_out
