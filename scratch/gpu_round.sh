#!/bin/bash
# full round: tests, smoke, both bench workloads, ncu launch list + ncu --set full of the tensor-core kernels
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
(timeout 500 python -m pytest tests -m gpu -x -q 2>&1 | tail -4) > gpurun_out/pytest_gpu.txt 2>&1
(timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3) > gpurun_out/smoke.txt 2>&1
timeout 500 python bench.py --steps 2 --warmup 3 > gpurun_out/bench_clip.json 2> gpurun_out/bench_clip.err
timeout 300 python bench.py --workload gl --steps 10 --warmup 3 > gpurun_out/bench_gl.json 2> gpurun_out/bench_gl.err
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_clip.csv python bench.py --steps 1 --warmup 3 --evals 1 --clips 8 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:'k_tc_gemm|k_flash_attn1' -s 250 -c 8 -o gpurun_out/prof_unet -f python scratch/unet_one.py 8 > gpurun_out/prof_unet.log 2>&1
cat gpurun_out/pytest_gpu.txt gpurun_out/smoke.txt; cat gpurun_out/bench_clip.json | cut -c1-1500; tail -2 gpurun_out/bench_clip.err; cat gpurun_out/bench_gl.json | cut -c1-600
