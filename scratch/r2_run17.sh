#!/bin/bash
# validation of the round's final build: full GPU suite, smoke, every bench workload, ncu of the Griffin-Lim kernels
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
(timeout -k 10 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -6) > gpurun_out/pytest_gpu.txt 2>&1
cat gpurun_out/pytest_gpu.txt
(timeout -k 10 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2) > gpurun_out/smoke.txt 2>&1
cat gpurun_out/smoke.txt
timeout -k 10 900 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_clip.json 2> gpurun_out/bench_clip.err
cut -c1-300 gpurun_out/bench_clip.json; tail -2 gpurun_out/bench_clip.err
timeout -k 10 300 python bench.py --workload gl --steps 10 --warmup 3 > gpurun_out/bench_gl.json 2> gpurun_out/bench_gl.err
cut -c1-200 gpurun_out/bench_gl.json
timeout -k 10 600 python bench.py --workload riffuse --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_riffuse38.json 2> gpurun_out/bench_riffuse38.err
cut -c1-250 gpurun_out/bench_riffuse38.json
timeout -k 10 900 python bench.py --workload roundtrip --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_roundtrip.json 2> gpurun_out/bench_roundtrip.err
cut -c1-250 gpurun_out/bench_roundtrip.json
timeout -k 10 600 ncu --set full --clock-control none --import-source on -k regex:'k_istft_half|k_stft_half|k_stft_edge|k_ola_assemble' -s 4 -c 5 -o gpurun_out/r02d_gl -f python scratch/prof_gl.py 64 4 > gpurun_out/ncu_gl.log 2>&1
tail -2 gpurun_out/ncu_gl.log
