#!/bin/bash
# final validation of the round: full GPU suite, smoke, every bench workload
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
(timeout -k 10 1800 python -m pytest tests -m gpu -q 2>&1 | tail -6) > gpurun_out/pytest_gpu.txt 2>&1
cat gpurun_out/pytest_gpu.txt
(timeout -k 10 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2) > gpurun_out/smoke.txt 2>&1
cat gpurun_out/smoke.txt
timeout -k 10 900 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_clip.json 2> gpurun_out/bench_clip.err
cut -c1-300 gpurun_out/bench_clip.json; tail -2 gpurun_out/bench_clip.err
timeout -k 10 600 python bench.py --workload riffuse --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_riffuse38.json 2> gpurun_out/bench_riffuse38.err
cut -c1-250 gpurun_out/bench_riffuse38.json
timeout -k 10 600 python bench.py --workload riffuse --denoising 1.0 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_riffuse50.json 2> gpurun_out/bench_riffuse50.err
cut -c1-250 gpurun_out/bench_riffuse50.json
timeout -k 10 900 python bench.py --workload roundtrip --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_roundtrip.json 2> gpurun_out/bench_roundtrip.err
cut -c1-250 gpurun_out/bench_roundtrip.json
