#!/bin/bash
# full GPU round without ncu: tests, smoke, GL bench, clip bench
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q -s 2>&1 | grep -v "^$" | tail -25 > gpurun_out/pytest_gpu.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.txt 2>&1
timeout 600 python bench.py --workload gl --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_gl.json 2> gpurun_out/bench_gl.err
timeout 900 python bench.py --steps 2 --warmup 3 > gpurun_out/bench_clip.json 2> gpurun_out/bench_clip.err
