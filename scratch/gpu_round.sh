#!/bin/bash
# one GPU session: tests, bench, ncu launch list, ncu full of the two FFT kernels
set -x
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/pytest_gpu.txt
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.txt 2>&1
python bench.py --steps 5 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 330 --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_istft_chunk|k_stft_pair' -s 2 -c 2 -o gpurun_out/prof_gl -f python scratch/prof_gl.py 32 2 > gpurun_out/prof_gl.log 2>&1
ls -la gpurun_out
