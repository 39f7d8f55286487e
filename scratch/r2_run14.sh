#!/bin/bash
# GL kernels with per-bin constant tables, batch loads with constant offsets, TMA bulk staging: parity + bench + ncu
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout -k 10 900 python -m pytest tests/test_audio_gpu.py tests/test_cli_gpu.py -q -x > gpurun_out/pytest_audio.txt 2>&1
echo "rc=$?" >> gpurun_out/pytest_audio.txt
tail -6 gpurun_out/pytest_audio.txt
timeout -k 10 300 python bench.py --workload gl --steps 10 --warmup 3 > gpurun_out/bench_gl.json 2> gpurun_out/bench_gl.err
cut -c1-1800 gpurun_out/bench_gl.json; tail -2 gpurun_out/bench_gl.err
timeout -k 10 600 ncu --set full --clock-control none --import-source on -k regex:'k_istft_half|k_stft_half' -s 2 -c 2 -o gpurun_out/r02b_gl -f python scratch/prof_gl.py 64 4 > gpurun_out/ncu_gl.log 2>&1
tail -2 gpurun_out/ncu_gl.log
