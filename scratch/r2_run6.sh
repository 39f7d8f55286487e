#!/bin/bash
# round 2, run 6: full GPU suite, ncu evidence (launch list of one clip step, --set full of the tensor-core / attention kernels,
# DRAM bytes of every k_tc_gemm launch of one evaluation)
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
(timeout -k 10 1800 python -m pytest tests -m gpu -q 2>&1 | tail -15) > gpurun_out/pytest_gpu.txt 2>&1
cat gpurun_out/pytest_gpu.txt
timeout -k 10 900 ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/r02_eval32_launches_dram.csv python scratch/prof_eval.py 32 ncu > gpurun_out/ncu_eval32.log 2>&1
timeout -k 10 900 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:'k_tc_gemm|k_flash_attn1|k_attn_short' -c 36 -o gpurun_out/r02_unet_full -f python scratch/prof_eval.py 32 ncu > gpurun_out/ncu_full.log 2>&1
tail -3 gpurun_out/ncu_full.log
timeout -k 10 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/r02_clip_launches.csv python bench.py --steps 1 --warmup 3 --evals 1 --clips 32 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
ls -la gpurun_out | tail -8
# Griffin-Lim after the 7-thread radix-49 pass and the edge / half-rate kernel split
timeout -k 10 600 python bench.py --workload gl --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_gl.json 2> gpurun_out/bench_gl.err
cut -c1-1800 gpurun_out/bench_gl.json; tail -2 gpurun_out/bench_gl.err
timeout -k 10 600 ncu --set full --clock-control none --import-source on -k regex:'k_istft_half|k_stft_half|k_istft_edge|k_stft_edge' -s 8 -c 4 -o gpurun_out/r02_gl -f python scratch/prof_gl.py 64 4 > gpurun_out/ncu_gl.log 2>&1
tail -2 gpurun_out/ncu_gl.log
