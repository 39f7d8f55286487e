#!/bin/bash
# round 2, run 7: Griffin-Lim A/B builds, attention FMA-exponential A/B, ncu evidence (kept under 64 MiB)
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
for v in A B C; do
  RF_B200_LIB=$PWD/scratch/variants/librf_$v.so timeout -k 10 300 python bench.py --workload gl --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_gl_$v.json 2> gpurun_out/bench_gl_$v.err
  python - <<PY
import json
d=json.loads(open("gpurun_out/bench_gl_$v.json").read().strip().splitlines()[-1])
print("GL variant $v:", round(d["value"],1), "clips/s", {k: round(x,2) for k,x in d["roofline"]["per_kernel_ms_per_step"].items()}, "loop frac", round(d["roofline"]["loop"]["frac"],3))
PY
done
timeout -k 10 600 python -m pytest tests/test_tc_gpu.py -q -k "attention" 2>&1 | tail -3
timeout -k 10 300 python scratch/prof_eval.py 32 > gpurun_out/eval32_poly.txt 2>&1
RF_ATTN_POLY=0 timeout -k 10 300 python scratch/prof_eval.py 32 > gpurun_out/eval32_nopoly.txt 2>&1
cat gpurun_out/eval32_poly.txt gpurun_out/eval32_nopoly.txt
timeout -k 10 900 ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/r02_eval32_launches_dram.csv python scratch/prof_eval.py 32 ncu > gpurun_out/ncu_eval32.log 2>&1
RF_ATTN_POLY=0 timeout -k 10 600 ncu --profile-from-start off --metrics gpu__time_duration.sum -k regex:k_flash_attn1 --clock-control none --csv --log-file gpurun_out/r02_attn_nopoly.csv python scratch/prof_eval.py 32 ncu > /dev/null 2>&1
timeout -k 10 900 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:'k_tc_gemm|k_flash_attn1' -s 3 -c 14 -o gpurun_out/r02_unet_full -f python scratch/prof_eval.py 32 ncu > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log
timeout -k 10 600 ncu --set full --clock-control none --import-source on -k regex:'k_istft_half|k_stft_half' -s 8 -c 2 -o gpurun_out/r02_gl -f python scratch/prof_gl.py 64 4 > gpurun_out/ncu_gl.log 2>&1
tail -2 gpurun_out/ncu_gl.log
timeout -k 10 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/r02_clip_launches.csv python bench.py --steps 1 --warmup 3 --evals 1 --clips 32 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
du -sh gpurun_out; ls -la gpurun_out | tail -12
