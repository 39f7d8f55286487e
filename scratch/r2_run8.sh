#!/bin/bash
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
for v in D E F B; do
  RF_B200_LIB=$PWD/scratch/variants/librf_$v.so timeout -k 10 300 python bench.py --workload gl --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_gl_$v.json 2> gpurun_out/bench_gl_$v.err
  python - <<PY
import json
d=json.loads(open("gpurun_out/bench_gl_$v.json").read().strip().splitlines()[-1])
print("GL variant $v:", round(d["value"],1), "clips/s", {k: round(x,2) for k,x in d["roofline"]["per_kernel_ms_per_step"].items()}, "loop frac", round(d["roofline"]["loop"]["frac"],3))
PY
done
timeout -k 10 600 python -m pytest tests/test_audio_gpu.py -q 2>&1 | tail -3
timeout -k 10 600 ncu --set full --clock-control none --import-source on -k regex:'k_istft_half|k_stft_half' -s 8 -c 2 -o gpurun_out/r02_gl -f python scratch/prof_gl.py 64 4 > gpurun_out/ncu_gl.log 2>&1
tail -2 gpurun_out/ncu_gl.log
