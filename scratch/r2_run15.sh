#!/bin/bash
# A/B builds of the Griffin-Lim kernels (scratch/variants): --workload gl on each
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
for v in main cg nopf stft4 stft4cg; do
  if [ $v = main ]; then unset RF_B200_LIB; else export RF_B200_LIB=$PWD/scratch/variants/librf_$v.so; fi
  timeout -k 10 200 python bench.py --workload gl --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/gl_$v.json 2> gpurun_out/gl_$v.err
  python - <<P
import json
d=json.load(open("gpurun_out/gl_$v.json"))
r=d["roofline"]
print("$v", round(d["value"],1), "e2e", round(d["e2e"]["value"],1), r["per_kernel_ms_per_step"], "loop", round(r["loop"]["frac"],3))
P
done
