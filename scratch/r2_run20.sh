#!/bin/bash
# 2-GPU sanity of the default bench line as the driver launches it
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout -k 10 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 2 --warmup 3 > gpurun_out/bench_clip_2gpu.json 2> gpurun_out/bench_clip_2gpu.err
cut -c1-400 gpurun_out/bench_clip_2gpu.json; tail -3 gpurun_out/bench_clip_2gpu.err
