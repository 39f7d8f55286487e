#!/bin/bash
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout -k 10 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 1 --warmup 3 > gpurun_out/bench_clip_2gpu.json 2> gpurun_out/bench_clip_2gpu.err
cut -c1-400 gpurun_out/bench_clip_2gpu.json; tail -3 gpurun_out/bench_clip_2gpu.err
timeout -k 10 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 1 --warmup 0 > gpurun_out/bench_ref_2gpu.json 2> gpurun_out/bench_ref_2gpu.err
cut -c1-900 gpurun_out/bench_ref_2gpu.json; tail -2 gpurun_out/bench_ref_2gpu.err
timeout -k 10 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --workload riffuse --gpus 2 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_riffuse_2gpu.json 2> gpurun_out/bench_riffuse_2gpu.err
cut -c1-300 gpurun_out/bench_riffuse_2gpu.json; tail -2 gpurun_out/bench_riffuse_2gpu.err
