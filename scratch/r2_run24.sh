#!/bin/bash
# compute-sanitizer memcheck over smoke(): small UNet evaluation on the tcgen05 kernels + the audio paths
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout -k 10 800 compute-sanitizer --tool memcheck --print-limit 20 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/sanitizer_smoke_memcheck.txt 2>&1
tail -8 gpurun_out/sanitizer_smoke_memcheck.txt
