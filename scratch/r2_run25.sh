#!/bin/bash
# last check after the inverse-mel change: benchmarked-path parity tests + smoke
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
(timeout -k 10 900 python -m pytest tests/test_parity_bench_gpu.py tests/test_cli_gpu.py tests/test_audio_gpu.py -q -x 2>&1 | tail -4) > gpurun_out/pytest_last.txt 2>&1
cat gpurun_out/pytest_last.txt
(timeout -k 10 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2)
