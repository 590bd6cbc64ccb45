#!/bin/bash
# Round-2 GPU call 30: row-ring conv (fused staging): op-level parity, then microbench
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_tc.py -m gpu -x -q -p no:cacheprovider -k "conv_rows" > gpurun_out/r2dd_tests.log 2>&1; echo "rows tests rc=$?"; tail -25 gpurun_out/r2dd_tests.log | cut -c1-300
