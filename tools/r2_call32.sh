#!/bin/bash
# Round-2 GPU call 32: row-ring conv with register prefetch: parity + microbench
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_tc.py -m gpu -x -q -p no:cacheprovider -k "conv_rows" > gpurun_out/r2ff_tests.log 2>&1; echo "rows tests rc=$?"; tail -5 gpurun_out/r2ff_tests.log | cut -c1-300
timeout 300 python tools/bench_tc.py 64 "18,18,320,3;3,18,320,3;18,3,320,3;8,8,320,3;2,8,320,3;16,16,160,3;18,18,160,3" > gpurun_out/r2ff_bench_tc_rows.txt 2>&1; cut -c1-40,96- gpurun_out/r2ff_bench_tc_rows.txt
