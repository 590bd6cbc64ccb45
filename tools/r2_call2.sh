#!/bin/bash
# Round-2 GPU call 2: whole GPU suite (FFT v2 now default), headline-config parity reports, default bench with the parity block.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -p no:cacheprovider --deselect tests/test_gpu_parity_full.py > gpurun_out/r2b_tests.log 2>&1; echo "gpu tests rc=$?"; tail -3 gpurun_out/r2b_tests.log
timeout 900 python -m pytest tests/test_gpu_parity_full.py -m gpu -q -p no:cacheprovider -s > gpurun_out/r2b_parity.log 2>&1; echo "parity rc=$?"; tail -5 gpurun_out/r2b_parity.log | cut -c1-400
timeout 600 python bench.py --steps 10 --warmup 3 --breakdown gpurun_out/r2b_breakdown.json > gpurun_out/r2b_bench.json 2> gpurun_out/r2b_bench.err; echo "bench rc=$?"; cut -c1-1500 gpurun_out/r2b_bench.json; tail -3 gpurun_out/r2b_bench.err
