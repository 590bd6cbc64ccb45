#!/bin/bash
# Round-2 GPU call 40: final validation after the lean-epilogue change: full GPU suite, smoke(), default bench
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q -p no:cacheprovider > gpurun_out/r2fin2_gpu_tests.log 2>&1; echo "gpu suite rc=$?"; tail -4 gpurun_out/r2fin2_gpu_tests.log | cut -c1-300
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2fin2_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/r2fin2_smoke.log | cut -c1-300
timeout 900 python bench.py --breakdown gpurun_out/r2fin2_breakdown_1gpu.json > gpurun_out/r2fin2_bench_1gpu.json 2> gpurun_out/r2fin2_bench_1gpu.err; echo "bench rc=$?"; python tools/jline.py gpurun_out/r2fin2_bench_1gpu.json || tail -5 gpurun_out/r2fin2_bench_1gpu.err
