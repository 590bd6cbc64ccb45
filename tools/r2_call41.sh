#!/bin/bash
# Round-2 GPU call 41: sanity of the clean-rebuilt library: tc + ops tests, smoke()
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_tc.py tests/test_gpu_ops.py -m gpu -x -q -p no:cacheprovider > gpurun_out/r2last_tests.log 2>&1; echo "tc+ops tests rc=$?"; tail -2 gpurun_out/r2last_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1 | cut -c1-200
