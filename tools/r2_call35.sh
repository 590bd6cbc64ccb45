#!/bin/bash
# Round-2 GPU call 35: full GPU suite after the graph-test bound fix
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q -p no:cacheprovider > gpurun_out/r2fin_gpu_tests.log 2>&1; echo "gpu suite rc=$?"; tail -4 gpurun_out/r2fin_gpu_tests.log | cut -c1-300
timeout 300 python -m pytest tests/test_gpu_models.py -m gpu -q -p no:cacheprovider -k graphed --count 1 > /dev/null 2>&1
for i in 1 2 3; do timeout 200 python -m pytest tests/test_gpu_models.py -m gpu -q -p no:cacheprovider -k graphed 2>&1 | tail -1; done
