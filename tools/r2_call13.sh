#!/bin/bash
# Round-2 GPU call 13: staged tensors without all-zero channel groups (ceil(C/8) planes): correctness + A/B numbers
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_tc.py -m gpu -x -q -p no:cacheprovider > gpurun_out/r2m_tc_tests.log 2>&1; echo "tc tests rc=$?"; tail -3 gpurun_out/r2m_tc_tests.log | cut -c1-300
SAN_TC_PAIR=0 timeout 400 python -m pytest tests/test_gpu_tc.py -m gpu -x -q -p no:cacheprovider > gpurun_out/r2m_tc_tests_pair0.log 2>&1; echo "tc tests PAIR=0 rc=$?"; tail -3 gpurun_out/r2m_tc_tests_pair0.log | cut -c1-300
timeout 600 python -m pytest tests/test_gpu_models.py tests/test_gpu_gan.py tests/test_gpu_ops.py tests/test_gpu_augment.py -m gpu -x -q -p no:cacheprovider > gpurun_out/r2m_model_tests.log 2>&1; echo "model+gan+ops tests rc=$?"; tail -3 gpurun_out/r2m_model_tests.log | cut -c1-300
timeout 600 python -m pytest tests/test_gpu_parity_full.py -m gpu -q -p no:cacheprovider -s -k cfg2 > gpurun_out/r2m_parity.log 2>&1; echo "parity rc=$?"; grep -o '"forward": {"img_rec": {[^}]*}' gpurun_out/r2m_parity.log; grep -o '"all_concatenated": {[^}]*}' gpurun_out/r2m_parity.log
timeout 300 python tools/bench_tc.py 64 > gpurun_out/r2m_bench_tc.txt 2>&1; cut -c1-40,76-165 gpurun_out/r2m_bench_tc.txt
timeout 400 python bench.py --steps 10 --warmup 3 --no-parity --breakdown gpurun_out/r2m_breakdown.json > gpurun_out/r2m_bench.json 2> gpurun_out/r2m_bench.err
echo "bench rc=$?"; python -c "import json; d=json.load(open('gpurun_out/r2m_bench.json')); print(d['value'], d['e2e']['value'], d['roofline']['frac'], d['peak_mem_gb'], d['kernel_time_shares'])" || tail -3 gpurun_out/r2m_bench.err
