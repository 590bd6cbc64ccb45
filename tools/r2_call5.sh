#!/bin/bash
# Round-2 GPU call 5: full GPU suite on the fp16-pair path (fused absmax, packed conversions), bench + breakdown, ncu of the DXN conv form
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -p no:cacheprovider > gpurun_out/r2e_tests.log 2>&1; echo "gpu tests rc=$?"; tail -4 gpurun_out/r2e_tests.log | cut -c1-300
SAN_TC_DXN=1 timeout 400 python -m pytest tests/test_gpu_tc.py -m gpu -x -q -p no:cacheprovider > gpurun_out/r2e_tc_tests_dxn1.log 2>&1; echo "tc tests (dxn on) rc=$?"; tail -3 gpurun_out/r2e_tc_tests_dxn1.log | cut -c1-300
timeout 400 python bench.py --steps 10 --warmup 3 --breakdown gpurun_out/r2e_breakdown.json > gpurun_out/r2e_bench.json 2> gpurun_out/r2e_bench.err
echo "bench rc=$?"; python -c "import json; d=json.load(open('gpurun_out/r2e_bench.json')); print(d['value'], d['e2e']['value'], d['roofline']['frac'], d['kernel_time_shares'], d['parity']['forward_rel_l2_vs_fp32'], d['parity']['grad_all_params_vs_fp64'], d['parity']['grad_cpu_fp32_oracle_vs_fp64'])"; tail -2 gpurun_out/r2e_bench.err
SAN_TC_DXN=1 timeout 300 ncu --set full --clock-control none --import-source on -k regex:conv_tc_kernel -c 1 -f -o gpurun_out/r2e_dxn python tools/bench_tc.py 64 "18,18,320,3" > gpurun_out/r2e_ncu_dxn.log 2>&1; tail -2 gpurun_out/r2e_ncu_dxn.log
