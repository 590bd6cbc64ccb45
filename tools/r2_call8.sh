#!/bin/bash
# Round-2 GPU call 8: weight gradient with the filter rows in the MMA N dimension: correctness, A/B microbench, bench; cfg5 step
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_tc.py -m gpu -x -q -p no:cacheprovider > gpurun_out/r2h_tc_tests.log 2>&1; echo "tc tests rc=$?"; tail -3 gpurun_out/r2h_tc_tests.log | cut -c1-300
for r in 1 0; do
  SAN_WG_ROWN=$r timeout 200 python tools/bench_tc.py 64 "3,18,320,3;18,18,320,3;36,18,320,3;18,36,160,3;36,36,160,3;72,36,160,3;36,72,80,3;2,32,320,3;32,32,320,3;64,64,160,3" > gpurun_out/r2h_bench_tc_rown$r.txt 2>&1
  echo "--- bench_tc ROWN=$r"; cut -c1-40,112-160 gpurun_out/r2h_bench_tc_rown$r.txt
done
timeout 400 python bench.py --steps 10 --warmup 3 --no-parity --breakdown gpurun_out/r2h_breakdown.json > gpurun_out/r2h_bench.json 2> gpurun_out/r2h_bench.err
echo "bench rc=$?"; python -c "import json; d=json.load(open('gpurun_out/r2h_bench.json')); print(d['value'], d['e2e']['value'], d['roofline']['frac'], d['kernel_time_shares'])"; tail -2 gpurun_out/r2h_bench.err
timeout 500 python bench.py --steps 5 --warmup 3 --no-parity --cpu-sample 2 --batch 16 --reg Mixed --mi-weight 1.0 --breakdown gpurun_out/r2h_breakdown_cfg5.json > gpurun_out/r2h_bench_cfg5_1gpu.json 2> gpurun_out/r2h_bench_cfg5.err
echo "cfg5 rc=$?"; python -c "import json; d=json.load(open('gpurun_out/r2h_bench_cfg5_1gpu.json')); print(d['value'], d['ms_per_step'], d['peak_mem_gb'], d['kernel_time_shares'], d['cpu_baseline'])" || tail -3 gpurun_out/r2h_bench_cfg5.err
timeout 300 python -m pytest tests/test_gpu_models.py tests/test_gpu_gan.py -m gpu -x -q -p no:cacheprovider > gpurun_out/r2h_model_tests.log 2>&1; echo "model+gan tests rc=$?"; tail -3 gpurun_out/r2h_model_tests.log | cut -c1-300
