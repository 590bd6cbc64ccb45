#!/bin/bash
# Round-2 GPU call 38: lean epilogue stores in conv_tc_kernel (hoisted addresses, bias branch outside the channel loop,
# predicated instead of branched stores): parity, microbench, bench
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_tc.py tests/test_gpu_models.py tests/test_gpu_gan.py -m gpu -x -q -p no:cacheprovider > gpurun_out/r2epi_tests.log 2>&1; echo "tc+model+gan tests rc=$?"; tail -3 gpurun_out/r2epi_tests.log | cut -c1-300
timeout 300 python tools/bench_tc.py 64 > gpurun_out/r2epi_bench_tc.txt 2>&1; cut -c1-40,96-170 gpurun_out/r2epi_bench_tc.txt | head -30
timeout 400 python bench.py --steps 8 --warmup 3 --no-parity --no-cpu-baseline --no-profile > gpurun_out/r2epi_bench.json 2> gpurun_out/r2epi_bench.err
echo "bench rc=$?"; python -c "import json; d=json.load(open('gpurun_out/r2epi_bench.json')); print(d['value'], d['e2e']['value'], d['ms_per_step'], d['clocks'])" || tail -3 gpurun_out/r2epi_bench.err
