#!/bin/bash
# Round-2 GPU call 29: InstanceNorm backward with two float4 groups in flight per thread: parity + microbench + bench
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_tc.py tests/test_gpu_models.py -m gpu -x -q -p no:cacheprovider -k "in_bwd or varnet or rec_step or fused_unet" > gpurun_out/r2cc_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/r2cc_tests.log | cut -c1-400
timeout 120 python tools/bench_norm.py 64 2>&1 | tee gpurun_out/r2cc_bench_norm.txt | cut -c1-60
timeout 400 python bench.py --steps 8 --warmup 3 --no-parity --no-cpu-baseline --breakdown gpurun_out/r2cc_breakdown.json > gpurun_out/r2cc_bench.json 2> gpurun_out/r2cc_bench.err
echo "bench rc=$?"; python -c "import json; d=json.load(open('gpurun_out/r2cc_bench.json')); b=json.load(open('gpurun_out/r2cc_breakdown.json')); print(d['value'], d['e2e']['value'], d['ms_per_step'], b['ops']['in_bwd_fused_map'])" || tail -5 gpurun_out/r2cc_bench.err
