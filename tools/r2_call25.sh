#!/bin/bash
# Round-2 GPU call 25: e2e leg after warm-up of the copy-stream buffers, strict vs lagged loss read
mkdir -p gpurun_out
timeout 400 python bench.py --steps 8 --warmup 3 --no-parity --no-cpu-baseline --no-profile > gpurun_out/r2y_bench.json 2> gpurun_out/r2y_bench.err
echo "bench rc=$?"; python -c "import json; d=json.load(open('gpurun_out/r2y_bench.json')); print(d['value'], d['ms_per_step'], d['e2e'])" || tail -5 gpurun_out/r2y_bench.err
