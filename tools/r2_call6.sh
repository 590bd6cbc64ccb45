#!/bin/bash
# Round-2 GPU call 6 (1 GPU): BASELINE cfg3 / cfg4 / cfg5 per-GPU steps + the reference's own batch size 4, each with its roofline block
mkdir -p gpurun_out
run() { name=$1; shift; timeout 500 python bench.py --steps 5 --warmup 3 --no-parity --cpu-sample 1 "$@" --breakdown gpurun_out/r2f_breakdown_$name.json > gpurun_out/r2f_bench_$name.json 2> gpurun_out/r2f_bench_$name.err; echo "$name rc=$?"; python -c "import json; d=json.load(open('gpurun_out/r2f_bench_$name.json')); print(d['value'], d['unit'], d['ms_per_step'], d['config']['workload'][:90], d['peak_mem_gb'], d['kernel_time_shares'])" 2>/dev/null || tail -3 gpurun_out/r2f_bench_$name.err; }
run cfg3_1gpu --batch 32 --mask standard --sparsity 0.125 --lncc-weight 1.0
run cfg4_1gpu --batch 4 --coils 15 --shape 640x368
run cfg5_1gpu --batch 16 --reg Mixed --mi-weight 1.0
run bs4 --batch 4
