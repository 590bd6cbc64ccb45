#!/bin/bash
# Round-2 GPU call 26 (1 GPU): final-tree numbers: default bench (value, e2e, roofline, parity, cpu baseline), reference arm,
# BASELINE cfg3 / cfg4 / cfg5 per-GPU steps, bs 4 eager + CUDA graph, ncu launch list of one step
mkdir -p gpurun_out
timeout 900 python bench.py --breakdown gpurun_out/r2z_breakdown_1gpu.json > gpurun_out/r2z_bench_1gpu.json 2> gpurun_out/r2z_bench_1gpu.err; echo "default bench rc=$?"; python tools/jline.py gpurun_out/r2z_bench_1gpu.json || tail -5 gpurun_out/r2z_bench_1gpu.err
timeout 600 python bench.py --impl reference > gpurun_out/r2z_bench_reference_arm.json 2> gpurun_out/r2z_bench_reference_arm.err; echo "reference arm rc=$?"; cut -c1-300 gpurun_out/r2z_bench_reference_arm.json | tail -1
run() {  # name, bench args...
  name=$1; shift
  timeout 500 python bench.py --steps 8 --warmup 3 --no-parity --no-cpu-baseline "$@" > gpurun_out/r2z_bench_$name.json 2> gpurun_out/r2z_bench_$name.err
  echo "bench $name rc=$?"; python -c "import json; d=json.load(open('gpurun_out/r2z_bench_$name.json')); print(d['value'], d['e2e']['value'], d['ms_per_step'], d.get('peak_mem_gb'), d['kernel_time_shares'])" || tail -5 gpurun_out/r2z_bench_$name.err
}
run cfg3_1gpu --batch 32 --mask standard --sparsity 0.125 --lncc-weight 1.0
run cfg4_1gpu --batch 4 --coils 15 --shape 640x368
run cfg5_1gpu --batch 16 --reg Mixed --mi-weight 1.0
run bs4_eager --batch 4 --no-profile
run bs4_graph --batch 4 --no-profile --graph
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 12000 --csv --log-file gpurun_out/r2z_launches.csv python bench.py --steps 1 --warmup 1 --min-warmup 1 --no-parity --no-cpu-baseline --no-profile > gpurun_out/r2z_ncu_bench.log 2>&1; tail -1 gpurun_out/r2z_ncu_bench.log | cut -c1-200; wc -l gpurun_out/r2z_launches.csv
