#!/bin/bash
# Round-2 GPU call 24: InstanceNorm backward with two CTAs of a cluster per big plane (DSMEM exchange of the partial sums):
# parity + microbench + bench A/B; host prefetcher test; e2e with the double-buffered input pipeline
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_tc.py tests/test_gpu_models.py -m gpu -x -q -p no:cacheprovider -k "in_bwd or prefetcher or varnet or rec_step" > gpurun_out/r2x_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/r2x_tests.log | cut -c1-400
for c in 1 0; do echo "--- SAN_IN_BWD_CL2=$c"; SAN_IN_BWD_CL2=$c timeout 120 python tools/bench_norm.py 64 2>&1 | tee gpurun_out/r2x_bench_norm_cl2_$c.txt | cut -c1-60; done
run() {  # name, env...
  name=$1; shift
  env "$@" timeout 400 python bench.py --steps 8 --warmup 3 --no-parity --no-cpu-baseline --breakdown gpurun_out/r2x_breakdown_$name.json > gpurun_out/r2x_bench_$name.json 2> gpurun_out/r2x_bench_$name.err
  echo "bench $name rc=$?"; python -c "import json; d=json.load(open('gpurun_out/r2x_bench_$name.json')); b=json.load(open('gpurun_out/r2x_breakdown_$name.json')); print(d['value'], d['e2e']['value'], d['ms_per_step'], d['e2e']['ms_per_step'], d.get('peak_mem_gb'), b['ops']['in_bwd_fused_map'])" || tail -5 gpurun_out/r2x_bench_$name.err
}
run cl2_1 SAN_IN_BWD_CL2=1
run cl2_0 SAN_IN_BWD_CL2=0
