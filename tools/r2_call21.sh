#!/bin/bash
# Round-2 GPU call 21: whole step as one CUDA graph with VarNet on 1 / 2 / 4 sub-batch streams (no host launch cost: the
# graph's parallel branches give the GPU the concurrency that eager launches cannot deliver at ~30 us of host time each)
mkdir -p gpurun_out
run() {  # name, env...
  name=$1; shift
  env "$@" timeout 500 python bench.py --graph --steps 8 --warmup 3 --no-parity --no-cpu-baseline --no-profile > gpurun_out/r2u_bench_$name.json 2> gpurun_out/r2u_bench_$name.err
  echo "bench $name rc=$?"; python -c "import json; d=json.load(open('gpurun_out/r2u_bench_$name.json')); print(d['value'], d['e2e']['value'], d['ms_per_step'], d.get('peak_mem_gb'), d.get('gpu_launches'))" || tail -5 gpurun_out/r2u_bench_$name.err
}
run g_s1 SAN_VARNET_STREAMS=1
run g_s2 SAN_VARNET_STREAMS=2
run g_s4 SAN_VARNET_STREAMS=4
run g_s2_nowg SAN_VARNET_STREAMS=2 SAN_WG_OVERLAP=0
SAN_VARNET_STREAMS=2 timeout 300 python -m pytest tests/test_gpu_models.py -m gpu -x -q -p no:cacheprovider -k graphed > gpurun_out/r2u_graph_test.log 2>&1; echo "graph test rc=$?"; tail -3 gpurun_out/r2u_graph_test.log | cut -c1-300
