#!/bin/bash
# Round-2 GPU call 19: full GPU suite on the current tree (statistics epilogue with fp64 flush, TMA FFT, staged-weight cache,
# wgrad overlap); A/B: staged-weight cache, pre-staging, in_bwd CTA size under the overlap
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q -p no:cacheprovider > gpurun_out/r2s_gpu_tests.log 2>&1; echo "gpu suite rc=$?"; tail -5 gpurun_out/r2s_gpu_tests.log | cut -c1-400
run() {  # name, env...
  name=$1; shift
  env "$@" timeout 400 python bench.py --steps 8 --warmup 3 --no-parity --no-cpu-baseline --breakdown gpurun_out/r2s_breakdown_$name.json > gpurun_out/r2s_bench_$name.json 2> gpurun_out/r2s_bench_$name.err
  echo "bench $name rc=$?"; python -c "import json; d=json.load(open('gpurun_out/r2s_bench_$name.json')); print(d['value'], d['e2e']['value'], d['ms_per_step'], d['kernel_time_shares'], d['roofline']['frac'], d['roofline_fft_dc'].get('soft_dc_launches'))" || tail -3 gpurun_out/r2s_bench_$name.err
}
run ws1
run ws0 SAN_WS_CACHE=0
run prestage0 SAN_WG_PRESTAGE=0
run inbwd512 SAN_IN_BWD_NT=512
run inbwd256 SAN_IN_BWD_NT=256
