#!/bin/bash
# Round-2 GPU call 16: batched read-only epilogue loads in the FFT column pass (soft-DC / coil-reduce): parity + microbench;
# weight-gradient chunk size under the side-stream overlap
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_ops.py -m gpu -x -q -p no:cacheprovider > gpurun_out/r2p_ops_tests.log 2>&1; echo "ops tests rc=$?"; tail -3 gpurun_out/r2p_ops_tests.log | cut -c1-300
timeout 200 python tools/bench_fft.py 64 20 > gpurun_out/r2p_bench_fft.txt 2>&1; cat gpurun_out/r2p_bench_fft.txt
run() {  # name, env...
  name=$1; shift
  env "$@" timeout 400 python bench.py --steps 8 --warmup 3 --no-parity --no-cpu-baseline --breakdown gpurun_out/r2p_breakdown_$name.json > gpurun_out/r2p_bench_$name.json 2> gpurun_out/r2p_bench_$name.err
  echo "bench $name rc=$?"; python -c "import json; d=json.load(open('gpurun_out/r2p_bench_$name.json')); print(d['value'], d['e2e']['value'], d['ms_per_step'], d.get('roofline_fft_dc'))" || tail -3 gpurun_out/r2p_bench_$name.err
}
run kc512 SAN_WG_KC_MAX=512
run kc256 SAN_WG_KC_MAX=256
run kc128 SAN_WG_KC_MAX=128
