#!/bin/bash
# Round-2 GPU call 37: one-lane barrier polls in the conv / weight-gradient epilogue warps (power / issue slots): A/B
mkdir -p gpurun_out
for v in 0 1 0 1; do
  SAN_TC_POLL1=$v timeout 400 python bench.py --steps 8 --warmup 3 --no-parity --no-cpu-baseline --no-profile > gpurun_out/r2poll_bench_$v.json 2> gpurun_out/r2poll_bench_$v.err
  echo "POLL1=$v rc=$?"; python -c "import json; d=json.load(open('gpurun_out/r2poll_bench_$v.json')); print(d['value'], d['e2e']['value'], d['ms_per_step'], d['clocks'])" || tail -3 gpurun_out/r2poll_bench_$v.err
done
SAN_TC_POLL1=1 timeout 300 python -m pytest tests/test_gpu_tc.py -m gpu -x -q -p no:cacheprovider > gpurun_out/r2poll_tests.log 2>&1; echo "tc tests POLL1=1 rc=$?"; tail -2 gpurun_out/r2poll_tests.log
