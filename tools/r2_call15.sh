#!/bin/bash
# Round-2 GPU call 15: weight gradient on a side stream next to the element-wise backward (SAN_WG_OVERLAP) + pre-staged dY:
# correctness (tc / model tests, headline parity), A/B bench over SAN_WG_OVERLAP and the wgrad shared-memory cap
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_tc.py tests/test_gpu_models.py -m gpu -x -q -p no:cacheprovider > gpurun_out/r2o_tests.log 2>&1; echo "tc+model tests rc=$?"; tail -3 gpurun_out/r2o_tests.log | cut -c1-300
run() {  # name, env...
  name=$1; shift
  env "$@" timeout 400 python bench.py --steps 8 --warmup 3 --no-parity --no-cpu-baseline --breakdown gpurun_out/r2o_breakdown_$name.json > gpurun_out/r2o_bench_$name.json 2> gpurun_out/r2o_bench_$name.err
  echo "bench $name rc=$?"; python -c "import json; d=json.load(open('gpurun_out/r2o_bench_$name.json')); print(d['value'], d['e2e']['value'], d['ms_per_step'], d['kernel_time_shares'])" || tail -3 gpurun_out/r2o_bench_$name.err
}
run ov0 SAN_WG_OVERLAP=0 SAN_WG_SMEM_KB=225
run ov1_209 SAN_WG_OVERLAP=1 SAN_WG_SMEM_KB=209
run ov1_225 SAN_WG_OVERLAP=1 SAN_WG_SMEM_KB=225
run ov1_177 SAN_WG_OVERLAP=1 SAN_WG_SMEM_KB=177
run ov0_209 SAN_WG_OVERLAP=0 SAN_WG_SMEM_KB=209
timeout 600 python -m pytest tests/test_gpu_parity_full.py -m gpu -q -p no:cacheprovider -s -k cfg2 > gpurun_out/r2o_parity.log 2>&1; echo "parity rc=$?"; grep -o '"forward": {"img_rec": {[^}]*}' gpurun_out/r2o_parity.log; grep -o '"all_concatenated": {[^}]*}' gpurun_out/r2o_parity.log
