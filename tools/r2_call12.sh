#!/bin/bash
# Round-2 GPU call 12: tap pairing: correctness (tc tests, model tests, headline parity) + A/B (microbench, bench)
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_tc.py -m gpu -x -q -p no:cacheprovider > gpurun_out/r2l_tc_tests.log 2>&1; echo "tc tests rc=$?"; tail -3 gpurun_out/r2l_tc_tests.log | cut -c1-300
timeout 400 python -m pytest tests/test_gpu_models.py tests/test_gpu_gan.py -m gpu -x -q -p no:cacheprovider > gpurun_out/r2l_model_tests.log 2>&1; echo "model+gan tests rc=$?"; tail -3 gpurun_out/r2l_model_tests.log | cut -c1-300
timeout 600 python -m pytest tests/test_gpu_parity_full.py -m gpu -q -p no:cacheprovider -s > gpurun_out/r2l_parity.log 2>&1; echo "parity rc=$?"; grep -o '"forward": {"img_rec": {[^}]*}' gpurun_out/r2l_parity.log; grep -o '"all_concatenated": {[^}]*}' gpurun_out/r2l_parity.log
SH="3,18,320,3;18,18,320,3;36,18,320,3;18,36,160,3;36,36,160,3;72,36,160,3;36,72,80,3;72,72,80,3;2,32,320,3;8,8,320,3"
for pr in 1 0; do
  SAN_TC_PAIR=$pr timeout 200 python tools/bench_tc.py 64 "$SH" > gpurun_out/r2l_bench_tc_pair$pr.txt 2>&1
  echo "--- bench_tc PAIR=$pr"; cut -c1-75 gpurun_out/r2l_bench_tc_pair$pr.txt | head -12
done
for pr in 1 0; do
  SAN_TC_PAIR=$pr timeout 400 python bench.py --steps 8 --warmup 3 --no-parity --no-cpu-baseline --breakdown gpurun_out/r2l_breakdown_pair$pr.json > gpurun_out/r2l_bench_pair$pr.json 2> gpurun_out/r2l_bench_pair$pr.err
  echo "bench PAIR=$pr rc=$?"; python -c "import json; d=json.load(open('gpurun_out/r2l_bench_pair$pr.json')); print(d['value'], d['e2e']['value'], d['roofline']['frac'], d['roofline']['avg_launch_ms'], d['kernel_time_shares'])" || tail -3 gpurun_out/r2l_bench_pair$pr.err
done
SAN_STAGE_SIMPLE=0 timeout 400 python bench.py --steps 8 --warmup 3 --no-parity --no-cpu-baseline --breakdown gpurun_out/r2l_breakdown_stage_general.json > gpurun_out/r2l_bench_stage_general.json 2> gpurun_out/r2l_bench_stage_general.err
echo "bench SAN_STAGE_SIMPLE=0 rc=$?"; python -c "import json; d=json.load(open('gpurun_out/r2l_bench_stage_general.json')); print(d['value'], d['kernel_time_shares'])"
SAN_STAGE_SIMPLE=0 timeout 100 python tools/bench_tc.py 64 "18,18,320,3;36,36,160,3" 2>&1 | cut -c1-40
