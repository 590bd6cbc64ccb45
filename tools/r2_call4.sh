#!/bin/bash
# Round-2 GPU call 4 (after the mixed-format fault: fp16 pairs with a dynamic scale for dY): fp16-pair operand formats (F16x3) + DXN conv form: correctness, microbench A/B, parity, bench A/B.
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_tc.py -m gpu -x -q -p no:cacheprovider > gpurun_out/r2d_tc_tests.log 2>&1; echo "tc tests (dxn on) rc=$?"; tail -4 gpurun_out/r2d_tc_tests.log | cut -c1-300
SAN_TC_DXN=0 timeout 400 python -m pytest tests/test_gpu_tc.py -m gpu -x -q -p no:cacheprovider > gpurun_out/r2d_tc_tests_dxn0.log 2>&1; echo "tc tests (dxn off) rc=$?"; tail -4 gpurun_out/r2d_tc_tests_dxn0.log | cut -c1-300
for cfg in "1 0" "0 0" "1 1" "1 2" "1 3"; do set -- $cfg
  SAN_TC_DXN=$1 SAN_TC_DXN_R=$2 timeout 200 python tools/bench_tc.py 64 > gpurun_out/r2d_bench_tc_dxn$1_r$2.txt 2>&1
  echo "--- bench_tc DXN=$1 R=$2"; cut -c1-75 gpurun_out/r2d_bench_tc_dxn$1_r$2.txt | head -30
done
timeout 600 python -m pytest tests/test_gpu_models.py tests/test_gpu_ops.py -m gpu -x -q -p no:cacheprovider > gpurun_out/r2d_model_tests.log 2>&1; echo "model tests rc=$?"; tail -4 gpurun_out/r2d_model_tests.log | cut -c1-300
timeout 600 python -m pytest tests/test_gpu_parity_full.py -m gpu -q -p no:cacheprovider -s > gpurun_out/r2d_parity.log 2>&1; echo "parity rc=$?"; grep -o '"forward": {"img_rec": {[^}]*}' gpurun_out/r2d_parity.log; grep -o '"all_concatenated": {[^}]*}' gpurun_out/r2d_parity.log
cp gpurun_out/parity_cfg2_n2_320_c12.json gpurun_out/r2d_parity_cfg2_f16.json; cp gpurun_out/parity_cfg4_n1_15c_640x368_c12.json gpurun_out/r2d_parity_cfg4_f16.json
SAN_TC_FMT=f16nomix timeout 400 python -m pytest tests/test_gpu_parity_full.py -m gpu -q -p no:cacheprovider -s -k cfg2 > gpurun_out/r2d_parity_nomix.log 2>&1; echo "parity nomix rc=$?"; grep -o '"forward": {"img_rec": {[^}]*}' gpurun_out/r2d_parity_nomix.log; grep -o '"all_concatenated": {[^}]*}' gpurun_out/r2d_parity_nomix.log
for d in 1 0; do
  SAN_TC_DXN=$d timeout 300 python bench.py --steps 5 --warmup 3 --no-parity --no-cpu-baseline --breakdown gpurun_out/r2d_breakdown_dxn$d.json > gpurun_out/r2d_bench_dxn$d.json 2> gpurun_out/r2d_bench_dxn$d.err
  echo "bench DXN=$d rc=$?"; python -c "import json; d=json.load(open('gpurun_out/r2d_bench_dxn$d.json')); print(d['value'], d['roofline']['frac'], d['roofline']['avg_launch_ms'], d['kernel_time_shares'])"; tail -2 gpurun_out/r2d_bench_dxn$d.err
done
