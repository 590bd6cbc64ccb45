#!/bin/bash
# Round-2 GPU call 10: hi/lo-stacked (HLS) conv form: correctness + A/B; CUDA-graph step test + bench
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_tc.py -m gpu -x -q -p no:cacheprovider > gpurun_out/r2j_tc_tests.log 2>&1; echo "tc tests (hls on) rc=$?"; tail -3 gpurun_out/r2j_tc_tests.log | cut -c1-300
SH="3,18,320,3;18,18,320,3;36,18,320,3;18,36,160,3;36,36,160,3;72,36,160,3;2,32,320,3;32,32,320,3;96,32,320,3;8,8,320,3;16,16,160,3"
for cfg in "1 0" "0 0" "1 1" "1 2" "1 3"; do set -- $cfg
  SAN_TC_HLS=$1 SAN_TC_HLS_R=$2 timeout 200 python tools/bench_tc.py 64 "$SH" > gpurun_out/r2j_bench_tc_hls$1_r$2.txt 2>&1
  echo "--- bench_tc HLS=$1 R=$2"; cut -c1-75 gpurun_out/r2j_bench_tc_hls$1_r$2.txt | head -14
done
timeout 300 python -m pytest tests/test_gpu_models.py -m gpu -x -q -p no:cacheprovider > gpurun_out/r2j_model_tests.log 2>&1; echo "model tests rc=$?"; tail -3 gpurun_out/r2j_model_tests.log | cut -c1-400
for h in 1 0; do
  SAN_TC_HLS=$h timeout 400 python bench.py --steps 8 --warmup 3 --no-parity --no-cpu-baseline --breakdown gpurun_out/r2j_breakdown_hls$h.json > gpurun_out/r2j_bench_hls$h.json 2> gpurun_out/r2j_bench_hls$h.err
  echo "bench HLS=$h rc=$?"; python -c "import json; d=json.load(open('gpurun_out/r2j_bench_hls$h.json')); print(d['value'], d['e2e']['value'], d['roofline']['frac'], d['roofline']['avg_launch_ms'], d['kernel_time_shares'])" || tail -3 gpurun_out/r2j_bench_hls$h.err
done
for b in 4 64; do
  timeout 400 python bench.py --steps 10 --warmup 3 --no-parity --no-cpu-baseline --no-profile --batch $b --graph > gpurun_out/r2j_bench_graph_bs$b.json 2> gpurun_out/r2j_bench_graph_bs$b.err
  echo "graph bs=$b rc=$?"; python -c "import json; d=json.load(open('gpurun_out/r2j_bench_graph_bs$b.json')); print(d['value'], d['e2e']['value'], d['ms_per_step'], d['gpu_launches'], d['peak_mem_gb'])" || tail -4 gpurun_out/r2j_bench_graph_bs$b.err
done
SAN_TC_HLS=1 timeout 300 ncu --set full --clock-control none --import-source on -k regex:"conv_tc_kernel" -c 1 -f -o gpurun_out/r2j_conv_hls python tools/bench_tc.py 64 "18,18,320,3" > gpurun_out/r2j_ncu.log 2>&1; tail -1 gpurun_out/r2j_ncu.log
