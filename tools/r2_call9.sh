#!/bin/bash
# Round-2 GPU call 9: tc tests, CUDA-graph step (test + bench at bs 4 / 64), norm microbench, ncu captures for profiles/
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_tc.py -m gpu -x -q -p no:cacheprovider > gpurun_out/r2i_tc_tests.log 2>&1; echo "tc tests rc=$?"; tail -3 gpurun_out/r2i_tc_tests.log | cut -c1-300
timeout 300 python -m pytest tests/test_gpu_models.py -m gpu -x -q -p no:cacheprovider -k "graphed or rec_step or train_and_eval" > gpurun_out/r2i_graph_test.log 2>&1; echo "graph test rc=$?"; tail -3 gpurun_out/r2i_graph_test.log | cut -c1-400
for b in 4 64; do
  timeout 400 python bench.py --steps 10 --warmup 3 --no-parity --no-cpu-baseline --batch $b --graph > gpurun_out/r2i_bench_graph_bs$b.json 2> gpurun_out/r2i_bench_graph_bs$b.err
  echo "graph bs=$b rc=$?"; python -c "import json; d=json.load(open('gpurun_out/r2i_bench_graph_bs$b.json')); print(d['value'], d['e2e']['value'], d['ms_per_step'], d['gpu_launches'], d['peak_mem_gb'])" || tail -4 gpurun_out/r2i_bench_graph_bs$b.err
done
timeout 200 python tools/bench_norm.py 64 > gpurun_out/r2i_bench_norm.txt 2>&1; cat gpurun_out/r2i_bench_norm.txt
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"stage_act_kernel|conv_tc_kernel|wgrad_tc_kernel" -c 6 -f -o gpurun_out/r2i_tc python tools/bench_tc.py 64 "18,18,320,3" > gpurun_out/r2i_ncu_tc.log 2>&1; tail -2 gpurun_out/r2i_ncu_tc.log
timeout 300 ncu --set full --clock-control none -k regex:"in_bwd_fused|act_bwd_reduce_map|act_bwd_apply_map" -c 3 -f -o gpurun_out/r2i_norm python tools/bench_norm.py 64 > gpurun_out/r2i_ncu_norm.log 2>&1; tail -2 gpurun_out/r2i_ncu_norm.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 12000 --csv --log-file gpurun_out/r2i_launches.csv python bench.py --steps 1 --warmup 1 --min-warmup 1 --no-parity --no-cpu-baseline --no-profile > gpurun_out/r2i_ncu_bench.log 2>&1; tail -1 gpurun_out/r2i_ncu_bench.log | cut -c1-200; wc -l gpurun_out/r2i_launches.csv
