#!/bin/bash
# Round-2 GPU call 17: statistics epilogue of the conv kernel (san_tc_conv_stats): parity + A/B; FFT load-before-barrier
# reorder: microbench + ncu capture of the FFT kernels
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_tc.py tests/test_gpu_ops.py tests/test_gpu_models.py -m gpu -x -q -p no:cacheprovider > gpurun_out/r2q_tests.log 2>&1; echo "tc+ops+model tests rc=$?"; tail -5 gpurun_out/r2q_tests.log | cut -c1-400
timeout 200 python tools/bench_fft.py 64 20 > gpurun_out/r2q_bench_fft.txt 2>&1; cat gpurun_out/r2q_bench_fft.txt
run() {  # name, env...
  name=$1; shift
  env "$@" timeout 400 python bench.py --steps 8 --warmup 3 --no-parity --no-cpu-baseline --breakdown gpurun_out/r2q_breakdown_$name.json > gpurun_out/r2q_bench_$name.json 2> gpurun_out/r2q_bench_$name.err
  echo "bench $name rc=$?"; python -c "import json; d=json.load(open('gpurun_out/r2q_bench_$name.json')); print(d['value'], d['e2e']['value'], d['ms_per_step'], d['kernel_time_shares'])" || tail -3 gpurun_out/r2q_bench_$name.err
}
run stats0 SAN_EPI_STATS=0
run stats1 SAN_EPI_STATS=1
run stats1_contig_all SAN_EPI_STATS=1 SAN_TC_CONTIG=1
run stats1_strided SAN_EPI_STATS=1 SAN_TC_CONTIG=0
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"fft_" -c 6 -f -o gpurun_out/r2q_fft python tools/bench_fft.py 64 1 > gpurun_out/r2q_ncu_fft.log 2>&1; tail -2 gpurun_out/r2q_ncu_fft.log
timeout 600 python -m pytest tests/test_gpu_parity_full.py -m gpu -q -p no:cacheprovider -s -k cfg2 > gpurun_out/r2q_parity.log 2>&1; echo "parity rc=$?"; grep -o '"forward": {"img_rec": {[^}]*}' gpurun_out/r2q_parity.log; grep -o '"all_concatenated": {[^}]*}' gpurun_out/r2q_parity.log
