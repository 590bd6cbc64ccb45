#!/bin/bash
# Round-2 GPU call 18: TMA-staged FFT column pass (soft DC / coil reduce): parity + microbench A/B; statistics-epilogue test
# after the stage fallback; bench with the corrected kernel classes
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_ops.py -m gpu -x -q -p no:cacheprovider -k "dc_block or fft" > gpurun_out/r2r_fft_tests.log 2>&1; echo "fft/dc tests rc=$?"; tail -5 gpurun_out/r2r_fft_tests.log | cut -c1-400
SAN_FFT_TMA=0 timeout 300 python -m pytest tests/test_gpu_ops.py -m gpu -x -q -p no:cacheprovider -k "dc_block" > gpurun_out/r2r_fft_tests_tma0.log 2>&1; echo "dc tests TMA=0 rc=$?"; tail -3 gpurun_out/r2r_fft_tests_tma0.log | cut -c1-400
for t in 1 0; do echo "--- SAN_FFT_TMA=$t"; SAN_FFT_TMA=$t timeout 120 python tools/bench_fft.py 64 20 2>&1 | tee gpurun_out/r2r_bench_fft_tma$t.txt; done
timeout 600 python -m pytest tests/test_gpu_tc.py tests/test_gpu_models.py -m gpu -x -q -p no:cacheprovider > gpurun_out/r2r_tests.log 2>&1; echo "tc+model tests rc=$?"; tail -3 gpurun_out/r2r_tests.log | cut -c1-400
timeout 400 python bench.py --steps 8 --warmup 3 --no-parity --no-cpu-baseline --breakdown gpurun_out/r2r_breakdown.json > gpurun_out/r2r_bench.json 2> gpurun_out/r2r_bench.err
echo "bench rc=$?"; python -c "import json; d=json.load(open('gpurun_out/r2r_bench.json')); print(d['value'], d['e2e']['value'], d['ms_per_step'], d['kernel_time_shares'], d['roofline']['frac'], d['roofline_fft_dc']['frac'], d['roofline_fft_dc']['avg_launch_ms'])" || tail -3 gpurun_out/r2r_bench.err
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"fft_" -c 6 -f -o gpurun_out/r2r_fft python tools/bench_fft.py 64 1 > gpurun_out/r2r_ncu_fft.log 2>&1; tail -2 gpurun_out/r2r_ncu_fft.log
