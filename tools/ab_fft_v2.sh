#!/bin/bash
# One gpurun call that decides whether the register-resident FFT (SAN_FFT_V2=1) becomes the default:
#   /usr/local/graft/bin/gpurun --timeout 420 -- 'bash tools/ab_fft_v2.sh'
# 1. parity: the existing FFT / DC / VarNet GPU tests with v2 switched on (the flag changes the kernels under
#    the same C ABI, so every 320-sized case exercises v2);
# 2. speed: tools/bench_fft.py and a short bench.py, v1 vs v2;
# 3. ncu --set full of one fft_expand_dc launch pair with v2.
mkdir -p gpurun_out
SAN_FFT_V2=1 timeout 200 python -m pytest tests/test_gpu_ops.py tests/test_gpu_models.py -m gpu -q --tb=short -p no:cacheprovider \
    -k "fft or dc or rss or varnet or rec_step" > gpurun_out/ab_v2_tests.log 2>&1
tail -3 gpurun_out/ab_v2_tests.log
for v in 0 1 2; do   # 0 = Stockham, 1 = register FFT with 8 columns per CTA, 2 = with 16
  SAN_FFT_V2=$v python tools/bench_fft.py 64 20 > gpurun_out/ab_fft_v$v.txt 2>&1
  echo "--- SAN_FFT_V2=$v"; cat gpurun_out/ab_fft_v$v.txt
  SAN_FFT_V2=$v python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ab_bench_v$v.json 2> gpurun_out/ab_bench_v$v.err
  python -c "import json,sys; d=json.load(open('gpurun_out/ab_bench_v$v.json')); print('bench', d['value'], d['roofline_fft_dc'])"
done
SAN_FFT_V2=1 timeout 120 ncu --set full --clock-control none --import-source on -k regex:fft_.*v2 -c 2 -f \
    -o gpurun_out/ab_fft_v2 python tools/bench_fft.py 64 1 > gpurun_out/ab_ncu.log 2>&1
tail -2 gpurun_out/ab_ncu.log
