#!/bin/bash
# Round-2 first GPU call: (1) FFT v2 A/B (tools/ab_fft_v2.sh), (2) the eager-PyTorch-on-B200 reference arm
# (SURVEY 8d "the real bar"): oracle port run on cuda with cuDNN/cuFFT, TF32 off and on, several batch sizes.
mkdir -p gpurun_out
bash tools/ab_fft_v2.sh
for tf in 0 1; do
  for bs in 64 32 16 4; do
    timeout 300 python bench.py --impl reference --ref-device cuda --tf32 $tf --cpu-sample $bs --steps 5 --warmup 1 \
      > gpurun_out/r2_gpuref_tf${tf}_bs${bs}.json 2> gpurun_out/r2_gpuref_tf${tf}_bs${bs}.err
    echo "gpuref tf32=$tf bs=$bs rc=$?"; tail -c 600 gpurun_out/r2_gpuref_tf${tf}_bs${bs}.json; tail -2 gpurun_out/r2_gpuref_tf${tf}_bs${bs}.err
  done
done
nvidia-smi --query-gpu=name,memory.total --format=csv
lscpu | grep -E "Model name|^CPU\(s\)"
