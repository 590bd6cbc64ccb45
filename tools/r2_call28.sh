#!/bin/bash
# Round-2 GPU call 28 (8 GPUs): final tree: cfg2 weak scaling (64 slices per GPU), cfg3 strong scaling (global 256)
mkdir -p gpurun_out
N=8
tr() {  # name, bench args...
  name=$1; shift
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus $N --steps 8 --warmup 3 --no-profile "$@" > gpurun_out/r2bb_${name}_8gpu.json 2> gpurun_out/r2bb_${name}_8gpu.err
  echo "$name rc=$?"; python tools/jline.py gpurun_out/r2bb_${name}_8gpu.json || tail -5 gpurun_out/r2bb_${name}_8gpu.err
}
tr cfg2_weak
tr cfg3_strong --scaling strong --batch 256 --mask standard --sparsity 0.125 --lncc-weight 1.0
