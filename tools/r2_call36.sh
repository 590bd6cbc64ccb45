#!/bin/bash
# Round-2 GPU call 36 (8 GPUs): final tree: cfg4 (15 coils 640x368) and cfg5 (Mixed, global 128, strong scaling)
mkdir -p gpurun_out
N=8
tr() {  # name, bench args...
  name=$1; shift
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus $N --steps 8 --warmup 3 --no-profile --no-parity "$@" > gpurun_out/r2cfg_${name}_8gpu.json 2> gpurun_out/r2cfg_${name}_8gpu.err
  echo "$name rc=$?"; python tools/jline.py gpurun_out/r2cfg_${name}_8gpu.json || tail -5 gpurun_out/r2cfg_${name}_8gpu.err
}
tr cfg4 --batch 4 --coils 15 --shape 640x368
tr cfg5_mixed --scaling strong --batch 128 --reg Mixed --mi-weight 1.0
