#!/bin/bash
# Round-2 GPU call 42 (8 GPUs): final tree (lean epilogue stores): cfg2 weak scaling, 64 slices per GPU
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29561 bench.py --gpus 8 --steps 8 --warmup 3 --no-profile --no-parity > gpurun_out/r2fin2_cfg2_weak_8gpu.json 2> gpurun_out/r2fin2_cfg2_weak_8gpu.err
echo "rc=$?"; python tools/jline.py gpurun_out/r2fin2_cfg2_weak_8gpu.json || tail -5 gpurun_out/r2fin2_cfg2_weak_8gpu.err
