#!/bin/bash
# Round-2 GPU call 23 (2 GPUs): final-tree multi-GPU validation: SyncBN test, cfg2 weak-scaling bench (overlapped buckets)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -x -q -p no:cacheprovider > gpurun_out/r2w_multi_tests.log 2>&1; echo "multi-gpu tests rc=$?"; tail -5 gpurun_out/r2w_multi_tests.log | cut -c1-400
N=2
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus $N --steps 8 --warmup 3 --no-profile > gpurun_out/r2w_cfg2_2gpu.json 2> gpurun_out/r2w_cfg2_2gpu.err; echo "bench rc=$?"; python tools/jline.py gpurun_out/r2w_cfg2_2gpu.json || tail -5 gpurun_out/r2w_cfg2_2gpu.err
