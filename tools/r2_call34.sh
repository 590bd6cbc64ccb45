#!/bin/bash
# Round-2 GPU call 34 (2 GPUs): final validation as the driver runs it: full GPU suite, smoke(), default bench at N = 1 (both arms)
# and N = 2 (torchrun, default flags)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q -p no:cacheprovider > gpurun_out/r2fin_gpu_tests.log 2>&1; echo "gpu suite rc=$?"; tail -4 gpurun_out/r2fin_gpu_tests.log | cut -c1-300
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2fin_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r2fin_smoke.log | cut -c1-400
CUDA_VISIBLE_DEVICES=0 timeout 900 python bench.py --gpus 1 --steps 5 --warmup 3 --breakdown gpurun_out/r2fin_breakdown_1gpu.json > gpurun_out/r2fin_bench_1gpu.json 2> gpurun_out/r2fin_bench_1gpu.err; echo "bench N=1 rc=$?"; python tools/jline.py gpurun_out/r2fin_bench_1gpu.json || tail -5 gpurun_out/r2fin_bench_1gpu.err
CUDA_VISIBLE_DEVICES=0 timeout 600 python bench.py --impl reference --gpus 1 --steps 5 --warmup 3 > gpurun_out/r2fin_bench_reference_arm.json 2> gpurun_out/r2fin_bench_reference_arm.err; echo "reference arm rc=$?"; cut -c1-200 gpurun_out/r2fin_bench_reference_arm.json | tail -1
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r2fin_bench_2gpu.json 2> gpurun_out/r2fin_bench_2gpu.err; echo "bench N=2 rc=$?"; python tools/jline.py gpurun_out/r2fin_bench_2gpu.json || tail -5 gpurun_out/r2fin_bench_2gpu.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29542 bench.py --impl reference --gpus 2 --steps 5 --warmup 3 > gpurun_out/r2fin_bench_reference_arm_2gpu.json 2> gpurun_out/r2fin_bench_reference_arm_2gpu.err; echo "reference arm N=2 rc=$?"; cut -c1-160 gpurun_out/r2fin_bench_reference_arm_2gpu.json | tail -1
