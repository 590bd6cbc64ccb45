#!/bin/bash
# Round-2 GPU call 11 (2 GPUs): weak-scaling step with the overlapped per-cascade gradient all-reduce vs the blocking flat one
mkdir -p gpurun_out
N=2
tr() { name=$1; shift; timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 --no-profile "$@" > gpurun_out/r2k_${name}_${N}gpu.json 2> gpurun_out/r2k_${name}_${N}gpu.err; echo "$name rc=$?"; python -c "import json; d=json.load(open('gpurun_out/r2k_${name}_${N}gpu.json')); print(d['value'], d['unit'], d['ms_per_step'], d['scaling'], d['config']['global_batch'], d.get('grad_allreduce'))" 2>/dev/null || tail -5 gpurun_out/r2k_${name}_${N}gpu.err; }
tr cfg2_overlap
tr cfg2_blocking --no-overlap
timeout 300 python bench.py --steps 10 --warmup 3 --no-parity --no-cpu-baseline --no-profile > gpurun_out/r2k_cfg2_1gpu.json 2>/dev/null; python -c "import json; d=json.load(open('gpurun_out/r2k_cfg2_1gpu.json')); print('1 gpu', d['value'], d['ms_per_step'])"
