#!/bin/bash
# Round-2 GPU call 7 (8 GPUs): weak scaling with / without the overlapped gradient all-reduce, strong scaling (cfg3: global 256),
# cfg4 and cfg5 on 8 GPUs.  torchrun, one rank per GPU.
mkdir -p gpurun_out
N=${1:-8}
tr() { name=$1; shift; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 8 --warmup 3 --no-profile "$@" > gpurun_out/r2g_${name}_${N}gpu.json 2> gpurun_out/r2g_${name}_${N}gpu.err; echo "$name rc=$?"; python tools/jline.py gpurun_out/r2g_${name}_${N}gpu.json || tail -3 gpurun_out/r2g_${name}_${N}gpu.err; }
tr cfg2_overlap
# (blocking vs overlapped all-reduce: A/B at N = 2 in tools/r2_call11.sh)
tr cfg3_strong --scaling strong --batch 256 --mask standard --sparsity 0.125 --lncc-weight 1.0
tr cfg4 --batch 4 --coils 15 --shape 640x368
tr cfg5_mixed --scaling strong --batch 128 --reg Mixed --mi-weight 1.0
