#!/bin/bash
# Round-2 GPU call 33: ncu of the row-ring conv kernel (stall sampling per source line)
mkdir -p gpurun_out
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"conv_rows_kernel" -c 1 -f -o gpurun_out/r2ii_rows python tools/bench_tc.py 64 "18,18,320,3" > gpurun_out/r2ii_ncu.log 2>&1; tail -2 gpurun_out/r2ii_ncu.log
