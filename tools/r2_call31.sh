#!/bin/bash
# Round-2 GPU call 31: row-ring conv microbench on the layers it supports
mkdir -p gpurun_out
timeout 300 python tools/bench_tc.py 64 "18,18,320,3;3,18,320,3;18,3,320,3;8,8,320,3;2,8,320,3;16,16,160,3;18,18,160,3" > gpurun_out/r2ee_bench_tc_rows.txt 2>&1; cut -c1-40,96- gpurun_out/r2ee_bench_tc_rows.txt
