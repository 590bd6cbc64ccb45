#!/bin/bash
# Round-2 GPU call 39: re-time the 36->18 @320 conv (outlier check) 
mkdir -p gpurun_out
for i in 1 2; do timeout 200 python tools/bench_tc.py 64 "36,18,320,3;18,18,320,3" 2>&1 | cut -c1-60; done
