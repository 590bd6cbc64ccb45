#!/bin/bash
# Round-2 GPU call 22: semi-final validation: new tc tests, graph test, smoke(), default bench (parity + cpu baseline + profile),
# keep-staged threshold A/B, conv microbench with the statistics epilogue, ncu of stage + conv(stats) on 18->18 @320
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_tc.py tests/test_gpu_models.py -m gpu -x -q -p no:cacheprovider > gpurun_out/r2v_tests.log 2>&1; echo "tc+model tests rc=$?"; tail -3 gpurun_out/r2v_tests.log | cut -c1-400
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2v_smoke.log 2>&1; echo "smoke rc=$?"; tail -4 gpurun_out/r2v_smoke.log | cut -c1-300
timeout 900 python bench.py --breakdown gpurun_out/r2v_breakdown.json > gpurun_out/r2v_bench.json 2> gpurun_out/r2v_bench.err; echo "default bench rc=$?"; python tools/jline.py gpurun_out/r2v_bench.json || tail -5 gpurun_out/r2v_bench.err
run() {  # name, env...
  name=$1; shift
  env "$@" timeout 400 python bench.py --steps 8 --warmup 3 --no-parity --no-cpu-baseline --no-profile > gpurun_out/r2v_bench_$name.json 2> gpurun_out/r2v_bench_$name.err
  echo "bench $name rc=$?"; python -c "import json; d=json.load(open('gpurun_out/r2v_bench_$name.json')); print(d['value'], d['e2e']['value'], d['ms_per_step'], d.get('peak_mem_gb'))" || tail -5 gpurun_out/r2v_bench_$name.err
}
run keep65 SAN_KEEP_STAGED_BELOW=0.65
run keep75 SAN_KEEP_STAGED_BELOW=0.75
timeout 300 python tools/bench_tc.py 64 > gpurun_out/r2v_bench_tc.txt 2>&1; cut -c1-40,96- gpurun_out/r2v_bench_tc.txt | head -30
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"stage_simple_kernel|conv_tc_kernel|wgrad_tc_kernel" -c 8 -f -o gpurun_out/r2v_tc python tools/bench_tc.py 64 "18,18,320,3" > gpurun_out/r2v_ncu_tc.log 2>&1; tail -2 gpurun_out/r2v_ncu_tc.log
