#!/bin/bash
# Round-2 GPU call 27: vectorised bias-gradient kernel + fork threshold: tests, bench, bs 4 eager
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_tc.py tests/test_gpu_models.py tests/test_gpu_gan.py -m gpu -x -q -p no:cacheprovider > gpurun_out/r2aa_tests.log 2>&1; echo "tc+model+gan tests rc=$?"; tail -3 gpurun_out/r2aa_tests.log | cut -c1-400
run() {  # name, bench args...
  name=$1; shift
  timeout 500 python bench.py --steps 8 --warmup 3 --no-parity --no-cpu-baseline --no-profile "$@" > gpurun_out/r2aa_bench_$name.json 2> gpurun_out/r2aa_bench_$name.err
  echo "bench $name rc=$?"; python -c "import json; d=json.load(open('gpurun_out/r2aa_bench_$name.json')); print(d['value'], d['e2e']['value'], d['ms_per_step'], d.get('peak_mem_gb'))" || tail -5 gpurun_out/r2aa_bench_$name.err
}
run cfg2
run bs4_eager --batch 4
