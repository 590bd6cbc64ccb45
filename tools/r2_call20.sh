#!/bin/bash
# Round-2 GPU call 20: VarNet on two sub-batch streams (conv of one chain next to the element-wise kernels of the other):
# correctness (model tests, headline parity) + A/B over the number of streams
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_models.py -m gpu -x -q -p no:cacheprovider > gpurun_out/r2t_tests.log 2>&1; echo "model tests rc=$?"; tail -3 gpurun_out/r2t_tests.log | cut -c1-400
run() {  # name, env...
  name=$1; shift
  env "$@" timeout 400 python bench.py --steps 8 --warmup 3 --no-parity --no-cpu-baseline --breakdown gpurun_out/r2t_breakdown_$name.json > gpurun_out/r2t_bench_$name.json 2> gpurun_out/r2t_bench_$name.err
  echo "bench $name rc=$?"; python -c "import json; d=json.load(open('gpurun_out/r2t_bench_$name.json')); print(d['value'], d['e2e']['value'], d['ms_per_step'], d.get('peak_mem_gb'), d['roofline']['frac'])" || tail -5 gpurun_out/r2t_bench_$name.err
}
run s1 SAN_VARNET_STREAMS=1
run s2 SAN_VARNET_STREAMS=2
run s4 SAN_VARNET_STREAMS=4
run s2_nowg SAN_VARNET_STREAMS=2 SAN_WG_OVERLAP=0
timeout 600 python -m pytest tests/test_gpu_parity_full.py -m gpu -q -p no:cacheprovider -s -k cfg2 > gpurun_out/r2t_parity.log 2>&1; echo "parity rc=$?"; grep -o '"forward": {"img_rec": {[^}]*}' gpurun_out/r2t_parity.log; grep -o '"all_concatenated": {[^}]*}' gpurun_out/r2t_parity.log
