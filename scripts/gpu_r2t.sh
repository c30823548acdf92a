#!/bin/bash
# Round 2 session t: spectral normalisation on a side stream under the generator's forward pass
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T=${TAG:-r2t}
echo "=== model / entry tests"
timeout -k 10 900 python -m pytest -q -m gpu -p no:cacheprovider tests/test_model_gpu.py tests/test_entry_gpu.py > gpurun_out/model_$T.log 2>&1
echo "exit $? : $(tail -n 2 gpurun_out/model_$T.log | tr '\n' ' ')"; grep -E "^(FAILED|ERROR)|^E  " gpurun_out/model_$T.log | head
for v in 1 0 1 0; do
  echo "=== bench FGC_SN_SIDE_STREAM=$v"
  FGC_SN_SIDE_STREAM=$v timeout -k 10 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_${T}_sn$v.json 2> gpurun_out/bench_${T}_sn$v.err
  python -c "import sys,json; d=json.loads(open('gpurun_out/bench_${T}_sn$v.json').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['clocks']['sm_mhz'])"; tail -n 3 gpurun_out/bench_${T}_sn$v.err
done
