#!/bin/bash
# Round 2 session ae: input gradient of the cell's Conv_2 from the low-resolution output gradient (phase launches of 2x2-tap convolutions)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T=${TAG:-r2ae}
timeout -k 10 900 python -m pytest -q -m gpu -p no:cacheprovider tests/test_ops_gpu.py -k "pooled_gy or (conv and tcgen05 and not gather)" > gpurun_out/ops_$T.log 2>&1
echo "exit $? : $(tail -n 2 gpurun_out/ops_$T.log | tr '\n' ' ')"; grep -E "^(FAILED|ERROR)|^E  " gpurun_out/ops_$T.log | head -20
timeout -k 10 900 python -m pytest -q -m gpu -p no:cacheprovider tests/test_model_gpu.py tests/test_production_shapes_gpu.py -x > gpurun_out/model_$T.log 2>&1
echo "exit $? : $(tail -n 2 gpurun_out/model_$T.log | tr '\n' ' ')"; grep -E "^(FAILED|ERROR)|^E  " gpurun_out/model_$T.log | head
for v in 1 0 1 0; do
  echo "=== bench FGC_PHASE_DGRAD=$v"
  FGC_PHASE_DGRAD=$v timeout -k 10 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_${T}_ph$v.json 2> gpurun_out/bench_${T}_ph$v.err
  python -c "import sys,json; d=json.loads(open('gpurun_out/bench_${T}_ph$v.json').read().strip().splitlines()[-1]); print('value', d['value'], 'ms', d['ms_per_step'], 'e2e', d['e2e']['value'], d['clocks']['sm_mhz'])"; tail -n 3 gpurun_out/bench_${T}_ph$v.err
done
