#!/bin/bash
# Round 2 session aa: gate + PReLU of the discriminator's cell as one pass
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T=${TAG:-r2aa}
timeout -k 10 600 python -m pytest -q -m gpu -p no:cacheprovider tests/test_ops_gpu.py -k "gate_prelu or gating or activations" > gpurun_out/ops_$T.log 2>&1
echo "exit $? : $(tail -n 2 gpurun_out/ops_$T.log | tr '\n' ' ')"; grep -E "^(FAILED|ERROR)|^E  " gpurun_out/ops_$T.log | head
timeout -k 10 900 python -m pytest -q -m gpu -p no:cacheprovider tests/test_model_gpu.py -x > gpurun_out/model_$T.log 2>&1
echo "exit $? : $(tail -n 2 gpurun_out/model_$T.log | tr '\n' ' ')"; grep -E "^(FAILED|ERROR)|^E  " gpurun_out/model_$T.log | head
for v in 1 0 1 0; do
  echo "=== bench FGC_GATE_PRELU=$v"
  FGC_GATE_PRELU=$v timeout -k 10 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_${T}_gp$v.json 2> gpurun_out/bench_${T}_gp$v.err
  python -c "import sys,json; d=json.loads(open('gpurun_out/bench_${T}_gp$v.json').read().strip().splitlines()[-1]); print('value', d['value'], 'ms', d['ms_per_step'], 'e2e', d['e2e']['value'], d['clocks']['sm_mhz'])"; tail -n 3 gpurun_out/bench_${T}_gp$v.err
done
