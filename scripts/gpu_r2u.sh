#!/bin/bash
# Round 2 session u: caption encoder's word part / weight gradients aside (side streams); 1x1 layers on the halo kernel for the matching model's trunk
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T=${TAG:-r2u}
echo "=== model / entry / callers tests"
timeout -k 10 900 python -m pytest -q -m gpu -p no:cacheprovider tests/test_model_gpu.py tests/test_entry_gpu.py tests/test_callers_gpu.py tests/test_pix2pix_gpu.py -x > gpurun_out/model_$T.log 2>&1
echo "exit $? : $(tail -n 2 gpurun_out/model_$T.log | tr '\n' ' ')"; grep -E "^(FAILED|ERROR)|^E  " gpurun_out/model_$T.log | head
for v in 1 0 1 0; do
  echo "=== bench FGC_SIDE_STREAMS=$v"
  FGC_SIDE_STREAMS=$v timeout -k 10 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_${T}_side$v.json 2> gpurun_out/bench_${T}_side$v.err
  python -c "import sys,json; d=json.loads(open('gpurun_out/bench_${T}_side$v.json').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['clocks']['sm_mhz'])"; tail -n 3 gpurun_out/bench_${T}_side$v.err
done
for v in 1 2; do
  echo "=== bench rmi FGC_HALO=$v"
  FGC_HALO=$v timeout -k 10 600 python bench.py --mode rmi --steps 5 --warmup 2 > gpurun_out/bench_rmi_${T}_halo$v.json 2> gpurun_out/bench_rmi_${T}_halo$v.err
  python -c "import sys,json; d=json.loads(open('gpurun_out/bench_rmi_${T}_halo$v.json').read().strip().splitlines()[-1]); print(d['value'], d['config']['single_pass_bf16'], d['config']['parity_mode'])"; tail -n 3 gpurun_out/bench_rmi_${T}_halo$v.err
done
