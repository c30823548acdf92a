#!/bin/bash
# Round 2 session ac: 24 x 24 layers (75 % tile efficiency) on the halo-reuse kernel instead of the per-tap gather kernel
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T=${TAG:-r2ac}
FGC_HALO_MIN_EFF=0.7 timeout -k 10 600 python -m pytest -q -m gpu -p no:cacheprovider tests/test_production_shapes_gpu.py tests/test_ops_gpu.py -k "production or (conv and tcgen05 and not gather) or prelu or gate_prelu" > gpurun_out/ops_$T.log 2>&1
echo "exit $? : $(tail -n 2 gpurun_out/ops_$T.log | tr '\n' ' ')"; grep -E "^(FAILED|ERROR)|^E  " gpurun_out/ops_$T.log | head
for v in 0.7 0.8 0.7 0.8; do
  echo "=== bench FGC_HALO_MIN_EFF=$v"
  FGC_HALO_MIN_EFF=$v timeout -k 10 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_${T}_eff$v.json 2> gpurun_out/bench_${T}_eff$v.err
  python -c "import sys,json; d=json.loads(open('gpurun_out/bench_${T}_eff$v.json').read().strip().splitlines()[-1]); print('value', d['value'], 'ms', d['ms_per_step'], 'e2e', d['e2e']['value'], d['clocks']['sm_mhz'])"; tail -n 3 gpurun_out/bench_${T}_eff$v.err
done
echo "=== rmi (24 x 24 batch form of group 5)"
for v in 0.7 0.8; do
FGC_HALO_MIN_EFF=$v timeout -k 10 600 python bench.py --mode rmi --steps 5 --warmup 2 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$v', d['value'], d['config']['single_pass_bf16']['trunk_ms'])"
done
