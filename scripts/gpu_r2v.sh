#!/bin/bash
# Round 2 session v: min-max gates in sample chunks (apply pass reads from L2)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T=${TAG:-r2v}
timeout -k 10 600 python -m pytest -q -m gpu -p no:cacheprovider tests/test_ops_gpu.py -k "minmax or cbn or gating" > gpurun_out/ops_$T.log 2>&1
echo "exit $? : $(tail -n 2 gpurun_out/ops_$T.log | tr '\n' ' ')"; grep -E "^(FAILED|ERROR)|^E  " gpurun_out/ops_$T.log | head
for v in 40 0 20 80 40 0; do
  echo "=== bench FGC_MINMAX_CHUNK_MB=$v"
  FGC_MINMAX_CHUNK_MB=$v timeout -k 10 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_${T}_mm$v.json 2> gpurun_out/bench_${T}_mm$v.err
  python -c "import sys,json; d=json.loads(open('gpurun_out/bench_${T}_mm$v.json').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['clocks']['sm_mhz'])"; tail -n 3 gpurun_out/bench_${T}_mm$v.err
done
