#!/bin/bash
# Round 2 session y (8 GPUs): final code with input prefetch; the bucketed all-reduce under the backward pass at 8 ranks
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T=${TAG:-r2y}
N=${NGPU:-8}
for v in 0 1; do
  FGC_OVERLAP_ALLREDUCE=$v timeout -k 10 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 \
      bench.py --gpus $N --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_${T}_n${N}_overlap$v.json 2> gpurun_out/bench_${T}_n${N}_overlap$v.err
  echo "exit $? (overlap $v)"
  python -c "import json; d=json.loads(open('gpurun_out/bench_${T}_n${N}_overlap$v.json').read().strip().splitlines()[-1]); print('N=$N overlap=$v: %.1f images/s, %.2f ms/iteration, e2e %.1f, %s' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['clocks']))"
  tail -n 2 gpurun_out/bench_${T}_n${N}_overlap$v.err
done
