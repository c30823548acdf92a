#!/bin/bash
# Round 2 session s (2 GPUs): the bucketed, overlapped gradient all-reduce inside the captured steps against the single one
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T=${TAG:-r2s}
N=${NGPU:-2}
run() {  # $1 = tag suffix, rest = env
  env "${@:2}" timeout -k 10 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 \
      bench.py --gpus $N --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_${T}_$1.json 2> gpurun_out/bench_${T}_$1.err
  echo "exit $? ($1)"; tail -c 400 gpurun_out/bench_${T}_$1.json | head -c 400; echo
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/bench_${T}_$1.json").read().strip().splitlines()[-1])
    print("$1: %.1f images/s, %.2f ms/iteration, e2e %.1f" % (d["value"], d["ms_per_step"], d["e2e"]["value"]))
except Exception as e:
    print("$1: no result line:", e)
PY
  tail -n 4 gpurun_out/bench_${T}_$1.err
}
run overlap FGC_OVERLAP_ALLREDUCE=1
run single FGC_OVERLAP_ALLREDUCE=0
run overlap2 FGC_OVERLAP_ALLREDUCE=1
