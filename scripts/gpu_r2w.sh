#!/bin/bash
# Round 2 session w (N GPUs): the final step (side streams inside the captured steps + the NCCL all-reduce) under torchrun, as the driver launches it
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T=${TAG:-r2w}
N=${NGPU:-2}
timeout -k 10 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29519 \
    bench.py --gpus $N --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_${T}_n$N.json 2> gpurun_out/bench_${T}_n$N.err
echo "exit $?"
python - <<PY
import json
d = json.loads(open("gpurun_out/bench_${T}_n$N.json").read().strip().splitlines()[-1])
print("N=$N: %.1f images/s, %.2f ms/iteration, e2e %.1f, clocks %s" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["clocks"]))
PY
tail -n 4 gpurun_out/bench_${T}_n$N.err
