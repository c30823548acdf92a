#!/bin/bash
# 2-GPU sanity of the bench through TrainSession (NCCL AVG all-reduce inside the captured graphs) + infer / bg bench lines
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T=${TAG:-r2e}
echo "=== bench 2 GPUs"
timeout -k 10 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 8 --warmup 3 > gpurun_out/bench2_$T.json 2> gpurun_out/bench2_$T.err
echo "rc $?"; tail -c 1500 gpurun_out/bench2_$T.json; tail -n 8 gpurun_out/bench2_$T.err
echo "=== bench --mode infer"; timeout -k 10 300 python bench.py --mode infer --steps 20 --warmup 3 > gpurun_out/bench_infer_$T.json 2> gpurun_out/bench_infer_$T.err; tail -c 2000 gpurun_out/bench_infer_$T.json; tail -n 5 gpurun_out/bench_infer_$T.err
echo "=== bench --mode bg"; timeout -k 10 300 python bench.py --mode bg --steps 10 --warmup 3 > gpurun_out/bench_bg_$T.json 2> gpurun_out/bench_bg_$T.err; tail -c 2000 gpurun_out/bench_bg_$T.json; tail -n 5 gpurun_out/bench_bg_$T.err
echo "=== reference arm under torchrun"
timeout -k 10 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 2>&1 | tail -c 600
