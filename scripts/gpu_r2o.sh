#!/bin/bash
# Round 2 session o: instance-matching model (BASELINE configs[4]) -- kernels, parity, throughput
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T=${TAG:-r2o}
echo "=== rmi ops + small models"
timeout -k 10 900 python -m pytest -q -m gpu -p no:cacheprovider -s tests/test_rmi_gpu.py -k "not resnet101" > gpurun_out/rmi_ops_$T.log 2>&1
echo "exit $? : $(tail -n 2 gpurun_out/rmi_ops_$T.log | tr '\n' ' ')"; grep -E "^(FAILED|ERROR)|^E  |^rmi " gpurun_out/rmi_ops_$T.log | head -30
echo "=== rmi published size"
timeout -k 10 1200 python -m pytest -q -m gpu -p no:cacheprovider -s tests/test_rmi_gpu.py -k "resnet101" > gpurun_out/rmi_full_$T.log 2>&1
echo "exit $? : $(tail -n 2 gpurun_out/rmi_full_$T.log | tr '\n' ' ')"; grep -E "^(FAILED|ERROR)|^E  |^rmi " gpurun_out/rmi_full_$T.log | head -30
echo "=== bench rmi"; timeout -k 10 900 python bench.py --mode rmi --steps 5 --warmup 2 > gpurun_out/bench_rmi_$T.json 2> gpurun_out/bench_rmi_$T.err; tail -c 1500 gpurun_out/bench_rmi_$T.json; tail -n 8 gpurun_out/bench_rmi_$T.err
