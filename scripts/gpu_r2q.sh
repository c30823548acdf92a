#!/bin/bash
# Round 2 session q: instance-matching model -- batch-norm + relu in the conv epilogues, aligned recurrent operand
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T=${TAG:-r2q}
echo "=== rmi + lstm op tests"
timeout -k 10 1200 python -m pytest -q -m gpu -p no:cacheprovider -s tests/test_rmi_gpu.py > gpurun_out/rmi_$T.log 2>&1
echo "exit $? : $(tail -n 2 gpurun_out/rmi_$T.log | tr '\n' ' ')"; grep -E "^(FAILED|ERROR)|^E  |^rmi " gpurun_out/rmi_$T.log | head -30
timeout -k 10 600 python -m pytest -q -m gpu -p no:cacheprovider tests/test_ops_gpu.py -k "lstm or text" > gpurun_out/lstm_$T.log 2>&1
echo "exit $? : $(tail -n 2 gpurun_out/lstm_$T.log | tr '\n' ' ')"
echo "=== op breakdown"; timeout -k 10 600 python scripts/prof_rmi.py > gpurun_out/prof_rmi_$T.log 2>&1; head -n 60 gpurun_out/prof_rmi_$T.log
echo "=== bench rmi"; timeout -k 10 900 python bench.py --mode rmi --steps 5 --warmup 2 > gpurun_out/bench_rmi_$T.json 2> gpurun_out/bench_rmi_$T.err; tail -c 1500 gpurun_out/bench_rmi_$T.json; tail -n 8 gpurun_out/bench_rmi_$T.err
