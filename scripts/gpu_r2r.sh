#!/bin/bash
# Round 2 session r: whole GPU suite as the driver runs it, smoke(), bench (train + rmi), launch list of the final state
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T=${TAG:-r2r}
echo "=== full gpu suite"
timeout -k 10 1800 python -m pytest tests/ -x -q -m gpu -p no:cacheprovider > gpurun_out/suite_$T.log 2>&1
echo "exit $? : $(tail -n 3 gpurun_out/suite_$T.log | tr '\n' ' ')"; grep -E "^(FAILED|ERROR)|^E  " gpurun_out/suite_$T.log | head -20
echo "=== smoke"; timeout -k 10 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -n 3
echo "=== bench rmi"; timeout -k 10 900 python bench.py --mode rmi --steps 5 --warmup 2 > gpurun_out/bench_rmi_$T.json 2> gpurun_out/bench_rmi_$T.err; tail -c 900 gpurun_out/bench_rmi_$T.json; tail -n 5 gpurun_out/bench_rmi_$T.err
timeout -k 10 600 python scripts/prof_rmi.py > gpurun_out/prof_rmi_$T.log 2>&1; head -n 12 gpurun_out/prof_rmi_$T.log
echo "=== bench"; timeout -k 10 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_$T.json 2> gpurun_out/bench_$T.err; tail -c 1500 gpurun_out/bench_$T.json; tail -n 5 gpurun_out/bench_$T.err
echo "=== ncu launches"
FGC_NCU_RANGE=1 timeout -k 10 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file gpurun_out/launches_$T.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench_$T.log 2>&1
wc -l gpurun_out/launches_$T.csv
