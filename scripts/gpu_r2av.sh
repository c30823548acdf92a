#!/bin/bash
# Round 2, session 3, final state: whole GPU suite as the driver runs it, smoke(), every bench line, launch list
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T=${TAG:-r2av}
echo "=== full gpu suite"
timeout -k 10 1800 python -m pytest tests/ -x -q -m gpu -p no:cacheprovider > gpurun_out/suite_$T.log 2>&1
echo "exit $? : $(tail -n 3 gpurun_out/suite_$T.log | tr '\n' ' ')"; grep -E "^(FAILED|ERROR)|^E  " gpurun_out/suite_$T.log | head -20
echo "=== smoke"; timeout -k 10 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -n 3
echo "=== bench"; timeout -k 10 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_$T.json 2> gpurun_out/bench_$T.err; tail -c 1600 gpurun_out/bench_$T.json; tail -n 5 gpurun_out/bench_$T.err
echo "=== bench reference arm"; timeout -k 10 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_$T.json 2> gpurun_out/bench_ref_$T.err; tail -c 600 gpurun_out/bench_ref_$T.json
for m in infer bg rmi; do
  echo "=== bench --mode $m"; timeout -k 10 900 python bench.py --mode $m --steps 10 --warmup 3 > gpurun_out/bench_${m}_$T.json 2> gpurun_out/bench_${m}_$T.err; tail -c 700 gpurun_out/bench_${m}_$T.json; tail -n 3 gpurun_out/bench_${m}_$T.err
done
echo "=== ncu launches"
FGC_NCU_RANGE=1 timeout -k 10 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file gpurun_out/launches_$T.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench_$T.log 2>&1
wc -l gpurun_out/launches_$T.csv
