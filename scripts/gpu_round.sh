#!/bin/bash
# One GPU session: tests, diagnostics, smoke, bench, ncu launch list, ncu full capture of the dominant conv.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
bash scripts/gpu_tests.sh ${TIERS:-ops_base model}
echo "=== diag"; timeout -k 10 300 python scripts/diag_grads.py > gpurun_out/diag.log 2>&1; tail -n 30 gpurun_out/diag.log
echo "=== smoke"; timeout -k 10 300 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; tail -n 3 gpurun_out/smoke.log
echo "=== prof_conv"; timeout -k 10 300 python scripts/prof_conv.py > gpurun_out/prof_conv.log 2>&1; cat gpurun_out/prof_conv.log
echo "=== bench"; timeout -k 10 900 python bench.py --steps ${STEPS:-3} --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -c 3000 gpurun_out/bench.json; tail -n 5 gpurun_out/bench.err
if [ -z "$NO_NCU" ]; then
echo "=== ncu launches"
timeout -k 10 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 12000 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
wc -l gpurun_out/launches.csv
echo "=== ncu full"
ONLY_FIRST=1 REPS=1 timeout -k 10 600 ncu --set full --clock-control none --import-source on -k regex:conv_ -c 6 -f -o gpurun_out/prof_conv \
    python scripts/prof_conv.py > gpurun_out/ncu_full.log 2>&1
ls -la gpurun_out/
fi
