#!/bin/bash
# Round-end check in the driver's own form: the whole GPU suite in ONE process, then the input-path timing, one ncu capture of
# its kernels, and smoke().
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T=${TAG:-cur}
echo "=== pytest -m gpu (single process)"; timeout -k 5 85 python -m pytest tests/ -x -q -m gpu -p no:cacheprovider > gpurun_out/gpu_suite_$T.log 2>&1
echo "exit $? : $(tail -n 3 gpurun_out/gpu_suite_$T.log | tr '\n' ' ')"; grep -E "^(FAILED|ERROR)|Error|diff" gpurun_out/gpu_suite_$T.log | head -20
echo "=== prof_input"; timeout -k 5 20 python scripts/prof_input.py > gpurun_out/prof_input_$T.log 2>&1; cat gpurun_out/prof_input_$T.log
echo "=== ncu full (input)"
ONLY_FAST=1 REPS=1 NBUF=1 timeout -k 5 25 ncu --set full --clock-control none --import-source on -k "regex:paired_" -c 2 -f -o gpurun_out/prof_input_$T \
    python scripts/prof_input.py > gpurun_out/ncu_input_$T.log 2>&1
ls -la gpurun_out/prof_input_$T.ncu-rep
echo "=== smoke"; timeout -k 5 25 python __graft_entry__.py --smoke > gpurun_out/smoke_$T.log 2>&1; tail -n 2 gpurun_out/smoke_$T.log
