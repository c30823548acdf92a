#!/bin/bash
# Round 2 session ab: generator conv-stack tensor-pipe list of the final kernels (the second half of BASELINE's metric)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T=${TAG:-r2ab}
timeout -k 10 900 ncu --profile-from-start off --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed --clock-control none --csv \
    --log-file gpurun_out/g_stack_$T.csv python scripts/g_conv_stack.py > gpurun_out/g_stack_$T.log 2>&1
wc -l gpurun_out/g_stack_$T.csv
python scripts/tensor_pipe_summary.py gpurun_out/g_stack_$T.csv > gpurun_out/g_conv_tensor_pipe_$T.txt; head -n 12 gpurun_out/g_conv_tensor_pipe_$T.txt
