#!/bin/bash
# Round 2 session ad: HBM evidence for the streaming kernels added this round (CUDA-event GB/s + ncu DRAM bytes / throughput)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T=${TAG:-r2ad}
timeout -k 10 300 python scripts/prof_new_elem.py > gpurun_out/prof_new_elem_$T.log 2>&1; cat gpurun_out/prof_new_elem_$T.log
ONCE=1 timeout -k 10 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,dram__throughput.avg.pct_of_peak_sustained_elapsed --clock-control none \
   -k "regex:gate_prelu|prelu_bwd_kernel|affine_act|maxpool|space_batch|lstm_cell_fwd_vec4|pad_cast" --csv --log-file gpurun_out/ncu_new_elem_$T.csv python scripts/prof_new_elem.py > /dev/null 2>&1
wc -l gpurun_out/ncu_new_elem_$T.csv
