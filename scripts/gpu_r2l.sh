#!/bin/bash
# Round 2 session l: LSTM kernels after the cp.async change, ncu evidence for the final kernel set, bench
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T=${TAG:-r2l}
echo "=== lstm ops + model"
timeout -k 10 600 python -m pytest -q -m gpu -p no:cacheprovider tests/test_ops_gpu.py tests/test_model_gpu.py -k "word_lstm or text_ops or inference_parity or training_graph_gradients" > gpurun_out/lstm_$T.log 2>&1
echo "exit $? : $(tail -n 2 gpurun_out/lstm_$T.log | tr '\n' ' ')"; grep -E "^(FAILED|ERROR)|^E  " gpurun_out/lstm_$T.log | head
echo "=== ncu full: dominant kernel"
ONLY_FIRST=1 REPS=1 timeout -k 10 300 ncu --set full --clock-control none --import-source on -k "regex:conv_halo_kernel" -c 2 -f -o gpurun_out/prof_dominant_$T python scripts/prof_conv.py > gpurun_out/ncu_dominant_$T.log 2>&1
echo "=== G conv stack tensor pipe"
timeout -k 10 900 ncu --profile-from-start off --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed --clock-control none --csv \
    --log-file gpurun_out/g_stack_$T.csv python scripts/g_conv_stack.py > gpurun_out/g_stack_$T.log 2>&1
wc -l gpurun_out/g_stack_$T.csv
echo "=== op breakdown"; timeout -k 10 300 python scripts/op_breakdown.py > gpurun_out/op_breakdown_$T.log 2>&1; head -n 40 gpurun_out/op_breakdown_$T.log
echo "=== bench"; timeout -k 10 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_$T.json 2> gpurun_out/bench_$T.err; tail -c 700 gpurun_out/bench_$T.json; tail -n 5 gpurun_out/bench_$T.err
