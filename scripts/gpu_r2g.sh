#!/bin/bash
# Round 2 session g: six-product mode (Residual / BG parity), caller tests on the device, clean bench, G conv-stack tensor pipe
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T=${TAG:-r2g}
run() {
  local name=$1 to=$2; shift 2
  echo "=== $name"
  timeout -k 10 "$to" python -m pytest -q -rP -m gpu -p no:cacheprovider "$@" > "gpurun_out/${name}_$T.log" 2>&1
  echo "exit $? : $(tail -n 2 gpurun_out/${name}_$T.log | tr '\n' ' ')"; grep -E "^(FAILED|ERROR)|^E  |rel err|rel-to-max|max-abs|Error|vs fp64|means|grey" "gpurun_out/${name}_$T.log" | head -30
}
run sixproduct 600 tests/test_residual_gpu.py tests/test_bg_gpu.py -k "inference"
run callers 400 tests/test_callers_gpu.py
echo "=== bench"; timeout -k 10 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_$T.json 2> gpurun_out/bench_$T.err; tail -c 1500 gpurun_out/bench_$T.json; tail -n 5 gpurun_out/bench_$T.err
echo "=== G conv stack tensor pipe"
timeout -k 10 900 ncu --profile-from-start off --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed --clock-control none --csv \
    --log-file gpurun_out/g_stack_$T.csv python scripts/g_conv_stack.py > gpurun_out/g_stack_$T.log 2>&1
wc -l gpurun_out/g_stack_$T.csv; python scripts/tensor_pipe_summary.py gpurun_out/g_stack_$T.csv | head -12
