#!/bin/bash
# Round 2 session c: word-LSTM sequence kernels -- op tests, model-level parity, op breakdown, bench, ncu launch list.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T=${TAG:-r2c}
run() {
  local name=$1 to=$2; shift 2
  echo "=== $name"
  timeout -k 10 "$to" python -m pytest -q -rP -m gpu -p no:cacheprovider "$@" > "gpurun_out/${name}_$T.log" 2>&1
  echo "exit $? : $(tail -n 2 gpurun_out/${name}_$T.log | tr '\n' ' ')"; grep -E "^(FAILED|ERROR)|rel err|rel-to-max|max-abs|Error|vs fp64|means" "gpurun_out/${name}_$T.log" | head -30
}
run lstmops 200 tests/test_ops_gpu.py -k "word_lstm or text_ops"
run model 600 tests/test_model_gpu.py
run variants 400 tests/test_pix2pix_gpu.py tests/test_residual_gpu.py tests/test_bg_gpu.py -k "inference or training or bf16 or generator"
echo "=== smoke"; timeout -k 10 300 python __graft_entry__.py --smoke 2>&1 | tail -n 3
echo "=== op_breakdown"; timeout -k 10 300 python scripts/op_breakdown.py > gpurun_out/op_breakdown_$T.log 2>&1; head -n 52 gpurun_out/op_breakdown_$T.log
echo "=== bench"; timeout -k 10 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_$T.json 2> gpurun_out/bench_$T.err; tail -c 3000 gpurun_out/bench_$T.json; tail -n 5 gpurun_out/bench_$T.err
echo "=== ncu launches"
FGC_NCU_RANGE=1 timeout -k 10 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file gpurun_out/launches_$T.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench_$T.log 2>&1
wc -l gpurun_out/launches_$T.csv
