#!/bin/bash
# FIRST GPU session of the next round: run everything that was written after round 1's GPU budget ended, then bench lines for
# the new variants.  Each tier in its own process with its own timeout; results under gpurun_out/.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T=${TAG:-r2a}
run() {  # name, timeout, args...
  local name=$1 to=$2; shift 2
  echo "=== $name"
  FGC_UNVERIFIED=1 timeout -k 10 "$to" python -m pytest -v -rP -m gpu -p no:cacheprovider "$@" > "gpurun_out/${name}_$T.log" 2>&1
  echo "exit $? : $(tail -n 3 gpurun_out/${name}_$T.log | tr '\n' ' ')"; grep -E "^(FAILED|ERROR)|rel err|max-abs err|Error|yardstick" "gpurun_out/${name}_$T.log" | head -20
}
run input_queue 200 tests/test_tfrecord_gpu.py
run pix2pix_model 300 tests/test_pix2pix_gpu.py -k "inference or training or bf16"
run residual 400 tests/test_residual_gpu.py
run bg 600 tests/test_bg_gpu.py
run optimizers 120 tests/test_optimizers_gpu.py
for bt in Pix2Pix Residual; do
  echo "=== bench --block-type $bt"
  timeout -k 10 600 python bench.py --steps 5 --warmup 3 --block-type $bt > gpurun_out/bench_${bt}_$T.json 2> gpurun_out/bench_${bt}_$T.err
  tail -c 2500 gpurun_out/bench_${bt}_$T.json; tail -n 5 gpurun_out/bench_${bt}_$T.err
done
echo "=== bench, e2e leg from TFRecord files (MRU)"
timeout -k 10 600 python bench.py --steps 5 --warmup 3 --input tfrecord --no-cpu-baseline > gpurun_out/bench_tfrecord_$T.json 2> gpurun_out/bench_tfrecord_$T.err
tail -c 2500 gpurun_out/bench_tfrecord_$T.json; tail -n 5 gpurun_out/bench_tfrecord_$T.err
echo "=== the same with the raw-batch copies on a side stream"
FGC_INPUT_SIDE_STREAM=1 timeout -k 10 600 python bench.py --steps 5 --warmup 3 --input tfrecord --no-cpu-baseline > gpurun_out/bench_tfrecord_side_$T.json 2> gpurun_out/bench_tfrecord_side_$T.err
tail -c 1200 gpurun_out/bench_tfrecord_side_$T.json; tail -n 5 gpurun_out/bench_tfrecord_side_$T.err
echo "=== bg 768 timing"; timeout -k 10 300 python scripts/prof_bg.py > gpurun_out/prof_bg_$T.log 2>&1; cat gpurun_out/prof_bg_$T.log
