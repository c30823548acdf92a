#!/bin/bash
# Round 2, first GPU session: every test that never ran on hardware (FGC_UNVERIFIED gates), then bench lines of the variants.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T=${TAG:-r2a}
run() {  # name, timeout, args...
  local name=$1 to=$2; shift 2
  echo "=== $name"
  FGC_UNVERIFIED=1 timeout -k 10 "$to" python -m pytest -v -rP -m gpu -p no:cacheprovider "$@" > "gpurun_out/${name}_$T.log" 2>&1
  echo "exit $? : $(tail -n 3 gpurun_out/${name}_$T.log | tr '\n' ' ')"; grep -E "^(FAILED|ERROR)|rel err|max-abs err|Error|yardstick" "gpurun_out/${name}_$T.log" | head -20
}
run input_queue 150 tests/test_tfrecord_gpu.py
run pix2pix_model 240 tests/test_pix2pix_gpu.py -k "inference or training or bf16"
run residual 300 tests/test_residual_gpu.py
run bg 300 tests/test_bg_gpu.py
run optimizers 100 tests/test_optimizers_gpu.py
for bt in Pix2Pix Residual; do
  echo "=== bench --block-type $bt"
  timeout -k 10 240 python bench.py --steps 5 --warmup 3 --block-type $bt --no-cpu-baseline > gpurun_out/bench_${bt}_$T.json 2> gpurun_out/bench_${bt}_$T.err
  tail -c 2500 gpurun_out/bench_${bt}_$T.json; tail -n 5 gpurun_out/bench_${bt}_$T.err
done
echo "=== bg 768 timing"; timeout -k 10 200 python scripts/prof_bg.py > gpurun_out/prof_bg_$T.log 2>&1; tail -20 gpurun_out/prof_bg_$T.log
