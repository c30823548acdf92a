#!/bin/bash
# Round 2 session d: operand-exchanged halo kernel (128 channels x 256 pixels) -- conv tests, production shapes, timings, bench
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T=${TAG:-r2d}
run() {
  local name=$1 to=$2; shift 2
  echo "=== $name"
  timeout -k 10 "$to" python -m pytest -q -rP -m gpu -p no:cacheprovider "$@" > "gpurun_out/${name}_$T.log" 2>&1
  echo "exit $? : $(tail -n 2 gpurun_out/${name}_$T.log | tr '\n' ' ')"; grep -E "^(FAILED|ERROR)|rel err|rel-to-max|max-abs|Error|vs fp64|means" "gpurun_out/${name}_$T.log" | head -30
}
run convops 400 tests/test_ops_gpu.py -k "conv and tcgen05 and not gather"
run prodshapes 400 tests/test_production_shapes_gpu.py
echo "=== prof_conv (swap on)"; timeout -k 10 300 python scripts/prof_conv.py > gpurun_out/prof_conv_$T.log 2>&1; cat gpurun_out/prof_conv_$T.log
echo "=== prof_conv (swap off)"; FGC_HALO_SWAP=0 ONLY_FIRST=3 timeout -k 10 300 python scripts/prof_conv.py 2>&1 | head -12
run model 600 tests/test_model_gpu.py -k "inference or gradients or graph"
echo "=== bench"; timeout -k 10 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_$T.json 2> gpurun_out/bench_$T.err; tail -c 1800 gpurun_out/bench_$T.json; tail -n 5 gpurun_out/bench_$T.err
