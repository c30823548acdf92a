#!/bin/bash
# Round 2, second GPU session: entry-point tests, production-shape parity, full-size gradients in both modes, bench through TrainSession.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T=${TAG:-r2b}
run() {
  local name=$1 to=$2; shift 2
  echo "=== $name"
  timeout -k 10 "$to" python -m pytest -v -rP -m gpu -p no:cacheprovider "$@" > "gpurun_out/${name}_$T.log" 2>&1
  echo "exit $? : $(tail -n 3 gpurun_out/${name}_$T.log | tr '\n' ' ')"; grep -E "^(FAILED|ERROR)|rel err|max-abs|Error|vs fp64|means|oracle" "gpurun_out/${name}_$T.log" | head -30
}
run entry 300 tests/test_entry_gpu.py
run prodshapes 400 tests/test_production_shapes_gpu.py
run fullsize 600 tests/test_model_gpu.py -k "full_size or stated or trains_like"
run optimizers 100 tests/test_optimizers_gpu.py
echo "=== bench"
timeout -k 10 300 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_$T.json 2> gpurun_out/bench_$T.err
tail -c 3000 gpurun_out/bench_$T.json; tail -n 5 gpurun_out/bench_$T.err
