#!/bin/bash
# Round 2 session f: tuned streaming / reduction kernels -- op tests, timings, model parity, bench, launch list
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T=${TAG:-r2f}
run() {
  local name=$1 to=$2; shift 2
  echo "=== $name"
  timeout -k 10 "$to" python -m pytest -q -rP -m gpu -p no:cacheprovider "$@" > "gpurun_out/${name}_$T.log" 2>&1
  echo "exit $? : $(tail -n 2 gpurun_out/${name}_$T.log | tr '\n' ' ')"; grep -E "^(FAILED|ERROR)|rel err|rel-to-max|max-abs|Error|vs fp64|means" "gpurun_out/${name}_$T.log" | head -30
}
run ops 400 tests/test_ops_gpu.py -k "not conv"
echo "=== prof_elem"; timeout -k 10 300 python scripts/prof_elem.py > gpurun_out/prof_elem_$T.log 2>&1; cat gpurun_out/prof_elem_$T.log
run model 600 tests/test_model_gpu.py
echo "=== bench"; timeout -k 10 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_$T.json 2> gpurun_out/bench_$T.err; tail -c 1800 gpurun_out/bench_$T.json; tail -n 5 gpurun_out/bench_$T.err
echo "=== ncu launches"
FGC_NCU_RANGE=1 timeout -k 10 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file gpurun_out/launches_$T.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench_$T.log 2>&1
wc -l gpurun_out/launches_$T.csv
