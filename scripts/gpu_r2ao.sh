#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T=${TAG:-r2ao}
echo "=== conv tests"
timeout -k 10 900 python -m pytest tests/test_ops_gpu.py tests/test_production_shapes_gpu.py -x -q -m gpu -p no:cacheprovider -k "conv or small or stem or production" > gpurun_out/tests_$T.log 2>&1
echo "exit $? : $(tail -n 3 gpurun_out/tests_$T.log | tr '\n' ' ')"; grep -E "^(FAILED|ERROR)|^E  " gpurun_out/tests_$T.log | head -20
echo "=== direct narrow convolutions"; timeout -k 10 300 python scripts/prof_small.py 2>&1 | tee gpurun_out/prof_small_$T.log | tail -n 6
echo "=== bench"
timeout -k 10 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_$T.json 2> gpurun_out/bench_$T.err
python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/bench_$T.json").read().strip().splitlines()[-1])
    print(d["value"], d["ms_per_step"], d["e2e"]["value"], d["roofline"]["achieved"], d["roofline"]["frac"], d["clocks"], d["config"].get("loss_d"), d["config"].get("loss_g"))
except Exception as e:
    print("failed", e); print(open("gpurun_out/bench_$T.err").read()[-1500:])
PY
