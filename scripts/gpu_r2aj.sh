#!/bin/bash
# Round 2, session 3: row-walking streaming kernels with grouped raw loads (walk_rows): op tests, family timing, step
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T=${TAG:-r2aj}
echo "=== op tests"
timeout -k 10 900 python -m pytest tests/test_ops_gpu.py -x -q -m gpu -p no:cacheprovider > gpurun_out/tests_$T.log 2>&1
echo "exit $? : $(tail -n 3 gpurun_out/tests_$T.log | tr '\n' ' ')"; grep -E "^(FAILED|ERROR)|^E  " gpurun_out/tests_$T.log | head -20
echo "=== elementwise family"; REPS=5 timeout -k 10 300 python scripts/prof_elem.py 2>&1 | tee gpurun_out/prof_elem_$T.log | tail -n 14
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
