#!/bin/bash
# Round 2, session 3: division-free tile staging of the direct narrow convolutions, per-pixel tapsum_w: tests, step
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T=${TAG:-r2an}
echo "=== op + production-shape tests"
timeout -k 10 900 python -m pytest tests/test_ops_gpu.py tests/test_production_shapes_gpu.py -x -q -m gpu -p no:cacheprovider > gpurun_out/tests_$T.log 2>&1
echo "exit $? : $(tail -n 3 gpurun_out/tests_$T.log | tr '\n' ' ')"; grep -E "^(FAILED|ERROR)|^E  " gpurun_out/tests_$T.log | head -20
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
echo "=== ncu launches"
FGC_NCU_RANGE=1 timeout -k 10 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file gpurun_out/launches_$T.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench_$T.log 2>&1
wc -l gpurun_out/launches_$T.csv
