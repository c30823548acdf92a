#!/bin/bash
# Round 2 session n: encoder-cell 1x1 skip at low resolution + column-folded narrow input gradients
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T=${TAG:-r2n}
echo "=== new op tests"
timeout -k 10 600 python -m pytest -q -m gpu -p no:cacheprovider tests/test_ops_gpu.py -k "conv_fwd_into or conv_dgrad" > gpurun_out/ops_$T.log 2>&1
echo "exit $? : $(tail -n 2 gpurun_out/ops_$T.log | tr '\n' ' ')"; grep -E "^(FAILED|ERROR)|^E  " gpurun_out/ops_$T.log | head
echo "=== model tests"
timeout -k 10 900 python -m pytest -q -m gpu -p no:cacheprovider tests/test_model_gpu.py tests/test_production_shapes_gpu.py > gpurun_out/model_$T.log 2>&1
echo "exit $? : $(tail -n 2 gpurun_out/model_$T.log | tr '\n' ' ')"; grep -E "^(FAILED|ERROR)|^E  " gpurun_out/model_$T.log | head
echo "=== bench"; timeout -k 10 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_$T.json 2> gpurun_out/bench_$T.err; tail -c 900 gpurun_out/bench_$T.json; tail -n 5 gpurun_out/bench_$T.err
echo "=== bench without the folded narrow gradients"; FGC_FOLD_DGRAD=0 timeout -k 10 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'])"
