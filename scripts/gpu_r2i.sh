#!/bin/bash
# Round 2 session i: CTA-pair halo kernel -- a quick production-shape check first (short timeouts: a wrong barrier hangs), then tests / timings
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T=${TAG:-r2i}
echo "=== prof_conv first layer, pair kernel"; ONLY_FIRST=1 timeout -k 5 60 python scripts/prof_conv.py 2>&1 | tail -5
echo "=== prof_conv first layer, single-CTA kernel"; FGC_HALO2=0 ONLY_FIRST=1 timeout -k 5 60 python scripts/prof_conv.py 2>&1 | tail -4
echo "=== production shapes"
timeout -k 10 300 python -m pytest -q -rP -m gpu -p no:cacheprovider tests/test_production_shapes_gpu.py > gpurun_out/prodshapes_$T.log 2>&1
echo "exit $? : $(tail -n 2 gpurun_out/prodshapes_$T.log | tr '\n' ' ')"; grep -E "^(FAILED|ERROR)|^E  " gpurun_out/prodshapes_$T.log | head -20
echo "=== conv op tests"
timeout -k 10 300 python -m pytest -q -m gpu -p no:cacheprovider tests/test_ops_gpu.py -k "conv and tcgen05 and not gather" > gpurun_out/convops_$T.log 2>&1
echo "exit $? : $(tail -n 2 gpurun_out/convops_$T.log | tr '\n' ' ')"; grep -E "^(FAILED|ERROR)|^E  " gpurun_out/convops_$T.log | head -20
echo "=== prof_conv"; timeout -k 10 300 python scripts/prof_conv.py > gpurun_out/prof_conv_$T.log 2>&1; cut -c1-60 gpurun_out/prof_conv_$T.log
echo "=== bench"; timeout -k 10 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_$T.json 2> gpurun_out/bench_$T.err; tail -c 900 gpurun_out/bench_$T.json; tail -n 5 gpurun_out/bench_$T.err
