#!/bin/bash
# Round 2 session k: whole GPU suite + bench with the operand-exchanged 128-wide halo kernel; pair kernel test via env
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T=${TAG:-r2k}
echo "=== full gpu suite"
timeout -k 10 1800 python -m pytest tests/ -x -q -m gpu -p no:cacheprovider > gpurun_out/suite_$T.log 2>&1
echo "exit $? : $(tail -n 3 gpurun_out/suite_$T.log | tr '\n' ' ')"; grep -E "^(FAILED|ERROR)|^E  " gpurun_out/suite_$T.log | head -20
echo "=== pair kernel (FGC_HALO2=1): production shapes + conv ops"
FGC_HALO2=1 timeout -k 10 400 python -m pytest -q -m gpu -p no:cacheprovider tests/test_production_shapes_gpu.py tests/test_ops_gpu.py -k "production or (conv and tcgen05 and not gather)" > gpurun_out/pair_$T.log 2>&1
echo "exit $? : $(tail -n 2 gpurun_out/pair_$T.log | tr '\n' ' ')"
echo "=== bench"; timeout -k 10 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_$T.json 2> gpurun_out/bench_$T.err; tail -c 1500 gpurun_out/bench_$T.json; tail -n 5 gpurun_out/bench_$T.err
echo "=== ncu launches"
FGC_NCU_RANGE=1 timeout -k 10 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file gpurun_out/launches_$T.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench_$T.log 2>&1
wc -l gpurun_out/launches_$T.csv
