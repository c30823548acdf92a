#!/bin/bash
# Round 2 session h: the whole GPU suite as the driver runs it, timings of the streaming kernels, bench, compute-sanitizer pass
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T=${TAG:-r2h}
echo "=== full gpu suite"
timeout -k 10 1500 python -m pytest tests/ -x -q -m gpu -p no:cacheprovider > gpurun_out/suite_$T.log 2>&1
echo "exit $? : $(tail -n 3 gpurun_out/suite_$T.log | tr '\n' ' ')"; grep -E "^(FAILED|ERROR)|^E  " gpurun_out/suite_$T.log | head -20
echo "=== prof_elem"; timeout -k 10 300 python scripts/prof_elem.py > gpurun_out/prof_elem_$T.log 2>&1; cat gpurun_out/prof_elem_$T.log
echo "=== bench"; timeout -k 10 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_$T.json 2> gpurun_out/bench_$T.err; tail -c 1200 gpurun_out/bench_$T.json; tail -n 5 gpurun_out/bench_$T.err
echo "=== compute-sanitizer memcheck (small op tests)"
timeout -k 10 900 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest -x -q -m gpu -p no:cacheprovider tests/test_ops_gpu.py -k "not conv_fwd and not conv_dgrad and not conv_wgrad and not spectral" > gpurun_out/sanitizer_memcheck_$T.log 2>&1
echo "exit $? : $(tail -n 4 gpurun_out/sanitizer_memcheck_$T.log | tr '\n' ' ')"
timeout -k 10 900 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest -x -q -m gpu -p no:cacheprovider tests/test_ops_gpu.py -k "conv_fwd and tcgen05 and not gather" > gpurun_out/sanitizer_memcheck_conv_$T.log 2>&1
echo "exit $? : $(tail -n 4 gpurun_out/sanitizer_memcheck_conv_$T.log | tr '\n' ' ')"
echo "=== compute-sanitizer racecheck (elementwise / reduction / LSTM kernels)"
timeout -k 10 900 compute-sanitizer --tool racecheck --error-exitcode 7 python -m pytest -x -q -m gpu -p no:cacheprovider tests/test_ops_gpu.py -k "cbn or minmax or text_ops or word_lstm or gating or pool" > gpurun_out/sanitizer_racecheck_$T.log 2>&1
echo "exit $? : $(tail -n 4 gpurun_out/sanitizer_racecheck_$T.log | tr '\n' ' ')"
