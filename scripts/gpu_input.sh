#!/bin/bash
# Short GPU session for the real-data input path: parity tests, timing, one ncu --set full capture of its kernels.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T=${TAG:-cur}
echo "=== tfrecord_gpu"; timeout -k 10 150 python -m pytest -q -m gpu -p no:cacheprovider tests/test_tfrecord_gpu.py > gpurun_out/tfrecord_gpu_$T.log 2>&1
echo "exit $? : $(tail -n 3 gpurun_out/tfrecord_gpu_$T.log | tr '\n' ' ')"; grep -E "^(FAILED|ERROR)|Error|diff" gpurun_out/tfrecord_gpu_$T.log | head -20
echo "=== prof_input"; timeout -k 10 100 python scripts/prof_input.py > gpurun_out/prof_input_$T.log 2>&1; cat gpurun_out/prof_input_$T.log
if [ -z "$NO_NCU" ]; then
echo "=== ncu full (input)"
REPS=1 NBUF=2 timeout -k 10 150 ncu --set full --clock-control none --import-source on -k "regex:paired_" -c 6 -f -o gpurun_out/prof_input_$T \
    python scripts/prof_input.py > gpurun_out/ncu_input_$T.log 2>&1
tail -n 3 gpurun_out/ncu_input_$T.log; ls -la gpurun_out/prof_input_$T.ncu-rep
fi
