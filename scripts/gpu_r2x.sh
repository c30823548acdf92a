#!/bin/bash
# Round 2 session x: input prefetch of the session loop (next batch over PCIe under the current step); sanitizer on the new kernels
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T=${TAG:-r2x}
echo "=== entry / model / input-queue tests"
timeout -k 10 900 python -m pytest -q -m gpu -p no:cacheprovider tests/test_entry_gpu.py tests/test_model_gpu.py tests/test_tfrecord_gpu.py -x > gpurun_out/entry_$T.log 2>&1
echo "exit $? : $(tail -n 2 gpurun_out/entry_$T.log | tr '\n' ' ')"; grep -E "^(FAILED|ERROR)|^E  " gpurun_out/entry_$T.log | head
for v in 1 0 1 0; do
  echo "=== bench FGC_INPUT_PREFETCH=$v"
  FGC_INPUT_PREFETCH=$v timeout -k 10 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_${T}_pf$v.json 2> gpurun_out/bench_${T}_pf$v.err
  python -c "import sys,json; d=json.loads(open('gpurun_out/bench_${T}_pf$v.json').read().strip().splitlines()[-1]); print('value', d['value'], 'ms', d['ms_per_step'], 'e2e', d['e2e']['value'], d['clocks']['sm_mhz'])"; tail -n 3 gpurun_out/bench_${T}_pf$v.err
done
echo "=== bench --input tfrecord"
timeout -k 10 900 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --input tfrecord > gpurun_out/bench_${T}_tfrecord.json 2> gpurun_out/bench_${T}_tfrecord.err
python -c "import sys,json; d=json.loads(open('gpurun_out/bench_${T}_tfrecord.json').read().strip().splitlines()[-1]); print('value', d['value'], 'e2e', d['e2e'])"; tail -n 3 gpurun_out/bench_${T}_tfrecord.err
echo "=== compute-sanitizer memcheck: trunk / LSTM cell / folded kernels"
timeout -k 10 900 compute-sanitizer --tool memcheck --error-exitcode 3 python -m pytest -q -m gpu -p no:cacheprovider -x tests/test_rmi_gpu.py -k "affine or maxpool or space_batch or resize or pad_cast or recurrent or small_64px" > gpurun_out/sanitizer_memcheck_rmi_$T.log 2>&1
echo "exit $? : $(tail -n 3 gpurun_out/sanitizer_memcheck_rmi_$T.log | tr '\n' ' ')"
timeout -k 10 900 compute-sanitizer --tool memcheck --error-exitcode 3 python -m pytest -q -m gpu -p no:cacheprovider -x tests/test_ops_gpu.py -k "conv_fwd_into or minmax_in_sample or (conv_dgrad and tcgen05 and not gather and (16 or 17)) or text_ops or lstm" > gpurun_out/sanitizer_memcheck_ops_$T.log 2>&1
echo "exit $? : $(tail -n 3 gpurun_out/sanitizer_memcheck_ops_$T.log | tr '\n' ' ')"
