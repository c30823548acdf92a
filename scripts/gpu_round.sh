#!/bin/bash
# One GPU session: tests, smoke, per-op breakdown, conv timings, bench, ncu launch list, ncu full capture of the dominant conv.
# env: TIERS (test tiers), STEPS, NO_NCU, NO_TESTS, TAG (suffix of the files written under gpurun_out/)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T=${TAG:-cur}
if [ -z "$NO_TESTS" ]; then
bash scripts/gpu_tests.sh conv_fwd conv_dgrad conv_wgrad
# a failing / hanging new conv kernel must not take the rest of the session down: fall back to the older kernels
if ! tail -n 1 gpurun_out/conv_fwd.log | grep -q passed || grep -q failed gpurun_out/conv_fwd.log || \
   ! tail -n 1 gpurun_out/conv_dgrad.log | grep -q passed || grep -q failed gpurun_out/conv_dgrad.log || \
   ! tail -n 1 gpurun_out/conv_wgrad.log | grep -q passed || grep -q failed gpurun_out/conv_wgrad.log; then
  echo "!!! conv tiers not green: details"; grep -E "^(FAILED|ERROR)|Error|rel-to-max" gpurun_out/conv_fwd.log gpurun_out/conv_dgrad.log gpurun_out/conv_wgrad.log | head -40
  for v in "FGC_HALO=0" "FGC_SMALL=0" "FGC_HALO=0 FGC_SMALL=0"; do
    echo "--- retry with $v"; env $v timeout -k 10 200 python -m pytest -q -m gpu -p no:cacheprovider tests/test_ops_gpu.py -k "tcgen05 and not gather and conv" -x 2>&1 | tail -n 3
  done
  export FGC_HALO=${FALLBACK_HALO:-0} FGC_SMALL=${FALLBACK_SMALL:-0}
  echo "continuing with FGC_HALO=$FGC_HALO FGC_SMALL=$FGC_SMALL"
fi
bash scripts/gpu_tests.sh ops_base model input variants
echo "=== smoke"; timeout -k 10 300 python __graft_entry__.py --smoke > gpurun_out/smoke_$T.log 2>&1; tail -n 3 gpurun_out/smoke_$T.log
fi
echo "=== prof_conv"; timeout -k 10 300 python scripts/prof_conv.py > gpurun_out/prof_conv_$T.log 2>&1; cat gpurun_out/prof_conv_$T.log
echo "=== prof_elem"; timeout -k 10 300 python scripts/prof_elem.py > gpurun_out/prof_elem_$T.log 2>&1; cat gpurun_out/prof_elem_$T.log
echo "=== op_breakdown"; timeout -k 10 300 python scripts/op_breakdown.py > gpurun_out/op_breakdown_$T.log 2>&1; head -n 60 gpurun_out/op_breakdown_$T.log
echo "=== bench"; timeout -k 10 900 python bench.py --steps ${STEPS:-5} --warmup 3 > gpurun_out/bench_$T.json 2> gpurun_out/bench_$T.err; tail -c 3000 gpurun_out/bench_$T.json; tail -n 5 gpurun_out/bench_$T.err
if [ -z "$NO_NCU" ]; then
echo "=== ncu launches (timed region only: cudaProfilerStart/Stop around it)"
FGC_NCU_RANGE=1 timeout -k 10 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file gpurun_out/launches_$T.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench_$T.log 2>&1
wc -l gpurun_out/launches_$T.csv
if [ "$(wc -l < gpurun_out/launches_$T.csv)" -lt 200 ]; then
  echo "graph replay gave no per-kernel list; repeating with --no-graphs"
  FGC_NCU_RANGE=1 timeout -k 10 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
      --log-file gpurun_out/launches_$T.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-graphs > gpurun_out/ncu_bench_$T.log 2>&1
  wc -l gpurun_out/launches_$T.csv
fi
echo "=== ncu full"
ONLY_FIRST=1 REPS=1 timeout -k 10 600 ncu --set full --clock-control none --import-source on -k "regex:conv_halo_kernel|conv_wgrad_halo_kernel|conv_igemm|conv_wgrad_kernel" -c 6 -f -o gpurun_out/prof_conv_$T \
    python scripts/prof_conv.py > gpurun_out/ncu_full_$T.log 2>&1
echo "=== ncu full (elementwise)"
ONLY=cbn_act_fwd,cbn_act_bwd,minmax_fwd,gate_fma_fwd,blend_fwd REPS=1 timeout -k 10 600 ncu --set full --clock-control none --import-source on \
    -k "regex:cbn_|minmax_|gate_fma|blend_" -c 12 -f -o gpurun_out/prof_elem_$T python scripts/prof_elem.py > gpurun_out/ncu_elem_$T.log 2>&1
ls -la gpurun_out/
fi
