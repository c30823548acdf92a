#!/bin/bash
# Round 2, session 3: row-tiled patch-tensor kernel (fgc_im2col_small) -- bit-exact tests, kernel timing old / new, step A/B
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T=${TAG:-r2ag}
echo "=== targeted tests"
timeout -k 10 900 python -m pytest tests/test_ops_gpu.py tests/test_production_shapes_gpu.py -x -q -m gpu -p no:cacheprovider \
    -k "patch or keep_packed or conv or head" > gpurun_out/tests_$T.log 2>&1
echo "exit $? : $(tail -n 3 gpurun_out/tests_$T.log | tr '\n' ' ')"; grep -E "^(FAILED|ERROR)|^E  " gpurun_out/tests_$T.log | head -20
echo "=== patch kernel timing"
for rows in 0 1; do
FGC_IM2COL_ROWS=$rows timeout -k 10 300 python - <<'PY' 2>&1 | tail -n 8
import os, torch
from sketchyscenecolorization_b200.cuda_ops import CudaOps
ops = CudaOps("cuda:0", torch.bfloat16)
print("FGC_IM2COL_ROWS =", os.environ.get("FGC_IM2COL_ROWS"))
for (N, H, W, C, k) in [(64, 192, 192, 3, 7), (128, 192, 192, 3, 3), (128, 192, 192, 8, 3), (64, 96, 96, 8, 3), (64, 96, 96, 3, 3)]:
    x = torch.randn(N, H, W, C, device="cuda").bfloat16()
    for _ in range(3):
        p = ops.small_patch(x, k)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        p = ops.small_patch(x, k)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    by = p.numel() * 2 + x.numel() * 2
    print("  N%d %dx%d C%d k%d: %.1f us, %.2f TB/s (%.0f MB)" % (N, H, W, C, k, ms * 1e3, by / ms / 1e9, by / 1e6))
PY
done
echo "=== bench A/B"
for rows in 0 1; do
  FGC_IM2COL_ROWS=$rows timeout -k 10 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_${T}_rows$rows.json 2> gpurun_out/bench_${T}_rows$rows.err
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/bench_${T}_rows$rows.json").read().strip().splitlines()[-1])
    print("rows=$rows", d["value"], d["ms_per_step"], d["e2e"]["value"], d["roofline"]["achieved"], d["roofline"]["frac"], d["roofline"].get("achieved_incl_weight_pack"), d["clocks"])
except Exception as e:
    print("rows=$rows failed", e); print(open("gpurun_out/bench_${T}_rows$rows.err").read()[-1500:])
PY
done
