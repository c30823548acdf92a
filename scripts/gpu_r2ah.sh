#!/bin/bash
# Round 2, session 3: (1) read-only streaming sweep (what a reduction pass needs to reach the HBM read rate), (2) patch kernel with
# vector staging + generic column sums: tests + timing, (3) elementwise family at the production shape
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T=${TAG:-r2ah}
echo "=== read sweep"; timeout -k 5 120 scripts/microbench/readbw 2>&1 | tee gpurun_out/readbw_$T.log | tail -n 25
echo "=== tests"
timeout -k 10 600 python -m pytest tests/test_ops_gpu.py -x -q -m gpu -p no:cacheprovider -k "patch or colsum or prelu" > gpurun_out/tests_$T.log 2>&1
echo "exit $? : $(tail -n 3 gpurun_out/tests_$T.log | tr '\n' ' ')"; grep -E "^(FAILED|ERROR)|^E  " gpurun_out/tests_$T.log | head -20
echo "=== patch kernel timing"
timeout -k 10 300 python - <<'PY' 2>&1 | tail -n 8
import os, torch
from sketchyscenecolorization_b200.cuda_ops import CudaOps
ops = CudaOps("cuda:0", torch.bfloat16)
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
g = torch.randn(128 * 192 * 192, 3, device="cuda").bfloat16()
out = torch.zeros(3, device="cuda")
for _ in range(3): ops.colsum_(g, out)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10): ops.colsum_(g, out)
e1.record(); torch.cuda.synchronize()
print("  colsum [128*192*192, 3] bf16: %.1f us" % (e0.elapsed_time(e1) * 100))
PY
echo "=== elementwise family"; REPS=5 timeout -k 10 300 python scripts/prof_elem.py 2>&1 | tee gpurun_out/prof_elem_$T.log | tail -n 14
