"""Runs the dominant convolution of the training step (3x3, 128->128, 192x192, bs 64, bf16) forward, dgrad and wgrad a few
times -- the target of the `ncu --set full` capture; also prints CUDA-event timings when run without a profiler."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import torch
from sketchyscenecolorization_b200.cuda_ops import CudaOps

bs = int(os.environ.get("BS", "64"))
reps = int(os.environ.get("REPS", "3"))
cases = [("128->128@192", 192, 128, 128), ("256->256@96", 96, 256, 256), ("512->512@48", 48, 512, 512), ("768->768@24", 24, 768, 768)]
if os.environ.get("ONLY_FIRST"):
    cases = cases[:1]
ops = CudaOps("cuda:0", torch.bfloat16)
for name, hw, cin, cout in cases:
    x = torch.randn(bs, hw, hw, cin, device="cuda").to(torch.bfloat16)
    gy = torch.randn(bs, hw, hw, cout, device="cuda").to(torch.bfloat16)
    w = (torch.randn(3, 3, cin, cout, device="cuda") * 0.02).contiguous()
    b = torch.zeros(cout, device="cuda")
    dw = torch.zeros_like(w)
    db = torch.zeros_like(b)
    flop = 2.0 * bs * hw * hw * 9 * cin * cout
    for what, fn in (("fwd", lambda: ops.conv_fwd([(x, False)], w, b)),
                     ("dgrad", lambda: ops.conv_dgrad(gy, w, 0, cin)),
                     ("wgrad", lambda: ops.conv_wgrad([(x, False)], gy, dw, db))):
        fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        print("%-14s %-6s %8.3f ms  %7.1f TFLOP/s" % (name, what, ms, flop / ms / 1e9), flush=True)
