"""Times the convolution shapes that dominate the training step (bs 64, bf16) forward / dgrad / wgrad with CUDA events,
A/B between the kernels (halo-reuse vs per-tap gather; direct narrow vs tensor path).  With ONLY_FIRST=1 it runs just the
dominant 3x3 128->128 @192x192 layer: the target of the `ncu --set full` capture."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import torch
from sketchyscenecolorization_b200.cuda_ops import CudaOps

bs = int(os.environ.get("BS", "64"))
reps = int(os.environ.get("REPS", "3"))
# name, H=W, source channels, Cout, k
cases = [("128->128@192", 192, [128], 128, 3), ("128+3->64@192", 192, [128, 3], 64, 3), ("64->64@192", 192, [64], 64, 3),
         ("64->3 k7@192", 192, [64], 3, 7), ("3->8 k7@192", 192, [3], 8, 7), ("8+3->8@192", 192, [8, 3], 8, 3),
         ("128+3+8->128@96", 96, [128, 3, 8], 128, 3), ("256->256@96", 96, [256], 256, 3), ("512->512@48", 48, [512], 512, 3),
         ("768->768@24", 24, [768], 768, 3)]
if os.environ.get("ONLY_FIRST"):
    cases = cases[:1]
ops = CudaOps("cuda:0", torch.bfloat16)
lib = ops.lib


def timeit(fn):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


for name, hw, cins, cout, k in cases:
    xs0 = [(torch.randn(bs, hw, hw, c, device="cuda").to(torch.bfloat16), False) for c in cins]
    # product routing: narrow sources next to wide ones travel with their patch tensors (as the MRU blocks pass them)
    xs = [(t, u, ops.small_patch(t, k) if (t.shape[-1] < 64 and len(cins) > 1) else None) for t, u in xs0]
    cin = sum(cins)
    gy = torch.randn(bs, hw, hw, cout, device="cuda").to(torch.bfloat16)
    w = (torch.randn(k, k, cin, cout, device="cuda") * 0.02).contiguous()
    b = torch.zeros(cout, device="cuda")
    dw = torch.zeros_like(w)
    db = torch.zeros_like(b)
    flop = 2.0 * bs * hw * hw * k * k * cin * cout
    # a narrow gradient under a large filter (the 7x7 head) travels with its patch tensors, as Generator.backward passes them
    head = cout < 8 and k >= 5 and len(cins) == 1 and cins[0] >= 32
    gp = ops.small_patch(gy, k) if head else None
    gpm = ops.small_patch(gy, k, mirror=True) if head else None
    fns = (("fwd", lambda: ops.conv_fwd(xs, w, b), lambda: ops.conv_fwd(xs0, w, b)),
           ("dgrad", lambda: ops.conv_dgrad(gy, w, 0, cins[0], gy_patch=gpm), lambda: ops.conv_dgrad(gy, w, 0, cins[0])),
           ("wgrad", lambda: ops.conv_wgrad(xs, gy, dw, db, gy_patch=gp), lambda: ops.conv_wgrad(xs0, gy, dw, db)))
    for what, fn, fn0 in fns:
        lib.fgc_set_conv_flags(1, 1)
        ms = timeit(fn)
        line = "%-16s %-6s %8.3f ms  %7.1f TFLOP/s" % (name, what, ms, flop * (cins[0] / cin if what == "dgrad" else 1.0) / ms / 1e9)
        if not os.environ.get("ONLY_FIRST"):
            lib.fgc_set_conv_flags(0, 0)
            ms0 = timeit(fn0)
            lib.fgc_set_conv_flags(1, 1)
            line += "   | gather/tensor-only kernels: %8.3f ms" % ms0
        print(line, flush=True)
