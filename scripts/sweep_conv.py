"""Timing sweep of the tcgen05 conv kernels (CUDA events): separates the per-CTA fixed cost from the per-slab cost."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import torch
from sketchyscenecolorization_b200.cuda_ops import CudaOps

ops = CudaOps("cuda:0", torch.bfloat16)
cases = [  # bs, hw, cin, cout, k
    (32, 192, 128, 128, 1), (32, 192, 128, 128, 3), (32, 192, 256, 128, 3), (32, 192, 512, 128, 3),
    (32, 192, 128, 64, 3), (32, 192, 128, 256, 3), (64, 96, 128, 128, 3), (64, 96, 256, 256, 3), (64, 48, 512, 512, 3),
    (64, 192, 8, 128, 3), (64, 192, 8, 8, 3), (64, 192, 128, 8, 1), (64, 192, 64, 64, 3),
]
which = os.environ.get("WHICH", "fwd,dgrad,wgrad").split(",")
if os.environ.get("CASES"):
    cases = [cases[int(i)] for i in os.environ["CASES"].split(",")]
for bs, hw, cin, cout, k in cases:
    x = torch.randn(bs, hw, hw, cin, device="cuda").to(torch.bfloat16)
    gy = torch.randn(bs, hw, hw, cout, device="cuda").to(torch.bfloat16)
    w = (torch.randn(k, k, cin, cout, device="cuda") * 0.02).contiguous()
    b = torch.zeros(cout, device="cuda")
    dw = torch.zeros_like(w)
    flop = 2.0 * bs * hw * hw * k * k * cin * cout
    fns = dict(fwd=lambda: ops.conv_fwd([(x, False)], w, b), dgrad=lambda: ops.conv_dgrad(gy, w, 0, cin),
               wgrad=lambda: ops.conv_wgrad([(x, False)], gy, dw, None))
    out = []
    for what in which:
        fn = fns[what]
        fn(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3):
            fn()
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 3
        out.append("%s %7.3f ms %6.1f TF" % (what, ms, flop / ms / 1e9))
    print("bs%d %3d^2 %4d->%-4d k%d | %s" % (bs, hw, cin, cout, k, " | ".join(out)), flush=True)
    del x, gy
