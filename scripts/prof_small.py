"""Times the direct narrow convolutions (conv_small_*: the discriminator's stem-level layers at 192 x 192 x 128 pictures)."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import torch
from sketchyscenecolorization_b200.cuda_ops import CudaOps

ops = CudaOps("cuda:0", torch.bfloat16)
dev = "cuda"
N = int(os.environ.get("BS", "128"))


def timeit(fn, reps=5):
    for _ in range(2):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3


x3 = torch.randn(N, 192, 192, 3, device=dev).bfloat16()
a8 = torch.randn(N, 192, 192, 8, device=dev).bfloat16()
g8 = torch.randn(N, 192, 192, 8, device=dev).bfloat16()
for name, srcs, k in (("stem 7x7 3->8", [(x3, False)], 7), ("update_gate 3x3 [8,3]->8", [(a8, False), (x3, False)], 3),
                      ("Conv 3x3 3->8", [(x3, False)], 3)):
    cin = sum(t.shape[-1] for t, _ in srcs)
    w = (torch.randn(k, k, cin, 8, device=dev) * 0.05).contiguous()
    b = torch.zeros(8, device=dev)
    dw, db = torch.zeros_like(w), torch.zeros_like(b)
    gmac = N * 192 * 192 * k * k * cin * 8 / 1e9
    t_f = timeit(lambda: ops.conv_fwd(srcs, w, b))
    t_w = timeit(lambda: ops.conv_wgrad(srcs, g8, dw, db))
    print("%-28s fwd %7.1f us (%5.1f TMAC/s)   wgrad %7.1f us (%5.1f TMAC/s)" % (name, t_f, gmac / t_f * 1e3, t_w, gmac / t_w * 1e3), flush=True)
w = (torch.randn(3, 3, 8, 8, device=dev) * 0.05).contiguous()
t_d = timeit(lambda: ops.conv_dgrad(g8, w, 0, 8))
print("dgrad 3x3 8->8               %7.1f us" % t_d)
