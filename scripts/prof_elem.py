"""Runs the HBM-bound elementwise / reduction kernels of the training step on the largest activation they see
([64,192,192,128] bf16 = 604 MB, > L2) and prints achieved algorithmic GB/s -- the target of the `ncu --set full`
capture of the elementwise family."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import torch
from sketchyscenecolorization_b200.cuda_ops import CudaOps

bs = int(os.environ.get("BS", "64"))
reps = int(os.environ.get("REPS", "3"))
Cc, hw = 128, 192
ops = CudaOps("cuda:0", torch.bfloat16)
dev = "cuda"
x = torch.randn(bs, hw, hw, Cc, device=dev).to(torch.bfloat16)
y = torch.randn(bs, hw, hw, Cc, device=dev).to(torch.bfloat16)
z = torch.rand(bs, hw, hw, Cc, device=dev).to(torch.bfloat16)
low = torch.randn(bs, hw // 2, hw // 2, Cc, device=dev).to(torch.bfloat16)
labels = torch.randint(0, 25, (bs,), device=dev).int()
scale = torch.rand(25, Cc, device=dev) + 0.5
offset = torch.randn(25, Cc, device=dev) * 0.1
dscale, doffset = torch.zeros_like(scale), torch.zeros_like(offset)
a = torch.full((1,), 0.2, device=dev)
da = torch.zeros(1, device=dev)
E = x.numel()
mean, rstd = ops.chan_stats(x)
_, mn, mx = ops.minmax_fwd(x)
cases = [  # name, fn, algorithmic bytes (bf16 storage: 2 B per element read or written)
    ("chan_stats", lambda: ops.chan_stats(x), 2 * E),
    ("cbn_act_fwd", lambda: ops.cbn_act_fwd(x, mean, rstd, scale, offset, labels), 4 * E),
    ("cbn_act_bwd", lambda: ops.cbn_act_bwd(y, x, mean, rstd, scale, offset, labels, dscale, doffset), 10 * E),
    ("minmax_fwd", lambda: ops.minmax_fwd(x), 6 * E),
    ("minmax_bwd", lambda: ops.minmax_bwd(y, x, mn, mx), 10 * E),
    ("gate_fma_fwd", lambda: ops.gate_fma_fwd(x, z, y), 8 * E),
    ("gate_fma_bwd", lambda: ops.gate_fma_bwd(x, z, y), 10 * E),
    ("blend_fwd", lambda: ops.blend_fwd(low, y, z), 6 * E + E // 2),          # reads h2, zg, sk/4; writes out
    ("blend_bwd", lambda: ops.blend_bwd(x, low, y, z), 11 * E),               # reads g, h2, zg, sk/4; writes g_h2, g_zg, g_sk/4
    ("mul_up_fwd", lambda: ops.mul_up_fwd(z, low), 4 * E + E // 2),
    ("addpool_fwd", lambda: ops.addpool_fwd(x, y), 4 * E + E // 2),
    ("prelu_fwd", lambda: ops.prelu_fwd(x, a), 4 * E),
    ("prelu_bwd", lambda: ops.prelu_bwd(y, x, a, da), 6 * E),
]
only = os.environ.get("ONLY")
for name, fn, nbytes in cases:
    if only and name not in only.split(","):
        continue
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    print("%-14s %8.3f ms  %7.1f GB/s (algorithmic bytes %.0f MB)" % (name, ms, nbytes / ms / 1e6, nbytes / 1e6), flush=True)
