"""Production-size launches of the streaming kernels added in round 2 (for an ncu pass with DRAM metrics, and CUDA-event GB/s):
gate + PReLU of the discriminator's cell, the accumulating PReLU backward, the matching model's end-of-unit pass, max pool,
space <-> batch, the 4-units-per-thread LSTM cell, pad + cast of the recurrent operand."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import torch
from sketchyscenecolorization_b200.cuda_ops import CudaOps

ops = CudaOps("cuda:0", torch.bfloat16)
bf, f32, dev = torch.bfloat16, torch.float32, "cuda:0"
R = lambda *s, dt=bf: torch.randn(*s, device=dev, dtype=f32).to(dt).contiguous()  # noqa: E731
a = torch.tensor(0.2, device=dev)
cases = []
x, rg, im, g = (R(128, 96, 96, 128) for _ in range(4))               # discriminator unit 2, 2N pictures: 302 MB each
gh = R(128, 96, 96, 128)
da = torch.zeros((), device=dev)
cases.append(("gate_prelu_fwd", 4 * x.numel() * 2, lambda: ops.gate_prelu_fwd(x, rg, im, a)))
cases.append(("gate_prelu_bwd (+ g_ht accumulate)", 8 * x.numel() * 2, lambda: ops.gate_prelu_bwd(g, x, rg, im, a, da, g_ht=gh, acc=True)))
cases.append(("prelu_bwd accumulating", 4 * x.numel() * 2, lambda: ops.prelu_bwd(g, x, a, da, acc_into=gh)))
t = R(128, 48, 48, 1024)                                              # matching model, group 4 in batch form: 604 MB
res = R(128, 48, 48, 1024)
sc, sh = torch.rand(1024, device=dev) + 0.5, torch.randn(1024, device=dev)
cases.append(("affine_act (bn + residual + relu)", 3 * t.numel() * 2, lambda: ops.affine_act(t, sc, sh, res=res, relu=True)))
cases.append(("affine_act (bn + shortcut bn + relu)", 3 * t.numel() * 2, lambda: ops.affine_act(t, sc, sh, res=res, rscale=sc, rshift=sh, relu=True)))
m = R(32, 384, 384, 64)
cases.append(("maxpool3x3s2", int(1.25 * m.numel() * 2), lambda: ops.maxpool3x3s2(m)))
s2 = R(32, 96, 96, 512)
cases.append(("space_to_batch r=2", 2 * s2.numel() * 2, lambda: ops.space_to_batch(s2, 2)))
Rr, D = 32 * 96 * 96, 500
ga, gp = R(Rr, 4 * D, dt=f32), R(Rr, 4 * D, dt=f32)
grow = R(32, 4 * D, dt=f32)
c0, h0 = R(Rr, D, dt=f32), R(Rr, D, dt=f32)
live = torch.ones(32, 15, dtype=torch.int32, device=dev)
cases.append(("lstm_cell_fwd, 4 units per thread, no pre kept", (2 * 4 * D + 4 * D) * Rr * 4, lambda: ops.lstm_cell_fwd(ga, gp, grow, c0, h0, live, 0, 96 * 96, save_pre=False)))
cases.append(("pad_cast_rows 500 -> 504 bf16", Rr * (D * 4 + 504 * 2), lambda: ops.pad_cast_rows(h0, 504, bf)))

for name, nbytes, fn in cases:
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 1 if os.environ.get("ONCE") else 10
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    print("%-46s %8.3f ms  %7.0f GB/s algorithmic (%.0f MB)" % (name, ms, nbytes / ms / 1e6, nbytes / 1e6))
