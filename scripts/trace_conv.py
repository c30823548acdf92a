"""Per-role event trace of CTA 0 of the persistent implicit-GEMM kernel (fgc_debug_set_trace)."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import torch
from sketchyscenecolorization_b200.cuda_ops import CudaOps
ops = CudaOps("cuda:0", torch.bfloat16)
bs, hw, cin, cout, k = [int(v) for v in os.environ.get("CASE", "32,192,128,128,1").split(",")]
x = torch.randn(bs, hw, hw, cin, device="cuda").to(torch.bfloat16)
w = (torch.randn(k, k, cin, cout, device="cuda") * 0.02).contiguous()
b = torch.zeros(cout, device="cuda")
ops.conv_fwd([(x, False)], w, b); torch.cuda.synchronize()
cap = 4000
buf = torch.zeros(4 + 4 * cap, dtype=torch.int64, device="cuda")
ops.lib.fgc_debug_set_trace(buf.data_ptr(), cap)
ops.conv_fwd([(x, False)], w, b); torch.cuda.synchronize()
ops.lib.fgc_debug_set_trace(None, 0)
n = min(int(buf[0].item()), cap)
ev = buf[4:4 + 4 * n].view(n, 4).cpu().tolist()
t0 = min(e[3] for e in ev)
names = {(0, 0): "P tile", (0, 1): "P got-empty", (1, 0): "M got-acc-empty", (1, 1): "M got-full", (2, 0): "E got-acc-full", (2, 1): "E ld-done"}
ev.sort(key=lambda e: e[3])
for role, e, t, c in ev[:int(os.environ.get("NEV", "120"))]:
    print("%9d  tile %5d  %s" % (c - t0, t, names[(role, e)]))
