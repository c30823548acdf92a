"""Per-operator time of one instance-matching forward pass (BASELINE configs[4]: 768x768, bs 32, bf16), op by op with CUDA events."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import numpy as np
import torch
from sketchyscenecolorization_b200.cuda_ops import CudaOps, enable_op_timing
from sketchyscenecolorization_b200.rmi import RMIModel

bs = int(os.environ.get("BS", "32"))
dt = torch.float32 if os.environ.get("FP32") else torch.bfloat16
ops = CudaOps("cuda:0", dt)
m = RMIModel(ops, "cuda:0")
m.initialize(seed=0)
rs = np.random.RandomState(0)
im = ops.cast((torch.rand(bs, 768, 768, 3, device="cuda") * 255.0 - 115.0).contiguous(), dt) if dt != torch.float32 else (torch.rand(bs, 768, 768, 3, device="cuda") * 255.0 - 115.0)
words, lengths = rs.randint(2, 59, size=(bs, 15)), rs.randint(4, 16, size=(bs,))
m.forward(im, words, lengths)
torch.cuda.synchronize()
enable_op_timing(ops)
m.forward(im, words, lengths)
torch.cuda.synchronize()
tot = sum(v[1] for v in ops.op_times.values())
print("total op time %.1f ms (bs %d, %s)" % (tot, bs, dt))
for k, (n, ms) in sorted(ops.op_times.items(), key=lambda kv: -kv[1][1]):
    print("%-26s %5d calls %9.2f ms %5.1f%%" % (k, n, ms, 100 * ms / tot))
print("---- convolutions by shape")
for k, (n, ms) in sorted(ops.op_detail.items(), key=lambda kv: -kv[1][1])[:40]:
    print("%7.2f ms %3d x  %s" % (ms, n, k))
