"""Per-operator time of one training iteration (bs 64, 192x192, bf16 mode), measured op by op with CUDA events."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import torch
import bench
from sketchyscenecolorization_b200.cuda_ops import CudaOps, enable_op_timing
from sketchyscenecolorization_b200.trainer import FgColorModel, FgColorTrainer

bs = int(os.environ.get("BS", "64"))
ops = CudaOps("cuda:0", torch.bfloat16)
m = FgColorModel(ops, "cuda:0", size=64, H=192, W=192)
m.initialize(seed=0)
tr = FgColorTrainer(m, max_iter=1000)
b = bench.synth_batch(bs, 1)
db = {k: v.cuda() for k, v in b.items() if k != "text"}
db["text"] = b["text"].numpy()
tr.d_step(db); tr.g_step(db)
torch.cuda.synchronize()
enable_op_timing(ops)
tr.d_step(db); tr.g_step(db)
torch.cuda.synchronize()
tot = sum(v[1] for v in ops.op_times.values())
print("total op time %.1f ms" % tot)
for k, (n, ms) in sorted(ops.op_times.items(), key=lambda kv: -kv[1][1]):
    print("%-22s %5d calls %9.2f ms %5.1f%%" % (k, n, ms, 100 * ms / tot))
print("---- convolutions by shape")
for k, (n, ms) in sorted(ops.op_detail.items(), key=lambda kv: -kv[1][1]):
    print("%7.2f ms %3d x  %s" % (ms, n, k))
