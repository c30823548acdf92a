"""Times the background-colorization generator at its published size (768 x 768, ngf 64, batch 1, 8-token caption; BASELINE.json
configs[3]) in fp32 parity mode and in bf16, CUDA events around whole forward passes, and prints pictures/s and the kernel count."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import numpy as np
import torch
from sketchyscenecolorization_b200.bg import BgColorModel
from sketchyscenecolorization_b200.cuda_ops import CudaOps

reps = int(os.environ.get("REPS", "5"))
ids = np.array([[0, 2, 3, 4, 5, 8, 3, 7]], dtype=np.int32)            # 'the sky is blue and the ground is green'
for name, dt in (("fp32 storage / bf16x3 convs (parity mode)", torch.float32), ("bf16 storage / single-pass bf16 convs", torch.bfloat16)):
    ops = CudaOps("cuda:0", dt)
    m = BgColorModel(ops, "cuda:0", ngf=64, vocab_size=18)
    m.initialize(seed=0)
    img = torch.rand(1, 768, 768, 3, device="cuda") * 2 - 1
    for _ in range(2):
        m.generate(img, ids)
    torch.cuda.synchronize()
    n0 = ops.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        out, reg = m.generate(img, ids)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    print("bg generator 768x768 bs1  %-44s %8.2f ms/picture  %6.2f pictures/s  %d launches/picture  finite=%s"
          % (name, ms, 1e3 / ms, (ops.launch_count() - n0) // reps, bool(torch.isfinite(out).all())), flush=True)
    del m, ops
    torch.cuda.empty_cache()
