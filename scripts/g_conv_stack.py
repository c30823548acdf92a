"""One generator forward + backward at the benchmarked configuration (bs 64, 192 x 192, bf16), nothing else: the target of
    ncu --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed --clock-control none \
        --profile-from-start off --csv --log-file gpurun_out/g_stack.csv python scripts/g_conv_stack.py
whose per-launch list scripts/tensor_pipe_summary.py reduces to the time-weighted tensor-pipe utilisation of the generator's
convolution stack (BASELINE.json metric: "gen conv tensor-pipe %")."""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import torch

import bench
from sketchyscenecolorization_b200.cuda_ops import CudaOps
from sketchyscenecolorization_b200.trainer import FgColorModel

bs = int(os.environ.get("BS", "64"))
ops = CudaOps("cuda:0", torch.bfloat16)
m = FgColorModel(ops, "cuda:0", size=64, H=192, W=192, with_discriminator=False)
m.initialize(seed=0)
b = bench.synth_batch(bs, 1)
sk, cls, noise = b["sketch"].cuda(), b["cls"].cuda(), b["noise"].cuda()
text = b["text"].cuda()
g = torch.randn(bs, 192, 192, 3, device="cuda").to(torch.bfloat16) * 0.01


def once():
    m.gstore.grad.zero_()
    out, ctx = m.G.forward(sk, text, cls, noise, save=True)
    m.G.backward(g.clone(), ctx)


once()
torch.cuda.synchronize()
torch.cuda.profiler.start()
once()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("done")
