"""Times fgc_paired_input (csrc/input.cu: raw uint8 records -> normalised NCHW fp32 batch) at the BASELINE batch, 64 samples of
384 x 384 x 3 -> 192 x 192, and prints achieved algorithmic GB/s.  Algorithmic bytes per sample: the sketch (442 368 B), the
even rows of the cartoon -- the sectors the BILINEAR pick touches (221 184 B) -- and the two fp32 outputs (2 x 442 368 B):
1 548 288 B.  The min-max pass re-reads the picked cartoon pixels (14 MB per batch, L2 resident).  Inputs of 8 batches are
rotated so that every call starts from HBM (8 x 113 MB of inputs+outputs > the 126 MB L2)."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import torch
from sketchyscenecolorization_b200.cuda_ops import CudaOps

bs = int(os.environ.get("BS", "64"))
reps = int(os.environ.get("REPS", "5"))
nbuf = int(os.environ.get("NBUF", "8"))
ops = CudaOps("cuda:0", torch.float32)
bufs = [(torch.randint(0, 256, (bs, 384, 384, 3), dtype=torch.uint8, device="cuda"),
         torch.randint(0, 256, (bs, 384, 384, 3), dtype=torch.uint8, device="cuda")) for _ in range(nbuf)]
per_sample = 442368 + 221184 + 2 * 442368
only_fast = bool(os.environ.get("ONLY_FAST"))
for name, hw in (("384->192 (fast path)", (192, 192)), ("384->64 (generic path)", (64, 64))):
    if only_fast and hw != (192, 192):
        continue
    if hw == (64, 64):
        per = 2 * 442368 + 2 * 3 * 64 * 64 * 4       # factor 6: every sector of both inputs is touched
    else:
        per = per_sample
    for c, s in bufs:
        ops.paired_input(c, s, hw, seed=1)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n0 = ops.launch_count()
    e0.record()
    for r in range(reps):
        for c, s in bufs:
            ops.paired_input(c, s, hw, seed=r)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / (reps * nbuf)
    print("paired_input %-24s bs %d  %8.4f ms/batch  %7.1f GB/s algorithmic (%.1f MB/batch)  %8.0f samples/s  %d launches/batch"
          % (name, bs, ms, per * bs / ms / 1e6, per * bs / 1e6, bs / ms * 1e3, (ops.launch_count() - n0) // (reps * nbuf)), flush=True)
if only_fast:
    sys.exit(0)
# end to end from pinned host memory (what a training step pays for one queue): H2D of the raw batch + the device pass
hc, hs = bufs[0][0].cpu().pin_memory(), bufs[0][1].cpu().pin_memory()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for r in range(reps):
    ops.paired_input(hc.to("cuda", non_blocking=True), hs.to("cuda", non_blocking=True), (192, 192), seed=r)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / reps
print("paired_input from pinned host memory   bs %d  %8.4f ms/batch  (%.1f MB H2D, %.1f GB/s over the bus)  %8.0f samples/s"
      % (bs, ms, 2 * hc.numel() / 1e6, 2 * hc.numel() / ms / 1e6, bs / ms * 1e3), flush=True)
