"""Diagnostic: per-tensor gradient error of the CUDA path vs the fp64 oracle, for both conv implementations."""
import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "tests"))
import torch
from oracle import fgcolor_oracle as O
from test_model_gpu import _model, _oracle_params, _dev_batch

size, H, W, N = 16, 64, 64, 3
gspecs, dspecs = O.generator_specs(size, 58, H, W), O.discriminator_specs(size)
b = O.make_batch(N, H, W, 5, torch.float64, n_pad=3)
b["text"][0, :7] = 0
for impl in (1, 0):
    m = _model(size, H, W, torch.float32)
    m.ops.lib.fgc_set_conv_impl(impl)
    gp, dp = _oracle_params(m, torch.float64)
    db = _dev_batch(b)
    r = m.g_step_grads(db)
    lg, _, u_new, _ = O.g_step_loss(gp, dp, gspecs, dspecs, b, size)
    gg = O.grads_of(lg, gp, gspecs)
    for s in m.gstore.specs:
        if s.trainable and s.reg > 0:
            m.gstore.g[s.name] += s.reg * m.gstore.p[s.name]
    gs = max(g.abs().max().item() for g in gg.values())
    rows = []
    for k, g in gg.items():
        a = (m.gstore.g[k].detach().cpu().double() - g).abs().max().item()
        rows.append((a / max(g.abs().max().item(), 1e-4 * gs), a, g.abs().max().item(), k))
    rows.sort(reverse=True)
    print("impl", impl, "loss", r["loss"].item(), lg.item(), "global max grad", gs)
    for rel, a, mx, k in rows[:12]:
        print("  rel %.3e abs %.3e max %.3e  %s" % (rel, a, mx, k))
m.ops.lib.fgc_set_conv_impl(0)
