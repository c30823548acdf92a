"""End-to-end GPU parity: generator inference and the two training graphs against the CPU oracle.

The bar (BASELINE.json north_star): inference pixels within 1e-3 max-abs (fp32) of the reference graph.
Gradients are compared against torch autograd on the fp64 oracle.
"""
import pytest
import torch

pytestmark = pytest.mark.gpu

INFER_TOL = 1e-3          # max-abs on the tanh output, north_star
GRAD_TOL = 5e-3           # max-abs error relative to the largest gradient entry of the tensor (bf16x3 convs, fp32 rest)


def _model(size, H, W, act_dtype, seed=3):
    from sketchyscenecolorization_b200.cuda_ops import CudaOps
    from sketchyscenecolorization_b200.trainer import FgColorModel
    ops = CudaOps("cuda:0", act_dtype)
    m = FgColorModel(ops, "cuda:0", size=size, H=H, W=W)
    m.initialize(seed=seed, perturb_tables=0.1)
    return m


def _oracle_params(m, dtype):
    gp = {k: v.detach().cpu().to(dtype).requires_grad_(True) for k, v in m.gstore.state_dict().items()}
    dp = {k: v.detach().cpu().to(dtype).requires_grad_(True) for k, v in m.dstore.state_dict().items()}
    return gp, dp


def _dev_batch(b):
    dev = "cuda:0"
    out = {k: (v.float().to(dev).contiguous() if v.is_floating_point() else v) for k, v in b.items()}
    out["cls"] = b["cls"].int().to(dev)
    out["cls_d"] = b["cls_d"].int().to(dev)
    out["text"] = b["text"].numpy()
    return out


@pytest.mark.parametrize("cfg", [(16, 64, 64, 3, 4), (64, 192, 192, 1, 12)], ids=["size16_64px_n3", "size64_192px_n1_cfg1"])
def test_generator_inference_parity(cfg):
    from oracle import fgcolor_oracle as O
    size, H, W, N, n_pad = cfg
    m = _model(size, H, W, torch.float32)
    gp, _ = _oracle_params(m, torch.float64)
    b = O.make_batch(N, H, W, 11, torch.float64, n_pad=n_pad)
    if N > 1:
        b["text"][0, :9] = 0
    if n_pad == 12:
        b["text"][0, 12:] = torch.tensor([24, 3, 6])     # 'the bus is orange' (tests/golden/text_ids.json)
        b["cls"][0] = 2
    with torch.no_grad():
        ref = O.generator_forward(gp, b["sketch"], b["text"], b["cls"], b["noise"], size)
    db = _dev_batch(b)
    out = m.generate(db["sketch"], db["text"], db["cls"], db["noise"])
    torch.cuda.synchronize()
    assert out.shape == ref.shape and out.dtype == torch.float32
    err = (out.cpu().double() - ref).abs().max().item()
    assert torch.isfinite(out).all()
    assert err <= INFER_TOL, "generator max-abs err %.3e > %.0e" % (err, INFER_TOL)


def _cmp_grads(store, ref):
    """(worst max-abs error relative to max(|g|_max, 1e-3 * global max), its tensor, worst relative L2 error, global cosine)."""
    worst, worst_k, worst_l2 = 0.0, None, 0.0
    gs = max(g.abs().max().item() for g in ref.values())
    dot = na = nb = 0.0
    for k, g in ref.items():
        mine = store.g[k].detach().cpu().double()
        d = (mine - g)
        rel = d.abs().max().item() / max(g.abs().max().item(), 1e-3 * gs)
        if rel > worst:
            worst, worst_k = rel, k
        if g.abs().max().item() > 1e-3 * gs:
            worst_l2 = max(worst_l2, d.norm().item() / g.norm().item())
        dot += (mine * g).sum().item()
        na += (mine * mine).sum().item()
        nb += (g * g).sum().item()
    return worst, worst_k, worst_l2, dot / (na ** 0.5 * nb ** 0.5)


@pytest.mark.parametrize("impl", [1, 0], ids=["cuda_core_conv", "tcgen05_conv"])
def test_training_graph_gradients(impl):
    """D-step and G-step gradients against torch autograd on the fp64 oracle.

    With the CUDA-core convolution (fp32 FMA) every tensor must agree entry by entry: this pins the hand-written
    backward passes.  With the tcgen05 convolution (bf16x3, ~1e-5 relative per product) the min-max gate
    normalisation (mru.py:415-416) can pick a different arg-max pixel than the oracle when two pixels are within
    rounding of each other -- a discontinuity of the reference function itself -- so that path is held to a relative
    L2 bound per tensor and a cosine bound overall instead of an entry-wise one."""
    from oracle import fgcolor_oracle as O
    size, H, W, N = 16, 64, 64, 3
    m = _model(size, H, W, torch.float32)
    m.ops.lib.fgc_set_conv_impl(impl)
    try:
        gp, dp = _oracle_params(m, torch.float64)
        gspecs, dspecs = O.generator_specs(size, 58, H, W), O.discriminator_specs(size)
        b = O.make_batch(N, H, W, 5, torch.float64, n_pad=3)
        b["text"][0, :7] = 0
        db = _dev_batch(b)

        def check(store, ref, tag):
            # the oracle loss includes the l2 decay, whose gradient the fused Adam kernel adds: add it for the comparison
            for s in store.specs:
                if s.trainable and s.reg > 0:
                    store.g[s.name] += s.reg * store.p[s.name]
            worst, k, l2, cos = _cmp_grads(store, ref)
            if impl == 1:
                assert worst <= GRAD_TOL, "%s grads: %s rel err %.3e" % (tag, k, worst)
            assert l2 <= 5e-2, "%s grads: worst relative L2 error %.3e" % (tag, l2)
            assert cos >= 0.9995, "%s grads: cosine %.6f" % (tag, cos)

        # ---- D step
        r = m.d_step_grads(db)
        ld, _, _ = O.d_step_loss(gp, dp, gspecs, dspecs, b, size)
        gd = O.grads_of(ld, dp, dspecs)
        torch.cuda.synchronize()
        assert abs(r["loss"].item() - ld.item()) <= 1e-3 * abs(ld.item())
        check(m.dstore, gd, "D")
        # ---- G step
        u_before = {k: v.clone() for k, v in m.dstore.state.items()}
        r = m.g_step_grads(db)
        lg, _, u_new, _ = O.g_step_loss(gp, dp, gspecs, dspecs, b, size)
        gg = O.grads_of(lg, gp, gspecs)
        torch.cuda.synchronize()
        assert abs(r["loss"].item() - lg.item()) <= 1e-3 * abs(lg.item())
        check(m.gstore, gg, "G")
        # spectral-norm u <- u' happens with the G step (graph_single.py:178-180)
        for k, v in m.dstore.state.items():
            assert (v.cpu().double() - u_new[k]).abs().max().item() <= 1e-4
            assert not torch.equal(v, u_before[k])
    finally:
        m.ops.lib.fgc_set_conv_impl(0)


def test_bf16_training_step_runs():
    """Training mode: bf16 NHWC activations, single-pass bf16 tensor-core convs, fp32 master weights/stats."""
    from oracle import fgcolor_oracle as O
    from sketchyscenecolorization_b200.trainer import FgColorTrainer
    size, H, W, N = 16, 64, 64, 4
    m = _model(size, H, W, torch.bfloat16)
    mref = _model(size, H, W, torch.float32)
    b = _dev_batch(O.make_batch(N, H, W, 7, torch.float64))
    r16 = m.d_step_grads(b)
    r32 = mref.d_step_grads(b)
    torch.cuda.synchronize()
    assert abs(r16["loss"].item() - r32["loss"].item()) <= 0.05 * abs(r32["loss"].item())
    g16, g32 = m.dstore.grad.double(), mref.dstore.grad.double()
    cos = torch.dot(g16, g32) / (g16.norm() * g32.norm())
    assert cos.item() > 0.98, "bf16 vs fp32 D-gradient cosine %.4f" % cos.item()
    r16 = m.g_step_grads(b)
    r32 = mref.g_step_grads(b)
    g16, g32 = m.gstore.grad.double(), mref.gstore.grad.double()
    cos = torch.dot(g16, g32) / (g16.norm() * g32.norm())
    assert cos.item() > 0.97, "bf16 vs fp32 G-gradient cosine %.4f" % cos.item()
    tr = FgColorTrainer(m, max_iter=100)
    for _ in range(2):
        od = tr.d_step(b)
        og = tr.g_step(b)
    torch.cuda.synchronize()
    assert torch.isfinite(od["loss"]) and torch.isfinite(og["loss"])
    assert torch.isfinite(m.gstore.flat).all() and torch.isfinite(m.dstore.flat).all()


def test_cuda_graph_replay_matches_eager():
    """FgColorTrainer(use_cuda_graphs=True): eager first call, captured second call, replays afterwards -- same
    parameters as the eager trainer fed the same batches (up to the ordering of fp32 atomics)."""
    from oracle import fgcolor_oracle as O
    from sketchyscenecolorization_b200.trainer import FgColorTrainer
    size, H, W, N = 16, 64, 64, 4
    me, mg = _model(size, H, W, torch.float32), _model(size, H, W, torch.float32)
    te = FgColorTrainer(me, max_iter=100)
    tg = FgColorTrainer(mg, max_iter=100, use_cuda_graphs=True)
    for it in range(4):
        b = O.make_batch(N, H, W, 100 + it, torch.float64, n_pad=it)       # different captions / pads every step
        be = _dev_batch(b)
        bg = dict(be)
        bg["text"] = b["text"].int().cuda()                               # graph mode: ids stay on the device
        oe_d, og_d = te.d_step(be), tg.d_step(bg)
        oe_g, og_g = te.g_step(be), tg.g_step(bg)
        torch.cuda.synchronize()
        assert abs(oe_d["loss"].item() - og_d["loss"].item()) <= 2e-3 * abs(oe_d["loss"].item()), it
        assert abs(oe_g["loss"].item() - og_g["loss"].item()) <= 2e-3 * abs(oe_g["loss"].item()), it
    assert "graph" in tg._g["d"] and "graph" in tg._g["g"] and tg.launches_per_step["d"] > 100
    assert (me.gstore.adam_t, me.dstore.adam_t, te.counter) == (mg.gstore.adam_t, mg.dstore.adam_t, tg.counter) == (4, 4, 4)
    for a, b_ in ((me.gstore.flat, mg.gstore.flat), (me.dstore.flat, mg.dstore.flat)):
        # Adam (beta1 = 0) moves every weight by ~lr/sqrt(1-beta2) per step whatever the gradient scale, so parameters
        # whose true gradient is zero (biases in front of a batch-norm) random-walk on fp32 atomic-ordering noise in
        # BOTH trainers: bound the difference by the total step budget and compare the bulk in relative L2.
        assert (a - b_).abs().max().item() <= 2 * 4 * 2e-4 * 3.2, (a - b_).abs().max().item()
        assert ((a - b_).norm() / a.norm()).item() <= 5e-3, ((a - b_).norm() / a.norm()).item()


# ------------------------------------------------------------------------------------------------------------------
# the benchmarked network -- size 64, 192 x 192 -- against the fp64 oracle, in both numeric modes
# ------------------------------------------------------------------------------------------------------------------
# Bounds of the single-pass bf16 TRAINING mode against the fp64 oracle at the full-size network (N = 2, random initialisation,
# cBN tables perturbed).  bf16 operands carry 2^-9 relative rounding per product; the 46-convolution generator with batch
# statistics and eps-free min-max gates amplifies it (2e-2 max-abs on the generator output), so the
# gradients are held to a direction bound (cosine) and a per-tensor relative-L2 bound instead of an entry-wise one.
BF16_LOSS_REL = 0.03
BF16_COS_D, BF16_COS_G = 0.98, 0.95
BF16_L2_D, BF16_L2_G = 0.60, 0.80      # the WORST single tensor (measured over five runs: 0.30 .. 0.41 and 0.39 .. 0.46)


@pytest.mark.parametrize("mode", ["fp32", "bf16"])
def test_full_size_training_graph_gradients(mode):
    """D-step and G-step of the size-64, 192 x 192 network (every production kernel instantiation: 16 x 8 / 32 x 8 / 64 x 8
    pixel tiles, 64..768-wide N tiles, patch-tensor sources) against torch autograd on the fp64 oracle."""
    from oracle import fgcolor_oracle as O
    size, H, W, N = 64, 192, 192, 2
    m = _model(size, H, W, torch.float32 if mode == "fp32" else torch.bfloat16)
    gp, dp = _oracle_params(m, torch.float64)
    gspecs, dspecs = O.generator_specs(size, 58, H, W), O.discriminator_specs(size)
    b = O.make_batch(N, H, W, 21, torch.float64, n_pad=2)
    b["text"][0, :9] = 0
    db = _dev_batch(b)

    def check(store, ref, tag, cos_min, l2_max):
        for s in store.specs:
            if s.trainable and s.reg > 0:
                store.g[s.name] += s.reg * store.p[s.name]
        worst, k, l2, cos = _cmp_grads(store, ref)
        print("%s %s grads vs fp64 oracle: worst entry err (rel. to tensor max) %.3e at %s, worst tensor rel-L2 %.3e, cosine %.6f"
              % (mode, tag, worst, k, l2, cos))
        assert l2 <= l2_max, "%s grads: worst relative L2 error %.3e" % (tag, l2)
        assert cos >= cos_min, "%s grads: cosine %.6f" % (tag, cos)

    loss_rel = 1e-3 if mode == "fp32" else BF16_LOSS_REL
    r = m.d_step_grads(db)
    ld, _, _ = O.d_step_loss(gp, dp, gspecs, dspecs, b, size)
    gd = O.grads_of(ld, dp, dspecs)
    torch.cuda.synchronize()
    print("%s loss_d %.6f (oracle %.6f)" % (mode, r["loss"].item(), ld.item()))
    assert abs(r["loss"].item() - ld.item()) <= loss_rel * abs(ld.item())
    check(m.dstore, gd, "D", 0.9995 if mode == "fp32" else BF16_COS_D, 5e-2 if mode == "fp32" else BF16_L2_D)
    r = m.g_step_grads(db)
    lg, _, _, _ = O.g_step_loss(gp, dp, gspecs, dspecs, b, size)
    gg = O.grads_of(lg, gp, gspecs)
    torch.cuda.synchronize()
    print("%s loss_g %.6f (oracle %.6f)" % (mode, r["loss"].item(), lg.item()))
    assert abs(r["loss"].item() - lg.item()) <= loss_rel * abs(lg.item())
    check(m.gstore, gg, "G", 0.9995 if mode == "fp32" else BF16_COS_G, 5e-2 if mode == "fp32" else BF16_L2_G)


def test_bf16_inference_error_is_stated():
    """The training-mode forward (single-pass bf16) against the fp64 oracle at the full-size network: NOT the parity mode
    (inference / val / test run bf16x3 and meet 1e-3, above) -- this pins how far the benchmarked arithmetic is from the
    reference's, so the number in DESIGN.md section 7 is a tested one."""
    from oracle import fgcolor_oracle as O
    size, H, W, N = 64, 192, 192, 1
    m = _model(size, H, W, torch.bfloat16)
    gp, _ = _oracle_params(m, torch.float64)
    b = O.make_batch(N, H, W, 11, torch.float64, n_pad=12)
    with torch.no_grad():
        ref = O.generator_forward(gp, b["sketch"], b["text"], b["cls"], b["noise"], size)
    db = _dev_batch(b)
    out = m.generate(db["sketch"], db["text"], db["cls"], db["noise"])
    torch.cuda.synchronize()
    d = (out.cpu().double() - ref).abs()
    print("single-pass bf16 generator vs fp64 oracle: max-abs %.3e, mean-abs %.3e" % (d.max().item(), d.mean().item()))
    assert d.max().item() <= 0.15 and d.mean().item() <= 1.5e-2


def test_bf16_trains_like_fp32():
    """200 iterations of the session loop at size 16 / 64 x 64 on eight cycling batches whose target pictures are a function of
    the sketch and the class (0.6 * sketch + 0.4 * class colour: learnable, unlike uniform noise), once in the split-precision
    mode (bf16x3 convolutions, fp32 activations) and once in the benchmarked mode (single-pass bf16): same data, same noise,
    same initial weights, learning rates 5x the defaults so that 200 iterations show the whole descent.  GAN training is
    chaotic, so the curves are compared as curves: the reconstruction loss must fall by 10x in both, and the bf16 run's window
    means must stay within a band of the fp32 run's."""
    from sketchyscenecolorization_b200.input_pipeline import SyntheticInput
    from sketchyscenecolorization_b200.main_procedure import TrainSession

    class Cycle:
        def __init__(self, seed, n=8):
            src = SyntheticInput(8, 64, 64, seed=seed)
            color = torch.rand(25, 3, generator=torch.Generator().manual_seed(99)) * 2 - 1
            self.b, self.i = [], 0
            for _ in range(n):
                b = next(src)
                for key, ck in (("images", "cls"), ("images_d", "cls_d")):
                    b[key] = (0.6 * b["sketch"] + 0.4 * color[b[ck].long()][:, :, None, None]).contiguous()
                self.b.append(b)

        def __iter__(self):
            return self

        def __next__(self):
            self.i += 1
            return self.b[(self.i - 1) % len(self.b)]

    curves = {}
    for mode, dt in (("fp32", torch.float32), ("bf16", torch.bfloat16)):
        torch.manual_seed(3)
        torch.cuda.manual_seed(3)
        m = _model(16, 64, 64, dt, seed=4)
        s = TrainSession(m, batch_size=8, max_iter=200, lr_g=1e-3, lr_d=5e-4, small=True, input_iter=Cycle(1), input_iter_d=Cycle(2))
        rows = []
        for _ in range(200):
            ld, lg, nd, ng = s.iteration()
            assert not nd and not ng
            rows.append((ld, lg, float(s.last_g["l1"])))
        curves[mode] = torch.tensor(rows)
    first = {k: v[:5].mean(0) for k, v in curves.items()}
    mid = {k: v[40:80].mean(0) for k, v in curves.items()}
    last = {k: v[-40:].mean(0) for k, v in curves.items()}
    print("means over iterations 0-4 / 40-79 / 160-199 of (loss_d, loss_g, l1): fp32 %s / %s / %s; bf16 %s / %s / %s"
          % tuple([[round(x, 3) for x in d[k].tolist()] for k in ("fp32", "bf16") for d in (first, mid, last)]))
    for mode in ("fp32", "bf16"):
        assert last[mode][2] < 0.1 * first[mode][2], "%s: the reconstruction loss did not fall" % mode
    assert abs(first["bf16"][2] - first["fp32"][2]) <= 0.02 * first["fp32"][2]          # same start
    for w in (mid, last):                                                               # same descent, within a band
        assert abs(w["bf16"][2] - w["fp32"][2]) <= 0.5 * w["fp32"][2]
        assert abs(w["bf16"][1] - w["fp32"][1]) <= 0.5 * w["fp32"][1]
        assert abs(w["bf16"][0] - w["fp32"][0]) <= 0.5 * max(w["fp32"][0], 0.5)
