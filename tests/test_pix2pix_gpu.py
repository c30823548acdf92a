"""GPU parity of `--block_type Pix2Pix`: the phase-form layout kernels (fgc_space_to_depth / fgc_depth_to_space / fgc_copy_rect /
fgc_phase_weights / fgc_phase_wgrad) against the plain-torch operators, the 4x4 layers built from them against the oracle's direct
convolutions, and the generator / training graphs against oracle/pix2pix_oracle.py.

Run on a B200 (profiles/r1u_pix2pix_ops_gpu_tests.log, 22 passed): the layout kernels, the filter scatter / gather and the three
phase-form layers (forward, input gradient, filter gradient); since round 2 also the three model-level tests at the bottom
(profiles/r2a_pix2pix_model_gpu_tests.log).  The host code they exercise is checked against autograd on the CPU
(tests/test_pix2pix_cpu.py)."""
import os

import pytest
import torch

pytestmark = pytest.mark.gpu

INFER_TOL = 1e-3
GRAD_TOL = 5e-3


@pytest.fixture(scope="module")
def env():
    from sketchyscenecolorization_b200.cuda_ops import CudaOps
    from torch_ops import TorchOps
    dev = torch.device("cuda:0")
    return dict(cu=CudaOps(dev, torch.float32), cub=CudaOps(dev, torch.bfloat16), ref=TorchOps(torch.float64, dev), dev=dev)


def _rnd(shape, seed, dev, dtype=torch.float32):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(shape, generator=g, dtype=torch.float64).to(dtype).to(dev)


@pytest.mark.parametrize("dt", [torch.float32, torch.bfloat16], ids=["f32", "bf16"])
@pytest.mark.parametrize("shape", [(2, 8, 12, 3), (3, 6, 10, 8), (2, 4, 6, 64), (1, 2, 2, 5), (2, 24, 24, 12)],
                         ids=lambda s: "x".join(map(str, s)))
def test_layout_kernels_are_exact_permutations(env, shape, dt):
    cu = env["cu"] if dt == torch.float32 else env["cub"]
    ref = env["ref"]
    x = _rnd(shape, 1, env["dev"], dt)
    N, H, W, C = shape
    d = cu.space_to_depth(x)
    assert d.shape == (N, H // 2, W // 2, 4 * C) and torch.equal(d, ref.space_to_depth(x))
    assert torch.equal(cu.depth_to_space(d), x)
    y = _rnd((N, H // 2, W // 2, 4 * C), 2, env["dev"], dt)
    assert torch.equal(cu.depth_to_space(y), ref.depth_to_space(y))
    for (hh, ww) in ((H - 1, W - 1), (H + 1, W + 1), (H, W), (H - 1, W + 2)):
        assert torch.equal(cu.copy_rect(x, hh, ww), ref.copy_rect(x, hh, ww))


@pytest.mark.parametrize("mode,ab", [("conv", (3, 16)), ("conv", (64, 128)), ("deconv", (16, 24)), ("deconv", (3, 128)),
                                     ("k5", (32, 64)), ("k5", (64, 1))], ids=str)
def test_phase_weights_and_adjoint(env, mode, ab):
    cu, ref, dev = env["cu"], env["ref"], env["dev"]
    f = _rnd((4, 4) + ab, 3, dev)
    w = cu.phase_weights(f, mode)
    assert torch.equal(w, ref.phase_weights(f, mode)) and int((w != 0).sum()) == f.numel()
    dw = _rnd(tuple(w.shape), 4, dev)
    df, df_ref = torch.full_like(f, 0.5), torch.full_like(f, 0.5)
    cu.phase_wgrad(dw, df, mode)
    ref.phase_wgrad(dw, df_ref, mode)
    assert torch.equal(df, df_ref)


@pytest.mark.parametrize("case", [("conv", 2, 16, 16, 3, 16), ("conv", 2, 24, 24, 64, 128), ("deconv", 2, 6, 6, 72, 64),
                                  ("deconv", 2, 24, 24, 32, 3), ("k5", 2, 24, 24, 32, 64), ("k5", 2, 23, 23, 64, 1)], ids=str)
def test_phase_form_layers_match_direct_convolutions(env, case):
    """The three 4x4 layers through the product's stride-1 SAME convolutions (bf16x3 tensor-core mode) against F.conv2d /
    F.conv_transpose2d in fp64: forward, input gradient, filter gradient."""
    import torch.nn.functional as F
    from sketchyscenecolorization_b200 import pix2pix as PX
    mode, N, H, W, cin, cout = case
    cu, dev = env["cu"], env["dev"]
    x = _rnd((N, H, W, cin), 5, dev)
    f = _rnd((4, 4, cout, cin) if mode == "deconv" else (4, 4, cin, cout), 6, dev) * 0.1

    class Store:
        p, g = {"f": f}, {"f": torch.zeros_like(f)}
    fl = PX._Filters(Store, cu)
    xd = x.double().permute(0, 3, 1, 2).contiguous().requires_grad_(True)
    fd = f.double().requires_grad_(True)
    if mode == "conv":
        y, ctx = PX.conv_s2_fwd(cu, fl, "f", x)
        want = F.conv2d(F.pad(xd, (1, 1, 1, 1)), fd.permute(3, 2, 0, 1), stride=2)
    elif mode == "deconv":
        y, y3 = PX.deconv_fwd(cu, fl, "f", [x[..., :cin // 2].contiguous(), x[..., cin // 2:].contiguous()])
        want = F.conv_transpose2d(xd, fd.permute(3, 2, 0, 1), stride=2, padding=1)
    else:
        y = PX.conv_s1_fwd(cu, fl, "f", x)
        want = F.conv2d(F.pad(xd, (1, 1, 1, 1)), fd.permute(3, 2, 0, 1))
    want_nhwc = want.permute(0, 2, 3, 1)
    scale = want_nhwc.abs().max().item()
    assert y.shape == want_nhwc.shape
    assert (y.double() - want_nhwc).abs().max().item() <= 2e-4 * scale
    gy = _rnd(tuple(y.shape), 7, dev)
    gx_want, gf_want = torch.autograd.grad(want, [xd, fd], gy.double().permute(0, 3, 1, 2).contiguous())
    if mode == "conv":
        gx = PX.conv_s2_bwd(cu, fl, "f", gy, ctx, need_x=True)
    elif mode == "deconv":
        srcs = [x[..., :cin // 2].contiguous(), x[..., cin // 2:].contiguous()]
        parts = PX.deconv_bwd(cu, fl, "f", cu.space_to_depth(gy), srcs, [True, True])
        gx = torch.cat(parts, dim=3)
    else:
        gx = PX.conv_s1_bwd(cu, fl, "f", gy, x)
    fl.finish_backward()
    assert (gx.double() - gx_want.permute(0, 2, 3, 1)).abs().max().item() <= 5e-4 * gx_want.abs().max().item()
    assert (Store.g["f"].double() - gf_want).abs().max().item() <= 5e-4 * gf_want.abs().max().item()


def _model(size, H, W, act_dtype, seed=3, **kw):
    from sketchyscenecolorization_b200.cuda_ops import CudaOps
    from sketchyscenecolorization_b200.trainer import FgColorModel
    m = FgColorModel(CudaOps("cuda:0", act_dtype), "cuda:0", size=size, H=H, W=W, block_type="Pix2Pix", **kw)
    m.initialize(seed=seed, perturb_tables=0.1)
    return m


def _dev_batch(b):
    out = {k: (v.float().to("cuda:0").contiguous() if v.is_floating_point() else v) for k, v in b.items()}
    out["cls"], out["cls_d"], out["text"] = b["cls"].int().to("cuda:0"), b["cls_d"].int().to("cuda:0"), b["text"].numpy()
    return out


@pytest.mark.parametrize("cfg", [(16, 64, 64, 3), (64, 192, 192, 2)], ids=["size16_64px_n3", "size64_192px_n2"])
def test_generator_inference_parity(cfg):
    from oracle import fgcolor_oracle as O
    from oracle import pix2pix_oracle as P
    size, H, W, N = cfg
    m = _model(size, H, W, torch.float32, with_discriminator=False)
    gp = {k: v.detach().cpu().double() for k, v in m.gstore.state_dict().items()}
    b = O.make_batch(N, H, W, 11, torch.float64, n_pad=4)
    b["text"][0, :9] = 0
    with torch.no_grad():
        ref = P.generator_forward(gp, b["sketch"], b["text"], b["cls"], b["noise"], size)
    db = _dev_batch(b)
    out = m.generate(db["sketch"], db["text"], db["cls"], db["noise"])
    torch.cuda.synchronize()
    err = (out.cpu().double() - ref).abs().max().item()
    assert out.shape == ref.shape and torch.isfinite(out).all()
    assert err <= INFER_TOL, "pix2pix generator max-abs err %.3e" % err


def test_training_graph_gradients():
    from oracle import fgcolor_oracle as O
    from oracle import pix2pix_oracle as P
    size, H, W, N = 16, 64, 64, 3
    m = _model(size, H, W, torch.float32)
    gp = {k: v.detach().cpu().double().requires_grad_(True) for k, v in m.gstore.state_dict().items()}
    dp = {k: v.detach().cpu().double().requires_grad_(True) for k, v in m.dstore.state_dict().items()}
    gspecs, dspecs = P.generator_specs(size, 58, H, W), P.discriminator_specs(size)
    b = O.make_batch(N, H, W, 5, torch.float64, n_pad=3)
    db = _dev_batch(b)

    def check(store, ref, tag):
        for s in store.specs:
            if s.trainable and s.reg > 0:
                store.g[s.name] += s.reg * store.p[s.name]
        gs = max(g.abs().max().item() for g in ref.values())
        for k, g in ref.items():
            rel = (store.g[k].detach().cpu().double() - g).abs().max().item() / max(g.abs().max().item(), 1e-3 * gs)
            assert rel <= GRAD_TOL, "%s grads: %s rel err %.3e" % (tag, k, rel)

    r = m.d_step_grads(db)
    ld, _, _ = P.d_step_loss(gp, dp, gspecs, dspecs, b, size)
    torch.cuda.synchronize()
    assert abs(r["loss"].item() - ld.item()) <= 1e-3 * abs(ld.item())
    check(m.dstore, O.grads_of(ld, dp, dspecs), "D")
    r = m.g_step_grads(db)
    lg, _, _, _ = P.g_step_loss(gp, dp, gspecs, dspecs, b, size)
    torch.cuda.synchronize()
    assert abs(r["loss"].item() - lg.item()) <= 1e-3 * abs(lg.item())
    check(m.gstore, O.grads_of(lg, gp, gspecs), "G")


def test_bf16_training_steps_run():
    """Training mode (bf16 activations, CUDA-graph replay): two iterations stay finite and move the weights."""
    from oracle import fgcolor_oracle as O
    from sketchyscenecolorization_b200.trainer import FgColorTrainer
    m = _model(16, 64, 64, torch.bfloat16)
    tr = FgColorTrainer(m, max_iter=100, use_cuda_graphs=True)
    before = m.gstore.flat.clone()
    for i in range(3):
        db = _dev_batch(O.make_batch(4, 64, 64, 20 + i, torch.float64))
        od, og = tr.d_step(db), tr.g_step(db)
    torch.cuda.synchronize()
    assert torch.isfinite(od["loss"]).item() and torch.isfinite(og["loss"]).item() and not torch.equal(before, m.gstore.flat)
