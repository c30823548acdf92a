"""GPU parity of `--block_type Residual` against oracle/residual_oracle.py.

Green on a B200 since round 2 (profiles/r2a_residual_gpu_tests.log).  The operators the model is made of have their own GPU
tests (tests/test_ops_gpu.py, tests/test_pix2pix_gpu.py); the host code is checked against autograd on the CPU
(tests/test_residual_cpu.py)."""
import pytest
import torch

pytestmark = [pytest.mark.gpu]

INFER_TOL = 1e-3
GRAD_TOL = 5e-3


def _model(size, H, W, act_dtype, seed=3, **kw):
    from sketchyscenecolorization_b200.cuda_ops import CudaOps
    from sketchyscenecolorization_b200.trainer import FgColorModel
    m = FgColorModel(CudaOps("cuda:0", act_dtype), "cuda:0", size=size, H=H, W=W, block_type="Residual", **kw)
    m.initialize(seed=seed, perturb_tables=0.1)
    return m


def _dev_batch(b):
    out = {k: (v.float().to("cuda:0").contiguous() if v.is_floating_point() else v) for k, v in b.items()}
    out["cls"], out["cls_d"], out["text"] = b["cls"].int().to("cuda:0"), b["cls_d"].int().to("cuda:0"), b["text"].numpy()
    return out


@pytest.mark.parametrize("dt", [torch.float32, torch.bfloat16], ids=["f32", "bf16"])
@pytest.mark.parametrize("shape", [(2, 8, 8, 3), (3, 5, 7, 64), (1, 1, 1, 5)], ids=lambda s: "x".join(map(str, s)))
def test_tanh_fwd(shape, dt):
    from sketchyscenecolorization_b200.cuda_ops import CudaOps
    cu = CudaOps("cuda:0", dt)
    g = torch.Generator().manual_seed(1)
    x = (torch.randn(shape, generator=g, dtype=torch.float64) * 2).to(dt).cuda()
    y = cu.tanh_fwd(x)
    want = torch.tanh(x.double())
    assert y.dtype == dt and (y.double() - want).abs().max().item() <= (1e-6 if dt == torch.float32 else 4e-3)


def _yardstick_and_bounds(ref64, ref32):
    """These networks are ~110 batch-normalised layers deep and amplify rounding by three to four orders of magnitude at
    random initialisation: the oracle itself, run in fp32 instead of fp64 on the CPU, moves the output by 2e-4 .. 4e-4
    (measured at several sizes; 3e-6 for the MRU generator, 4e-7 for Pix2Pix).  So the 1e-3 bar of the MRU path is at the
    noise floor of an fp32 reference here, and the bounds are stated against that yardstick: the fp32 CUDA-core convolutions
    (every other kernel as in the product) within 30 yardsticks, the bf16x3 tensor-core convolutions (unit round-off 2^-16
    per product against 2^-24) within 100 (emulating the
    operand split in the oracle predicts 4e-3 here, 1.3e-2 for the background generator)."""
    yard = (ref32.double() - ref64).abs().max().item()
    return yard, max(INFER_TOL, 30 * yard), max(INFER_TOL, 100 * yard)


@pytest.mark.parametrize("cfg", [(8, 64, 64, 2), (64, 192, 192, 1)], ids=["size8_64px_n2", "size64_192px_n1"])
def test_generator_inference_parity(cfg):
    from oracle import fgcolor_oracle as O
    from oracle import residual_oracle as R
    size, H, W, N = cfg
    m = _model(size, H, W, torch.float32, with_discriminator=False)
    gp = {k: v.detach().cpu().double() for k, v in m.gstore.state_dict().items()}
    b = O.make_batch(N, H, W, 11, torch.float64, n_pad=4)
    b["text"][0, :9] = 0
    with torch.no_grad():
        ref = R.generator_forward(gp, b["sketch"], b["text"], b["cls"], b["noise"], size)
        ref32 = R.generator_forward({k: v.float() for k, v in gp.items()}, b["sketch"].float(), b["text"], b["cls"],
                                    b["noise"].float(), size)
    yard, bound_fp32, bound_tc = _yardstick_and_bounds(ref, ref32)
    db = _dev_batch(b)
    try:
        for impl, bound, tag in ((1, bound_fp32, "fp32 CUDA-core convolutions"), (0, bound_tc, "bf16x3 tensor-core convolutions")):
            m.ops.lib.fgc_set_conv_impl(impl)
            out = m.generate(db["sketch"], db["text"], db["cls"], db["noise"])
            torch.cuda.synchronize()
            err = (out.cpu().double() - ref).abs().max().item()
            print("residual generator, %s: max-abs err %.3e (fp32-oracle yardstick %.3e, bound %.3e)" % (tag, err, yard, bound))
            assert out.shape == ref.shape and torch.isfinite(out).all()
            assert err <= bound, "residual generator, %s: max-abs err %.3e > %.3e (yardstick %.3e)" % (tag, err, bound, yard)
    finally:
        m.ops.lib.fgc_set_conv_impl(0)
    # CudaOps(conv_terms=3): six-product split (the three bf16 terms of both operands).  Measured on a B200
    # (profiles/r2g_sixproduct_gpu_tests.log): it halves the bf16x3 error here (5.3e-3 -> 2.7e-3, 3.6e-3 -> 1.8e-3) but does
    # not reach the fp32 CUDA-core figure (3e-4 .. 6e-4) -- operand rounding is no longer the floor, the tensor core's fp32
    # accumulation is (tests/test_ops_gpu.py::test_six_product_convolution).  Bound: the bf16x3 one.
    from sketchyscenecolorization_b200.cuda_ops import CudaOps
    from sketchyscenecolorization_b200.trainer import FgColorModel
    m6 = FgColorModel(CudaOps("cuda:0", torch.float32, conv_terms=3), "cuda:0", size=size, H=H, W=W, with_discriminator=False,
                      block_type="Residual")
    m6.gstore.load_state_dict(m.gstore.state_dict())
    out = m6.generate(db["sketch"], db["text"], db["cls"], db["noise"])
    torch.cuda.synchronize()
    err = (out.cpu().double() - ref).abs().max().item()
    print("residual generator, six-product tensor-core convolutions: max-abs err %.3e (fp32-oracle yardstick %.3e)" % (err, yard))
    assert torch.isfinite(out).all() and err <= bound_tc


def test_training_graph_gradients():
    from oracle import fgcolor_oracle as O
    from oracle import residual_oracle as R
    size, H, W, N = 8, 64, 64, 2
    m = _model(size, H, W, torch.float32)
    m.ops.lib.fgc_set_conv_impl(1)          # fp32 CUDA-core convolutions: see _yardstick_and_bounds; every other kernel as in the product
    gp = {k: v.detach().cpu().double().requires_grad_(True) for k, v in m.gstore.state_dict().items()}
    dp = {k: v.detach().cpu().double().requires_grad_(True) for k, v in m.dstore.state_dict().items()}
    gspecs, dspecs = R.generator_specs(size, 58, H, W), R.discriminator_specs(size)
    b = O.make_batch(N, H, W, 5, torch.float64, n_pad=3)
    db = _dev_batch(b)

    def check(store, ref, tag):
        for s in store.specs:
            if s.trainable and s.reg > 0:
                store.g[s.name] += s.reg * store.p[s.name]
        gs = max(g.abs().max().item() for g in ref.values())
        worst_l2, dot, na, nb = 0.0, 0.0, 0.0, 0.0
        for k, g in ref.items():
            mine = store.g[k].detach().cpu().double()
            if g.abs().max().item() > 1e-3 * gs:
                worst_l2 = max(worst_l2, (mine - g).norm().item() / g.norm().item())
            dot, na, nb = dot + (mine * g).sum().item(), na + (mine * mine).sum().item(), nb + (g * g).sum().item()
        # ~50 batch-normalised layers deep with N = 2: held to a relative L2 bound per tensor and a global cosine
        assert worst_l2 <= 5e-2, "%s grads: worst relative L2 error %.3e" % (tag, worst_l2)
        assert dot / (na ** 0.5 * nb ** 0.5) >= 0.9995, "%s grads: cosine" % tag

    r = m.d_step_grads(db)
    ld, _, _ = R.d_step_loss(gp, dp, gspecs, dspecs, b, size)
    torch.cuda.synchronize()
    assert abs(r["loss"].item() - ld.item()) <= 1e-3 * abs(ld.item())
    check(m.dstore, O.grads_of(ld, dp, dspecs), "D")
    r = m.g_step_grads(db)
    lg, _, _, _ = R.g_step_loss(gp, dp, gspecs, dspecs, b, size)
    torch.cuda.synchronize()
    assert abs(r["loss"].item() - lg.item()) <= 1e-3 * abs(lg.item())
    check(m.gstore, O.grads_of(lg, gp, gspecs), "G")
    m.ops.lib.fgc_set_conv_impl(0)


def test_bf16_training_steps_run():
    from oracle import fgcolor_oracle as O
    from sketchyscenecolorization_b200.trainer import FgColorTrainer
    m = _model(8, 64, 64, torch.bfloat16)
    tr = FgColorTrainer(m, max_iter=100, use_cuda_graphs=True)
    before = m.gstore.flat.clone()
    for i in range(3):
        db = _dev_batch(O.make_batch(4, 64, 64, 20 + i, torch.float64))
        od, og = tr.d_step(db), tr.g_step(db)
    torch.cuda.synchronize()
    assert torch.isfinite(od["loss"]).item() and torch.isfinite(og["loss"]).item() and not torch.equal(before, m.gstore.flat)
