"""CPU tests of `--block_type Pix2Pix` (pix2pix.py: hand-derived backward passes over the phase-form 4x4 layers) against torch
autograd on oracle/pix2pix_oracle.py, with the plain-torch operator set (fp64) -- the same chain as tests/test_host_cpu.py."""
import pytest
import torch

from oracle import fgcolor_oracle as O
from oracle import pix2pix_oracle as P
from sketchyscenecolorization_b200.params import ParamStore, pix2pix_discriminator_vars, pix2pix_generator_vars
from sketchyscenecolorization_b200.trainer import FgColorModel, FgColorTrainer
from torch_ops import TorchOps

SIZE, H, W, N = 16, 64, 64, 3


@pytest.fixture(scope="module")
def setup():
    ops = TorchOps(torch.float64)
    m = FgColorModel(ops, "cpu", size=SIZE, H=H, W=W, param_dtype=torch.float64, block_type="Pix2Pix")
    m.initialize(seed=3, perturb_tables=0.1)
    gp = {k: v.clone().requires_grad_(True) for k, v in m.gstore.state_dict().items()}
    dp = {k: v.clone().requires_grad_(True) for k, v in m.dstore.state_dict().items()}
    b = O.make_batch(N, H, W, 5, torch.float64, n_pad=3)
    b["text"][0, :7] = 0
    bb = dict(b)
    bb["cls"], bb["cls_d"], bb["text"] = b["cls"].int(), b["cls_d"].int(), b["text"].numpy()
    return dict(ops=ops, m=m, gp=gp, dp=dp, b=b, bb=bb, gspecs=P.generator_specs(SIZE, 58, H, W), dspecs=P.discriminator_specs(SIZE))


def _worst(store, ref, ops):
    ops.add_reg_grad(store)
    gs = max(g.abs().max().item() for g in ref.values())
    return max((store.g[k] - g).abs().max().item() / max(g.abs().max().item(), 1e-6 * gs) for k, g in ref.items())


def test_phase_forms_equal_the_4x4_layers():
    """space_to_depth + 3x3 SAME == 4x4 stride 2 pad 1; 3x3 SAME + depth_to_space == conv2d_transpose(4x4, stride 2, SAME);
    5x5 SAME cropped == 4x4 stride 1 pad 1; phase_wgrad is the adjoint of phase_weights."""
    ops = TorchOps(torch.float64)
    g = torch.Generator().manual_seed(0)
    x = torch.randn(2, 8, 12, 5, generator=g, dtype=torch.float64)
    f = torch.randn(4, 4, 5, 7, generator=g, dtype=torch.float64)
    ft = torch.randn(4, 4, 7, 5, generator=g, dtype=torch.float64)
    nchw, nhwc = (lambda t: t.permute(0, 3, 1, 2)), (lambda t: t.permute(0, 2, 3, 1))
    got = ops.conv_fwd([(ops.space_to_depth(x), False)], ops.phase_weights(f, "conv"), None)
    assert (got - nhwc(P.nchw_conv(nchw(x), f, 2))).abs().max() < 1e-12
    got = ops.depth_to_space(ops.conv_fwd([(x, False)], ops.phase_weights(ft, "deconv"), None))
    assert (got - nhwc(P.nchw_deconv(nchw(x), ft))).abs().max() < 1e-12
    got = ops.copy_rect(ops.conv_fwd([(x, False)], ops.phase_weights(f, "k5"), None), 7, 11)
    assert (got - nhwc(P.nchw_conv(nchw(x), f, 1))).abs().max() < 1e-12
    assert torch.equal(ops.depth_to_space(ops.space_to_depth(x)), x)
    assert torch.equal(ops.copy_rect(ops.copy_rect(x, 7, 11), 8, 12)[:, :7, :11], x[:, :7, :11])
    for mode, filt in (("conv", f), ("deconv", ft), ("k5", f)):
        w = ops.phase_weights(filt, mode)
        assert int((w != 0).sum()) == filt.numel()                     # every tap lands exactly once
        r = torch.randn(w.shape, generator=g, dtype=torch.float64)
        df = torch.zeros_like(filt)
        ops.phase_wgrad(r, df, mode)
        assert abs(float((w * r).sum()) - float((filt * df).sum())) < 1e-10    # <P f, r> == <f, P^T r>


def test_parameter_inventory():
    """Variable names / shapes of the two networks agree with the oracle's reading of the reference scopes; pix2pix at 192^2:
    generator filters 4x4 (3-64-128-256-512-512 down, 576-512 / 1024-256 / 512-128 / 256-64 / 128-3 up)."""
    ospec = {s.name: s.shape for s in P.generator_specs(64, 58, 192, 192) + P.discriminator_specs(64)}
    mine = {s.name: tuple(s.shape) for s in pix2pix_generator_vars(64, 58, 192, 192) + pix2pix_discriminator_vars(64)}
    assert ospec == mine
    g = ParamStore(pix2pix_generator_vars(64, 58, 192, 192), "cpu")
    assert g.p["generator/decoder_5/deconv/filter"].shape == (4, 4, 512, 576)
    assert g.p["generator/decoder_1/deconv/filter"].shape == (4, 4, 3, 128)
    assert g.p["generator/fully_connected/weights"].shape == (256, 64 * 6 * 6)
    d = ParamStore(pix2pix_discriminator_vars(64), "cpu")
    assert d.p["discriminator/layer_4/conv/filter"].shape == (4, 4, 256, 512) and len(d.state) == 1


def test_generator_forward_matches_oracle(setup):
    s = setup
    out = s["m"].generate(s["b"]["sketch"], s["bb"]["text"], s["bb"]["cls"], s["b"]["noise"])
    ref = P.generator_forward(s["gp"], s["b"]["sketch"], s["b"]["text"], s["b"]["cls"], s["b"]["noise"], SIZE)
    assert out.shape == ref.shape == (N, 3, H, W)
    assert (out - ref).abs().max().item() < 1e-10


def test_discriminator_forward_matches_oracle(setup):
    s = setup
    m, ops = s["m"], s["ops"]
    wv = m.D.new_weight_view(need_wgrad=False)
    d, lg, _ = m.D.forward(ops.nchw_to_nhwc(s["b"]["sketch"]), ops.nchw_to_nhwc(s["b"]["images_d"]), wv, save=False)
    rd, rl = P.discriminator_forward(s["dp"], s["b"]["sketch"], s["b"]["images_d"], SIZE)
    assert d.shape == (N, H // 8 - 2, W // 8 - 2, 1) and rd.shape == (N, 1, H // 8 - 2, W // 8 - 2)
    assert (d.permute(0, 3, 1, 2) - rd).abs().max().item() < 1e-10
    assert (lg.reshape(N, -1) - rl).abs().max().item() < 1e-10


def test_d_step_gradients_match_autograd(setup):
    s = setup
    r = s["m"].d_step_grads(s["bb"])
    ld, _, _ = P.d_step_loss(s["gp"], s["dp"], s["gspecs"], s["dspecs"], s["b"], SIZE)
    assert abs(r["loss"].item() - ld.item()) < 1e-10
    assert _worst(s["m"].dstore, O.grads_of(ld, s["dp"], s["dspecs"]), s["ops"]) < 1e-7


def test_g_step_gradients_and_u_update(setup):
    s = setup
    m = s["m"]
    saved = {k: v.clone() for k, v in m.dstore.state.items()}
    r = m.g_step_grads(s["bb"])
    lg, _, u_new, _ = P.g_step_loss(s["gp"], s["dp"], s["gspecs"], s["dspecs"], s["b"], SIZE)
    assert abs(r["loss"].item() - lg.item()) < 1e-10
    assert _worst(m.gstore, O.grads_of(lg, s["gp"], s["gspecs"]), s["ops"]) < 1e-7
    for k, v in m.dstore.state.items():
        assert (v - u_new[k]).abs().max().item() < 1e-12
        v.copy_(saved[k])


def test_generator_without_text_and_training_steps(setup):
    """--lstm_hybrid 0 (the bottleneck feeds the decoder directly) and two optimiser steps of the alternating loop."""
    ops = setup["ops"]
    m = FgColorModel(ops, "cpu", size=SIZE, H=H, W=W, param_dtype=torch.float64, block_type="Pix2Pix", lstm_hybrid=False)
    m.initialize(seed=4)
    gp = {k: v.clone().requires_grad_(True) for k, v in m.gstore.state_dict().items()}
    dp = {k: v.clone().requires_grad_(True) for k, v in m.dstore.state_dict().items()}
    b, bb = setup["b"], setup["bb"]
    r = m.g_step_grads(bb)
    fake = P.generator_forward(gp, b["sketch"], b["text"], b["cls"], b["noise"], SIZE, lstm_hybrid=False)
    rd, rl = P.discriminator_forward(dp, b["sketch"], b["images_d"], SIZE)
    fd, fl = P.discriminator_forward(dp, b["sketch"], fake, SIZE)
    gspecs = P.generator_specs(SIZE, 58, H, W)
    lg, _, _ = O.losses(rd, rl, fd, fl, b["cls_d"], b["cls"], b["images"], fake, O.reg_loss(gp, gspecs), O.reg_loss(dp, setup["dspecs"]))
    assert abs(r["loss"].item() - lg.item()) < 1e-10
    ref = {k: g for k, g in O.grads_of(lg, gp, gspecs).items() if "TextLSTM" not in k}
    assert _worst(m.gstore, ref, ops) < 1e-7
    tr = FgColorTrainer(m, max_iter=10)
    before = m.gstore.flat.clone()
    od, og = tr.d_step(bb), tr.g_step(bb)
    assert torch.isfinite(od["loss"]) and torch.isfinite(og["loss"]) and not torch.equal(before, m.gstore.flat)


def test_snapshot_roundtrip_and_graph_builder(setup, tmp_path):
    """Snapshots of a Pix2Pix model carry the reference's variable names (TF V2 bundle) and restore bit for bit; the graph
    builder refuses a block type the bound model was not built with."""
    from sketchyscenecolorization_b200 import checkpoint, graph_single, tf_bundle
    m, ops, b, bb = setup["m"], setup["ops"], setup["b"], setup["bb"]
    checkpoint.save(m, str(tmp_path), 7, 8)
    prefix = checkpoint.latest_checkpoint(str(tmp_path))
    keys = set(tf_bundle.read_bundle(prefix))
    for k in ("generator/encoder_1/conv/filter", "generator/decoder_5/deconv/filter/Adam_1", "generator/encoder_3/scale",
              "discriminator/layer_4/offset", "discriminator/fully_connected/discriminator/fully_connected/u", "Variable"):
        assert k in keys, k
    m2 = FgColorModel(ops, "cpu", size=SIZE, H=H, W=W, param_dtype=torch.float64, block_type="Pix2Pix")
    m2.initialize(seed=99)
    assert checkpoint.restore(m2, prefix) == 8
    assert torch.equal(m2.gstore.flat.float(), m.gstore.flat.float()) and torch.equal(m2.dstore.flat.float(), m.dstore.flat.float())
    ret = graph_single.build_single_graph(b["images"], b["sketch"], None, bb["cls"], None, bb["text"], batch_size=N, training=False,
                                          LSTM_hybrid=True, vocab_size=58, block_type="Pix2Pix", model=m, noise=b["noise"])
    ref = P.generator_forward(setup["gp"], b["sketch"], b["text"], b["cls"], b["noise"], SIZE)
    assert (ret[0].double() - ref).abs().max().item() < 1e-5          # the builder feeds fp32 tensors
    with pytest.raises(ValueError):
        graph_single.build_single_graph(b["images"], b["sketch"], None, bb["cls"], None, bb["text"], batch_size=N, training=False,
                                        LSTM_hybrid=True, vocab_size=58, block_type="MRU", model=m, noise=b["noise"])
    with pytest.raises(NotImplementedError):
        FgColorModel(ops, "cpu", size=SIZE, H=H, W=W, block_type="Unet")
