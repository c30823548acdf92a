"""CPU tests of `--block_type Residual` (residual.py: bottleneck blocks with hand-derived backward passes over direct and
phase-form convolutions) against torch autograd on oracle/residual_oracle.py, with the plain-torch operator set (fp64)."""
import pytest
import torch

from oracle import fgcolor_oracle as O
from oracle import residual_oracle as R
from sketchyscenecolorization_b200.params import ParamStore, residual_discriminator_vars, residual_generator_vars
from sketchyscenecolorization_b200.trainer import FgColorModel, FgColorTrainer
from torch_ops import TorchOps

SIZE, H, W, N = 8, 64, 64, 2


@pytest.fixture(scope="module")
def setup():
    ops = TorchOps(torch.float64)
    m = FgColorModel(ops, "cpu", size=SIZE, H=H, W=W, param_dtype=torch.float64, block_type="Residual")
    m.initialize(seed=3, perturb_tables=0.1)
    gp = {k: v.clone().requires_grad_(True) for k, v in m.gstore.state_dict().items()}
    dp = {k: v.clone().requires_grad_(True) for k, v in m.dstore.state_dict().items()}
    b = O.make_batch(N, H, W, 5, torch.float64, n_pad=3)
    b["text"][0, :7] = 0
    bb = dict(b)
    bb["cls"], bb["cls_d"], bb["text"] = b["cls"].int(), b["cls_d"].int(), b["text"].numpy()
    return dict(ops=ops, m=m, gp=gp, dp=dp, b=b, bb=bb, gspecs=R.generator_specs(SIZE, 58, H, W), dspecs=R.discriminator_specs(SIZE))


def _worst(store, ref, ops):
    ops.add_reg_grad(store)
    gs = max(g.abs().max().item() for g in ref.values())
    return max((store.g[k] - g).abs().max().item() / max(g.abs().max().item(), 1e-6 * gs) for k, g in ref.items())


def test_parameter_inventory():
    """Names / shapes agree with the oracle's reading of the reference scopes; 16 + 16 bottleneck blocks in the generator
    (units 3, 4, 6, 3 per level, models_collection.py:609), five stride-2 blocks in the discriminator."""
    ospec = {s.name: s.shape for s in R.generator_specs(64, 58, 192, 192) + R.discriminator_specs(64)}
    gv, dv = residual_generator_vars(64, 58, 192, 192), residual_discriminator_vars(64)
    assert ospec == {s.name: tuple(s.shape) for s in gv + dv}
    g = ParamStore(gv, "cpu")
    assert g.p["generator/encoder_1/conv_ex/filter"].shape == (7, 7, 3, 64)
    assert g.p["generator/encoder_4_5/block_1/conv_ex/filter"].shape == (4, 4, 512, 128)       # the sixth unit of level 3
    assert g.p["generator/decoder_5_0/block_add/deconv/filter"].shape == (4, 4, 512, 576)
    assert g.p["generator/decoder_1/deconv/filter"].shape == (4, 4, 3, 128)
    assert sum(1 for s in gv if s.name.endswith("block_3/conv_ex/filter")) == 32
    d = ParamStore(dv, "cpu")
    assert d.p["discriminator/layer_5/conv_ex/filter"].shape == (4, 4, 512, 1)
    assert d.p["discriminator/layer_1/block_1/conv/filter"].shape == (4, 4, 6, 16)


def test_generator_and_discriminator_forward_match_oracle(setup):
    s = setup
    m, ops, b = s["m"], s["ops"], s["b"]
    out = m.generate(b["sketch"], s["bb"]["text"], s["bb"]["cls"], b["noise"])
    ref = R.generator_forward(s["gp"], b["sketch"], b["text"], b["cls"], b["noise"], SIZE)
    assert out.shape == ref.shape == (N, 3, H, W)
    assert (out - ref).abs().max().item() < 1e-9
    wv = m.D.new_weight_view(need_wgrad=False)
    d, lg, _ = m.D.forward(ops.nchw_to_nhwc(b["sketch"]), ops.nchw_to_nhwc(b["images_d"]), wv, save=False)
    rd, rl = R.discriminator_forward(s["dp"], b["sketch"], b["images_d"], SIZE)
    assert d.shape == (N, H // 32, W // 32, 1)
    assert (d.permute(0, 3, 1, 2) - rd).abs().max().item() < 1e-9 and (lg.reshape(N, -1) - rl).abs().max().item() < 1e-9


def test_d_step_gradients_match_autograd(setup):
    s = setup
    r = s["m"].d_step_grads(s["bb"])
    ld, _, _ = R.d_step_loss(s["gp"], s["dp"], s["gspecs"], s["dspecs"], s["b"], SIZE)
    assert abs(r["loss"].item() - ld.item()) < 1e-9
    assert _worst(s["m"].dstore, O.grads_of(ld, s["dp"], s["dspecs"]), s["ops"]) < 1e-6


def test_g_step_gradients_and_u_update(setup):
    s = setup
    m = s["m"]
    saved = {k: v.clone() for k, v in m.dstore.state.items()}
    r = m.g_step_grads(s["bb"])
    lg, _, u_new, _ = R.g_step_loss(s["gp"], s["dp"], s["gspecs"], s["dspecs"], s["b"], SIZE)
    assert abs(r["loss"].item() - lg.item()) < 1e-9
    assert _worst(m.gstore, O.grads_of(lg, s["gp"], s["gspecs"]), s["ops"]) < 1e-6
    for k, v in m.dstore.state.items():
        assert (v - u_new[k]).abs().max().item() < 1e-12
        v.copy_(saved[k])


def test_training_steps_move_the_weights(setup):
    m = setup["m"]
    tr = FgColorTrainer(m, max_iter=10)
    before_g, before_d = m.gstore.flat.clone(), m.dstore.flat.clone()
    od, og = tr.d_step(setup["bb"]), tr.g_step(setup["bb"])
    assert torch.isfinite(od["loss"]) and torch.isfinite(og["loss"])
    assert not torch.equal(before_g, m.gstore.flat) and not torch.equal(before_d, m.dstore.flat)
