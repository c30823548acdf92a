"""CPU tests of the host-side network code (hand-derived backward passes, no autograd) against torch autograd on
the oracle, using the plain-torch operator set in tests/torch_ops.py (fp64).  This validates
blocks.py / generator.py / discriminator.py / text_fusion.py / trainer.py before any GPU time is spent;
the GPU tests then swap in the CUDA operator set under the same host code."""
import pytest
import torch

from oracle import fgcolor_oracle as O
from sketchyscenecolorization_b200.params import ParamStore, discriminator_vars, generator_vars
from sketchyscenecolorization_b200.trainer import FgColorModel, FgColorTrainer, lr_decay
from torch_ops import TorchOps

SIZE, H, W, N = 16, 64, 64, 3


@pytest.fixture(scope="module")
def setup():
    ops = TorchOps(torch.float64)
    m = FgColorModel(ops, "cpu", size=SIZE, H=H, W=W, param_dtype=torch.float64)
    m.initialize(seed=3, perturb_tables=0.1)
    gp = {k: v.clone().requires_grad_(True) for k, v in m.gstore.state_dict().items()}
    dp = {k: v.clone().requires_grad_(True) for k, v in m.dstore.state_dict().items()}
    b = O.make_batch(N, H, W, 5, torch.float64, n_pad=3)
    b["text"][0, :7] = 0
    bb = dict(b)
    bb["cls"], bb["cls_d"], bb["text"] = b["cls"].int(), b["cls_d"].int(), b["text"].numpy()
    return dict(ops=ops, m=m, gp=gp, dp=dp, b=b, bb=bb, gspecs=O.generator_specs(SIZE, 58, H, W), dspecs=O.discriminator_specs(SIZE))


def _worst(store, ref, ops):
    ops.add_reg_grad(store)      # the oracle differentiates the l2 decay; the product adds it inside the Adam kernel
    gs = max(g.abs().max().item() for g in ref.values())
    return max((store.g[k] - g).abs().max().item() / max(g.abs().max().item(), 1e-6 * gs) for k, g in ref.items())


def test_parameter_inventory_matches_survey():
    """G = 153 trainable tensors / 30 306 499 params, D = 60 / 17 159 752 (+23 SN u vectors)."""
    g = ParamStore(generator_vars(64, 58, 192, 192), "cpu")
    d = ParamStore(discriminator_vars(64), "cpu")
    assert (g.num_trainable_tensors(), g.num_params()) == (153, 30306499)
    assert (d.num_trainable_tensors(), d.num_params()) == (60, 17159752)
    assert len(d.state) == 23
    ospec = {s.name: s.shape for s in O.generator_specs(64, 58, 192, 192) + O.discriminator_specs(64)}
    mine = {s.name: tuple(s.shape) for s in generator_vars(64, 58, 192, 192) + discriminator_vars(64)}
    assert ospec == mine


def test_generator_forward_matches_oracle(setup):
    s = setup
    out = s["m"].generate(s["b"]["sketch"], s["bb"]["text"], s["bb"]["cls"], s["b"]["noise"])
    ref = O.generator_forward(s["gp"], s["b"]["sketch"], s["b"]["text"], s["b"]["cls"], s["b"]["noise"], SIZE)
    assert (out - ref).abs().max().item() < 1e-10


def test_d_step_gradients_match_autograd(setup):
    s = setup
    r = s["m"].d_step_grads(s["bb"])
    ld, _, _ = O.d_step_loss(s["gp"], s["dp"], s["gspecs"], s["dspecs"], s["b"], SIZE)
    assert abs(r["loss"].item() - ld.item()) < 1e-10
    assert _worst(s["m"].dstore, O.grads_of(ld, s["dp"], s["dspecs"]), s["ops"]) < 1e-7


def test_g_step_gradients_and_u_update(setup):
    s = setup
    m = s["m"]
    saved = {k: v.clone() for k, v in m.dstore.state.items()}
    r = m.g_step_grads(s["bb"])
    lg, _, u_new, _ = O.g_step_loss(s["gp"], s["dp"], s["gspecs"], s["dspecs"], s["b"], SIZE)
    assert abs(r["loss"].item() - lg.item()) < 1e-10
    assert _worst(m.gstore, O.grads_of(lg, s["gp"], s["gspecs"]), s["ops"]) < 1e-7
    for k, v in m.dstore.state.items():
        assert (v - u_new[k]).abs().max().item() < 1e-12
        v.copy_(saved[k])            # leave the fixture as it was


def test_all_pad_caption_and_lr_decay(setup):
    s = setup
    bb = dict(s["bb"])
    bb["text"] = bb["text"].copy()
    bb["text"][:] = 0
    b = dict(s["b"])
    b["text"] = torch.zeros_like(b["text"])
    out = s["m"].generate(b["sketch"], bb["text"], bb["cls"], b["noise"])
    ref = O.generator_forward(s["gp"], b["sketch"], b["text"], b["cls"], b["noise"], SIZE)
    assert (out - ref).abs().max().item() < 1e-10
    assert lr_decay(0, 100) == 1.0 and abs(lr_decay(50, 100) - 0.55) < 1e-12 and lr_decay(100, 100) == 0.2


def test_adam_matches_oracle_update(setup):
    s = setup
    m, ops = s["m"], s["ops"]
    tr = FgColorTrainer(m, lr_g=2e-4, lr_d=1e-4, max_iter=1000)
    before = m.dstore.flat.clone()
    tr.d_step(s["bb"])
    g = m.dstore.grad.clone()       # includes the decay term after the step
    want, _ = O.adam_update(before, g, torch.zeros_like(g), 1e-4 * lr_decay(0, 1000), 1)
    assert (m.dstore.flat - want).abs().max().item() < 1e-12
    m.dstore.flat.copy_(before)
    m.dstore.adam_v.zero_()
    m.dstore.adam_t = 0


@pytest.mark.parametrize("name", ["RMSprop", "AdaDelta", "AdaGrad"])
def test_other_optimizers_match_oracle_updates(setup, name, tmp_path):
    """--optimizer RMSprop | AdaDelta | AdaGrad (graph_single.get_optimizer, :584-593): two D steps against the oracle's
    restatement of the TF-1 update rules (slot initialisers included), and the slots survive a snapshot under TF's slot names."""
    from sketchyscenecolorization_b200 import checkpoint, tf_bundle
    s = setup
    ops = s["ops"]
    m = FgColorModel(ops, "cpu", size=SIZE, H=H, W=W, param_dtype=torch.float64)
    m.initialize(seed=3, perturb_tables=0.1)
    tr = FgColorTrainer(m, lr_g=2e-4, lr_d=1e-4, max_iter=1000, optimizer=name)
    kind = name.lower()
    p = m.dstore.flat.clone()
    s1 = torch.full_like(p, {"rmsprop": 1.0, "adadelta": 0.0, "adagrad": 0.1}[kind])
    s2 = torch.zeros_like(p)
    assert torch.equal(m.dstore.adam_v, s1)
    for step in range(2):
        tr.d_step(s["bb"])
        g = m.dstore.grad.clone()                   # includes the decay term after the step
        lr = 1e-4 * lr_decay(tr.counter, 1000)
        if kind == "rmsprop":
            p, s1 = O.rmsprop_update(p, g, s1, lr)
        elif kind == "adadelta":
            p, s1, s2 = O.adadelta_update(p, g, s1, s2, lr)
        else:
            p, s1 = O.adagrad_update(p, g, s1, lr)
        assert (m.dstore.flat - p).abs().max().item() < 1e-12, (name, step)
        m.dstore.flat.copy_(p)                      # keep the two trajectories on the same point
    prefix = checkpoint.save(m, str(tmp_path), 1, 2)
    keys = set(tf_bundle.read_bundle(prefix))
    slot = {"rmsprop": "RMSProp", "adadelta": "Adadelta", "adagrad": "Adagrad"}[kind]
    assert "discriminator/Conv/weights/" + slot in keys and "discriminator/Conv/weights/Adam" not in keys
    m2 = FgColorModel(ops, "cpu", size=SIZE, H=H, W=W, param_dtype=torch.float64)
    FgColorTrainer(m2, optimizer=name)
    checkpoint.restore(m2, prefix)
    for k, v in m.dstore.p.items():
        o = m.dstore.offsets[k]
        assert torch.equal(m2.dstore.adam_v[o:o + v.numel()].float(), m.dstore.adam_v[o:o + v.numel()].float()), k
        if kind == "adadelta":
            assert torch.equal(m2.dstore.opt_s2[o:o + v.numel()].float(), m.dstore.opt_s2[o:o + v.numel()].float()), k


def test_session_loop_snapshots_resume_and_nan_status(tmp_path, monkeypatch):
    """main_procedure.train (main_procedure.py:62-242) on a resident CPU model: alternating D / G steps on fresh batches,
    summaries every `summary_write_freq`, snapshots at i % save_model_freq == save_model_freq - 1 in the reference's layout,
    resume from the latest one (iter_from = its step + 1, obj_colorization_main.py:58-62), status -1 on a NaN loss."""
    import json
    import os
    import obj_colorization_main as M
    from sketchyscenecolorization_b200 import checkpoint, main_procedure
    from sketchyscenecolorization_b200.config import Config
    monkeypatch.chdir(tmp_path)
    ops = TorchOps(torch.float32)
    m = FgColorModel(ops, "cpu", size=16, H=64, W=64)
    m.initialize(seed=1)
    params = dict(dataset_type="train", resume_from="", batch_size=2, max_iter_step=3, disc_iterations=1, optimizer="Adam", lr_G=2e-4,
                  lr_D=1e-4, num_gpu=1, small_img=1, distance_map=0, LSTM_hybrid=1, block_type="MRU", vocab_size=58, ld=10,
                  extra_info="", summary_write_freq=1, save_model_freq=2, count_left_time_freq=1, count_inception_score_freq=-1,
                  infer_name="", instruction="")
    before = m.gstore.flat.clone()
    status, stamp = M.launch_training(model=m, synthetic_input=True, **params)
    assert status == 0 and len(stamp.split("-")) == 6
    run = tmp_path / "outputs" / stamp
    assert json.load(open(run / "log" / "param_0.json"))["batch_size"] == 2
    assert checkpoint.latest_checkpoint(str(run / "snapshot")).endswith("model_1.ckpt-1")            # saved at i = 1 only
    recs = [json.loads(ln) for ln in open(run / "log" / "summaries.jsonl")]
    assert [r["step"] for r in recs] == [0, 1, 2] and all(k in recs[0] for k in ("GAN_loss/G", "ACGAN_loss/D", "l1_perceptual_loss"))
    assert not torch.equal(before, m.gstore.flat)
    # resume: a fresh model picks up weights, optimiser slots and the counter from the snapshot and continues at iteration 2
    m2 = FgColorModel(ops, "cpu", size=16, H=64, W=64)
    m2.initialize(seed=99)
    params2 = dict(params, resume_from=stamp, max_iter_step=4)
    status, stamp2 = M.launch_training(model=m2, synthetic_input=True, **params2)
    assert status == 0 and stamp2 == stamp and os.path.exists(run / "log" / "param_2.json")
    assert checkpoint.latest_checkpoint(str(run / "snapshot")).endswith("model_3.ckpt-3")
    assert m2.dstore.adam_t == 4                                                                    # 2 restored + 2 new steps
    # NaN: the loop returns -1 (the caller restarts from the latest snapshot)
    from sketchyscenecolorization_b200.input_pipeline import SyntheticInput

    class Poisoned(SyntheticInput):
        def __next__(self):
            b = super().__next__()
            b["images_d"][0, 0, 0, 0] = float("nan")
            return b
    Config.set_from_dict(dict(params2, log_dir=str(run / "log"), ckpt_dir=str(run / "snapshot"), max_iter_step=6))
    assert main_procedure.train(iter_from=4, model=m2, synthetic_input=True, input_iter_d=Poisoned(2, 64, 64, seed=5)) == -1
    # no dataset and no explicit request for synthetic batches: fail like the reference's os.listdir (input_pipeline.py:133)
    with pytest.raises(FileNotFoundError):
        main_procedure.train(iter_from=4, model=m2)


def test_width_preserving_encoder_block(setup):
    """size 8: generator unit 1 maps 8 -> 8 channels, so the block has no 1x1 skip convolution (mru.py:446-452 is guarded by
    cin != cout) and the hidden state joins the sum as it is -- forward and both gradient sets against the oracle."""
    ops = setup["ops"]
    m = FgColorModel(ops, "cpu", size=8, H=64, W=64, param_dtype=torch.float64)
    m.initialize(seed=3, perturb_tables=0.1)
    assert "generator/mru_conv_unit_t_1_layer_0/Conv_3/weights" not in m.gstore.p
    gp = {k: v.clone().requires_grad_(True) for k, v in m.gstore.state_dict().items()}
    dp = {k: v.clone().requires_grad_(True) for k, v in m.dstore.state_dict().items()}
    gspecs, dspecs = O.generator_specs(8, 58, 64, 64), O.discriminator_specs(8)
    b, bb = setup["b"], setup["bb"]
    r = m.d_step_grads(bb)
    ld, _, _ = O.d_step_loss(gp, dp, gspecs, dspecs, b, 8)
    assert abs(r["loss"].item() - ld.item()) < 1e-10
    assert _worst(m.dstore, O.grads_of(ld, dp, dspecs), ops) < 1e-7
    r = m.g_step_grads(bb)
    lg, _, _, _ = O.g_step_loss(gp, dp, gspecs, dspecs, b, 8)
    assert abs(r["loss"].item() - lg.item()) < 1e-10
    assert _worst(m.gstore, O.grads_of(lg, gp, gspecs), ops) < 1e-7


def test_generator_without_caption(setup):
    """--lstm_hybrid 0 (models_collection.py:298-302): the bottleneck feeds the decoder as it is; forward and generator
    gradients against the oracle (the caption-encoder variables receive no gradient)."""
    ops, b, bb = setup["ops"], setup["b"], setup["bb"]
    m = FgColorModel(ops, "cpu", size=SIZE, H=H, W=W, param_dtype=torch.float64, lstm_hybrid=False)
    m.initialize(seed=3, perturb_tables=0.1)
    gp = {k: v.clone().requires_grad_(True) for k, v in m.gstore.state_dict().items()}
    dp = {k: v.clone().requires_grad_(True) for k, v in m.dstore.state_dict().items()}
    out = m.generate(b["sketch"], bb["text"], bb["cls"], b["noise"])
    ref = O.generator_forward(gp, b["sketch"], b["text"], b["cls"], b["noise"], SIZE, lstm_hybrid=False)
    assert (out - ref).abs().max().item() < 1e-10
    r = m.g_step_grads(bb)
    rd, rl = O.discriminator_forward(dp, b["images_d"], SIZE)
    fd, fl = O.discriminator_forward(dp, ref, SIZE)
    lg, _, _ = O.losses(rd, rl, fd, fl, b["cls_d"], b["cls"], b["images"], ref, O.reg_loss(gp, setup["gspecs"]),
                        O.reg_loss(dp, setup["dspecs"]))
    assert abs(r["loss"].item() - lg.item()) < 1e-10
    grads = O.grads_of(lg, gp, setup["gspecs"])
    assert all(float(g.abs().max()) == 0.0 for k, g in grads.items() if "TextLSTM" in k)
    assert _worst(m.gstore, {k: g for k, g in grads.items() if "TextLSTM" not in k}, ops) < 1e-7
    assert float(m.gstore.g["generator/TextLSTM/embedding"].abs().max()) == 0.0


def test_word_lstm_fallback_path_equals_sequence_operator(setup):
    """text_fusion steps the word LSTM a cell at a time where the operator set declines the size (lstm_seq_supported False:
    hidden sizes beyond 512, the background generator): same forward and same gradients as the one-launch operator."""
    from sketchyscenecolorization_b200 import text_fusion as TF

    class NoSeq(TorchOps):
        def lstm_seq_supported(self, N, D):
            return False

    outs = []
    for cls in (TorchOps, NoSeq):
        ops = cls(torch.float64)
        m = FgColorModel(ops, "cpu", size=8, H=64, W=64, param_dtype=torch.float64)
        m.initialize(seed=3, perturb_tables=0.1)
        g = torch.Generator().manual_seed(0)
        e4 = torch.randn(3, 2, 2, 64, generator=g, dtype=torch.float64)
        ids = torch.tensor([[0, 0, 5, 7, 9], [0, 0, 0, 0, 0], [2, 4, 6, 8, 10]], dtype=torch.int32)
        out, ctx = TF.text_fusion_fwd(ops, m.gstore, e4, ids.numpy())
        m.gstore.grad.zero_()
        g_e4 = TF.text_fusion_bwd(ops, m.gstore, torch.randn(out.shape, generator=g, dtype=torch.float64), ctx)
        outs.append((out, g_e4, m.gstore.grad.clone()))
    for a, b in zip(*outs):
        assert (a - b).abs().max().item() <= 1e-12 * max(1.0, b.abs().max().item())
