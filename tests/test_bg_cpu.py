"""CPU tests of the background-colorization generator (bg.py, forward only) against oracle/bg_oracle.py with the plain-torch
operator set (fp64), and of the bg_colorization_main.py test loop on a small synthetic data directory."""
import json
import os

import numpy as np
import pytest
import torch

from oracle import bg_oracle as B
from oracle import fgcolor_oracle as O
from sketchyscenecolorization_b200.bg import BgColorModel
from sketchyscenecolorization_b200.params import ParamStore, bg_generator_vars
from torch_ops import TorchOps

NGF, S, N, T = 4, 64, 2, 8


def test_parameter_inventory():
    """Names / shapes agree with the oracle's reading of the reference scopes at the published size (ngf 64, vocab 18)."""
    ospec = {s.name: s.shape for s in B.generator_specs(64, 18)}
    gv = bg_generator_vars(64, 18)
    assert ospec == {s.name: tuple(s.shape) for s in gv}
    g = ParamStore(gv, "cpu")
    assert g.p["generator/encoder_1/conv_ex/filter"].shape == (7, 7, 3, 64)
    assert g.p["generator/encoder_5_2/block_3/conv_ex/filter"].shape == (1, 1, 256, 1024)
    assert g.p["generator/mLSTM_G/RNN/ALSTM/multi_rnn_cell/cell_0/basic_lstm_cell/kernel"].shape == (4096, 4096)
    assert g.p["generator/decoder_5_0/block_1/deconv/filter"].shape == (4, 4, 128, 1024)
    assert g.p["generator/region_br_3/deconv/filter"].shape == (4, 4, 3, 3)
    assert g.p["generator/decoder_1/batchnorm/scale"].shape == (3,)


def test_generator_forward_matches_oracle():
    ops = TorchOps(torch.float64)
    m = BgColorModel(ops, "cpu", ngf=NGF, vocab_size=18, param_dtype=torch.float64)
    m.initialize(seed=2, perturb_tables=0.1)
    gp = m.gstore.state_dict()
    g = torch.Generator().manual_seed(4)
    img = torch.rand(N, 3, S, S, generator=g, dtype=torch.float64) * 2 - 1
    ids = torch.randint(2, 18, (N, T), generator=g)
    ids[0, :5] = 0
    out, reg = m.generate(img.permute(0, 2, 3, 1).contiguous(), ids.numpy())
    ref_out, ref_reg = B.generator_forward(gp, img, ids)
    assert out.shape == (N, S, S, 3) and reg.shape == (N, S, S, 3)
    assert (out.permute(0, 3, 1, 2) - ref_out).abs().max().item() < 1e-8
    assert (reg.permute(0, 3, 1, 2) - ref_reg).abs().max().item() < 1e-8
    assert float(reg.min()) >= 0.0 and float(out.abs().max()) < 1.0


def test_uint8_boundary_follows_convert_image_dtype():
    ops = TorchOps(torch.float32)
    m = BgColorModel(ops, "cpu", ngf=NGF, vocab_size=18)
    m.initialize(seed=5)
    rng = np.random.default_rng(0)
    u8 = rng.integers(0, 256, (1, 64, 64, 3), dtype=np.uint8)
    ids = np.array([[0, 0, 0, 2, 3, 4, 5, 7]], dtype=np.int32)
    pic, seg = m.colorize_u8(u8, ids)
    assert pic.dtype == np.uint8 and pic.shape == (64, 64, 3) and seg.shape == (64, 64) and set(np.unique(seg)) <= {0, 1, 2}
    out, _ = m.generate(torch.from_numpy(u8).float() / 255 * 2 - 1, ids)
    want = np.floor(np.clip((out[0].numpy() + 1) / 2 * 255.5, 0, 255)).astype(np.uint8)
    assert np.array_equal(pic, want)


def test_test_mode_end_to_end(tmp_path, monkeypatch):
    """bg_colorization_main.bg_colorization(mode='test'): data/{foreground,background,segment}/test + captions/test.json ->
    outputs/<ts>/results/<name>_{inputs,outputs,targets}.png, foreground pasted back through the segment mask."""
    import cv2
    import bg_colorization_main as M
    base = tmp_path / "data"
    for d in ("foreground/test", "background/test", "segment/test", "captions"):
        os.makedirs(base / d)
    rng = np.random.default_rng(1)
    fg = rng.integers(0, 256, (64, 64, 3), dtype=np.uint8)
    bgp = rng.integers(0, 256, (64, 64, 3), dtype=np.uint8)
    seg = np.full((64, 64, 3), 255, np.uint8)
    seg[10:30, 10:30] = 0                                  # 0 = foreground (:873-874)
    cv2.imwrite(str(base / "foreground/test/a.png"), fg[:, :, ::-1])
    cv2.imwrite(str(base / "background/test/a_bg.png"), bgp[:, :, ::-1])
    cv2.imwrite(str(base / "segment/test/a.png"), seg)
    json.dump([dict(fg_name="a.png", bg_name="a_bg.png", color_text="the sky is blue and the ground is green")],
              open(base / "captions/test.json", "w"))
    monkeypatch.chdir(tmp_path)
    m = BgColorModel(TorchOps(torch.float32), "cpu", ngf=NGF, vocab_size=18)
    m.initialize(seed=6)
    n = M.bg_colorization(mode="test", resume_from="run1", data_base_dir=str(base), image_size=64, text_len=8, vocab_size=18,
                          vocab_file="data/bg_vocab.txt", ngf=NGF, model=m)
    assert n == 1
    res = tmp_path / "outputs" / "run1" / "results"
    for kind in ("inputs", "outputs", "targets"):
        assert (res / ("a_bg_%s.png" % kind)).exists()
    out = cv2.imread(str(res / "a_bg_outputs.png"))[:, :, ::-1]
    assert np.array_equal(out[10:30, 10:30], fg[10:30, 10:30])               # foreground pasted back
    assert np.array_equal(cv2.imread(str(res / "a_bg_inputs.png"))[:, :, ::-1], fg)
    with pytest.raises(Exception):
        M.bg_colorization(mode="test", resume_from="", data_base_dir=str(base), image_size=64, model=m)
    ids = M.preprocess_sentence("the sky is blue and the ground is green", M.bg_vocab_dict(), 8)
    assert ids == [0, 2, 3, 4, 5, 8, 3, 7]


def test_snapshot_names_load_from_a_tf_bundle(tmp_path):
    """A `snapshot-<step>` bundle keyed by the reference's variable names (tf.train.Saver layout, :812-823) restores strictly."""
    from sketchyscenecolorization_b200 import checkpoint, tf_bundle
    m = BgColorModel(TorchOps(torch.float32), "cpu", ngf=NGF, vocab_size=18)
    m.initialize(seed=8)
    d = str(tmp_path / "snapshot")
    os.makedirs(d)
    tensors = {k: v.numpy() for k, v in m.gstore.state_dict().items()}
    tensors["global_step"] = np.int64(100)
    tf_bundle.write_bundle(os.path.join(d, "snapshot-100"), tensors)
    open(os.path.join(d, "checkpoint"), "w").write('model_checkpoint_path: "snapshot-100"\nall_model_checkpoint_paths: "snapshot-100"\n')
    prefix = checkpoint.latest_checkpoint(d)
    assert prefix.endswith("snapshot-100")
    m2 = BgColorModel(TorchOps(torch.float32), "cpu", ngf=NGF, vocab_size=18)
    m2.gstore.load_state_dict(tf_bundle.read_bundle(prefix), strict=True)
    assert torch.equal(m2.gstore.flat, m.gstore.flat)
