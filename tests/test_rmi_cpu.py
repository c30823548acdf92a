"""Instance-matching model (BASELINE.json configs[4]) on the CPU: the oracle against vectors produced by the reference's own
helpers, the oracle's two evaluation orders against each other, and the host code (sketchyscenecolorization_b200/rmi.py) on
the plain-torch operator set against the oracle in fp64."""
import json
import os

import numpy as np
import pytest
import torch

from oracle import rmi_oracle as O
from oracle.fgcolor_oracle import init_params
from sketchyscenecolorization_b200 import rmi
from torch_ops import TorchOps

GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "rmi_helpers.json")))
UNITS, FILTERS = (2, 2, 3, 2), (8, 16, 32, 48, 64)
DIMS = dict(vocab_size=30, w_emb=12, v_emb=20, m_rnn=10, w_rnn=14)


def test_helpers_against_reference_vectors():
    vocab = {w: i for i, w in enumerate(GOLD["vocab"])}
    for c in GOLD["sentences"]:
        assert O.preprocess_sentence(c["sentence"], vocab, GOLD["T"]) == (c["ids"], c["len"])
        assert rmi.preprocess_sentence(c["sentence"], vocab, GOLD["T"]) == (c["ids"], c["len"])
    for g in GOLD["spatial"]:
        want = np.asarray(g["values"], dtype=np.float32).reshape(g["N"], g["h"], g["w"], 8)
        assert np.array_equal(O.generate_spatial_batch(g["N"], g["h"], g["w"]), want)
        assert np.array_equal(rmi.spatial_rows(g["N"], g["h"], g["w"]).reshape(want.shape), want)


def _randomise_moments(P, seed):
    """Non-trivial stored moments: at their initial values (0, 1, factor 1) the batch norms are almost the identity."""
    g = torch.Generator().manual_seed(seed)
    for k, v in P.items():
        if k.endswith("/mean") or k.endswith("/beta"):
            v.copy_(torch.randn(v.shape, generator=g, dtype=torch.float64) * 0.1)
        elif k.endswith("/variance") or k.endswith("/gamma"):
            v.copy_(0.5 + torch.rand(v.shape, generator=g, dtype=torch.float64))
        elif k.endswith("/factor"):
            v.fill_(1.3)
        elif k.endswith("/bias") or k.endswith("/biases"):
            v.copy_(torch.randn(v.shape, generator=g, dtype=torch.float64) * 0.1)


def _setup(N=2, S=64, T=6, seed=5):
    P = init_params(O.model_specs(UNITS, FILTERS, **DIMS), seed, torch.float64)
    _randomise_moments(P, seed + 1)
    g = torch.Generator().manual_seed(seed + 2)
    im = torch.randn(N, S, S, 3, generator=g, dtype=torch.float64) * 60.0
    words = torch.randint(1, DIMS["vocab_size"], (N, T), generator=g)
    lengths = torch.tensor([3, T][:N] + [0] * max(0, N - 2))
    return P, im, words, lengths


def test_parameter_inventory():
    from sketchyscenecolorization_b200.params import rmi_vars
    a = [(s.name, tuple(s.shape)) for s in rmi_vars()]
    b = [(s.name, tuple(s.shape)) for s in O.model_specs()]
    assert a == b
    convs = [n for n, _ in a if n.endswith("/DW") and n.startswith("ResNet/")]
    assert len(convs) == 1 + 3 * 33 + 4           # ResNet-101: stem + 33 bottlenecks x 3 + 4 projection shortcuts
    assert dict(a)["text_sketchyscene/mLSTM/lstm_cell/kernel"] == (3508, 2000)
    assert dict(a)["text_sketchyscene/wLSTM/lstm_cell/kernel"] == (2000, 4000)


def test_oracle_hoisted_equals_literal():
    P, im, words, lengths = _setup(N=3)
    with torch.no_grad():
        feat = O.trunk_forward(P, im, UNITS, FILTERS)
        a = O.fusion_forward(P, feat, words, lengths, 64, 64, hoisted=False)
        b = O.fusion_forward(P, feat, words, lengths, 64, 64, hoisted=True)
    for x, y in zip(a, b):
        assert (x - y).abs().max().item() <= 1e-12
    # sample 2 has length 0: dynamic_rnn leaves the zero state, the prediction is the projection bias
    assert torch.allclose(a[0][2], P["text_sketchyscene/m_lstm_output_projection/biases"].expand_as(a[0][2]))


def test_atrous_as_space_to_batch():
    """tf.nn.atrous_conv2d(rate) == batch_to_space(SAME conv(space_to_batch)): the form the trunk's dilated groups run in."""
    ops = TorchOps(torch.float64)
    g = torch.Generator().manual_seed(0)
    x = torch.randn(2, 8, 12, 5, generator=g, dtype=torch.float64)
    w = torch.randn(3, 3, 5, 7, generator=g, dtype=torch.float64)
    for r in (2, 4):
        want = O.conv(x.permute(0, 3, 1, 2), w, 1, r).permute(0, 2, 3, 1)
        xb = ops.space_to_batch(x, r)
        assert xb.shape == (2 * r * r, 8 // r, 12 // r, 5)
        got = ops.batch_to_space(ops.conv_fwd([(xb, False)], w, None), r)
        assert (got - want).abs().max().item() <= 1e-12
        assert torch.equal(ops.batch_to_space(xb, r), x)


@pytest.mark.parametrize("N", [1, 3])
def test_host_model_against_oracle(N):
    P, im, words, lengths = _setup(N=N)
    with torch.no_grad():
        feat_ref = O.trunk_forward(P, im, UNITS, FILTERS)
        pred_ref, up_ref, sg_ref = O.fusion_forward(P, feat_ref, words, lengths, 64, 64)
    m = rmi.RMIModel(TorchOps(torch.float64), "cpu", units=UNITS, filters=FILTERS, param_dtype=torch.float64, **DIMS)
    m.load_state_dict(P)
    feat = m.trunk(im)
    assert feat.shape == feat_ref.shape and (feat - feat_ref).abs().max().item() <= 1e-9 * max(1.0, feat_ref.abs().max().item())
    pred, up, sg = m.fuse(feat, words.numpy(), lengths.numpy(), 64, 64)
    assert (pred - pred_ref).abs().max().item() <= 1e-9
    assert (up - up_ref).abs().max().item() <= 1e-9 and (sg - sg_ref).abs().max().item() <= 1e-9
    up2, sg2 = m.forward(im, words, lengths)
    assert torch.equal(up2, up) and up2.shape == (N, 64, 64, 1)


def test_predict_mask_boundary():
    P, _, _, _ = _setup(N=1)
    m = rmi.RMIModel(TorchOps(torch.float64), "cpu", units=UNITS, filters=FILTERS, param_dtype=torch.float64, **DIMS)
    m.load_state_dict(P)
    vocab = {w: i for i, w in enumerate(GOLD["vocab"][:DIMS["vocab_size"]])}
    rs = np.random.RandomState(0)
    sk = np.where(rs.rand(64, 64, 1) < 0.1, 0, 255).astype(np.uint8).repeat(3, axis=2)
    mask = m.predict_mask(sk, "the dog on the right", vocab, T=6)
    assert mask.shape == (64, 64) and set(np.unique(mask)).issubset({0.0, 1.0})
    assert (mask[sk[:, :, 0] == 255] == 0).all()
