"""Whole-pipeline instance matching on the CPU: the mask -> instances rule against vectors produced by the reference's own
get_pred_instance_mask (tests/golden/make_match_golden.py), and build_instance_matching end to end on the plain-torch operator
set with a snapshot written in TensorFlow's bundle format."""
import json
import os

import numpy as np
import torch

from sketchyscenecolorization_b200 import pipeline_match as PM
from sketchyscenecolorization_b200 import rmi, tf_bundle
from torch_ops import TorchOps

G = os.path.join(os.path.dirname(__file__), "golden")
CASES = json.load(open(os.path.join(G, "match_cases.json")))
SEG = os.path.join(G, "match_seg_data.npz")


def test_mask_to_instances_matches_reference_vectors():
    S = CASES["size"]
    packed = np.load(os.path.join(G, "match_overall_masks.npz"))["overall"]
    for c, bits in zip(CASES["cases"], packed):
        overall = np.unpackbits(bits)[:S * S].reshape(S, S).astype(np.float32)
        masks, scores, boxes, cls, idx = PM.get_pred_instance_mask(SEG, overall)
        assert idx == c["matched"]
        if idx:
            assert np.allclose(scores, c["scores"], rtol=0, atol=1e-12) and [int(x) for x in cls] == c["class_ids"]
            assert list(masks.shape) == c["masks_shape"] and int(masks.sum()) == c["masks_sum"] and len(boxes) == len(idx)
        else:
            assert masks.size == 0 and scores.size == 0


def test_build_instance_matching_end_to_end(tmp_path):
    """Snapshot (TF V2 bundle + `checkpoint` state file) -> RMIModel -> stroke mask -> matched indices; the picture is built so
    that the model-independent part is checkable: whatever the random network predicts, the indices must be exactly those the
    rule gives for its own mask."""
    units, filters = (1, 1, 1, 1), (8, 16, 32, 48, 64)
    dims = dict(vocab_size=76, w_emb=12, v_emb=20, m_rnn=10, w_rnn=14)
    ops = TorchOps(torch.float32)
    m = rmi.RMIModel(ops, "cpu", units=units, filters=filters, **dims)
    m.initialize(seed=2)
    m.store.p["text_sketchyscene/m_lstm_output_projection/biases"].fill_(0.05)       # a positive score somewhere
    snap = tmp_path / "snapshots"
    snap.mkdir()
    tf_bundle.write_bundle(str(snap / "deeplab_RMI_iter_7.tfmodel"), {k: v.numpy() for k, v in m.store.state_dict().items()})
    (snap / "checkpoint").write_text('model_checkpoint_path: "deeplab_RMI_iter_7.tfmodel"\n')
    vocab = tmp_path / "vocab.txt"
    vocab.write_text("\n".join(["<pad>", "<unk>", "the", "dog", "on", "right"] + ["w%d" % i for i in range(70)]) + "\n")
    from PIL import Image
    rs = np.random.RandomState(0)
    sk = np.where(rs.rand(768, 768, 1) < 0.08, 0, 255).astype(np.uint8).repeat(3, axis=2)
    Image.fromarray(sk).save(tmp_path / "scene.png")

    class Small(rmi.RMIModel):                      # load_matching_model builds the published widths; the test uses small ones
        def __init__(self, ops_, device, vocab_size):
            super().__init__(ops_, device, units=units, filters=filters, **dict(dims, vocab_size=vocab_size))
    orig = PM.rmi.RMIModel
    PM.rmi.RMIModel = Small
    try:
        got = PM.build_instance_matching(str(tmp_path), str(tmp_path / "scene.png"), "the dog on the right", SEG, str(vocab), 76,
                                         str(snap), 15, ops=ops)
    finally:
        PM.rmi.RMIModel = orig
    mask = m.predict_mask(sk, "the dog on the right", {w.strip(): i for i, w in enumerate(open(vocab))}, T=15)
    assert got == PM.get_pred_instance_mask(SEG, mask)[4] and isinstance(got, list)
