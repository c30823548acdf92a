"""Whole-pipeline caller (pipeline_fg.build_instance_colorization): helper semantics against vectors produced by the
reference's own functions (tests/golden/pipeline_fg.json, made by tests/golden/make_pipeline_golden.py), and an end-to-end
run on a synthetic scene whose coloured pixels are checked against the CPU oracle's generator."""
import json
import os
import sys

import numpy as np
import pytest
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
from pipeline_cases import SENTENCES, road_sketch          # noqa: E402

from oracle import fgcolor_oracle as O                     # noqa: E402
from sketchyscenecolorization_b200 import checkpoint, pipeline_fg as P      # noqa: E402
from sketchyscenecolorization_b200.text_processing import DEFAULT_VOCAB, preprocess_sentence, default_vocab_dict   # noqa: E402
from sketchyscenecolorization_b200.trainer import FgColorModel                # noqa: E402
from torch_ops import TorchOps                              # noqa: E402

GOLD = json.load(open(os.path.join(HERE, "golden", "pipeline_fg.json")))


def test_caption_segmentation_matches_reference():
    assert set(GOLD["segment_user_input_text"]) == set(SENTENCES)
    for s, want in GOLD["segment_user_input_text"].items():
        assert P.segment_user_input_text(s) == want, s


def test_road_check_matches_reference():
    for name, want in GOLD["is_road_not_single_line"].items():
        assert P.is_road_not_single_line(road_sketch(name)) is want, name


def test_small_helpers():
    # thicken: a single black pixel grows into the 2x2 square up/left of it (skimage dilation with square(2))
    img = np.full((6, 6, 3), 255, dtype=np.uint8)
    img[3, 3] = 0
    t = P.thicken_drawings(img)
    assert t.shape == (6, 6, 3) and (t[:, :, 0] == 0).sum() == 4 and (t[2:4, 2:4, 0] == 0).all()
    # masks: boxes are inclusive on both ends
    m = P.expand_small_segmentation_mask([np.ones((3, 2), np.uint8)], np.array([[5, 7, 7, 8]]))
    assert m.shape == (1, 768, 768) and m.sum() == 6 and m[0, 5:8, 7:9].all()
    # reverse resize: padding cut on the correct axis, margin removed
    inst = np.zeros((192, 192, 3), np.uint8)
    inst[:, 48:144] = 200                                   # tall box (h > w): content in the middle columns
    back = P.reverse_resize_image(inst, 100, 40, margin_size=10)
    assert back.shape == (100, 40, 3) and back[50, 20, 0] == 200
    assert P.instance_result_postprocessing(np.zeros((1, 3, 192, 192), np.float32), [10, 20, 110, 60], 'NCHW', 12).shape == (100, 40, 3)
    assert (P.instance_result_postprocessing(np.full((1, 3, 192, 192), 0.999, np.float32), [0, 0, 50, 50], 'NCHW', P.ROAD_LABEL) == 254).all()


def _scene(tmp):
    """768x768 sketch with a 'bus' (class 12) and a two-edge 'road' (class 36) + the files the caller reads."""
    import scipy.io
    from PIL import Image
    S = 768
    sk = np.full((S, S, 3), 255, np.uint8)
    inner = np.zeros((S, S), np.int32)
    boxes, masks, classes = [], [], []
    # instance 0: bus outline
    y1, x1, y2, x2 = 100, 150, 260, 450
    m = np.zeros((y2 - y1 + 1, x2 - x1 + 1), np.uint8)
    m[0:3, :] = m[-3:, :] = 1
    m[:, 0:3] = m[:, -3:] = 1
    m[60:63, :] = 1
    boxes.append([y1, x1, y2, x2]); masks.append(m); classes.append(12)
    inner[y1 + 3:y2 - 2, x1 + 3:x2 - 2] = 1
    # instance 1: road = two long horizontal edges
    y1, x1, y2, x2 = 500, 40, 640, 740
    m = np.zeros((y2 - y1 + 1, x2 - x1 + 1), np.uint8)
    m[10:14, :] = 1
    m[120:124, :] = 1
    boxes.append([y1, x1, y2, x2]); masks.append(m); classes.append(36)
    inner[y1 + 14:y1 + 120, x1:x2] = 2
    # instance 2: a class the fg model does not know (class 1) -- must be refused
    boxes.append([10, 10, 40, 40]); masks.append(np.ones((31, 31), np.uint8)); classes.append(1)
    for b, m in zip(boxes, masks):
        sk[b[0]:b[2] + 1, b[1]:b[3] + 1][m == 1] = 0
    Image.fromarray(sk).save(os.path.join(tmp, "scene.png"))
    scipy.io.savemat(os.path.join(tmp, "inner.mat"), {"inner_masks": inner})
    np.savez(os.path.join(tmp, "seg.npz"), pred_class_ids=np.array(classes), pred_boxes=np.array(boxes),
             pred_masks=np.array(masks, dtype=object))
    names = np.empty((46, 1), dtype=object)
    for i in range(46):
        names[i, 0] = np.array(["cls%d" % i])
    scipy.io.savemat(os.path.join(tmp, "colorMapC46.mat"), {"colorMap": names})
    with open(os.path.join(tmp, "vocab.txt"), "w") as f:
        f.write("\n".join(DEFAULT_VOCAB) + "\n")
    return sk, inner, boxes, masks, classes


def test_build_instance_colorization_end_to_end(tmp_path):
    tmp = str(tmp_path)
    sk, inner, boxes, masks, classes = _scene(tmp)
    size = 16                                                # a narrow generator keeps the CPU run short
    ops = TorchOps(torch.float64)
    model = FgColorModel(ops, "cpu", size=size, H=192, W=192, param_dtype=torch.float64, with_discriminator=False)
    model.initialize(seed=5, perturb_tables=0.1)
    text = "the bus on the left is yellow with blue windows"
    args = dict(data_base_dir=tmp, image_id=7, input_text=text, sketch_path=os.path.join(tmp, "scene.png"),
                inner_masks_mat_path=os.path.join(tmp, "inner.mat"), segm_data_npz_path=os.path.join(tmp, "seg.npz"),
                results_base_dir=tmp, fgcolor_vocab_size=58, fgcolor_max_len=15, fgcolor_vocab_path=os.path.join(tmp, "vocab.txt"),
                fgcolor_snapshot_root=os.path.join(tmp, "snapshot"))
    out = P.build_instance_colorization(inst_indices=[0, 1], new_result_image_name="r1.png", last_result_image_name="",
                                        model=model, noise_seed=11, **args)
    from PIL import Image
    saved = np.array(Image.open(os.path.join(tmp, "results", "7", "r1.png")).convert("RGB"))
    assert saved.shape == (768, 768, 3) and np.array_equal(saved, out)

    # expected picture: oracle generator on the same prepared sketches, pasted by the reference's rules
    gp = {k: v.detach().clone() for k, v in model.gstore.state_dict().items()}
    ids = torch.tensor([preprocess_sentence(P.segment_user_input_text(text), default_vocab_dict(), 15)])
    gen = torch.Generator().manual_seed(11)
    pm = P.expand_small_segmentation_mask(masks, np.array(boxes))
    want = sk.copy()
    for i in (0, 1):
        sketch = torch.from_numpy(P.prepare_instance_sketch(pm[i], boxes[i], classes[i])).double()
        noise = torch.randn(1, 256, generator=gen).double()
        with torch.no_grad():
            img = O.generator_forward(gp, sketch, ids, torch.tensor([P.SKE_TO_FG_CLASS[classes[i]]]), noise, size)
        col = P.instance_result_postprocessing(img.numpy(), boxes[i], 'NCHW', classes[i])
        y1, x1, y2, x2 = boxes[i]
        sel = inner[y1:y2, x1:x2] == i + 1
        want[y1:y2, x1:x2][sel] = col[sel]
    moved = sk.copy()
    moved[1:, 1:] = sk[:-1, :-1]
    want[moved[:, :, 0] == 0] = moved[moved[:, :, 0] == 0]
    diff = np.abs(out.astype(int) - want.astype(int))
    assert diff.max() <= 1, "pipeline picture differs from the oracle-based expectation by %d grey levels" % diff.max()
    assert (out[inner == 1] != 255).any() and (out[inner == 2] != 255).any()            # both instances were coloured
    assert np.array_equal(out[inner == 0], want[inner == 0])                               # nothing else was touched

    # second call paints on top of the first result; an unknown class is refused before anything is written
    out2 = P.build_instance_colorization(inst_indices=[1], new_result_image_name="r2.png", last_result_image_name="r1.png",
                                         model=model, noise_seed=12, **args)
    assert np.array_equal(out2[inner == 1], out[inner == 1]) and os.path.exists(os.path.join(tmp, "results", "7", "r2.png"))
    with pytest.raises(Exception, match="Wrong matching instance"):
        P.build_instance_colorization(inst_indices=[2], new_result_image_name="r3.png", last_result_image_name="r2.png",
                                      model=model, **args)
    assert not os.path.exists(os.path.join(tmp, "results", "7", "r3.png"))

    # the snapshot route: restore a generator from a TF-format snapshot directory (narrow model through a patched default)
    full = FgColorModel(ops, "cpu", size=size, H=192, W=192, param_dtype=torch.float64, with_discriminator=False)
    full.initialize(seed=5, perturb_tables=0.1)
    checkpoint.save(full, os.path.join(tmp, "snapshot"), 3, counter=4)
    import sketchyscenecolorization_b200.trainer as T
    orig = T.FgColorModel
    try:
        T.FgColorModel = lambda *a, **k: orig(*a, size=size, param_dtype=torch.float64, **k)
        out3 = P.build_instance_colorization(inst_indices=[0, 1], new_result_image_name="r4.png", last_result_image_name="",
                                             ops=ops, noise_seed=11, **args)
    finally:
        T.FgColorModel = orig
    assert np.abs(out3.astype(int) - out.astype(int)).max() <= 1      # fp32 snapshot of the same weights
