"""Whole-pipeline background caller (pipeline_bg.py), editing records (customization_util.py) and the pipeline entry point:
helpers pinned to vectors produced by the reference's own functions (tests/golden/pipeline_bg.json, made by
tests/golden/make_pipeline_bg_golden.py), then one BG and one FG instruction end to end on a synthetic scene with small models on
the plain-torch operator set."""
import json
import os
import sys

import numpy as np
import pytest
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
from pipeline_bg_cases import gradient_case                                         # noqa: E402
from sketchyscenecolorization_b200 import customization_util as CU                  # noqa: E402
from sketchyscenecolorization_b200 import pipeline_bg as PB                         # noqa: E402

GOLD = json.load(open(os.path.join(HERE, "golden", "pipeline_bg.json")))


def test_text_helpers_match_reference():
    for s, want in GOLD["get_text_type"].items():
        assert PB.get_text_type(s) == want, s
    for s, want in GOLD["judge_colorize_type"].items():
        assert CU.judge_colorize_type(s) == want, s
    for new, prev, want in GOLD["combine_bg_input_text"]:
        if want.startswith("EXC:"):
            with pytest.raises(Exception) as e:
                PB.combine_bg_input_text(new, prev)
            if want != "EXC:AssertionError" and e.type is not AssertionError:
                assert str(e.value.args[0]) == want[4:], (new, prev)
        else:
            assert PB.combine_bg_input_text(new, prev) == want, (new, prev)


def test_hsv_maps_against_colorsys():
    import colorsys
    rng = np.random.default_rng(0)
    rgb = rng.random((50, 3))
    rgb[:5] = rgb[:5, :1]                      # greys: hue 0, saturation 0
    rgb[5] = 0.0
    hsv = PB.rgb2hsv(rgb)
    want = np.array([colorsys.rgb_to_hsv(*p) for p in rgb])
    assert np.abs(hsv - want).max() < 1e-12
    back = PB.hsv2rgb(hsv)
    assert np.abs(back - rgb).max() < 1e-12
    assert np.abs(PB.hsv2rgb(want) - np.array([colorsys.hsv_to_rgb(*p) for p in want])).max() < 1e-12


@pytest.mark.parametrize("name", ["blue_sky", "two_tone", "low_horizon"])
def test_color_gradient_matches_reference(name):
    img, mask = gradient_case(name)
    got = PB.add_color_gradient(img, mask)
    want = np.array(GOLD["add_color_gradient"][name]["full"], dtype=np.uint8)
    diff = np.abs(got.astype(np.int32) - want.astype(np.int32))
    assert diff.max() <= 1 and (diff == 0).mean() > 0.999          # float64 HSV round trip: at most a knife-edge truncation
    assert np.array_equal(got[mask != 0], img[mask != 0])         # the foreground is untouched
    assert not np.array_equal(got[0], img[0])                      # the top row was faded


def test_records_roundtrip(tmp_path):
    base = str(tmp_path)
    new, last, bg, summary = CU.fetch_records(7, base)
    assert (new, last, bg, summary) == ("7_1.png", "", "", [])
    os.makedirs(os.path.join(base, "results", "7"))
    open(os.path.join(base, "results", "7", "7_1.png"), "wb").write(b"x")
    CU.update_records(7, "the sky is blue and the ground is green", base, "BG", new, "the sky is blue and the ground is green", summary)
    new2, last2, bg2, summary2 = CU.fetch_records(7, base)
    assert (new2, last2, bg2) == ("7_2.png", "7_1.png", "the sky is blue and the ground is green") and len(summary2) == 1
    open(os.path.join(base, "results", "7", "7_2.png"), "wb").write(b"y")
    CU.update_records(7, "the bus is red", base, "FG", new2, bg2, summary2)
    rec = json.load(open(os.path.join(base, "update_records", "7_records.json")))
    assert [r["colorization_type"] for r in rec] == ["BG", "FG"] and list(rec[0]) == ["colorization_type", "result_name", "input_text", "proc_bg_text"]
    CU.withdraw_records(7, base)
    assert not os.path.exists(os.path.join(base, "results", "7", "7_2.png"))
    assert len(json.load(open(os.path.join(base, "update_records", "7_records.json")))) == 1
    CU.withdraw_records(7, base)
    assert not os.path.exists(os.path.join(base, "update_records", "7_records.json"))
    with pytest.raises(Exception):
        CU.withdraw_records(7, base)


def _scene(base, image_id, S):
    """A synthetic scene: sketch with two boxes and a horizon line, inner masks (instance 1 = bus box, instance 2 = grass box),
    Mask-RCNN style segmentation data."""
    import scipy.io
    from PIL import Image
    for d in ("sketches", "seg_data", "inner_masks"):
        os.makedirs(os.path.join(base, d), exist_ok=True)
    sk = np.full((S, S, 3), 255, np.uint8)
    sk[S // 2, :] = 0
    sk[20:50, 16] = sk[20:50, 44] = sk[20, 16:45] = sk[50, 16:45] = 0
    Image.fromarray(sk, "RGB").save(os.path.join(base, "sketches", "%d.png" % image_id))
    inner = np.zeros((S, S), np.uint8)
    inner[21:50, 17:44] = 1
    inner[S - 14:S - 4, 8:40] = 2
    scipy.io.savemat(os.path.join(base, "inner_masks", "%d.mat" % image_id), {"inner_masks": inner})
    np.savez(os.path.join(base, "seg_data", "%d_datas.npz" % image_id), pred_class_ids=np.array([12, 27]),
             pred_boxes=np.array([[20, 16, 51, 45], [S - 15, 7, S - 3, 41]]))
    return sk, inner


def test_background_instruction_end_to_end(tmp_path):
    """BG instruction -> one generator call -> foreground / strokes / sky gradient laid over it, records updated; then a
    one-sided follow-up is completed from the recorded caption."""
    from PIL import Image
    import sketchyscene_colorization_main as M
    from sketchyscenecolorization_b200.bg import BgColorModel
    from torch_ops import TorchOps
    S = 64
    data, res = str(tmp_path / "examples"), str(tmp_path / "outputs")
    sk, inner = _scene(data, 3, S)
    m = BgColorModel(TorchOps(torch.float32), "cpu", ngf=4, vocab_size=18)
    m.initialize(seed=1)
    args = (data, res, "", 76, "", 15, "", 58, "", 15, "no_such_vocab.txt", 18, "", 8)
    import sketchyscenecolorization_b200.pipeline_bg as PBm
    orig = PBm.build_background_colorization
    try:
        PBm_build = lambda *a, **k: orig(*a, image_size=S, **k)       # noqa: E731  (the reference is fixed at 768)
        M.build_background_colorization = PBm_build
        kind, name = M.colorization_main(3, "the sky is blue and the ground is green", *args, bg_model=m)
        assert (kind, name) == ("BG", "3_1.png")
        out = np.array(Image.open(os.path.join(res, "results", "3", "3_1.png")).convert("RGB"))
        assert out.shape == (S, S, 3)
        assert np.array_equal(out[30, 30], sk[30, 30])                 # inside the instance: the previous (sketch) pixels
        assert tuple(out[S // 2 + 1, 60]) == (0, 0, 0)                 # the horizon stroke, shifted by one pixel, drawn on top
        assert os.path.exists(os.path.join(res, "results", "3", "3_fg.png"))
        kind, name = M.colorization_main(3, "the sky is purple", *args, bg_model=m)
        rec = json.load(open(os.path.join(res, "update_records", "3_records.json")))
        assert name == "3_2.png" and rec[1]["proc_bg_text"] == "the sky is purple and the ground is green"
        # an FG instruction goes to the instance-matching model; without a snapshot under match_snapshot_root that is an error
        # that says so (its own tests: tests/test_pipeline_match_cpu.py)
        with pytest.raises(FileNotFoundError):
            M.colorization_main(3, "the bus is red", *args, bg_model=m, ops=m.ops)
    finally:
        M.build_background_colorization = orig
