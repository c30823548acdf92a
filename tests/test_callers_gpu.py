"""The callers on either side of the hot path, on the CUDA operator set (a17's `val` / `test` modes):
`--mode val` over TFRecord files, `--mode test` over the captions / images directory contract, and the whole-pipeline caller
`build_instance_colorization` on a synthetic scene -- each against the CPU oracle's generator on the same prepared inputs
(<= 1 grey level; the inference mode is fp32 storage + bf16x3 tensor-core products), snapshots restored from TF-format bundles
written by this package.  The CPU twins of these tests (tests/test_tfrecord_cpu.py, tests/test_pipeline_cpu.py) run the same
host code on the plain-torch operator set; fixtures are shared."""
import json
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

SIZE = 16


def _model(H, W):
    from sketchyscenecolorization_b200.cuda_ops import CudaOps
    from sketchyscenecolorization_b200.trainer import FgColorModel
    m = FgColorModel(CudaOps("cuda:0", torch.float32), "cuda:0", size=SIZE, H=H, W=W, with_discriminator=False)
    m.initialize(seed=1, perturb_tables=0.1)
    return m


def test_validation_mode_on_device(tmp_path):
    """`--mode val` (main_procedure.py:245-358): TFRecord files -> device queue -> generator -> three PNGs per sample."""
    import cv2
    from oracle import fgcolor_oracle as O
    from sketchyscenecolorization_b200 import main_procedure
    from sketchyscenecolorization_b200.config import Config
    from test_tfrecord_cpu import _write_split
    base = str(tmp_path)
    _write_split(base, "val", 2, files=("bus",))
    Config.set_from_dict(dict(dataset_type="val", batch_size=2, ckpt_dir=os.path.join(base, "snapshot"),
                              results_dir=os.path.join(base, "validation_results"), data_format="NCHW", distance_map=0, small_img=1,
                              LSTM_hybrid=1, block_type="MRU", vocab_size=58))
    model = _model(64, 64)
    noise = torch.zeros(2, 256, device="cuda")
    assert main_procedure.validation(model=model, data_base_dir=base, noise=noise) == 1
    out = os.path.join(base, "validation_results", "with_text")
    pics = {}
    for i in range(2):
        for kind in ("output", "target", "input"):
            p = os.path.join(out, "bus_img%03d_%s.png" % (i, kind))
            assert os.path.exists(p)
            pics[i, kind] = cv2.imread(p)
            assert pics[i, kind].shape == (64, 64, 3)
    assert pics[0, "input"].max() >= 254 and pics[0, "input"].min() < 140
    # the written picture is the oracle generator's on the batch the queue delivered (conditional BN sees both samples)
    from sketchyscenecolorization_b200.tfrecord_input import PairedEvalInput
    (b,) = list(PairedEvalInput("val", 2, model.ops, data_base_dir=base, small=True))
    gp = {k: v.detach().cpu().double() for k, v in model.gstore.state_dict().items()}
    with torch.no_grad():
        ref = O.generator_forward(gp, b["sketch"].cpu().double(), b["text"].cpu().long(), b["cls"].cpu().long(),
                                  torch.zeros(2, 256, dtype=torch.float64), SIZE)
    want = (((ref.permute(0, 2, 3, 1).numpy() + 1) / 2) * 255)[:, :, :, ::-1].astype(np.uint8)
    for i in range(2):
        assert np.abs(pics[i, "output"].astype(int) - want[i].astype(int)).max() <= 1


def test_test_mode_on_device_from_a_snapshot(tmp_path):
    """`--mode test` (main_procedure.py:361-492) with the generator restored from a TF-format snapshot written by checkpoint.py."""
    import cv2
    from oracle import fgcolor_oracle as O
    from sketchyscenecolorization_b200 import checkpoint, main_procedure
    from sketchyscenecolorization_b200.config import Config
    from sketchyscenecolorization_b200.pipeline_fg import thicken_drawings
    from sketchyscenecolorization_b200.text_processing import default_vocab_dict, preprocess_sentence
    base = tmp_path / "data"
    sketches = {}
    for cate, key, text in (("bus", "228_1.png", "A yellow bus with blue window"), ("road", "7_3.png", "the road is gray")):
        os.makedirs(base / "captions" / cate)
        os.makedirs(base / "images" / cate / "sketch")
        sk = np.full((64, 64, 3), 255, np.uint8)
        sk[20:22, 8:56] = 0
        sk[40:42, 8:56] = 0
        cv2.imwrite(str(base / "images" / cate / "sketch" / key), sk)
        json.dump([dict(key=key, color_text=text)], open(base / "captions" / cate / "test.json", "w"))
        sketches[cate] = (sk, key, text)
    res, snap = str(tmp_path / "test_results"), str(tmp_path / "snapshot")
    src = _model(64, 64)
    checkpoint.save(src, snap, 7, counter=8)
    model = _model(64, 64)
    model.gstore.flat.zero_()
    assert checkpoint.restore(model, checkpoint.latest_checkpoint(snap)) == 8
    assert torch.equal(model.gstore.flat, src.gstore.flat)
    Config.set_from_dict(dict(dataset_type="test", batch_size=1, ckpt_dir=snap, results_dir=res, data_format="NCHW", distance_map=0,
                              small_img=1, LSTM_hybrid=1, block_type="MRU", vocab_size=58))
    noise = torch.zeros(1, 256, device="cuda")
    assert main_procedure.test(model=model, noise=noise, data_base_dir=str(base)) == 2
    gp = {k: v.detach().cpu().double() for k, v in src.gstore.state_dict().items()}
    for ci, cate in enumerate(("bus", "road")):
        sk, key, text = sketches[cate]
        out = cv2.imread(os.path.join(res, "%s_%s_output.png" % (cate, key[:-4])))
        inp = cv2.imread(os.path.join(res, "%s_%s_input.png" % (cate, key[:-4])))
        want_in = thicken_drawings(sk.astype(np.float32)) if cate == "road" else sk
        assert np.array_equal(inp, want_in)
        x = torch.from_numpy(want_in.astype(np.float64) / 255 * 2 - 1).permute(2, 0, 1)[None]
        ids = torch.tensor([preprocess_sentence(text, default_vocab_dict(), 15)])
        with torch.no_grad():
            ref = O.generator_forward(gp, x, ids, torch.tensor([ci]), torch.zeros(1, 256, dtype=torch.float64), SIZE)
        want = (((ref[0].permute(1, 2, 0).numpy() + 1) / 2) * 255)[:, :, ::-1].astype(np.uint8)
        assert np.abs(out.astype(int) - want.astype(int)).max() <= 1


def test_instance_colorization_pipeline_on_device(tmp_path):
    """Pipeline_utils/fg_color_utils.build_instance_colorization (:188-363) with the resident generator on the B200."""
    from oracle import fgcolor_oracle as O
    from sketchyscenecolorization_b200 import pipeline_fg as P
    from sketchyscenecolorization_b200.text_processing import default_vocab_dict, preprocess_sentence
    from test_pipeline_cpu import _scene
    tmp = str(tmp_path)
    sk, inner, boxes, masks, classes = _scene(tmp)
    model = _model(192, 192)
    text = "the bus on the left is yellow with blue windows"
    args = dict(data_base_dir=tmp, image_id=7, input_text=text, sketch_path=os.path.join(tmp, "scene.png"),
                inner_masks_mat_path=os.path.join(tmp, "inner.mat"), segm_data_npz_path=os.path.join(tmp, "seg.npz"),
                results_base_dir=tmp, fgcolor_vocab_size=58, fgcolor_max_len=15, fgcolor_vocab_path=os.path.join(tmp, "vocab.txt"),
                fgcolor_snapshot_root=os.path.join(tmp, "snapshot"))
    out = P.build_instance_colorization(inst_indices=[0, 1], new_result_image_name="r1.png", last_result_image_name="",
                                        model=model, noise_seed=11, **args)
    gp = {k: v.detach().cpu().double() for k, v in model.gstore.state_dict().items()}
    ids = torch.tensor([preprocess_sentence(P.segment_user_input_text(text), default_vocab_dict(), 15)])
    gen = torch.Generator().manual_seed(11)
    pm = P.expand_small_segmentation_mask(masks, np.array(boxes))
    want = sk.copy()
    for i in (0, 1):
        sketch = torch.from_numpy(P.prepare_instance_sketch(pm[i], boxes[i], classes[i])).double()
        noise = torch.randn(1, 256, generator=gen).double()
        with torch.no_grad():
            img = O.generator_forward(gp, sketch, ids, torch.tensor([P.SKE_TO_FG_CLASS[classes[i]]]), noise, SIZE)
        col = P.instance_result_postprocessing(img.numpy(), boxes[i], 'NCHW', classes[i])
        y1, x1, y2, x2 = boxes[i]
        sel = inner[y1:y2, x1:x2] == i + 1
        want[y1:y2, x1:x2][sel] = col[sel]
    moved = sk.copy()
    moved[1:, 1:] = sk[:-1, :-1]
    want[moved[:, :, 0] == 0] = moved[moved[:, :, 0] == 0]
    diff = np.abs(out.astype(int) - want.astype(int))
    # the un-pad / resize back into the instance box interpolates between generated pixels: a 1e-4 difference of the generator
    # moves a handful of interpolated pixels by one more grey level at most
    assert diff.max() <= 2 and (diff > 1).mean() < 1e-4, "pipeline picture differs by %d grey levels" % diff.max()
    assert (out[inner == 1] != 255).any() and (out[inner == 2] != 255).any()
