"""Whole-pipeline consumer of the background generator: `build_background_colorization`.

Same arguments, files and results as the reference's Pipeline_utils/bg_utils.py (:169-325): the previous result (or the bare
sketch) is cut down to its foreground through the inner mask, the instruction is merged with the last background caption
(`combine_bg_input_text`, :59-93), ONE 768 x 768 generator call paints the background, then the foreground, the sketch strokes
(shifted by one pixel, grass excepted) and a sky gradient (`add_color_gradient`, :96-166) are laid over it.  The generator is
bg.BgColorModel, restored from the newest TensorFlow-format snapshot under `bg_snapshot_root` or handed in resident (`model=`).
The text helpers and the gradient are pinned to vectors produced by the reference's own functions
(tests/golden/pipeline_bg.json, make_pipeline_bg_golden.py); skimage is not needed: the HSV maps are restated here.
"""
from __future__ import annotations

import os
import re

import numpy as np

from .text_processing import bg_vocab_dict, load_vocab_dict_from_file, preprocess_sentence

GRASS_LABEL = 27          # bg_utils.py:17
IMAGE_SIZE = 768
INPUT_TEXT_TYPES = ['None', 'ground', 'sky', 'both']
ALL_COLOR = ['blue', 'green', 'cyan', 'red', 'orange', 'yellow', 'brown', 'purple', 'pink', 'black', 'gray']
_SPLIT = re.compile(r'(\W+)')


def _words(text):
    return [w.lower() for w in _SPLIT.split(text.strip()) if len(w.strip()) > 0]


def get_text_type(text):
    """:24-37 -- which of sky / ground an instruction talks about."""
    words = _words(text)
    sky = 1 if 'sky' in words else 0
    ground = 1 if ('ground' in words or 'floor' in words or 'land' in words) else 0
    return INPUT_TEXT_TYPES[2 * sky + ground]


def check_duplicated_color(text):
    """:40-56 -- the first two colour words must differ."""
    colors = [w for w in _words(text) if w in ALL_COLOR][:2]
    sky_color = colors[0] if colors else ''
    ground_color = colors[1] if len(colors) > 1 else ''
    if sky_color == ground_color:
        raise Exception('It is not recommended to use the same sky and ground color.')


def combine_bg_input_text(new_text, previous_text):
    """:59-93 -- complete a one-sided instruction with the other half of the previous caption."""
    input_text_type, previous_text_type = get_text_type(new_text), get_text_type(previous_text)
    assert input_text_type != 'None'
    if input_text_type == 'both':
        rst_text = new_text
    elif input_text_type == 'sky':
        if previous_text_type in ('None', 'sky'):
            raise Exception('No ground infomation provided and found in records.')
        if previous_text_type == 'ground':
            rst_text = new_text + ' and ' + previous_text
        else:
            rst_text = new_text + ' ' + previous_text[previous_text.index('and'):]
    else:
        if previous_text_type in ('None', 'ground'):
            raise Exception('No sky infomation provided and found in records.')
        if previous_text_type == 'sky':
            rst_text = previous_text + ' and ' + new_text
        else:
            rst_text = previous_text[:previous_text.index('and')] + 'and ' + new_text
    assert rst_text != ''
    check_duplicated_color(rst_text)
    return rst_text


def rgb2hsv(rgb):
    """skimage.color.rgb2hsv: float [...,3] in [0,1] -> h, s, v in [0,1]."""
    rgb = np.asarray(rgb, dtype=np.float64)
    v = rgb.max(-1)
    delta = rgb.max(-1) - rgb.min(-1)
    with np.errstate(invalid='ignore', divide='ignore'):
        s = np.where(delta == 0, 0.0, delta / v)
        r, g, b = rgb[..., 0], rgb[..., 1], rgb[..., 2]
        h = np.where(r == v, (g - b) / delta, np.where(g == v, 2.0 + (b - r) / delta, 4.0 + (r - g) / delta))
    h = np.where(delta == 0, 0.0, (h / 6.0) % 1.0)
    return np.stack([h, np.nan_to_num(s), v], -1)


def hsv2rgb(hsv):
    """skimage.color.hsv2rgb."""
    hsv = np.asarray(hsv, dtype=np.float64)
    h, s, v = hsv[..., 0], hsv[..., 1], hsv[..., 2]
    hi = np.floor(h * 6)
    f = h * 6 - hi
    p, q, t = v * (1 - s), v * (1 - f * s), v * (1 - (1 - f) * s)
    hi = (hi.astype(np.int64) % 6)[..., None]
    cands = np.stack([np.stack(c, -1) for c in ((v, t, p), (q, v, p), (p, v, t), (p, q, v), (t, p, v), (v, p, q))], 0)
    return np.take_along_axis(cands, np.broadcast_to(hi[None], (1,) + hi.shape[:-1] + (3,)), 0)[0]


def add_color_gradient(color_image, inner_mask, search_height=2, search_from=5):
    """:96-166 -- fade the sky towards the top: the most frequent background colour of rows `search_from` .. is the sky colour,
    its lowest row in the upper half the horizon; from 3/4 of that height up to row 0 saturation falls to a third and value
    rises by half (in HSV); the foreground is laid back on top."""
    img_h = color_image.shape[0]
    img_bg = np.full(color_image.shape, 255, dtype=np.uint8)
    img_bg[inner_mask == 0] = color_image[inner_mask == 0]
    counts = {}                                              # insertion ordered: ties go to the colour seen first (np.argmax)
    for i in range(search_height):
        row, free = img_bg[i + search_from], inner_mask[i + search_from] == 0
        for rgb in map(tuple, row[free].tolist()):
            counts[rgb] = counts.get(rgb, 0) + 1
    sky_color = max(counts, key=lambda c: counts[c])          # first maximum in insertion order
    sky = np.array(sky_color, dtype=np.uint8)
    sky_bottom = -1
    for i in range(int(img_h / 2), -1, -1):
        if (img_bg[i] == sky).all(axis=1).any():
            sky_bottom = i
            break
    assert sky_bottom != -1
    start_height = int(sky_bottom / 4 * 3)
    sky_hsv = rgb2hsv(np.array(sky_color, dtype=np.float32)[None, None] / 255.)[0][0]
    grad = rgb2hsv(img_bg / 255.)
    end_s, end_v = sky_hsv[1] / 3., min(1., sky_hsv[2] * 1.5)
    for i in range(start_height, -1, -1):                    # start_height == 0 divides by zero, as in the reference
        grad[i, :, 1] = (start_height - i) / start_height * end_s + i / start_height * sky_hsv[1]
        grad[i, :, 2] = (start_height - i) / start_height * end_v + i / start_height * sky_hsv[2]
    out = np.array(hsv2rgb(grad) * 255., dtype=np.uint8)
    out[inner_mask != 0] = color_image[inner_mask != 0]
    return out


def _load_generator(snapshot_root, vocab_size, ops=None, device=None):
    import torch
    from . import checkpoint, tf_bundle
    from .bg import BgColorModel
    if ops is None:
        from .cuda_ops import CudaOps
        device = device or "cuda:%d" % int(os.environ.get("LOCAL_RANK", "0"))
        ops = CudaOps(device, torch.float32)
    model = BgColorModel(ops, device or getattr(ops, "device", "cpu"), ngf=64, vocab_size=vocab_size)
    prefix = checkpoint.latest_checkpoint(snapshot_root)
    print("loading model from checkpoint", prefix)
    if prefix is None:
        raise RuntimeError("no snapshot in %s" % snapshot_root)
    model.gstore.load_state_dict(tf_bundle.read_bundle(prefix), strict=True)
    return model


def build_background_colorization(image_id, input_text, sketch_path, inner_masks_mat_path, segm_data_npz_path, results_base_dir,
                                  bg_vocab_size, bg_max_len, bg_vocab_path, bg_snapshot_root, new_result_image_name,
                                  last_result_image_name, last_bg_text, color_gradient=True, *, model=None, ops=None,
                                  image_size=IMAGE_SIZE):
    """Background colorization of one scene and its result file; returns the processed caption for the records (:169-325).
    `image_size` (keyword only) exists for tests; the reference is fixed at 768."""
    import scipy.io
    from PIL import Image
    vocab_dict = load_vocab_dict_from_file(bg_vocab_path) if os.path.exists(bg_vocab_path) else bg_vocab_dict()
    sketch = Image.open(sketch_path).convert("RGB").resize((image_size, image_size), resample=Image.NEAREST)
    sketch = np.array(sketch, dtype=np.uint8)
    results_dir = os.path.join(results_base_dir, 'results', str(image_id))
    os.makedirs(results_dir, exist_ok=True)
    if last_result_image_name == '':                        # empty records: start from the bare sketch
        assert last_bg_text == ""
        last_bg_text = "the sky is blue and the ground is green"
        previous = sketch.copy()
    else:
        previous = np.array(Image.open(os.path.join(results_dir, last_result_image_name)).convert('RGB'), dtype=np.uint8)
    pred_class_ids = np.load(segm_data_npz_path)['pred_class_ids']
    grass = [i for i in range(len(pred_class_ids)) if pred_class_ids[i] == GRASS_LABEL]
    inner_mask = scipy.io.loadmat(inner_masks_mat_path)['inner_masks']
    fg_image = np.full(previous.shape, 255, dtype=np.uint8)
    fg_image[inner_mask != 0] = previous[inner_mask != 0]
    fg_image_temp = fg_image.copy()
    proc_input_text = combine_bg_input_text(input_text, last_bg_text)
    print('proc_input_text:', proc_input_text)
    if model is None:
        model = _load_generator(bg_snapshot_root, bg_vocab_size, ops)
    print("parameter_count =", model.gstore.num_params())
    fg_data = fg_image[None]
    ids = np.array(preprocess_sentence(proc_input_text, vocab_dict, bg_max_len), dtype=np.int32)[None]
    background, _ = model.colorize_u8(fg_data, ids)
    assert inner_mask.shape[0] == fg_data.shape[1] and inner_mask.shape[1] == fg_data.shape[2]
    background[inner_mask != 0] = fg_image[inner_mask != 0]
    no_grass = np.zeros(inner_mask.shape, dtype=np.int32)   # 1 where a grass instance lies: its strokes are not redrawn
    for gi in grass:
        no_grass[inner_mask == gi + 1] = 1
    moved = sketch.copy()
    moved[1:image_size, 1:image_size] = sketch[0:image_size - 1, 0:image_size - 1]
    drawings = np.logical_and(moved[:, :, 0] == 0, no_grass != 1)
    background[drawings] = moved[drawings]
    fg_image_temp[drawings] = moved[drawings]
    Image.fromarray(fg_image_temp, 'RGB').save(os.path.join(results_dir, str(image_id) + '_fg.png'), 'PNG')
    if color_gradient:
        background = add_color_gradient(background, inner_mask)
        background[drawings] = moved[drawings]
    Image.fromarray(background, 'RGB').save(os.path.join(results_dir, new_result_image_name), 'PNG')
    return proc_input_text
