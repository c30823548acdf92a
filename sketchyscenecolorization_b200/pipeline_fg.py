"""Whole-pipeline caller of the fg-colorization path: colour the matched instances of a 768x768 scene sketch.

Reference: Pipeline_utils/fg_color_utils.py -- `build_instance_colorization` (:188-363) and its helpers
`segment_user_input_text` (:50-77), `is_road_not_single_line` (:80-134), `reverse_resize_image` (:137-163),
`instance_result_postprocessing` (:166-185); from the matching module it borrows `load_image2`,
`expand_small_segmentation_mask` (Instance_Matching/data_processing/sketch_data_processing.py:24-29,202-214) and
`search_for_self_category` / `search_for_color` (Instance_Matching/data_processing/text_processing.py:44-78); from the
fg module `thicken_drawings` (obj_lib/input_pipeline.py:242-256).

Same arguments, same files read and written.  What differs is the engine: the reference builds a fresh TF graph and
session per call and feeds the instances one `sess.run` at a time; here ONE resident generator (restored from the
latest snapshot, or handed in by the caller) runs the instances back to back.  The instances are NOT stacked into one
batch: the generator's conditional batch-norm uses batch statistics (models_collection.py:22-34), so a batch-k call
would change every pixel -- batch 1 per instance is part of the reference's result.  The generator noise
(`tf.random_normal`, models_collection.py:310) is drawn from a seeded generator so that runs are reproducible.
"""
from __future__ import annotations

import os
import re

import numpy as np

from .input_pipeline import resize_and_padding_mask_image
from .text_processing import load_vocab_dict_from_file, preprocess_sentence

# 46 scene-sketch classes -> the 25 fg-colorization classes (fg_color_utils.py:18-21)
SKE_TO_FG_CLASS = {7: 0, 9: 1, 12: 2, 13: 3, 14: 4, 15: 5, 16: 6, 17: 7, 18: 8, 19: 9, 22: 10, 23: 11, 27: 12, 28: 13,
                   29: 14, 30: 15, 32: 16, 34: 17, 35: 18, 36: 19, 37: 20, 39: 21, 41: 22, 43: 23, 44: 24}
ROAD_LABEL, GRASS_LABEL = 36, 27
INSTANCE_SIZE, IMAGE_SIZE = 192, 768

_SPLIT = re.compile(r'(\W+)')
_SIMPLE_COLORS = ('brown gray black red green blue yellow orange pink purple cyan white').split()
_CATEGORIES = ('bench bird bus butterfly car cat chair chicken cloud cow dog duck horse house grass moon person pig rabbit '
               'road sheep star sun tree truck').split()
_PLURALS = ('benches birds buses butterflies cars cats chairs chickens clouds cows dogs ducks horses houses grasses moons '
            'people pigs rabbits roads sheep stars suns trees trucks').split()


def _words(text, drop_dash=False):
    w = [t.lower() for t in _SPLIT.split(text.strip()) if len(t.strip()) > 0]
    return [t for t in w if t != '-'] if drop_dash else w


def _self_category(caption):
    """First category word of the caption (singular form) -- text_processing.search_for_self_category."""
    for w in _words(caption, drop_dash=True):
        if w in _CATEGORIES:
            return w
        if w in _PLURALS:
            return _CATEGORIES[_PLURALS.index(w)]
    return None


def _has_color(caption):
    return any(w in _SIMPLE_COLORS for w in _words(caption, drop_dash=True))


def _verb_leads(text, verb):
    """False when 'with' comes before the verb ('a man with blue pants has red shirt' must not be split at 'has')."""
    words = _words(text)
    return not ('with' in words and words.index('with') < words.index(verb.lower()))


def segment_user_input_text(user_text):
    """'the bus on the left is yellow with blue windows' -> 'the bus is yellow with blue windows': drop the locating phrase
    in front of the verb when the colours come after it (fg_color_utils.py:50-77; substring tests as in the reference)."""
    cate = _self_category(user_text)
    for verb in ('has', 'have', 'is', 'are'):
        if verb in user_text and _verb_leads(user_text, verb):
            split_idx = user_text.index(verb)
            break
    else:
        return user_text
    head, tail = user_text[:split_idx], user_text[split_idx:]
    if _has_color(head) or not _has_color(tail):
        return user_text
    return 'the ' + cate + ' ' + tail


def is_road_not_single_line(road_sketch_, parallel_width=25):
    """A road drawn as two edges crosses most scan lines an even number of times; a single stroke does not.  Columns first,
    then rows; True as soon as `parallel_width` lines had a positive even number of stroke runs (fg_color_utils.py:80-134)."""
    s = np.array(road_sketch_, dtype=np.uint8, copy=True)
    s[(s >= 235).all(axis=2)] = [255, 255, 255]
    s[(s != 255).all(axis=2)] = [0, 0, 0]
    p = s[:, :, 0].astype(np.int64)
    p[p == 0] = 1
    p[p == 255] = 0
    for axis in (0, 1):                                   # 0: walk down each column; 1: walk along each row
        nxt_is_one = np.zeros(p.shape, dtype=bool)
        if axis == 0:
            nxt_is_one[:-1, :] = p[1:, :] == 1
        else:
            nxt_is_one[:, :-1] = p[:, 1:] == 1
        ends = np.where(nxt_is_one, 0, p)                 # a stroke pixel survives only where its successor is not stroke
        crossings = ends.sum(axis=axis)
        valid = np.cumsum((crossings > 0) & (crossings % 2 == 0))
        if valid.size and valid[-1] >= parallel_width:
            return True
    return False


def thicken_drawings(image):
    """White-background drawing [H,W,3] -> strokes dilated by a 2x2 square (skimage.morphology.dilation(255 - img,
    square(2)) in the reference, obj_lib/input_pipeline.py:242-256; restated with numpy: skimage is not needed)."""
    img = 255 - np.array(image[:, :, 0], dtype=np.uint8)
    p = np.pad(img, ((0, 1), (0, 1)), mode='edge')
    d = np.maximum(np.maximum(p[:-1, :-1], p[:-1, 1:]), np.maximum(p[1:, :-1], p[1:, 1:]))
    return np.repeat((255 - d)[:, :, None], 3, axis=2)


def reverse_resize_image(cartoon_instance, box_h, box_w, h_w_ratio=1, margin_size=10):
    """Undo resize_and_padding_mask_image: cut the padding, scale to the (margin-extended) box, cut the margin
    (fg_color_utils.py:137-163; scipy.misc.imresize == PIL bilinear resize of the uint8 image)."""
    from PIL import Image
    ori = cartoon_instance.shape[0]
    bh, bw = box_h + 2 * margin_size, box_w + 2 * margin_size
    if bh * h_w_ratio > bw:
        pad = int(round(ori * (bh * h_w_ratio - bw) / (bh * h_w_ratio) / 2.))
        cut = cartoon_instance[:, pad: ori - pad]
    else:
        pad = int(round(ori * (bw - bh * h_w_ratio) / bw / 2.))
        cut = cartoon_instance[pad: ori - pad, :]
    rev = np.array(Image.fromarray(np.ascontiguousarray(cut)).resize((bw, bh), resample=Image.BILINEAR))
    return rev[margin_size: margin_size + box_h, margin_size: margin_size + box_w]


def instance_result_postprocessing(generated_img, bbox, data_format, class_id46):
    """[1,3,H,W] in [-1,1] -> uint8 [box_h, box_w, 3] (fg_color_utils.py:166-185; the uint8 cast truncates)."""
    if data_format == 'NCHW':
        assert generated_img.shape[1] == 3
        generated_img = np.transpose(generated_img, (0, 2, 3, 1))
    img = (((generated_img + 1) / 2.) * 255).astype(np.uint8)[0]
    margin = 0 if class_id46 == ROAD_LABEL else 10
    return reverse_resize_image(img, bbox[2] - bbox[0], bbox[3] - bbox[1], margin_size=margin)


def load_scene_sketch(path):
    """uint8 [768,768,3] (nearest-neighbour resize when the file has another size; sketch_data_processing.load_image2)."""
    from PIL import Image
    im = Image.open(path).convert("RGB")
    if im.width != IMAGE_SIZE or im.height != IMAGE_SIZE:
        im = im.resize((IMAGE_SIZE, IMAGE_SIZE), resample=Image.NEAREST)
    return np.array(im, dtype=np.uint8)


def expand_small_segmentation_mask(masks_small, boxes):
    """per-instance box-sized masks -> [N,768,768] (sketch_data_processing.expand_small_segmentation_mask; boxes inclusive)."""
    out = np.zeros((len(masks_small), IMAGE_SIZE, IMAGE_SIZE), dtype=np.uint8)
    for i, m in enumerate(masks_small):
        y1, x1, y2, x2 = boxes[i]
        out[i, y1: y2 + 1, x1: x2 + 1] = m
    return out


def prepare_instance_sketch(inst_mask768, bbox, class_id46):
    """Crop the instance mask, draw it black on white, resize + pad to 192 -> float32 [1,3,192,192] in [-1,1]
    (fg_color_utils.py:292-319).  Raises for a road drawn as a single line, as the reference does."""
    from PIL import Image
    y1, x1, y2, x2 = bbox
    m = inst_mask768[y1: y2, x1: x2]
    img = np.full((m.shape[0], m.shape[1], 3), 255, dtype=np.uint8)
    img[m == 1] = [0, 0, 0]
    pil = Image.fromarray(img, 'RGB')
    if pil.width != INSTANCE_SIZE or pil.height != INSTANCE_SIZE:
        sk = resize_and_padding_mask_image(pil, INSTANCE_SIZE, margin_size=0 if class_id46 == ROAD_LABEL else 10)
    else:
        sk = np.array(pil, dtype=np.uint8)
    assert sk.shape[0] == INSTANCE_SIZE and sk.shape[1] == INSTANCE_SIZE
    if class_id46 == ROAD_LABEL and not is_road_not_single_line(sk.copy()):
        raise Exception('Road is single line')
    if class_id46 == GRASS_LABEL:
        sk = thicken_drawings(sk)
    sk = sk.astype(np.float32) / 255. * 2. - 1
    return np.transpose(sk[None], [0, 3, 1, 2])


def _load_generator(snapshot_root, vocab_size, ops=None, device=None):
    import torch
    from . import checkpoint
    from .trainer import FgColorModel
    if ops is None:
        from .cuda_ops import CudaOps                     # product path: CUDA only (raises without a GPU / the built library)
        device = device or "cuda:%d" % int(os.environ.get("LOCAL_RANK", "0"))
        ops = CudaOps(device, torch.float32)              # fp32 storage + bf16x3 tensor-core mode: the parity mode
    model = FgColorModel(ops, device or getattr(ops, "device", "cpu"), H=INSTANCE_SIZE, W=INSTANCE_SIZE, vocab_size=vocab_size,
                         lstm_hybrid=True, with_discriminator=False)
    prefix = checkpoint.latest_checkpoint(snapshot_root)
    if prefix is None:
        raise FileNotFoundError("no snapshot under %s" % snapshot_root)
    print('Restore trained model:', prefix)
    checkpoint.restore(model, prefix, strict=True)
    return model


def build_instance_colorization(data_base_dir, image_id, input_text, inst_indices, sketch_path,
                                inner_masks_mat_path, segm_data_npz_path, results_base_dir,
                                fgcolor_vocab_size, fgcolor_max_len, fgcolor_vocab_path, fgcolor_snapshot_root,
                                new_result_image_name, last_result_image_name, *, model=None, ops=None, noise_seed=0):
    """Instance colorization of the listed instances on top of the last result image; writes
    <results_base_dir>/results/<image_id>/<new_result_image_name> and returns it as uint8 [768,768,3].
    Keyword-only extras: `model` (a resident FgColorModel: skips building / restoring), `ops` (operator set to build the
    model on), `noise_seed`."""
    import scipy.io
    import torch
    from PIL import Image
    assert type(inst_indices) is list
    vocab = load_vocab_dict_from_file(fgcolor_vocab_path)
    color_map = scipy.io.loadmat(os.path.join(data_base_dir, 'colorMapC46.mat'))['colorMap']
    categories46 = [str(color_map[i][0][0]) for i in range(46)]

    sketch_image = load_scene_sketch(sketch_path)
    inner_mask = scipy.io.loadmat(inner_masks_mat_path)['inner_masks']
    results_dir = os.path.join(results_base_dir, 'results', str(image_id))
    os.makedirs(results_dir, exist_ok=True)
    if last_result_image_name == '':
        base_image = sketch_image.copy()
    else:
        base_image = np.array(Image.open(os.path.join(results_dir, last_result_image_name)).convert('RGB'), dtype=np.uint8)
    new_result_image = base_image.copy()

    npz = np.load(segm_data_npz_path, allow_pickle=True)
    pred_class_ids = np.array(npz['pred_class_ids'], dtype=np.int32)
    pred_boxes = np.array(npz['pred_boxes'], dtype=np.int32)
    pred_masks = expand_small_segmentation_mask(npz['pred_masks'], pred_boxes)
    grass = [i for i in range(len(pred_class_ids)) if pred_class_ids[i] == GRASS_LABEL]

    inst_color_text = segment_user_input_text(input_text)
    print('## segment_user_input_text: ', inst_color_text)
    ids = np.array(preprocess_sentence(inst_color_text, vocab, fgcolor_max_len), dtype=np.int32)[None]

    # every instance is prepared (and validated) before the generator is touched
    jobs = []
    for inst_idx in inst_indices:
        class_id46 = int(pred_class_ids[inst_idx])
        if class_id46 not in SKE_TO_FG_CLASS:
            raise Exception('Wrong matching instance: %s' % categories46[class_id46])
        bbox = pred_boxes[inst_idx]
        jobs.append((inst_idx, class_id46, bbox, prepare_instance_sketch(pred_masks[inst_idx], bbox, class_id46)))

    if model is None:
        model = _load_generator(fgcolor_snapshot_root, fgcolor_vocab_size, ops=ops)
    dev = model.device
    gen = torch.Generator().manual_seed(noise_seed)
    for inst_idx, class_id46, bbox, sketch in jobs:
        y1, x1, y2, x2 = bbox
        labels = torch.tensor([SKE_TO_FG_CLASS[class_id46]], dtype=torch.int32, device=dev)
        noise = torch.randn(1, 256, generator=gen).to(dev)
        sk = torch.from_numpy(np.ascontiguousarray(sketch)).to(dev)      # the prepared sketch is a transposed view
        # batch 1 per instance (see module docstring); on the CUDA operator set every instance after the second replays one
        # captured graph instead of ~830 launches from Python
        gen_fn = getattr(model, 'generate_replay', None) if getattr(model.ops, 'supports_cuda_graphs', False) else None
        out = (gen_fn or model.generate)(sk.to(noise.dtype).contiguous(), ids, labels, noise)        # [1,3,192,192]
        color = instance_result_postprocessing(out.detach().float().cpu().numpy(), bbox, 'NCHW', class_id46)
        box = new_result_image[y1: y2, x1: x2]
        sel = inner_mask[y1: y2, x1: x2] == inst_idx + 1
        box[sel] = color[sel]
        new_result_image[y1: y2, x1: x2] = box

    # grass strokes stay coloured; every other drawing is laid back over the result, shifted by one pixel as in the reference
    no_grass = np.zeros(inner_mask.shape, dtype=np.int32)
    for gi in grass:
        no_grass[inner_mask == gi + 1] = 1
    moved = sketch_image.copy()
    moved[1: IMAGE_SIZE, 1: IMAGE_SIZE] = sketch_image[0: IMAGE_SIZE - 1, 0: IMAGE_SIZE - 1]
    drawn = np.logical_and(moved[:, :, 0] == 0, no_grass != 1)
    new_result_image[drawn] = moved[drawn]
    Image.fromarray(new_result_image, 'RGB').save(os.path.join(results_dir, new_result_image_name), 'PNG')
    return new_result_image
