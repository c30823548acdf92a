"""Whole-pipeline caller of the instance-matching model: caption -> indices of the segmented instances it refers to.

Reference: Pipeline_utils/fg_matching_utils.build_instance_matching (:14-77; the same feed / fetch as
Instance_Matching/matching_main.inference, :420-466) and Instance_Matching/data_processing/sketch_data_processing --
load_image2 (:24-29), get_pred_instance_mask (:254-281), compute_mask_occupied_percentage (:241-251),
expand_small_segmentation_mask (:202-214).  `sketchyscene_colorization_main.colorization_main` calls it for an FG instruction
(reference: sketchyscene_colorization_main.py:33-37) and hands the indices to `build_instance_colorization`.

The model is `rmi.RMIModel` (restored from the TensorFlow snapshot under `match_snapshot_root`, or handed in); the step from its
binary stroke mask to instances is byte work on the host, as in the reference: an instance of the segmentation file is matched
when more than half of its mask pixels lie inside the predicted mask.
"""
from __future__ import annotations

import os

import numpy as np

from . import rmi

IMAGE_SIZE = 768                    # sketch_data_processing.py:12
OCCUPIED_THRESHOLD = 0.5            # get_pred_instance_mask's default


def load_sketch(path):
    """load_image2: RGB, NEAREST-resized to 768 x 768 when it has another size; uint8 [768,768,3]."""
    from PIL import Image
    im = Image.open(path).convert("RGB")
    if im.size != (IMAGE_SIZE, IMAGE_SIZE):
        im = im.resize((IMAGE_SIZE, IMAGE_SIZE), resample=Image.NEAREST)
    return np.asarray(im, dtype=np.uint8)


def get_pred_instance_mask(segm_data_path, pred_overall_mask, mask_occupied_threshold=OCCUPIED_THRESHOLD):
    """-> (masks [H,W,K] uint8, scores [K], boxes [K,4], class ids [K], matched instance indices) -- empty arrays and [] when
    nothing matches.  The instance masks of the file are box-sized; the share of an instance inside the predicted mask is
    evaluated on its box (what the reference computes after pasting every mask into a full-size canvas)."""
    npz = np.load(segm_data_path, allow_pickle=True)
    small = npz['pred_masks']
    class_ids = np.asarray(npz['pred_class_ids'], dtype=np.int32)
    boxes = np.asarray(npz['pred_boxes'], dtype=np.int32)
    overall = np.asarray(pred_overall_mask) != 0
    H, W = overall.shape
    picked, scores = [], []
    for i in range(len(small)):
        y1, x1, y2, x2 = (int(v) for v in boxes[i])
        inst = np.asarray(small[i]) != 0
        area = int(inst.sum())
        if area == 0:
            continue                                 # the reference divides 0 by 0 here and nan never passes the threshold
        share = float(np.logical_and(overall[y1:y2 + 1, x1:x2 + 1], inst).sum()) / area
        if share > mask_occupied_threshold:
            picked.append(i)
            scores.append(share)
    if not picked:
        return np.array(()), np.array(()), np.array(()), np.array(()), []
    masks = np.zeros((H, W, len(picked)), dtype=np.uint8)
    for k, i in enumerate(picked):
        y1, x1, y2, x2 = (int(v) for v in boxes[i])
        masks[y1:y2 + 1, x1:x2 + 1, k] = np.asarray(small[i], dtype=np.uint8)
    return masks, np.asarray(scores), boxes[picked], class_ids[picked], picked


def load_matching_model(match_snapshot_root, match_vocab_size, ops=None, device="cuda:0"):
    """RMIModel with the variables of the latest TensorFlow snapshot under `match_snapshot_root` (tf.train.Saver files, read
    by tf_bundle without TensorFlow; the reference's `snapshot_restorer.restore`, fg_matching_utils.py:34-38)."""
    import torch

    from . import checkpoint, tf_bundle
    if ops is None:
        from .cuda_ops import CudaOps
        ops = CudaOps(device, torch.float32)
    prefix = checkpoint.latest_checkpoint(match_snapshot_root)
    if prefix is None:
        raise FileNotFoundError("no instance-matching snapshot under %s (a `checkpoint` state file and the V2 bundle it names)"
                                % match_snapshot_root)
    print('Restore:', prefix)
    model = rmi.RMIModel(ops, ops.device, vocab_size=match_vocab_size)
    tensors = tf_bundle.read_bundle(prefix)
    model.load_state_dict({k: torch.from_numpy(np.ascontiguousarray(v)) for k, v in tensors.items() if k in model.store.p})
    return model


def build_instance_matching(data_base_dir, sketch_path, input_text, segm_data_npz_path, match_vocab_path, match_vocab_size,
                            match_snapshot_root, match_max_len, *, model=None, ops=None):
    """Same positional arguments as the reference function; returns the list of matched instance indices."""
    with open(match_vocab_path) as f:
        vocab = {w.strip(): i for i, w in enumerate(f.readlines())}
    if model is None:
        model = load_matching_model(match_snapshot_root, match_vocab_size, ops)
    sketch = load_sketch(sketch_path)
    predicts = model.predict_mask(sketch, input_text, vocab, T=match_max_len)          # [768,768] in {0, 1}
    masks, scores, boxes, class_ids, matched = get_pred_instance_mask(segm_data_npz_path, predicts)
    print('pred_masks', masks.shape)
    print('pred_scores', scores.shape, scores)
    print('pred_class_ids', class_ids.shape, class_ids)
    return matched
