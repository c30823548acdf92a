"""Generates tests/golden/match_seg_data.npz, match_overall_masks.npz (bit-packed input masks) and match_cases.json by running the REFERENCE's own
Instance_Matching/data_processing/sketch_data_processing.get_pred_instance_mask on a small synthetic segmentation file.
Run in the build container only (the reference tree does not exist on the GPU box); the outputs are committed.
matplotlib (imported by the reference module for its plotting helpers, not used here) is absent from this image: stubbed."""
import json
import os
import sys
import types

import numpy as np

sys.modules.setdefault("matplotlib", types.ModuleType("matplotlib"))
sys.modules.setdefault("matplotlib.pyplot", types.ModuleType("matplotlib.pyplot"))
sys.path.insert(0, "/root/reference/Instance_Matching/data_processing")
import sketch_data_processing as ref  # noqa: E402

_np_load = np.load
np.load = lambda *a, **k: _np_load(*a, **{**k, "allow_pickle": True})     # the reference predates numpy's allow_pickle=False default
HERE = os.path.dirname(os.path.abspath(__file__))
rs = np.random.RandomState(3)
S = ref.IMAGE_SIZE
boxes, masks, cls = [], [], []
for i in range(9):
    h, w = rs.randint(20, 200), rs.randint(20, 200)
    y1, x1 = rs.randint(0, S - h), rs.randint(0, S - w)
    boxes.append([y1, x1, y1 + h - 1, x1 + w - 1])
    masks.append((rs.rand(h, w) < 0.35).astype(np.uint8))
    cls.append(rs.randint(1, 47))
boxes.append([5, 5, 9, 9]); masks.append(np.zeros((5, 5), np.uint8)); cls.append(3)        # an empty instance mask
pm = np.empty(len(masks), dtype=object)
for i, m in enumerate(masks):
    pm[i] = m
npz_path = os.path.join(HERE, "match_seg_data.npz")
np.savez_compressed(npz_path, pred_masks=pm, pred_boxes=np.asarray(boxes, np.int32), pred_class_ids=np.asarray(cls, np.int32))

cases, overalls = [], []
for picks, noise, drop in (((0,), 0.0, 0.0), ((2, 5), 0.01, 0.2), ((1, 3, 7), 0.02, 0.45), ((), 0.05, 0.0), ((4, 8), 0.0, 0.55),
                           (tuple(range(9)), 0.0, 0.3)):
    overall = (rs.rand(S, S) < noise).astype(np.float32)
    for i in picks:
        y1, x1, y2, x2 = boxes[i]
        keep = masks[i] * (rs.rand(*masks[i].shape) >= drop)
        overall[y1:y2 + 1, x1:x2 + 1] = np.maximum(overall[y1:y2 + 1, x1:x2 + 1], keep)
    with np.errstate(invalid="ignore", divide="ignore"):
        m, s, b, c, idx = ref.get_pred_instance_mask(npz_path, overall.copy())
    overalls.append(np.packbits(overall.astype(np.uint8).reshape(-1)))
    cases.append({"seed_picks": list(picks), "matched": [int(i) for i in idx],
                  "scores": [float(x) for x in np.atleast_1d(s)] if len(idx) else [],
                  "class_ids": [int(x) for x in np.atleast_1d(c)] if len(idx) else [],
                  "masks_shape": list(m.shape), "masks_sum": int(m.sum()) if len(idx) else 0})
np.savez_compressed(os.path.join(HERE, "match_overall_masks.npz"), overall=np.stack(overalls))
json.dump({"size": S, "cases": cases}, open(os.path.join(HERE, "match_cases.json"), "w"))
print("wrote", len(cases), "cases;", [c["matched"] for c in cases])
