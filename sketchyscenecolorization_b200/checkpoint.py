"""Snapshots in the reference's directory / file-name layout (main_procedure.py:141,235-237,550,559;
obj_colorization_main.py:58-62):

    <ckpt_dir>/model_<i>.ckpt-<i>.index                  JSON: variable name -> dtype, shape, byte offset
    <ckpt_dir>/model_<i>.ckpt-<i>.data-00000-of-00001    raw little-endian tensors, concatenated
    <ckpt_dir>/checkpoint                                'model_checkpoint_path: "model_<i>.ckpt-<i>"' (+ history)

All global variables are saved under the reference's TF variable names: weights, Adam second moments
(`<name>/Adam_1`), SN `u` vectors, optimiser step counts and `counter`.  The byte format of .index/.data is
this package's own (TF's SSTable/protobuf tensor bundle is listed under "next" in DESIGN.md)."""
from __future__ import annotations

import json
import os
import re

import numpy as np
import torch


def latest_checkpoint(ckpt_dir):
    """tf.train.latest_checkpoint: path prefix of the newest snapshot or None."""
    state = os.path.join(ckpt_dir, "checkpoint")
    if not os.path.exists(state):
        return None
    m = re.search(r'^model_checkpoint_path:\s*"([^"]+)"', open(state).read(), flags=re.M)
    if not m:
        return None
    prefix = os.path.join(ckpt_dir, m.group(1))
    return prefix if os.path.exists(prefix + ".index") else None


def _collect(model, counter):
    out = {}
    for store, tag in ((model.gstore, "generator"), (model.dstore, "discriminator")):
        if store is None:
            continue
        for k, v in store.p.items():
            out[k] = v
            o = store.offsets[k]
            out[k + "/Adam_1"] = store.adam_v[o:o + v.numel()].view(v.shape)
        for k, v in store.state.items():
            out[k] = v
        out["beta2_power/" + tag] = torch.tensor(float(store.adam_t))
    out["counter"] = torch.tensor(float(counter))
    return out


def save(model, ckpt_dir, step, counter, max_to_keep=100):
    os.makedirs(ckpt_dir, exist_ok=True)
    name = "model_%d.ckpt-%d" % (step, step)
    index, off = {}, 0
    with open(os.path.join(ckpt_dir, name + ".data-00000-of-00001"), "wb") as f:
        for k, v in _collect(model, counter).items():
            a = v.detach().float().cpu().numpy().astype("<f4")
            index[k] = {"dtype": "float32", "shape": list(a.shape), "offset": off, "nbytes": a.nbytes}
            f.write(a.tobytes())
            off += a.nbytes
    with open(os.path.join(ckpt_dir, name + ".index"), "w") as f:
        json.dump(index, f)
    state = os.path.join(ckpt_dir, "checkpoint")
    hist = []
    if os.path.exists(state):
        hist = re.findall(r'^all_model_checkpoint_paths:\s*"([^"]+)"', open(state).read(), flags=re.M)
    hist = (hist + [name])[-max_to_keep:]
    with open(state, "w") as f:
        f.write('model_checkpoint_path: "%s"\n' % name)
        for h in hist:
            f.write('all_model_checkpoint_paths: "%s"\n' % h)
    return os.path.join(ckpt_dir, name)


def restore(model, prefix, strict=True):
    """Loads a snapshot written by `save`; returns the saved `counter`."""
    index = json.load(open(prefix + ".index"))
    blob = np.fromfile(prefix + ".data-00000-of-00001", dtype=np.uint8)

    def get(k):
        e = index[k]
        return torch.from_numpy(blob[e["offset"]:e["offset"] + e["nbytes"]].view("<f4").reshape(e["shape"]).copy())

    for store, tag in ((model.gstore, "generator"), (model.dstore, "discriminator")):
        if store is None:
            continue
        for k, v in list(store.p.items()) + list(store.state.items()):
            if k in index:
                v.copy_(get(k).to(v.dtype))
            elif strict:
                raise KeyError("snapshot %s has no variable %s" % (prefix, k))
        for k, v in store.p.items():
            if k + "/Adam_1" in index:
                o = store.offsets[k]
                store.adam_v[o:o + v.numel()].copy_(get(k + "/Adam_1").reshape(-1))
        if "beta2_power/" + tag in index:
            store.adam_t = int(get("beta2_power/" + tag).item())
    return int(get("counter").item()) if "counter" in index else 0
