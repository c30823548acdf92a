"""Snapshots in the reference's directory / file-name layout AND byte format (main_procedure.py:141,235-237,550,559;
obj_colorization_main.py:58-62):

    <ckpt_dir>/model_<i>.ckpt-<i>.index                  TensorFlow V2 tensor-bundle index (SSTable of BundleEntryProto)
    <ckpt_dir>/model_<i>.ckpt-<i>.data-00000-of-00001    the tensors' bytes, little endian, in key order
    <ckpt_dir>/checkpoint                                text CheckpointState: model_checkpoint_path + history

written / read by tf_bundle.py (no TensorFlow needed).  Everything `tf.train.Saver()` of the reference graph saves is
present under the reference's variable names, so that a TF-1 reader of the reference finds every key it asks for:

    <var>                       weights (fp32, the reference's HWIO / [25,C] layouts)
    <var>/Adam, <var>/Adam_1    AdamOptimizer slots m and v.  beta1 = 0 (graph_single.py:588), so m is just the last
                                gradient and never influences a later step: it is written as zeros.
    beta1_power, beta2_power    non-slot Adam accumulators of the generator's optimiser (created first, graph_single.py:208),
    beta1_power_1, beta2_power_1   and of the discriminator's: beta^(t+1) after t steps
    discriminator/<scope>/discriminator/<scope>/u   spectral-norm vectors (sn.py:17-18; doubled path, provisional -- SURVEY 8a)
    Variable                    the int32 iteration `counter` (main_procedure.py:105, an unnamed tf.Variable)

`restore` also still reads the JSON-indexed snapshots written by earlier versions of this package."""
from __future__ import annotations

import json
import math
import os
import re

import numpy as np
import torch

from . import tf_bundle

_BETA2 = 0.9


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
    """name -> numpy array of everything the reference's Saver would save."""
    out = {}
    for store, suffix in ((model.gstore, ""), (model.dstore, "_1")):
        if store is None:
            continue
        opt = getattr(store, "optimizer", "adam")
        slot = lambda t, k, v: t[store.offsets[k]:store.offsets[k] + v.numel()].view(v.shape).detach().float().cpu().numpy()  # noqa: E731
        for k, v in store.p.items():
            out[k] = v.detach().float().cpu().numpy()
            if opt == "adam":
                out[k + "/Adam"] = np.zeros(tuple(v.shape), dtype=np.float32)
                out[k + "/Adam_1"] = slot(store.adam_v, k, v)
            elif opt == "rmsprop":                          # slots `rms` and `momentum` (zero: momentum = 0.0)
                out[k + "/RMSProp"] = slot(store.adam_v, k, v)
                out[k + "/RMSProp_1"] = np.zeros(tuple(v.shape), dtype=np.float32)
            elif opt == "adadelta":                         # slots `accum` and `accum_update`
                out[k + "/Adadelta"] = slot(store.adam_v, k, v)
                out[k + "/Adadelta_1"] = slot(store.opt_s2, k, v)
            else:
                out[k + "/Adagrad"] = slot(store.adam_v, k, v)
        for k, v in store.state.items():
            out[k] = v.detach().float().cpu().numpy()
        if opt == "adam":
            out["beta1_power" + suffix] = np.float32(0.0)
            out["beta2_power" + suffix] = np.float32(_BETA2 ** (store.adam_t + 1))
    out["Variable"] = np.int32(counter)
    return out


def save(model, ckpt_dir, step, counter, max_to_keep=100):
    os.makedirs(ckpt_dir, exist_ok=True)
    name = "model_%d.ckpt-%d" % (step, step)
    tf_bundle.write_bundle(os.path.join(ckpt_dir, name), _collect(model, counter))
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


def _load_legacy(prefix):
    index = json.load(open(prefix + ".index"))
    blob = np.fromfile(prefix + ".data-00000-of-00001", dtype=np.uint8)
    out = {}
    for k, e in index.items():
        out[k] = blob[e["offset"]:e["offset"] + e["nbytes"]].view("<f4").reshape(e["shape"]).copy()
    return out


def default_aliases(key):
    """Other names a TensorFlow-written snapshot may hold `key` under.  Two names of this package's inventory are provisional
    (checking them needs one of the published checkpoints, which cannot be fetched here): the spectral-norm vector --
    sn.py:17-18 opens a nested variable_scope named after the weight's own scope, which yields either the doubled path
    `discriminator/<s>/discriminator/<s>/u` or, when TF re-enters the existing scope, `discriminator/<s>/u` -- and the LSTM
    variables, `kernel` / `bias` since TF 1.2, `weights` / `biases` before."""
    out = []
    m = re.match(r"^(.*)/\1/u$", key)
    if m:
        out.append(m.group(1) + "/u")
    if key.endswith("/basic_lstm_cell/kernel"):
        out.append(key[:-len("kernel")] + "weights")
    if key.endswith("/basic_lstm_cell/bias"):
        out.append(key[:-len("bias")] + "biases")
    return out


def restore(model, prefix, strict=True, key_map=None):
    """Loads a snapshot (TensorFlow V2 bundle, e.g. one written by `save` or by the reference's tf.train.Saver; or the JSON
    format of earlier versions of this package).  Returns the saved iteration counter.

    key_map: optional dict {this package's variable name: name in the snapshot} or callable name -> name | [names], tried
    before the name itself; `default_aliases` is tried after it."""
    with open(prefix + ".index", "rb") as f:
        legacy = f.read(1) == b"{"
    t = _load_legacy(prefix) if legacy else tf_bundle.read_bundle(prefix)
    if key_map is not None or any(a in t and k not in t for store in (model.gstore, model.dstore) if store is not None
                                  for k in list(store.p) + list(store.state) for a in default_aliases(k)):
        t = dict(t)
        for store in (model.gstore, model.dstore):
            if store is None:
                continue
            for k in list(store.p) + list(store.state):
                cands = []
                if key_map is not None:
                    c = key_map(k) if callable(key_map) else key_map.get(k)
                    cands += [c] if isinstance(c, str) else list(c or [])
                cands += [k] + default_aliases(k)
                src = next((c for c in cands if c in t), None)
                if src is not None and src != k:
                    t[k] = t[src]
                    for slot in ("/Adam", "/Adam_1", "/RMSProp", "/RMSProp_1", "/Adadelta", "/Adadelta_1", "/Adagrad"):
                        if src + slot in t:
                            t[k + slot] = t[src + slot]

    for store, suffix, tag in ((model.gstore, "", "generator"), (model.dstore, "_1", "discriminator")):
        if store is None:
            continue
        for k, v in list(store.p.items()) + list(store.state.items()):
            if k in t:
                v.copy_(torch.from_numpy(np.asarray(t[k], dtype=np.float32)).reshape(v.shape).to(v.dtype))
            elif strict:
                raise KeyError("snapshot %s has no variable %s" % (prefix, k))
        s1_name = {"adam": "/Adam_1", "rmsprop": "/RMSProp", "adadelta": "/Adadelta", "adagrad": "/Adagrad"}[
            getattr(store, "optimizer", "adam")]
        for k, v in store.p.items():
            o = store.offsets[k]
            if k + s1_name in t:
                store.adam_v[o:o + v.numel()].copy_(torch.from_numpy(np.asarray(t[k + s1_name], dtype=np.float32)).reshape(-1))
            if getattr(store, "opt_s2", None) is not None and k + "/Adadelta_1" in t:
                store.opt_s2[o:o + v.numel()].copy_(torch.from_numpy(np.asarray(t[k + "/Adadelta_1"], dtype=np.float32)).reshape(-1))
        if "beta2_power" + suffix in t:
            b2 = float(t["beta2_power" + suffix])
            store.adam_t = max(0, int(round(math.log(max(b2, 1e-300)) / math.log(_BETA2))) - 1)
        elif "beta2_power/" + tag in t:                      # legacy: the step count itself
            store.adam_t = int(t["beta2_power/" + tag])
    if "Variable" in t:
        return int(t["Variable"])
    return int(t["counter"]) if "counter" in t else 0
