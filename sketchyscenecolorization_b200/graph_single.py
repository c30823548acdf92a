"""`build_single_graph` -- the call both `main_procedure.{validation,test,inference}` and the whole-pipeline
caller (Pipeline_utils/fg_color_utils.py:258-265) use to get the generator.

Reference: obj_lib/graph_single.py:221-314.  The reference returns symbolic TF tensors to `sess.run`; here the
graph executes eagerly on the bound `FgColorModel`, with the same argument list and the same return convention:
`[image_gens, images, sketches]` when `training=False`, `(loss_g, loss_d, grad_g, grad_d)` when `training=True`
(gradients are left in the model's flat gradient buffers; the flat tensors are returned).
Tensors at this boundary are NCHW float32 (Config.data_format); ids are int32."""
from __future__ import annotations

import numpy as np
import torch

_default_model = None


def bind_model(model):
    """Select the FgColorModel (weights + kernels on one GPU) that subsequent build_single_graph calls run on."""
    global _default_model
    _default_model = model
    return model


def build_single_graph(images, sketches, images_d, image_data_class_id, image_data_class_id_d, text_vocab_indiceses,
                       batch_size, training, LSTM_hybrid, vocab_size, ld=10, data_format='NCHW', distance_map=True,
                       optim_g=None, optim_d=None, block_type='MRU', model=None, noise=None):
    m = model or _default_model
    if m is None:
        raise RuntimeError("build_single_graph: no model bound (call graph_single.bind_model(FgColorModel(...)) first)")
    if block_type not in ('MRU', 'Pix2Pix', 'Residual'):
        raise ValueError("block_type %r (MRU, Pix2Pix, Residual)" % (block_type,))
    if block_type != getattr(m, 'block_type', 'MRU'):
        raise ValueError("build_single_graph(block_type=%r) on a model built as %r" % (block_type, m.block_type))
    if data_format != 'NCHW':
        raise ValueError("the boundary layout is NCHW (reference config.py:6)")
    dev = m.device
    f = lambda t: None if t is None else torch.as_tensor(np.asarray(t) if not torch.is_tensor(t) else t).float().to(dev).contiguous()  # noqa: E731
    i32 = lambda t: None if t is None else torch.as_tensor(np.asarray(t) if not torch.is_tensor(t) else t).int().to(dev).contiguous()  # noqa: E731
    sketches_d, images_dev = f(sketches), f(images)
    N = sketches_d.shape[0]
    assert N == batch_size, "batch_size %d does not match the fed tensors (%d)" % (batch_size, N)
    ids = np.asarray(text_vocab_indiceses.cpu() if torch.is_tensor(text_vocab_indiceses) else text_vocab_indiceses, dtype=np.int32)
    if noise is None:            # the reference draws tf.random_normal([N,256]) inside the graph, also at inference
        noise = torch.randn(N, 256, device=dev)
    noise = f(noise)
    cls = i32(image_data_class_id)
    m.G.lstm_hybrid = bool(LSTM_hybrid)
    if not training:
        gen = m.generate(sketches_d, ids, cls, noise)
        return [gen, images_dev, sketches_d]
    batch = dict(sketch=sketches_d, images=images_dev, images_d=f(images_d), cls=cls, cls_d=i32(image_data_class_id_d),
                 text=ids, noise=noise)
    out_d = m.d_step_grads(batch)
    out_g = m.g_step_grads(batch)
    return out_g["loss"], out_d["loss"], m.gstore.grad, m.dstore.grad
