"""Caption encoder: word LSTM + multimodal ("convolutional") LSTM fused with the 6x6 bottleneck.

Reference: models_collection.encode_feat_with_text (:150-248).  The reference unrolls B x T copies of
two BasicLSTMCells in the TF graph and runs each sample separately; here all samples advance together,
pad tokens (id 0, tf.cond :235) are handled with a per-sample mask, and the time-invariant part of the
mLSTM input product is hoisted (SURVEY 7.2 "mLSTM algebra"):

    [vis, tile(e), tile(l2n(h_w)), h_a] @ K_a
        = vis @ K_a[0:D]                      (once per forward)
        + tile([e, l2n(h_w)] @ K_a[D:3D])     (one [N,2D]x[2D,4D] product per step, broadcast over positions)
        + h_a @ K_a[3D:4D]                    (the only per-step [N*P, D] x [D, 4D] product)

All state is fp32 (LSTM state, gates, l2-norms); matrices are multiplied by the conv kernels with k=1.
BPTT is written out by hand.
"""
from __future__ import annotations

import numpy as np
import torch

_PRE = "generator/TextLSTM"
_KW = _PRE + "/RNN/WLSTM/multi_rnn_cell/cell_0/basic_lstm_cell/kernel"
_BW = _PRE + "/RNN/WLSTM/multi_rnn_cell/cell_0/basic_lstm_cell/bias"
_KA = _PRE + "/RNN/ALSTM/multi_rnn_cell/cell_0/basic_lstm_cell/kernel"
_BA = _PRE + "/RNN/ALSTM/multi_rnn_cell/cell_0/basic_lstm_cell/bias"
_EMB = _PRE + "/embedding"


def _mat(w2d):
    """[K,N] matrix as a 1x1 HWIO conv weight."""
    return w2d.view(1, 1, w2d.shape[0], w2d.shape[1])


def _rows(t):
    """[R,C] rows as NHWC [R,1,1,C]."""
    return t.view(t.shape[0], 1, 1, t.shape[1])


def text_fusion_fwd(ops, store, e4, ids_host, save=True):
    """e4: [N,h,w,D] activation; ids_host: int array [N,T] on the HOST (numpy / CPU tensor), or an int32 DEVICE tensor.
    With host ids, time steps at which every caption is <pad> are skipped outright (the reference's tf.cond, :235);
    with device ids nothing on the host depends on the data (CUDA-graph capture): every step runs and <pad> samples
    are masked inside the cell kernel -- same result.  Returns ([N,h,w,D], ctx)."""
    on_device = torch.is_tensor(ids_host) and ids_host.is_cuda
    ids_np = None if on_device else np.asarray(ids_host, dtype=np.int32)
    N, hh, ww, D = e4.shape
    P = hh * ww
    R = N * P
    T = ids_host.shape[1]
    dev = e4.device
    emb, kw, bw, ka, ba = store.p[_EMB], store.p[_KW], store.p[_BW], store.p[_KA], store.p[_BA]
    # [N,T] int32 on the device; pad mask = (id == 0)
    ids_dev = ids_host.to(torch.int32).contiguous() if on_device else torch.as_tensor(ids_np, device=dev).contiguous()
    f32 = torch.float32

    e4r = ops.cast(e4, f32).view(R, D)
    vis, inv_v = ops.l2norm_rows_fwd(e4r)                                           # :201-202
    gv = ops.conv_fwd([(_rows(vis), False)], _mat(ka[0:D]), None, out_dtype=f32).view(R, 4 * D)
    cw = ops.zeros_f32((N, D)); hw = ops.zeros_f32((N, D))                          # :190
    ca = ops.zeros_f32((R, D)); ha = ops.zeros_f32((R, D))                          # :196,204
    steps = []
    for t in range(T):
        if ids_np is not None and not (ids_np[:, t] != 0).any():      # every sample is <pad> here: state passes through
            continue
        e_t = ops.embedding_fwd(emb, ids_dev, t)                                         # :182,211
        gw = ops.conv_fwd([(_rows(e_t), False), (_rows(hw), False)], _mat(kw), bw, out_dtype=f32).view(N, 4 * D)
        cw2, hw2, pre_w = ops.lstm_cell_fwd(gw, None, None, cw, hw, ids_dev, t, 1)        # :212-213
        lang, inv_l = ops.l2norm_rows_fwd(hw2)                                      # :215-216
        r_t = ops.conv_fwd([(_rows(e_t), False), (_rows(lang), False)], _mat(ka[D:3 * D]), None,
                           out_dtype=f32).view(N, 4 * D)
        ga = ops.conv_fwd([(_rows(ha), False)], _mat(ka[3 * D:4 * D]), ba, out_dtype=f32).view(R, 4 * D)
        ca2, ha2, pre_a = ops.lstm_cell_fwd(ga, gv, r_t, ca, ha, ids_dev, t, P)           # :225-226
        if save:
            steps.append(dict(t=t, e=e_t, cw_prev=cw, hw_prev=hw, cw=cw2, hw=hw2, pre_w=pre_w,
                              lang=lang, inv_l=inv_l, ca_prev=ca, ha_prev=ha, ca=ca2, pre_a=pre_a))
        cw, hw, ca, ha = cw2, hw2, ca2, ha2
    out = ops.atanh_relu_fwd(ha)                                                    # :239-241
    out = ops.cast(out.view(N, hh, ww, D), e4.dtype)
    ctx = dict(steps=steps, ids=ids_dev, vis=vis, inv_v=inv_v, ha=ha, shape=(N, hh, ww, D), in_dtype=e4.dtype) if save else None
    return out, ctx


def text_fusion_bwd(ops, store, g_out, ctx):
    """Returns g_e4 [N,h,w,D]; accumulates gradients of embedding / both LSTM kernels / biases."""
    N, hh, ww, D = ctx["shape"]
    P = hh * ww
    R = N * P
    f32 = torch.float32
    kw, ka = store.p[_KW], store.p[_KA]
    dkw, dbw, dka, dba, demb = store.g[_KW], store.g[_BW], store.g[_KA], store.g[_BA], store.g[_EMB]

    g_ha = ops.atanh_relu_bwd(ops.cast(g_out, f32).view(R, D), ctx["ha"])
    g_ca = ops.zeros_f32((R, D))
    g_hw = ops.zeros_f32((N, D))
    g_cw = ops.zeros_f32((N, D))
    g_gv = None
    ids = ctx["ids"]
    for s in reversed(ctx["steps"]):
        t = s["t"]
        # ---- mLSTM cell
        g_pre_a, g_ca, g_ha_pass = ops.lstm_cell_bwd(g_ca, g_ha, s["pre_a"], s["ca_prev"], s["ca"], ids, t, P)
        gpa4 = _rows(g_pre_a)
        ops.conv_wgrad([(_rows(s["ha_prev"]), False)], gpa4, _mat(dka[3 * D:4 * D]), dba)
        g_ha = ops.conv_dgrad(gpa4, _mat(ka[3 * D:4 * D]), 0, D, out_dtype=f32).view(R, D)
        ops.add_(g_ha, g_ha_pass)
        if g_gv is None:
            g_gv = g_pre_a
        else:
            ops.add_(g_gv, g_pre_a)
        g_r = ops.rows_group_sum(g_pre_a, P)                                         # [N,4D]
        gr4 = _rows(g_r)
        ops.conv_wgrad([(_rows(s["e"]), False), (_rows(s["lang"]), False)], gr4, _mat(dka[D:3 * D]), None)
        g_e = ops.conv_dgrad(gr4, _mat(ka[D:3 * D]), 0, D, out_dtype=f32).view(N, D)
        g_lang = ops.conv_dgrad(gr4, _mat(ka[D:3 * D]), D, D, out_dtype=f32).view(N, D)
        ops.add_(g_hw, ops.l2norm_rows_bwd(g_lang, s["lang"], s["inv_l"]))
        # ---- word LSTM cell
        g_pre_w, g_cw, g_hw_pass = ops.lstm_cell_bwd(g_cw, g_hw, s["pre_w"], s["cw_prev"], s["cw"], ids, t, 1)
        gpw4 = _rows(g_pre_w)
        ops.conv_wgrad([(_rows(s["e"]), False), (_rows(s["hw_prev"]), False)], gpw4, _mat(dkw), dbw)
        ops.add_(g_e, ops.conv_dgrad(gpw4, _mat(kw), 0, D, out_dtype=f32).view(N, D))
        g_hw = ops.conv_dgrad(gpw4, _mat(kw), D, D, out_dtype=f32).view(N, D)
        ops.add_(g_hw, g_hw_pass)
        ops.embedding_bwd(g_e, ids, t, demb)
    if g_gv is None:          # all-pad batch: output is relu(0) = 0, no gradient reaches e4
        return ops.cast(ops.zeros_f32((N, hh, ww, D)), ctx["in_dtype"])
    ops.conv_wgrad([(_rows(ctx["vis"]), False)], _rows(g_gv), _mat(dka[0:D]), None)
    g_vis = ops.conv_dgrad(_rows(g_gv), _mat(ka[0:D]), 0, D, out_dtype=f32).view(R, D)
    g_e4 = ops.l2norm_rows_bwd(g_vis, ctx["vis"], ctx["inv_v"])
    return ops.cast(g_e4.view(N, hh, ww, D), ctx["in_dtype"])
