"""Caption encoder: word LSTM + multimodal ("convolutional") LSTM fused with the 6x6 bottleneck.

Reference: models_collection.encode_feat_with_text (:150-248).  The reference unrolls B x T copies of
two BasicLSTMCells in the TF graph and runs each sample separately; here all samples advance together,
pad tokens (id 0, tf.cond :235) are handled with a per-sample mask, and the time-invariant part of the
mLSTM input product is hoisted (SURVEY 7.2 "mLSTM algebra"):

    [vis, tile(e), tile(l2n(h_w)), h_a] @ K_a
        = vis @ K_a[0:D]                      (once per forward)
        + tile([e, l2n(h_w)] @ K_a[D:3D])     (one [N,2D]x[2D,4D] product per step, broadcast over positions)
        + h_a @ K_a[3D:4D]                    (the only per-step [N*P, D] x [D, 4D] product)

All state is fp32 (LSTM state, gates, l2-norms); matrices are multiplied by the conv kernels with k=1 (per-sample
[N, .] products run on the skinny-product kernel of conv_small.cu).  BPTT is written out by hand.  The per-step operands
of the weight gradients (embeddings, l2n(h_w), h_w, h_a and the gate gradients) are written into per-time-step stacks, so
each LSTM kernel gets ONE weight-gradient product over all steps instead of one per step, and the gradient to the
embedding rows -- which is not part of the recurrence -- is computed once at the end.
"""
from __future__ import annotations

import numpy as np
import torch

_PRE = "generator/TextLSTM"           # the fg networks; the background generator keeps the same cells under generator/mLSTM_G


def _names(prefix):
    """(embedding, WLSTM kernel, WLSTM bias, ALSTM kernel, ALSTM bias) variable names under `prefix`."""
    cell = prefix + "/RNN/%s/multi_rnn_cell/cell_0/basic_lstm_cell/%s"
    return (prefix + "/embedding", cell % ("WLSTM", "kernel"), cell % ("WLSTM", "bias"), cell % ("ALSTM", "kernel"),
            cell % ("ALSTM", "bias"))


def _mat(w2d):
    """[K,N] matrix as a 1x1 HWIO conv weight."""
    return w2d.view(1, 1, w2d.shape[0], w2d.shape[1])


def _rows(t):
    """[R,C] rows as NHWC [R,1,1,C]."""
    return t.view(t.shape[0], 1, 1, t.shape[1])


def _op(ops, t):
    """Matrix-product operand in the operator set's activation dtype: in bf16 training mode the LSTM products run as
    single-pass bf16 tensor-core GEMMs like every convolution (state, gates and accumulation stay fp32); in fp32 mode
    (inference, parity tests) this is the identity and the products run in the bf16x3 split mode."""
    return _rows(t if ops.act_dtype != torch.bfloat16 else ops.cast(t, torch.bfloat16))


def text_fusion_fwd(ops, store, e4, ids_host, save=True, prefix=_PRE):
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
    emb, kw, bw, ka, ba = (store.p[n] for n in _names(prefix))
    # [N,T] int32 on the device; pad mask = (id == 0)
    ids_dev = ids_host.to(torch.int32).contiguous() if on_device else torch.as_tensor(ids_np, device=dev).contiguous()
    f32 = torch.float32
    # steps at which at least one caption has a real token (all of them when the ids live on the device)
    ts = [t for t in range(T) if ids_np is None or (ids_np[:, t] != 0).any()]
    S = len(ts)

    e4r = ops.cast(e4, f32).view(R, D)
    vis, inv_v = ops.l2norm_rows_fwd(e4r)                                           # :201-202
    gv = ops.conv_fwd([(_op(ops, vis), False)], _mat(ka[0:D]), None, out_dtype=f32).view(R, 4 * D)
    # stacks over the executed steps: slot j holds the operand of step ts[j]; the h stacks hold the INPUT state of step j
    # in slot j (slot 0 = the zero initial state, :190,:196,:204) and receive the output state in slot j+1
    e_all = ops.zeros_f32((max(S, 1), N, D))
    lang_all = ops.zeros_f32((max(S, 1), N, D))
    hw_all = ops.zeros_f32((S + 1, N, D))
    ha_all = ops.zeros_f32((S + 1, R, D))
    cw = ops.zeros_f32((N, D))
    ca = ops.zeros_f32((R, D))
    steps = []
    for j, t in enumerate(ts):
        hw, ha = hw_all[j], ha_all[j]
        e_t = ops.embedding_fwd(emb, ids_dev, t, out=e_all[j])                           # :182,211
        e_op = _op(ops, e_t)
        gw = ops.conv_fwd([(e_op, False), (_op(ops, hw), False)], _mat(kw), bw, out_dtype=f32).view(N, 4 * D)
        cw2, hw2, pre_w = ops.lstm_cell_fwd(gw, None, None, cw, hw, ids_dev, t, 1, out_h=hw_all[j + 1])    # :212-213
        lang, inv_l = ops.l2norm_rows_fwd(hw2, out=lang_all[j])                     # :215-216
        r_t = ops.conv_fwd([(e_op, False), (_op(ops, lang), False)], _mat(ka[D:3 * D]), None,
                           out_dtype=f32).view(N, 4 * D)
        ga = ops.conv_fwd([(_op(ops, ha), False)], _mat(ka[3 * D:4 * D]), ba, out_dtype=f32).view(R, 4 * D)
        ca2, ha2, pre_a = ops.lstm_cell_fwd(ga, gv, r_t, ca, ha, ids_dev, t, P, out_h=ha_all[j + 1])       # :225-226
        if save:
            steps.append(dict(t=t, cw_prev=cw, cw=cw2, pre_w=pre_w, inv_l=inv_l, ca_prev=ca, ca=ca2, pre_a=pre_a))
        cw, ca = cw2, ca2
    ha = ha_all[S]
    out = ops.atanh_relu_fwd(ha)                                                    # :239-241
    out = ops.cast(out.view(N, hh, ww, D), e4.dtype)
    ctx = None
    if save:
        ctx = dict(steps=steps, ids=ids_dev, vis=vis, inv_v=inv_v, ha=ha, shape=(N, hh, ww, D), in_dtype=e4.dtype,
                   e_all=e_all, lang_all=lang_all, hw_all=hw_all, ha_all=ha_all)
    return out, ctx


def text_fusion_bwd(ops, store, g_out, ctx, prefix=_PRE):
    """Returns g_e4 [N,h,w,D]; accumulates gradients of embedding / both LSTM kernels / biases."""
    N, hh, ww, D = ctx["shape"]
    P = hh * ww
    R = N * P
    f32 = torch.float32
    n_emb, n_kw, n_bw, n_ka, n_ba = _names(prefix)
    kw, ka = store.p[n_kw], store.p[n_ka]
    dkw, dbw, dka, dba, demb = store.g[n_kw], store.g[n_bw], store.g[n_ka], store.g[n_ba], store.g[n_emb]
    steps = ctx["steps"]
    S = len(steps)
    if S == 0:                # all-pad batch: output is relu(0) = 0, no gradient reaches e4
        return ops.cast(ops.zeros_f32((N, hh, ww, D)), ctx["in_dtype"])
    e_all, lang_all, hw_all, ha_all = ctx["e_all"], ctx["lang_all"], ctx["hw_all"], ctx["ha_all"]
    ids = ctx["ids"]

    g_ha = ops.atanh_relu_bwd(ops.cast(g_out, f32).view(R, D), ctx["ha"])
    g_ca = ops.zeros_f32((R, D))
    g_hw = ops.zeros_f32((N, D))
    g_cw = ops.zeros_f32((N, D))
    g_gv = ops.zeros_f32((R, 4 * D))
    gpa_all = ops.zeros_f32((S, R, 4 * D))       # gate gradients of every step, operands of the batched weight gradients
    gr_all = ops.zeros_f32((S, N, 4 * D))
    gpw_all = ops.zeros_f32((S, N, 4 * D))
    for j in range(S - 1, -1, -1):
        s = steps[j]
        t = s["t"]
        # ---- mLSTM cell
        g_pre_a, g_ca, g_ha_pass = ops.lstm_cell_bwd(g_ca, g_ha, s["pre_a"], s["ca_prev"], s["ca"], ids, t, P, out_gpre=gpa_all[j])
        g_ha = ops.conv_dgrad(_op(ops, g_pre_a), _mat(ka[3 * D:4 * D]), 0, D, out_dtype=f32).view(R, D)
        ops.add_(g_ha, g_ha_pass)
        ops.add_(g_gv, g_pre_a)
        g_r = ops.rows_group_sum(g_pre_a, P, out=gr_all[j])                          # [N,4D]
        g_lang = ops.conv_dgrad(_op(ops, g_r), _mat(ka[D:3 * D]), D, D, out_dtype=f32).view(N, D)
        ops.add_(g_hw, ops.l2norm_rows_bwd(g_lang, lang_all[j], s["inv_l"]))
        # ---- word LSTM cell
        g_pre_w, g_cw, g_hw_pass = ops.lstm_cell_bwd(g_cw, g_hw, s["pre_w"], s["cw_prev"], s["cw"], ids, t, 1, out_gpre=gpw_all[j])
        g_hw = ops.conv_dgrad(_op(ops, g_pre_w), _mat(kw), D, D, out_dtype=f32).view(N, D)
        ops.add_(g_hw, g_hw_pass)
    # ---- weight gradients: one product per kernel block over all steps (rows = step x sample [x position])
    gpa, gr, gpw = _op(ops, gpa_all.view(S * R, 4 * D)), _op(ops, gr_all.view(S * N, 4 * D)), _op(ops, gpw_all.view(S * N, 4 * D))
    e_rows, lang_rows = _op(ops, e_all[:S].view(S * N, D)), _op(ops, lang_all[:S].view(S * N, D))
    ops.conv_wgrad([(_op(ops, ha_all[:S].view(S * R, D)), False)], gpa, _mat(dka[3 * D:4 * D]), dba)
    ops.conv_wgrad([(e_rows, False), (lang_rows, False)], gr, _mat(dka[D:3 * D]), None)
    ops.conv_wgrad([(e_rows, False), (_op(ops, hw_all[:S].view(S * N, D)), False)], gpw, _mat(dkw), dbw)
    # ---- embedding rows: d/d(e_t) through both LSTMs, all steps at once, then one scatter-add per step
    g_e = ops.conv_dgrad(gr, _mat(ka[D:3 * D]), 0, D, out_dtype=f32)
    ops.conv_dgrad(gpw, _mat(kw), 0, D, out=g_e, acc=True)
    g_e = g_e.view(S, N, D)
    for j in range(S):
        ops.embedding_bwd(g_e[j], ids, steps[j]["t"], demb)
    g_gv_op = _op(ops, g_gv)
    ops.conv_wgrad([(_op(ops, ctx["vis"]), False)], g_gv_op, _mat(dka[0:D]), None)
    g_vis = ops.conv_dgrad(g_gv_op, _mat(ka[0:D]), 0, D, out_dtype=f32).view(R, D)
    g_e4 = ops.l2norm_rows_bwd(g_vis, ctx["vis"], ctx["inv_v"])
    return ops.cast(g_e4.view(N, hh, ww, D), ctx["in_dtype"])
