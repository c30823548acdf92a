"""Caption encoder: word LSTM + multimodal ("convolutional") LSTM fused with the 6x6 bottleneck.

Reference: models_collection.encode_feat_with_text (:150-248).  The reference unrolls B x T copies of
two BasicLSTMCells in the TF graph and runs each sample separately; here all samples advance together,
pad tokens (id 0, tf.cond :235) are handled with a per-sample mask, and everything that does not depend on a
recurrence is hoisted out of the time loop (SURVEY 7.2 "mLSTM algebra"):

  word LSTM     [e_t, h_w] @ K_w = e_t @ K_w[0:D]  (all T steps: ONE [T*N, D] x [D, 4D] product, bias folded in)
                                 + h_w @ K_w[D:2D] (the recurrence: ONE persistent launch for the whole sequence,
                                                    ops.lstm_seq_fwd -- fp32 GEMV + gates fused, weights resident in
                                                    shared memory, a grid barrier per step)
  mLSTM         [vis, tile(e), tile(l2n(h_w)), h_a] @ K_a
                    = vis @ K_a[0:D] + b_a                     (once per forward)
                    + tile([e_t, l2n(h_w(t))] @ K_a[D:3D])     (all T steps: ONE [T*N, 2D] x [2D, 4D] product -- the word LSTM
                                                                does not depend on the mLSTM, so it runs to the end first)
                    + h_a @ K_a[3D:4D]                         (the only per-step [N*P, D] x [D, 4D] product)

All state is fp32 (LSTM state, gates, l2-norms); the hoisted matrices are multiplied by the conv kernels with k=1.  BPTT is
written out by hand and mirrors the forward: the mLSTM steps backwards first (one input-gradient product per step), its
row-term gradients of all steps go through K_a[D:3D] and the l2-normalisation in one product / one pass, and the word LSTM's
BPTT is again one persistent launch (ops.lstm_seq_bwd).  Every LSTM kernel block gets ONE weight-gradient product over all
steps, and the embedding gradient is one scatter-add.
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


def _word_lstm_fwd(ops, gx, kh, ids, T, N, D):
    """(hw_all [T+1,N,D], cw_all [T+1,N,D], pre_w [T,N,4D]) -- the persistent kernel, or a cell at a time where the operator
    set does not take the size (hidden sizes beyond 512: the background generator)."""
    if ops.lstm_seq_supported(N, D):
        return ops.lstm_seq_fwd(gx, kh, ids)
    f32 = torch.float32
    hw_all, cw_all = ops.zeros_f32((T + 1, N, D)), ops.zeros_f32((T + 1, N, D))
    pre_w = ops.zeros_f32((T, N, 4 * D))
    for t in range(T):
        gh = ops.conv_fwd([(_op(ops, hw_all[t]), False)], _mat(kh), None, out_dtype=f32).view(N, 4 * D)
        c2, _, pre = ops.lstm_cell_fwd(gh, gx[t], None, cw_all[t], hw_all[t], ids, t, 1, out_h=hw_all[t + 1])
        cw_all[t + 1].copy_(c2)
        pre_w[t].copy_(pre)
    return hw_all, cw_all, pre_w


def _word_lstm_bwd(ops, g_hext, pre_w, cw_all, kw, ids, T, N, D):
    """Gate gradients of every word-LSTM step [T,N,4D] given the gradient reaching each h_w(t) from the mLSTM."""
    if ops.lstm_seq_supported(N, D):
        return ops.lstm_seq_bwd(g_hext, pre_w, cw_all, kw[D:2 * D], ids)
    f32 = torch.float32
    gpw_all = ops.zeros_f32((T, N, 4 * D))
    g_hw, g_cw = ops.zeros_f32((N, D)), ops.zeros_f32((N, D))
    for t in range(T - 1, -1, -1):
        ops.add_(g_hw, g_hext[t])
        g_pre_w, g_cw, g_hw_pass = ops.lstm_cell_bwd(g_cw, g_hw, pre_w[t], cw_all[t], cw_all[t + 1], ids, t, 1, out_gpre=gpw_all[t])
        g_hw = ops.conv_dgrad(_op(ops, g_pre_w), _mat(kw), D, D, out_dtype=f32).view(N, D)
        ops.add_(g_hw, g_hw_pass)
    return gpw_all


def text_words_fwd(ops, store, ids_host, prefix=_PRE):
    """Everything of the caption encoder that does not depend on the picture (:182,211-224): embeddings, the word LSTM over all
    steps, l2n(h_w), and the mLSTM's spatially constant input term of every step.  A caller may evaluate it aside
    (ops.run_aside) while the encoder convolutions run, and hand the result to text_fusion_fwd(words=)."""
    on_device = torch.is_tensor(ids_host) and ids_host.is_cuda
    ids_np = None if on_device else np.asarray(ids_host, dtype=np.int32)
    emb, kw, bw, ka, ba = (store.p[n] for n in _names(prefix))
    D = emb.shape[1]
    N, T = ids_host.shape
    f32 = torch.float32
    # [N,T] int32 on the device; pad mask = (id == 0)
    ids_dev = ids_host.to(torch.int32).contiguous() if on_device else torch.as_tensor(ids_np, device=emb.device).contiguous()
    # steps at which at least one caption has a real token (all of them when the ids live on the device)
    ts = [t for t in range(T) if ids_np is None or (ids_np[:, t] != 0).any()]
    # ---- word LSTM (:182,211-213): embeddings and their gate products for all steps, then the recurrence in one launch
    e_all = ops.embedding_all_fwd(emb, ids_dev)                                      # [T,N,D]
    e_rows = _op(ops, e_all.view(T * N, D))
    gx = ops.conv_fwd([(e_rows, False)], _mat(kw[0:D]), bw, out_dtype=f32).view(T, N, 4 * D)
    hw_all, cw_all, pre_w = _word_lstm_fwd(ops, gx, kw[D:2 * D], ids_dev, T, N, D)
    lang_all, inv_l = ops.l2norm_rows_fwd(hw_all[1:].view(T * N, D))                 # :215-216, every step's h_w
    lang_rows = _op(ops, lang_all)
    # the mLSTM's spatially constant input term of every step (:218-224)
    r_all = ops.conv_fwd([(e_rows, False), (lang_rows, False)], _mat(ka[D:3 * D]), None, out_dtype=f32).view(T, N, 4 * D)
    return dict(ids=ids_dev, ts=ts, T=T, e_rows=e_rows, hw_all=hw_all, cw_all=cw_all, pre_w=pre_w, lang_all=lang_all, inv_l=inv_l,
                lang_rows=lang_rows, r_all=r_all)


def text_fusion_fwd(ops, store, e4, ids_host, save=True, prefix=_PRE, words=None):
    """e4: [N,h,w,D] activation; ids_host: int array [N,T] on the HOST (numpy / CPU tensor), or an int32 DEVICE tensor.
    With host ids, mLSTM time steps at which every caption is <pad> are skipped outright (the reference's tf.cond, :235);
    with device ids nothing on the host depends on the data (CUDA-graph capture): every step runs and <pad> samples
    are masked inside the cell kernels -- same result.  words: the result of text_words_fwd when the caller has evaluated it
    already.  Returns ([N,h,w,D], ctx)."""
    N, hh, ww, D = e4.shape
    P = hh * ww
    R = N * P
    if words is None:
        words = text_words_fwd(ops, store, ids_host, prefix)
    T, ts, ids_dev, r_all = words["T"], words["ts"], words["ids"], words["r_all"]
    emb, kw, bw, ka, ba = (store.p[n] for n in _names(prefix))
    f32 = torch.float32
    S = len(ts)

    e4r = ops.cast(e4, f32).view(R, D)
    vis, inv_v = ops.l2norm_rows_fwd(e4r)                                           # :201-202
    gv = ops.conv_fwd([(_op(ops, vis), False)], _mat(ka[0:D]), ba, out_dtype=f32).view(R, 4 * D)
    # ---- mLSTM recurrence over the executed steps: slot j of ha_all holds the INPUT state of step ts[j] (slot 0 = zeros, :204)
    ha_all = ops.zeros_f32((S + 1, R, D))
    ca = ops.zeros_f32((R, D))
    steps = []
    for j, t in enumerate(ts):
        ha = ha_all[j]
        ga = ops.conv_fwd([(_op(ops, ha), False)], _mat(ka[3 * D:4 * D]), None, out_dtype=f32).view(R, 4 * D)
        ca2, ha2, pre_a = ops.lstm_cell_fwd(ga, gv, r_all[t], ca, ha, ids_dev, t, P, out_h=ha_all[j + 1])       # :225-226
        if save:
            steps.append(dict(t=t, ca_prev=ca, ca=ca2, pre_a=pre_a))
        ca = ca2
    ha = ha_all[S]
    out = ops.atanh_relu_fwd(ha)                                                    # :239-241
    out = ops.cast(out.view(N, hh, ww, D), e4.dtype)
    ctx = None
    if save:
        ctx = dict(steps=steps, ids=ids_dev, vis=vis, inv_v=inv_v, ha=ha, shape=(N, hh, ww, D), in_dtype=e4.dtype, T=T,
                   e_rows=words["e_rows"], lang_all=words["lang_all"], lang_rows=words["lang_rows"], inv_l=words["inv_l"],
                   hw_all=words["hw_all"], cw_all=words["cw_all"], pre_w=words["pre_w"], ha_all=ha_all)
    return out, ctx


def text_fusion_bwd(ops, store, g_out, ctx, prefix=_PRE, aside=False):
    """Returns g_e4 [N,h,w,D]; accumulates gradients of embedding / both LSTM kernels / biases.
    aside=True: returns (g_e4, join) instead -- only the chain that ends in g_e4 is on the caller's critical path; the word
    LSTM's BPTT, the batched weight gradients and the embedding scatter are handed to ops.run_aside and join() must be called
    before the gradients are consumed (they then overlap with the encoder's backward pass)."""
    N, hh, ww, D = ctx["shape"]
    P = hh * ww
    R = N * P
    T = ctx["T"]
    f32 = torch.float32
    n_emb, n_kw, n_bw, n_ka, n_ba = _names(prefix)
    kw, ka = store.p[n_kw], store.p[n_ka]
    dkw, dbw, dka, dba, demb = store.g[n_kw], store.g[n_bw], store.g[n_ka], store.g[n_ba], store.g[n_emb]
    steps = ctx["steps"]
    S = len(steps)
    if S == 0:                # all-pad batch: output is relu(0) = 0, no gradient reaches e4
        z = ops.cast(ops.zeros_f32((N, hh, ww, D)), ctx["in_dtype"])
        return (z, (lambda: None)) if aside else z
    ha_all, ids = ctx["ha_all"], ctx["ids"]

    # ---- mLSTM, backwards through the executed steps
    g_ha = ops.atanh_relu_bwd(ops.cast(g_out, f32).view(R, D), ctx["ha"])
    g_ca = ops.zeros_f32((R, D))
    g_gv = ops.zeros_f32((R, 4 * D))
    gpa_all = ops.zeros_f32((S, R, 4 * D))       # gate gradients of every step, operands of the batched weight gradients
    gr_all = ops.zeros_f32((T, N, 4 * D))        # row-term gradients, indexed by time step (zero where a step did not run)
    for j in range(S - 1, -1, -1):
        s = steps[j]
        t = s["t"]
        g_pre_a, g_ca, g_ha_pass = ops.lstm_cell_bwd(g_ca, g_ha, s["pre_a"], s["ca_prev"], s["ca"], ids, t, P, out_gpre=gpa_all[j])
        g_ha = ops.conv_dgrad(_op(ops, g_pre_a), _mat(ka[3 * D:4 * D]), 0, D, out_dtype=f32).view(R, D)
        ops.add_(g_ha, g_ha_pass)
        ops.add_(g_gv, g_pre_a)
        ops.rows_group_sum(g_pre_a, P, out=gr_all[t])                                # [N,4D]
    g_gv_op = _op(ops, g_gv)

    def off_path():
        # ---- all steps at once: row terms -> l2n(h_w(t)) -> h_w(t); then the word LSTM's BPTT in one launch
        gr = _op(ops, gr_all.view(T * N, 4 * D))
        g_lang = ops.conv_dgrad(gr, _mat(ka[D:3 * D]), D, D, out_dtype=f32).view(T * N, D)
        g_hext = ops.l2norm_rows_bwd(g_lang, ctx["lang_all"], ctx["inv_l"]).view(T, N, D)
        gpw_all = _word_lstm_bwd(ops, g_hext, ctx["pre_w"], ctx["cw_all"], kw, ids, T, N, D)
        # ---- weight gradients: one product per kernel block over all steps (rows = step x sample [x position])
        gpa, gpw = _op(ops, gpa_all.view(S * R, 4 * D)), _op(ops, gpw_all.view(T * N, 4 * D))
        e_rows, lang_rows = ctx["e_rows"], ctx["lang_rows"]
        ops.conv_wgrad([(_op(ops, ha_all[:S].view(S * R, D)), False)], gpa, _mat(dka[3 * D:4 * D]), dba)
        ops.conv_wgrad([(e_rows, False), (lang_rows, False)], gr, _mat(dka[D:3 * D]), None)
        ops.conv_wgrad([(e_rows, False), (_op(ops, ctx["hw_all"][:T].view(T * N, D)), False)], gpw, _mat(dkw), dbw)
        # ---- embedding rows: d/d(e_t) through both LSTMs, all steps at once, then one scatter-add
        g_e = ops.conv_dgrad(gr, _mat(ka[D:3 * D]), 0, D, out_dtype=f32)
        ops.conv_dgrad(gpw, _mat(kw), 0, D, out=g_e, acc=True)
        ops.embedding_all_bwd(g_e.view(T, N, D), ids, demb)
        ops.conv_wgrad([(_op(ops, ctx["vis"]), False)], g_gv_op, _mat(dka[0:D]), None)

    join = lambda: None
    if aside:
        _, join_side = ops.run_aside(off_path)
        # the side work reads buffers allocated on the caller's stream: they must not return to its allocator (and be handed
        # to the encoder's backward pass) before the join
        keep = [gr_all, gpa_all, g_gv, g_gv_op]

        def join():
            join_side()
            keep.clear()
    # ---- the caller's critical path: gate gradients -> visual rows -> e4
    g_vis = ops.conv_dgrad(g_gv_op, _mat(ka[0:D]), 0, D, out_dtype=f32).view(R, D)
    g_e4 = ops.l2norm_rows_bwd(g_vis, ctx["vis"], ctx["inv_v"])
    g_e4 = ops.cast(g_e4.view(N, hh, ww, D), ctx["in_dtype"])
    if not aside:
        off_path()
        return g_e4
    return g_e4, join
