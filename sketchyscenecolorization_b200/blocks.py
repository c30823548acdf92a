"""MRU cells with hand-written backward passes (no autograd anywhere in this package).

Reference: obj_lib/mru.py -- `mru_conv_block_v3` (:353-461, encoder + discriminator cell) and
`mru_deconv_block_v2` (:527-591, decoder cell); normalisers/activations from
models_collection.py:22-34 (conditional BN), :56-65 (prelu, miu_relu).

Two exact algebraic rewrites are used:
  * the decoder's hidden state is never materialised at full resolution: convs read it through the
    `ups` flag, and the 1x1 skip conv + cBN + miu_relu run at LOW resolution (all three commute with the
    nearest-neighbour x2 upsample, batch statistics included);
  * `mean_pool(sk + h2)` is one fused pass.
"""
from __future__ import annotations

from .ops_base import ACT_LRELU, ACT_MIU, ACT_NONE


class WeightView:
    """How a network hands (possibly spectrally-normalised) weights to the cells.

    get(scope) -> (w_hwio, bias_1d); after the backward pass `gw(scope)` is the tensor wgrad accumulates into
    (the raw grad view for plain weights, a temporary dL/dW_bar for SN weights)."""

    def __init__(self, store, ops, sn=False, need_wgrad=True):
        self.store, self.ops, self.sn, self.need_wgrad = store, ops, sn, need_wgrad
        self.cache = {}
        self.sn_ctx = {}
        self.gwbar = {}

    def get(self, scope):
        if scope in self.cache:
            return self.cache[scope]
        w = self.store.p[scope + "/weights"]
        b = self.store.p[scope + "/biases"].reshape(-1)
        if self.sn:
            u = self.store.state[scope + "/" + scope + "/u"]
            wbar2d, ctx = self.ops.sn_fwd(w.reshape(-1, w.shape[-1]), u)
            self.sn_ctx[scope] = ctx
            w = wbar2d.view(w.shape)
        self.cache[scope] = (w, b)
        return w, b

    def grads(self, scope):
        """(dw, db) accumulation targets for `scope`."""
        db = self.store.g[scope + "/biases"].reshape(-1)
        if not self.sn:
            return self.store.g[scope + "/weights"], db
        if scope not in self.gwbar:
            w = self.store.p[scope + "/weights"]
            self.gwbar[scope] = self.ops.zeros_f32(w.shape)
        return self.gwbar[scope], db

    def finish_backward(self, prefixes=None):
        """Push dL/dW_bar through the spectral normalisation (sn.py) into the raw weight gradients -- of every weight, or
        (prefixes given) of the scopes under them only: a unit whose backward pass is over can be finished at once, which
        is what lets its slice of the flat gradient buffer go to the all-reduce while the earlier units still run."""
        for scope in list(self.gwbar):
            if prefixes is not None and not any(scope == q or scope.startswith(q + "/") for q in prefixes):
                continue
            gwb = self.gwbar.pop(scope)
            w = self.store.p[scope + "/weights"]
            dw = self.store.g[scope + "/weights"]
            self.ops.sn_bwd(gwb.reshape(-1, w.shape[-1]), w.reshape(-1, w.shape[-1]), self.sn_ctx[scope],
                            dw.reshape(-1, w.shape[-1]))

    def commit_u(self):
        """u <- u' (the SPECTRAL_NORM_UPDATE_OPS run as a control dependency of the G step,
        graph_single.py:178-180,208-210)."""
        for scope, ctx in self.sn_ctx.items():
            self.store.state[scope + "/" + scope + "/u"].copy_(ctx["u_new"])


# ------------------------------------------------------------------------------------------------
# norm + activation  (mru.py:367-376 `norm_activ`)
# ------------------------------------------------------------------------------------------------
def norm_act_fwd(ops, store, scope, x, labels, kind):
    if kind == "cbn":
        mean, rstd = ops.chan_stats(x)
        y = ops.cbn_act_fwd(x, mean, rstd, store.p[scope + "/scale"], store.p[scope + "/offset"], labels, ACT_MIU)
        return y, (x, mean, rstd)
    a = store.p[scope + "/prelu/param"]
    return ops.prelu_fwd(x, a), (x,)


def norm_act_bwd(ops, store, scope, gy, ctx, labels, kind, need_wgrad=True, dbias=None, acc_into=None):
    """dbias: bias-gradient accumulator of the convolution whose output was normalised here (fused column sums).
    acc_into: tensor the result is added to (and returned) -- one pass less than add_ afterwards where the operator has that form."""
    if kind != "cbn" and acc_into is not None and dbias is None:
        (x,) = ctx
        return ops.prelu_bwd(gy, x, store.p[scope + "/prelu/param"], store.g[scope + "/prelu/param"] if need_wgrad else None,
                             acc_into=acc_into)
    if kind == "cbn":
        x, mean, rstd = ctx
        g = ops.cbn_act_bwd(gy, x, mean, rstd, store.p[scope + "/scale"], store.p[scope + "/offset"], labels,
                            store.g[scope + "/scale"], store.g[scope + "/offset"], ACT_MIU, dbias=dbias)
    else:
        (x,) = ctx
        g = ops.prelu_bwd(gy, x, store.p[scope + "/prelu/param"],
                          store.g[scope + "/prelu/param"] if need_wgrad else None, dbias=dbias)
    return g if acc_into is None else ops.add_(acc_into, g)


# ------------------------------------------------------------------------------------------------
# encoder / discriminator cell: mru_conv_block_v3 (mru.py:353-461), stride 2, norm_input=True
# ------------------------------------------------------------------------------------------------
def enc_block_fwd(ops, wv, scope, x, ht, labels, kind, save=True):
    st = wv.store
    a, c_a = norm_act_fwd(ops, st, scope + "/norm_activation_in", ht, labels, kind)
    xs_ = (x, False, ops.small_patch(x, 3))      # the 3-channel image level: flattened once, read by 2 convs + 2 wgrads
    w, b = wv.get(scope + "/update_gate")
    rg_raw = ops.conv_fwd([(a, False), xs_], w, b, act=ACT_LRELU)                 # mru.py:408-414
    rg, mn, mx = ops.minmax_fwd(rg_raw)                                          # mru.py:415-416
    w, b = wv.get(scope + "/Conv")
    im = ops.conv_fwd([xs_], w, b)                                               # mru.py:419-424
    fused = kind == "prelu" and getattr(ops, "fuse_gate_prelu", True)
    if fused:                  # discriminator: gate + PReLU in one pass, the sum is recomputed by the backward pass
        p, c_p = ops.gate_prelu_fwd(ht, rg, im, st.p[scope + "/norm_activation_merge_1/prelu/param"]), None
    else:
        hp = ops.gate_fma_fwd(ht, rg, im)                                        # mru.py:426
        p, c_p = norm_act_fwd(ops, st, scope + "/norm_activation_merge_1", hp, labels, kind)
    w, b = wv.get(scope + "/Conv_1")
    ps_ = (p, False, ops.small_patch(p, 3) if p.shape[-1] < 64 else None)     # unit 1: 8 channels at full resolution
    h1_raw = ops.conv_fwd([ps_], w, b)                                           # mru.py:431-436
    h1, c_h1 = norm_act_fwd(ops, st, scope + "/Conv_1", h1_raw, labels, kind)
    w, b = wv.get(scope + "/Conv_2")
    h2 = ops.conv_fwd([(h1, False)], w, b)                                       # mru.py:437-442
    ht_p = None
    if scope + "/Conv_3/weights" in st.p:
        # mru.py:446-453,457: mean_pool(conv1x1(ht) + h2).  A 1x1 convolution (+ bias) commutes with the 2x2 mean, so the
        # skip runs on the POOLED hidden state -- a quarter of the products, and the full-resolution skip tensor (1.2 GB
        # per pass in the discriminator's first unit) is never written; it accumulates into mean_pool(h2).
        w, b = wv.get(scope + "/Conv_3")
        ht_p = ops.meanpool_fwd(ht)
        out = ops.meanpool_fwd(h2)
        ops.conv_fwd([(ht_p, False)], w, b, out=out, acc=True)
    else:
        out = ops.addpool_fwd(ht, h2)                                            # same width: the hidden state passes as it is
    ctx = None
    if save:
        ctx = dict(x=x, xs_=xs_, ps_=ps_, ht=ht, ht_p=ht_p, a=a, c_a=c_a, rg_raw=rg_raw, rg=rg, mn=mn, mx=mx, im=im, c_p=c_p, p=p,
                   c_h1=c_h1, h1=h1)
    return out, ctx


def enc_block_bwd(ops, wv, scope, g_out, ctx, labels, kind, *, need_x_grad=False, need_ht_grad=True):
    """Returns (g_ht, g_x or None).  Weight gradients accumulate through `wv.grads` when wv.need_wgrad."""
    st, nw = wv.store, wv.need_wgrad
    x, ht = ctx["x"], ctx["ht"]
    cin = ht.shape[-1]
    # d/d(h2) = up2(g_out) / 4.  With ops.phase_dgrad only Conv_2's weight gradient wants it in memory: the input gradient is
    # then evaluated from g_out directly (ops.conv_dgrad_pooled_gy: per output phase a 2x2-tap convolution of the low-resolution
    # gradient -- 2.25x fewer products; measured neutral on a B200, so it is an option)
    phase = bool(getattr(ops, "phase_dgrad", False))
    g_full = ops.unpool_bwd(g_out) if (nw or not phase) else None
    # bias gradients: column sums of the gradient at a conv output.  They come out of the kernel that produces that
    # gradient (dbias= of the norm / activation / min-max backward); for Conv_2 / Conv_3 the gradient is the un-pooled
    # g_out / 4 replicated 2x2, whose column sums equal those of the (4x smaller) g_out itself.
    # Conv_3 (1x1 skip, evaluated on the pooled hidden state -- see the forward pass; identity when the block keeps its width)
    cs = None
    if scope + "/Conv_3/weights" in st.p:
        w3, _ = wv.get(scope + "/Conv_3")
        if nw:
            dw3, db3 = wv.grads(scope + "/Conv_3")
            cs = ops.zeros_f32(db3.shape)          # Conv_2 and Conv_3 see the same output gradient: one column-sum pass for both
            ops.colsum_(g_out, cs)
            ops.add_(db3, cs)
            ops.conv_wgrad([(ctx["ht_p"], False)], g_out, dw3, None)
        g_ht = ops.unpool_bwd(ops.conv_dgrad(g_out, w3, 0, cin)) if need_ht_grad else None
    else:
        g_ht = ops.unpool_bwd(g_out) if need_ht_grad else None     # its own buffer: the branches below accumulate into it
    # Conv_2
    w2, _ = wv.get(scope + "/Conv_2")
    if nw:
        dw2, db2 = wv.grads(scope + "/Conv_2")
        if cs is not None:
            ops.add_(db2, cs)
        else:
            ops.colsum_(g_out, db2)
        ops.conv_wgrad([(ctx["h1"], False)], g_full, dw2, None)
    g_h1 = ops.conv_dgrad_pooled_gy(g_out, w2, 0, w2.shape[2]) if phase else ops.conv_dgrad(g_full, w2, 0, w2.shape[2])
    del g_full
    dw1, db1 = wv.grads(scope + "/Conv_1") if nw else (None, None)
    g_h1raw = norm_act_bwd(ops, st, scope + "/Conv_1", g_h1, ctx["c_h1"], labels, kind, nw, dbias=db1)
    del g_h1
    # Conv_1
    w1, _ = wv.get(scope + "/Conv_1")
    if nw:
        ops.conv_wgrad([ctx["ps_"]], g_h1raw, dw1, None)
    g_p = ops.conv_dgrad(g_h1raw, w1, 0, cin)
    del g_h1raw
    if kind == "prelu" and ctx["c_p"] is None:
        sc = scope + "/norm_activation_merge_1/prelu/param"
        g_rg, g_im = ops.gate_prelu_bwd(g_p, ht, ctx["rg"], ctx["im"], st.p[sc], st.g[sc] if nw else None,
                                        g_ht=g_ht if need_ht_grad else None, acc=True)
        del g_p
    else:
        g_hp = norm_act_bwd(ops, st, scope + "/norm_activation_merge_1", g_p, ctx["c_p"], labels, kind, nw)
        del g_p
        if need_ht_grad:
            ops.add_(g_ht, g_hp)
        g_rg, g_im = ops.gate_fma_bwd(g_hp, ctx["rg"], ctx["im"])
        del g_hp
    # Conv (image branch)
    wc, _ = wv.get(scope + "/Conv")
    if nw:
        ops.conv_wgrad([ctx["xs_"]], g_im, *wv.grads(scope + "/Conv"))
    g_x = ops.conv_dgrad(g_im, wc, 0, x.shape[-1]) if need_x_grad else None
    del g_im
    # update gate
    dwu, dbu = wv.grads(scope + "/update_gate") if nw else (None, None)
    g_rgraw = ops.minmax_bwd(g_rg, ctx["rg_raw"], ctx["mn"], ctx["mx"], dbias=dbu)
    del g_rg
    wu, _ = wv.get(scope + "/update_gate")
    if nw:
        ops.conv_wgrad([(ctx["a"], False), ctx["xs_"]], g_rgraw, dwu, None)
    if need_x_grad:
        ops.conv_dgrad(g_rgraw, wu, cin, x.shape[-1], out=g_x, acc=True)
    if need_ht_grad:
        g_a = ops.conv_dgrad(g_rgraw, wu, 0, cin)
        del g_rgraw
        norm_act_bwd(ops, st, scope + "/norm_activation_in", g_a, ctx["c_a"], labels, kind, nw, acc_into=g_ht)
    return g_ht, g_x


# ------------------------------------------------------------------------------------------------
# decoder cell: mru_deconv_block_v2 (mru.py:527-591), stride 2
# ------------------------------------------------------------------------------------------------
def dec_block_fwd(ops, wv, scope, xs, ht_low, cout, labels, save=True):
    """xs: list of full-resolution extra sources (sketch level first); ht_low: hidden state at half resolution."""
    st = wv.store
    chid = ht_low.shape[-1]
    # H = upsample(ht) is written once and read by both gate convolutions (and by their weight gradients): tensor-map
    # TMA, which feeds the halo-reuse conv kernel, cannot replicate pixels.  The 1x1 skip below still runs at low resolution.
    h_up = ops.upsample_fwd(ht_low)
    xsrc = [(x, False, ops.small_patch(x, 3) if x.shape[-1] < 64 else None) for x in xs]
    f = [(h_up, False)] + xsrc
    w, b = wv.get(scope + "/Conv")
    rg_raw = ops.conv_fwd(f, w, b, act=ACT_LRELU)                                # mru.py:555-559
    rg, mn0, mx0 = ops.minmax_fwd(rg_raw)
    w, b = wv.get(scope + "/Conv_1")
    zg_raw = ops.conv_fwd(f, w, b, act=ACT_LRELU)                                # mru.py:563-567
    zg, mn1, mx1 = ops.minmax_fwd(zg_raw)
    gh = ops.mul_up_fwd(rg, ht_low)                                              # rg * ht   (mru.py:572)
    w, b = wv.get(scope + "/Conv_2")
    h1_raw = ops.conv_fwd([(gh, False)] + xsrc, w, b)
    h1, c_h1 = norm_act_fwd(ops, st, scope + "/Conv_2", h1_raw, labels, "cbn")
    w, b = wv.get(scope + "/Conv_3")
    h2_raw = ops.conv_fwd([(h1, False)], w, b)                                   # mru.py:577-581
    h2, c_h2 = norm_act_fwd(ops, st, scope + "/Conv_3", h2_raw, labels, "cbn")
    c_sk = None
    if chid != cout:
        w, b = wv.get(scope + "/Conv_4")
        sk_raw = ops.conv_fwd([(ht_low, False)], w, b)                           # mru.py:585-588 at low resolution
        sk, c_sk = norm_act_fwd(ops, st, scope + "/Conv_4", sk_raw, labels, "cbn")
    else:
        sk = ht_low
    out = ops.blend_fwd(sk, h2, zg)                                              # mru.py:589
    ctx = None
    if save:
        ctx = dict(xs=xs, ht_low=ht_low, rg_raw=rg_raw, rg=rg, mn0=mn0, mx0=mx0, zg_raw=zg_raw, zg=zg, mn1=mn1,
                   mx1=mx1, gh=gh, c_h1=c_h1, h1=h1, c_h2=c_h2, h2=h2, c_sk=c_sk, sk=sk, cout=cout, h_up=h_up, xsrc=xsrc)
    return out, ctx


def dec_block_bwd(ops, wv, scope, g_out, ctx, labels, xs_need_grad):
    """Returns (g_ht_low, [g_x or None for x in xs])."""
    st = wv.store
    xs, ht_low, cout = ctx["xs"], ctx["ht_low"], ctx["cout"]
    chid = ht_low.shape[-1]
    offs, o = [], chid
    for x in xs:
        offs.append(o)
        o += x.shape[-1]
    g_sk, g_h2, g_zg = ops.blend_bwd(g_out, ctx["sk"], ctx["h2"], ctx["zg"])
    # skip path
    if chid != cout:
        dw4, db4 = wv.grads(scope + "/Conv_4")
        g_skraw = norm_act_bwd(ops, st, scope + "/Conv_4", g_sk, ctx["c_sk"], labels, "cbn", dbias=db4)
        w4, _ = wv.get(scope + "/Conv_4")
        ops.conv_wgrad([(ht_low, False)], g_skraw, dw4, None)
        g_ht = ops.conv_dgrad(g_skraw, w4, 0, chid)
        del g_skraw
    else:
        g_ht = g_sk
    # Conv_3
    dw3, db3 = wv.grads(scope + "/Conv_3")
    g_h2raw = norm_act_bwd(ops, st, scope + "/Conv_3", g_h2, ctx["c_h2"], labels, "cbn", dbias=db3)
    del g_h2
    w3, _ = wv.get(scope + "/Conv_3")
    ops.conv_wgrad([(ctx["h1"], False)], g_h2raw, dw3, None)
    g_h1 = ops.conv_dgrad(g_h2raw, w3, 0, cout)
    del g_h2raw
    # Conv_2
    dw2, db2 = wv.grads(scope + "/Conv_2")
    g_h1raw = norm_act_bwd(ops, st, scope + "/Conv_2", g_h1, ctx["c_h1"], labels, "cbn", dbias=db2)
    del g_h1
    w2, _ = wv.get(scope + "/Conv_2")
    ops.conv_wgrad([(ctx["gh"], False)] + ctx["xsrc"], g_h1raw, dw2, None)
    g_gh = ops.conv_dgrad(g_h1raw, w2, 0, chid)
    g_xs = [ops.conv_dgrad(g_h1raw, w2, offs[i], xs[i].shape[-1]) if xs_need_grad[i] else None
            for i in range(len(xs))]
    del g_h1raw
    g_rg, g_ht_mul = ops.mul_up_bwd(g_gh, ctx["rg"], ht_low)
    del g_gh
    ops.add_(g_ht, g_ht_mul)
    del g_ht_mul
    # gates: the weight gradients read the upsampled hidden state kept by the forward pass
    f_w = [(ctx["h_up"], False)] + ctx["xsrc"]
    for (gg, raw, mn, mx, sc) in ((g_rg, ctx["rg_raw"], ctx["mn0"], ctx["mx0"], scope + "/Conv"),
                                  (g_zg, ctx["zg_raw"], ctx["mn1"], ctx["mx1"], scope + "/Conv_1")):
        dwg, dbg = wv.grads(sc)
        g_raw = ops.minmax_bwd(gg, raw, mn, mx, dbias=dbg)
        w, _ = wv.get(sc)
        ops.conv_wgrad(f_w, g_raw, dwg, None)
        ops.conv_dgrad(g_raw, w, 0, chid, ups=True, out=g_ht, acc=True)
        for i in range(len(xs)):
            if xs_need_grad[i]:
                ops.conv_dgrad(g_raw, w, offs[i], xs[i].shape[-1], out=g_xs[i], acc=True)
        del g_raw
    del f_w
    return g_ht, g_xs
