"""Spectrally-normalised MRU discriminator with ACGAN head, forward and hand-written backward.

Reference: models_collection.discriminate_mru (:676-786).  The MRU variant ignores the sketch argument
and looks at the image only; every weight goes through sn.spectral_normed_weight (Config.sn=True,
config.py:8), activations are PReLU, there is no normaliser (set_param, :910-911).
"""
from __future__ import annotations

from . import blocks
from .params import disc_channels


class Discriminator:
    def __init__(self, ops, store, size=64):
        self.ops, self.store, self.size = ops, store, size

    def new_weight_view(self, need_wgrad=True):
        """One WeightView per training step: W_bar = W/sigma is computed once and shared by the real and
        the fake pass (both TF instantiations evaluate the same u, W)."""
        return blocks.WeightView(self.store, self.ops, sn=True, need_wgrad=need_wgrad)

    def prefetch_weights(self, wv):
        """Evaluate the spectral normalisation of every weight now (23 x 4 microsecond-sized kernels that depend on nothing but
        the parameters): a trainer can do this on a side stream while the generator's forward pass runs."""
        for s in self.store.specs:
            if s.sn and s.name.endswith("/weights") and not s.name.endswith("fully_connected/weights"):
                wv.get(s.name[:-len("/weights")])
        self._fc(wv)

    def forward(self, img, wv, save=True):
        """img NHWC [N,H,W,3] -> (patch logits [N,h,w,1], class logits [N,1,1,25], ctx)."""
        ops, st, p = self.ops, self.store, "discriminator"
        ch = disc_channels(self.size)
        X = [img]
        for _ in range(3):                                       # :693-700 (only 4 levels are consumed)
            X.append(ops.meanpool_fwd(X[-1]))
        w, b = wv.get(p + "/Conv")
        h0_raw = ops.conv_fwd([(img, False)], w, b)              # :710  7x7 s1
        h, c0 = blocks.norm_act_fwd(ops, st, p + "/Conv", h0_raw, None, "prelu")
        ectx = []
        for u in range(1, 5):
            h, c = blocks.enc_block_fwd(ops, wv, "%s/mru_conv_unit_t_%d_layer_0" % (p, u), X[u - 1], h, None, "prelu", save)
            ectx.append(c)
        hl, c_last = blocks.norm_act_fwd(ops, st, p + "/mru_conv_unit_last_norm", h, None, "prelu")
        w, b = wv.get(p + "/Conv_1")
        disc = ops.conv_fwd([(hl, False)], w, b)                 # :767  1x1 -> 1 channel
        pooled = ops.spatial_mean_fwd(hl)                        # :783
        wfc = st.p[p + "/fully_connected/weights"]
        wf, bf = self._fc(wv)
        logits = ops.conv_fwd([(pooled, False)], wf, bf)         # :784
        ctx = dict(X=X, c0=c0, ectx=ectx, c_last=c_last, hl=hl, pooled=pooled, wv=wv) if save else None
        return disc, logits, ctx

    def _fc(self, wv):
        """fully_connected weights [768,25] as a 1x1 HWIO conv (biases are 1-D here, mru.py:78)."""
        p = "discriminator/fully_connected"
        if p in wv.cache:
            return wv.cache[p]
        w = self.store.p[p + "/weights"]
        u = self.store.state[p + "/" + p + "/u"]
        wbar, sctx = self.ops.sn_fwd(w, u)
        wv.sn_ctx[p] = sctx
        wv.cache[p] = (wbar.view(1, 1, *w.shape), self.store.p[p + "/biases"])
        return wv.cache[p]

    def backward(self, g_disc, g_logits, ctx, need_x_grad, grads_ready=None):
        """Returns dL/d(img) (NHWC) if need_x_grad else None; accumulates weight grads when wv.need_wgrad.
        grads_ready(lo, hi): optional; called as soon as the flat-gradient range [lo, hi) is final (the heads and unit 4 after
        unit 4's backward pass, then unit 3, unit 2), so that a data-parallel trainer can start its all-reduce under the
        remaining units -- the early ones at 192 x 192 / 96 x 96 are most of the pass and hold almost no parameters."""
        ops, st, p = self.ops, self.store, "discriminator"
        wv = ctx["wv"]
        nw = wv.need_wgrad
        hl = ctx["hl"]
        # heads
        g_hl = None
        if g_logits is not None:
            wf, _ = self._fc(wv)
            if nw:
                fc = p + "/fully_connected"
                if fc not in wv.gwbar:
                    wv.gwbar[fc] = ops.zeros_f32(st.p[fc + "/weights"].shape)
                ops.conv_wgrad([(ctx["pooled"], False)], g_logits, wv.gwbar[fc].view(1, 1, *wv.gwbar[fc].shape),
                               st.g[fc + "/biases"])
            g_pool = ops.conv_dgrad(g_logits, wf, 0, wf.shape[2])
            g_hl = ops.spatial_mean_bwd(g_pool, hl.shape[1], hl.shape[2])
        w, _ = wv.get(p + "/Conv_1")
        if nw:
            ops.conv_wgrad([(hl, False)], g_disc, *wv.grads(p + "/Conv_1"))
        if g_hl is None:
            g_hl = ops.conv_dgrad(g_disc, w, 0, w.shape[2])
        else:
            ops.conv_dgrad(g_disc, w, 0, w.shape[2], out=g_hl, acc=True)
        g = blocks.norm_act_bwd(ops, st, p + "/mru_conv_unit_last_norm", g_hl, ctx["c_last"], None, "prelu", nw)
        X = ctx["X"]
        g_X = [None] * 4
        for u in (4, 3, 2, 1):
            scope = "%s/mru_conv_unit_t_%d_layer_0" % (p, u)
            g, g_X[u - 1] = blocks.enc_block_bwd(ops, wv, scope, g, ctx["ectx"][u - 1],
                                                 None, "prelu", need_x_grad=need_x_grad, need_ht_grad=True)
            if grads_ready is not None and nw and u >= 2:
                wv.finish_backward([scope] + ([p + "/Conv_1", p + "/fully_connected"] if u == 4 else []))
                lo = st.offsets[scope + "/norm_activation_in/prelu/param"]            # first variable of the unit
                hi = st.n_flat if u == 4 else st.offsets["%s/mru_conv_unit_t_%d_layer_0/norm_activation_in/prelu/param" % (p, u + 1)]
                grads_ready(lo, hi)
        g = blocks.norm_act_bwd(ops, st, p + "/Conv", g, ctx["c0"], None, "prelu", nw)
        w, _ = wv.get(p + "/Conv")
        if nw:
            ops.conv_wgrad([(X[0], False)], g, *wv.grads(p + "/Conv"))
        if not need_x_grad:
            return None
        # image pyramid backward (models_collection.py:693-700): X[k+1] = mean_pool(X[k])
        for k in (3, 2, 1):
            ops.add_(g_X[k - 1], ops.unpool_bwd(g_X[k]))
        ops.conv_dgrad(g, w, 0, 3, out=g_X[0], acc=True)
        return g_X[0]
