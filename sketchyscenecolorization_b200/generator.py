"""Text-conditioned MRU U-Net generator, forward and hand-written backward.

Reference: models_collection.generate_mru (:251-377) = image_encoder_mru (:68-147) +
encode_feat_with_text (:150-248) + noise FC (:310-316) + five mru_deconv units (:323-371) +
7x7 tanh head (:372-374).  Inputs/outputs at this boundary use the reference's NCHW fp32 layout;
everything in between is NHWC in `ops.act_dtype`.
"""
from __future__ import annotations

import torch

from . import blocks, text_fusion
from .ops_base import ACT_MIU, ACT_NONE, ACT_TANH
from .params import decoder_plan, encoder_channels


class Generator:
    def __init__(self, ops, store, size=64, lstm_hybrid=True):
        self.ops, self.store, self.size, self.lstm_hybrid = ops, store, size, lstm_hybrid

    # ------------------------------------------------------------------
    def forward(self, sketch_nchw, text_ids_host, labels, noise, save=True):
        """sketch [N,3,H,W] fp32 in [-1,1]; text ids [N,15] (host ints); labels int32 [N] (device);
        noise [N,256] fp32 (the reference draws it inside the graph, models_collection.py:310; it is an explicit
        input here so that outputs are reproducible).  Returns (image NHWC [N,H,W,3] act dtype in [-1,1], ctx)."""
        ops, st, p = self.ops, self.store, "generator"
        wv = blocks.WeightView(st, ops, sn=False)
        ch = encoder_channels(self.size)
        N = sketch_nchw.shape[0]
        # the caption's word LSTM does not depend on the picture: aside, under the encoder convolutions
        words, join_words = (ops.run_aside(lambda: text_fusion.text_words_fwd(ops, st, text_ids_host)) if self.lstm_hybrid
                             else (None, None))
        s0 = ops.nchw_to_nhwc(sketch_nchw)                         # [N,H,W,3]
        # sketch pyramid: cascaded 2x2 means (encoder, :84-86) == AREA resize (decoder, :268-272) at 2^k factors
        S = [s0]
        for _ in range(5):
            S.append(ops.meanpool_fwd(S[-1]))
        w, b = wv.get(p + "/Conv")
        h0 = ops.conv_fwd([(s0, False)], w, b, stride=2)           # :93-97  7x7 s2, no norm / activation
        enc, ectx = [h0], []
        ht = h0
        for u in range(1, 5):
            ht, c = blocks.enc_block_fwd(ops, wv, "%s/mru_conv_unit_t_%d_layer_0" % (p, u), S[u], ht, labels, "cbn", save)
            if u == 4:                                             # mru.py:651-653
                ht, c_last = blocks.norm_act_fwd(ops, st, p + "/mru_conv_unit_last_norm", ht, labels, "cbn")
            enc.append(ht)
            ectx.append(c)
        tctx = None
        if self.lstm_hybrid:
            join_words()
            feat, tctx = text_fusion.text_fusion_fwd(ops, st, enc[4], text_ids_host, save, words=words)   # :298
        else:
            feat = enc[4]
        # noise FC (:310-316): [N,256] -> miu_relu -> reshape NCHW [N,C/8,2h,2w]
        nc = ch[4] // 8
        nh, nw = enc[4].shape[1] * 2, enc[4].shape[2] * 2
        wfc = st.p[p + "/fully_connected/weights"]
        bfc = st.p[p + "/fully_connected/biases"]
        nz = noise.view(N, 1, 1, noise.shape[1])
        fc = ops.conv_fwd([(nz, False)], wfc.view(1, 1, *wfc.shape), bfc, act=ACT_MIU, out_dtype=torch.float32)
        nzf = ops.nchw_to_nhwc(fc.view(N, nc, nh, nw))             # [N,2h,2w,C/8] act dtype
        extras = {0: [S[4], nzf], 2: [S[3], enc[2]], 4: [S[2], enc[1]], 6: [S[1], enc[0]], 8: [S[0]]}
        dctx = []
        ht = feat
        for (u, cx, chid, cout) in decoder_plan(self.size):
            ht, c = blocks.dec_block_fwd(ops, wv, "%s/mru_deconv_unit_t_%d_layer_0" % (p, u), extras[u], ht, cout,
                                         labels, save)
            dctx.append(c)
        w, b = wv.get(p + "/Conv_1")
        out = ops.conv_fwd([(ht, False)], w, b, act=ACT_TANH)      # :372-374
        ctx = None
        if save:
            ctx = dict(wv=wv, s0=s0, ectx=ectx, c_last=c_last, tctx=tctx, nz=nz, fc=fc, dctx=dctx, ht_last=ht, out=out,
                       labels=labels, N=N, nc=nc, nh=nh, nw=nw)
        return out, ctx

    # ------------------------------------------------------------------
    def backward(self, g_out, ctx, grads_ready=None):
        """g_out: dL/d(image) NHWC.  Accumulates all generator weight gradients into store.grad.
        grads_ready(lo, hi): optional; called when the flat-gradient range [lo, hi) is final -- noise FC + decoder + head
        (70 % of the parameters) after the decoder, the caption encoder after its BPTT, encoder unit 4 after its backward pass."""
        ops, st, p = self.ops, self.store, "generator"
        wv, labels = ctx["wv"], ctx["labels"]
        plan = decoder_plan(self.size)
        # head
        g = ops.act_bwd(g_out, ctx["out"], ACT_TANH)
        w, _ = wv.get(p + "/Conv_1")
        # 3-channel gradient under a 7x7 filter: its flattened (tap, channel) patch tensors let both products run on the
        # TMA-fed kernels (K = 147 instead of gathering the 3 channels 49 times)
        ops.conv_wgrad([(ctx["ht_last"], False)], g, *wv.grads(p + "/Conv_1"), gy_patch=ops.small_patch(g, w.shape[0]))
        g_ht = ops.conv_dgrad(g, w, 0, w.shape[2], gy_patch=ops.small_patch(g, w.shape[0], mirror=True))
        del g
        g_enc = {}          # gradients flowing into encoder outputs through the skip connections
        g_nz = None
        for i in reversed(range(len(plan))):
            u, cx, chid, cout = plan[i]
            need = [False] + ([True] if cx else [])
            g_ht, g_xs = blocks.dec_block_bwd(ops, wv, "%s/mru_deconv_unit_t_%d_layer_0" % (p, u), g_ht, ctx["dctx"][i],
                                              labels, need)
            if cx:
                if u == 0:
                    g_nz = g_xs[1]
                else:
                    g_enc[{2: 2, 4: 1, 6: 0}[u]] = g_xs[1]
        # noise FC
        N, nc, nh, nw = ctx["N"], ctx["nc"], ctx["nh"], ctx["nw"]
        g_fc = ops.nhwc_to_nchw(g_nz, out_dtype=torch.float32).view(N, 1, 1, nc * nh * nw)
        g_fc = ops.act_bwd(g_fc, ctx["fc"], ACT_MIU)
        wfc = st.p[p + "/fully_connected/weights"]
        ops.conv_wgrad([(ctx["nz"], False)], g_fc, st.g[p + "/fully_connected/weights"].view(1, 1, *wfc.shape),
                       st.g[p + "/fully_connected/biases"])
        del g_fc
        off = st.offsets
        if grads_ready is not None:
            grads_ready(off[p + "/fully_connected/weights"], st.n_flat)
        # text fusion: its weight gradients and the word LSTM's BPTT run aside, under the encoder's backward pass
        g_e4, join_text = text_fusion.text_fusion_bwd(ops, st, g_ht, ctx["tctx"], aside=True) if self.lstm_hybrid else (g_ht, None)
        # encoder
        g_cur = blocks.norm_act_bwd(ops, st, p + "/mru_conv_unit_last_norm", g_e4, ctx["c_last"], labels, "cbn")
        for u in (4, 3, 2, 1):
            if u != 4 and u in g_enc:
                ops.add_(g_cur, g_enc[u])
            g_cur, _ = blocks.enc_block_bwd(ops, wv, "%s/mru_conv_unit_t_%d_layer_0" % (p, u), g_cur, ctx["ectx"][u - 1],
                                            labels, "cbn", need_x_grad=False, need_ht_grad=True)
            if grads_ready is not None and u == 4:
                first = "%s/mru_conv_unit_t_4_layer_0/norm_activation_in/offset" % p
                grads_ready(off[first], off.get(p + "/TextLSTM/embedding", off[p + "/fully_connected/weights"]))
        ops.add_(g_cur, g_enc[0])
        ops.conv_wgrad([(ctx["s0"], False)], g_cur, *wv.grads(p + "/Conv"), stride=2)
        if join_text is not None:
            join_text()
            if grads_ready is not None:
                grads_ready(off[p + "/TextLSTM/embedding"], off[p + "/fully_connected/weights"])
