"""`--block_type Residual`: bottleneck-residual U-Net generator and discriminator, forward and hand-written backward.

Reference: residual_util.py (conv :16-25, conv_ex :28-34, batchnorm :56-68, deconv :71-80, bottleneck_residual_en / _de / _pu
:83-175); models_collection.image_encoder_residual (:541-576), generate_residual (:579-672), discriminate_residual (:844-893).

Every unit is convolution -> batch norm -> (lrelu | relu | nothing); a block is three such units plus a shortcut.  The
convolutions are the stride-1 SAME tensor-core kernels of the MRU path: 7x7 stride 2 / 3x3 / 1x1 directly, the 4x4 ones in
phase form (pix2pix.py: stride 2 with pad 1 through space_to_depth, conv2d_transpose through depth_to_space, 4x4 SAME stride 1
as a 5x5 SAME filter whose first row and column are zero -- TF puts the odd SAME pad pixel at the bottom / right).
"""
from __future__ import annotations

import torch

from . import blocks, text_fusion
from .ops_base import ACT_MIU, ACT_NONE, ACT_TANH
from .pix2pix import _Filters, _slopes, bn_bwd, bn_fwd

UNITS = [3, 4, 6, 3]                       # models_collection.py:609


class Layers:
    """conv -> batch norm -> activation units over one parameter store, with their backward passes."""

    def __init__(self, ops, store, need_wgrad=True):
        self.ops, self.store, self.need_wgrad = ops, store, need_wgrad
        self.fl = _Filters(store, ops, need_wgrad)
        self.a_lrelu, self.a_relu = _slopes(store)
        self.zeros = None

    def begin(self, N, device):
        self.zeros = torch.zeros(N, dtype=torch.int32, device=device)

    # ---- convolutions.  kind: 'c7s2' | 'c3' | 'c1' (direct SAME), 'c4s2' | 'c4' | 'deconv' (phase form)
    def conv_fwd(self, kind, scope, srcs, xs=None):
        ops, fl = self.ops, self.fl
        if kind == "c4s2":
            xs = ops.space_to_depth(srcs[0]) if xs is None else xs
            return ops.conv_fwd([(xs, False)], fl.get(scope + "/conv/filter", "conv"), None), xs
        if kind == "deconv":
            y3 = ops.conv_fwd([(s, False) for s in srcs], fl.get(scope + "/deconv/filter", "deconv"), None)
            return ops.depth_to_space(y3), srcs
        if kind == "c4":
            return ops.conv_fwd([(srcs[0], False)], fl.get(scope + "/conv_ex/filter", "k5"), None), srcs[0]
        w = self.store.p[scope + "/conv_ex/filter"]
        return ops.conv_fwd([(srcs[0], False)], w, None, stride=2 if kind == "c7s2" else 1), srcs[0]

    def conv_bwd(self, kind, scope, gy, c, need_x=True):
        """-> list of source gradients (one entry except for 'deconv')."""
        ops, fl, nw = self.ops, self.fl, self.need_wgrad
        if kind == "c4s2":
            name = scope + "/conv/filter"
            w = fl.get(name, "conv")
            if nw:
                ops.conv_wgrad([(c, False)], gy, fl.grad(name), None)
            return [ops.depth_to_space(ops.conv_dgrad(gy, w, 0, w.shape[2])) if need_x else None]
        if kind == "deconv":
            name = scope + "/deconv/filter"
            w = fl.get(name, "deconv")
            g3 = ops.space_to_depth(gy)
            if nw:
                ops.conv_wgrad([(s, False) for s in c], g3, fl.grad(name), None)
            out, off = [], 0
            for s in c:
                out.append(ops.conv_dgrad(g3, w, off, s.shape[3]) if need_x else None)
                off += s.shape[3]
            return out
        if kind == "c4":
            name = scope + "/conv_ex/filter"
            w = fl.get(name, "k5")
            if nw:
                ops.conv_wgrad([(c, False)], gy, fl.grad(name), None)
            return [ops.conv_dgrad(gy, w, 0, w.shape[2]) if need_x else None]
        name = scope + "/conv_ex/filter"
        w = self.store.p[name]
        if nw:
            ops.conv_wgrad([(c, False)], gy, self.store.g[name], None, stride=2 if kind == "c7s2" else 1)
        if not need_x:
            return [None]
        assert kind != "c7s2", "the 7x7 stride-2 stem only ever sees the sketch: no input gradient"
        return [ops.conv_dgrad(gy, w, 0, w.shape[2])]

    # ---- batch norm with table gradients optional (the G step does not want the discriminator's)
    def bn_fwd(self, scope, x):
        return bn_fwd(self.ops, self.store, scope, x, self.zeros)

    def bn_bwd(self, scope, gy, c):
        if self.need_wgrad:
            return bn_bwd(self.ops, self.store, scope, gy, c, self.zeros)
        ops, st = self.ops, self.store
        x, mean, rstd = c
        scale, offset = st.p[scope + "/scale"].view(1, -1), st.p[scope + "/offset"].view(1, -1)
        return ops.cbn_act_bwd(gy, x, mean, rstd, scale, offset, self.zeros, ops.zeros_f32(scale.shape), ops.zeros_f32(offset.shape),
                               ACT_NONE)

    def _slope(self, act):
        return self.a_lrelu if act == "lrelu" else self.a_relu

    # ---- conv -> bn -> act
    def unit_fwd(self, kind, scope, srcs, act, xs=None, bn_scope=None):
        y, cc = self.conv_fwd(kind, scope, srcs, xs)
        yb, bc = self.bn_fwd(bn_scope or scope + "/batchnorm", y)
        out = yb if act is None else self.ops.prelu_fwd(yb, self._slope(act))
        return out, dict(kind=kind, scope=scope, cc=cc, bc=bc, yb=yb, act=act, bn_scope=bn_scope or scope + "/batchnorm")

    def unit_bwd(self, g, c, need_x=True):
        if c["act"] is not None:
            g = self.ops.prelu_bwd(g, c["yb"], self._slope(c["act"]), None)
        g = self.bn_bwd(c["bn_scope"], g, c["bc"])
        return self.conv_bwd(c["kind"], c["scope"], g, c["cc"], need_x)

    # ---- blocks (residual_util.py:83-175)
    def block_fwd(self, kind, scope, srcs, act):
        """kind 'en' (stride 2), 'de' (x2 up, `srcs` = the concatenated inputs), 'pu' (same size, identity shortcut)."""
        ops = self.ops
        first = {"en": "c4s2", "de": "deconv", "pu": "c4"}[kind]
        h, c1 = self.unit_fwd(first, scope + "/block_1", srcs, act)
        h, c2 = self.unit_fwd("c3", scope + "/block_2", [h], act)
        h, c3 = self.unit_fwd("c1", scope + "/block_3", [h], None)
        ca = None
        if kind == "pu":
            ops.add_(h, srcs[0])
        else:           # the shortcut convolves the same input: share its space_to_depth copy
            o, ca = self.unit_fwd(first, scope + "/block_add", srcs, None, xs=c1["cc"] if kind == "en" else None)
            ops.add_(h, o)
        out = ops.prelu_fwd(h, self._slope(act))
        return out, dict(kind=kind, c1=c1, c2=c2, c3=c3, ca=ca, s=h, act=act)

    def block_bwd(self, g, c, need_x=True):
        """-> list of source gradients."""
        ops = self.ops
        g_s = ops.prelu_bwd(g, c["s"], self._slope(c["act"]), None)
        g_h = self.unit_bwd(g_s, c["c3"])[0]
        g_h = self.unit_bwd(g_h, c["c2"])[0]
        g_x = self.unit_bwd(g_h, c["c1"], need_x)
        if c["kind"] == "pu":
            ops.add_(g_x[0], g_s)
            return g_x
        g_o = self.unit_bwd(g_s, c["ca"], need_x)
        if need_x:
            for a, b in zip(g_x, g_o):
                ops.add_(a, b)
        return g_x

    def finish_backward(self):
        if self.need_wgrad:
            self.fl.finish_backward()


# ------------------------------------------------------------------------------------------------
# generator
# ------------------------------------------------------------------------------------------------
class ResidualGenerator:
    def __init__(self, ops, store, size=64, lstm_hybrid=True):
        self.ops, self.store, self.size, self.lstm_hybrid = ops, store, size, lstm_hybrid

    def forward(self, sketch_nchw, text_ids_host, labels, noise, save=True):
        """Same boundary as generator.Generator.forward; `labels` is unused (plain batch norm), as in the reference."""
        ops, st, p = self.ops, self.store, "generator"
        L = Layers(ops, st)
        N = sketch_nchw.shape[0]
        L.begin(N, sketch_nchw.device)
        s0 = ops.nchw_to_nhwc(sketch_nchw)
        h, c_stem = L.unit_fwd("c7s2", p + "/encoder_1", [s0], "lrelu", bn_scope=p + "/encoder_1")       # :555-559
        z, ectx = [h], []
        for lvl in range(4):                                                                              # :568-574
            h, c = L.block_fwd("en", "%s/encoder_%d_0" % (p, lvl + 2), [z[-1]], "lrelu")
            cs = [c]
            for u in range(1, UNITS[lvl]):
                h, c = L.block_fwd("pu", "%s/encoder_%d_%d" % (p, lvl + 2, u), [h], "lrelu")
                cs.append(c)
            z.append(h)
            ectx.append(cs)
        tctx = None
        if self.lstm_hybrid:
            feat, tctx = text_fusion.text_fusion_fwd(ops, st, z[4], text_ids_host, save)                 # :620-625
        else:
            feat = z[4]
        nc, nh, nw = z[4].shape[3] // 8, z[4].shape[1], z[4].shape[2]                                    # :627-634
        wfc, bfc = st.p[p + "/fully_connected/weights"], st.p[p + "/fully_connected/biases"]
        nz = noise.view(N, 1, 1, noise.shape[1])
        fc = ops.conv_fwd([(nz, False)], wfc.view(1, 1, *wfc.shape), bfc, act=ACT_MIU, out_dtype=torch.float32)
        nzf = ops.nchw_to_nhwc(fc.view(N, nc, nh, nw))
        dctx = []
        d = None
        for i in range(4):                                                                                # :644-658
            skip = 4 - i
            srcs = [feat, nzf] if i == 0 else [d, z[skip]]
            d, c = L.block_fwd("de", "%s/decoder_%d_0" % (p, skip + 1), srcs, "relu")
            cs = [c]
            for u in range(1, UNITS[skip - 1]):
                d, c = L.block_fwd("pu", "%s/decoder_%d_%d" % (p, skip + 1, u), [d], "relu")
                cs.append(c)
            dctx.append(cs)
        y, c_head = L.unit_fwd("deconv", p + "/decoder_1", [d, z[0]], None, bn_scope=p + "/decoder_1")   # :661-666
        out = ops.tanh_fwd(y)
        ctx = None
        if save:
            ctx = dict(L=L, c_stem=c_stem, ectx=ectx, tctx=tctx, nz=nz, fc=fc, dctx=dctx, c_head=c_head, out=out, N=N, nc=nc,
                       nh=nh, nw=nw)
        return out, ctx

    def backward(self, g_out, ctx):
        ops, st, p = self.ops, self.store, "generator"
        L = ctx["L"]
        g_z = [None] * 5                       # gradients reaching the encoder outputs through the skip connections
        g, g_z[0] = L.unit_bwd(ops.act_bwd(g_out, ctx["out"], ACT_TANH), ctx["c_head"])
        g_feat = g_nz = None
        for i in reversed(range(4)):
            skip = 4 - i
            cs = ctx["dctx"][i]
            for c in reversed(cs[1:]):
                g = L.block_bwd(g, c)[0]
            g_a, g_b = L.block_bwd(g, cs[0])
            if i == 0:
                g_feat, g_nz = g_a, g_b
            else:
                g, g_z[skip] = g_a, g_b
        N, nc, nh, nw = ctx["N"], ctx["nc"], ctx["nh"], ctx["nw"]
        g_fc = ops.nhwc_to_nchw(g_nz, out_dtype=torch.float32).view(N, 1, 1, nc * nh * nw)
        g_fc = ops.act_bwd(g_fc, ctx["fc"], ACT_MIU)
        wfc = st.p[p + "/fully_connected/weights"]
        ops.conv_wgrad([(ctx["nz"], False)], g_fc, st.g[p + "/fully_connected/weights"].view(1, 1, *wfc.shape),
                       st.g[p + "/fully_connected/biases"])
        g = text_fusion.text_fusion_bwd(ops, st, g_feat, ctx["tctx"]) if self.lstm_hybrid else g_feat
        for lvl in (3, 2, 1, 0):
            if g_z[lvl + 1] is not None:
                ops.add_(g, g_z[lvl + 1])
            cs = ctx["ectx"][lvl]
            for c in reversed(cs[1:]):
                g = L.block_bwd(g, c)[0]
            g = L.block_bwd(g, cs[0])[0]
        ops.add_(g, g_z[0])
        L.unit_bwd(g, ctx["c_stem"], need_x=False)
        L.finish_backward()


# ------------------------------------------------------------------------------------------------
# discriminator
# ------------------------------------------------------------------------------------------------
class ResidualDiscriminator:
    FC = "discriminator/fully_connected"

    def __init__(self, ops, store, size=64):
        self.ops, self.store, self.size = ops, store, size

    def new_weight_view(self, need_wgrad=True):
        wv = blocks.WeightView(self.store, self.ops, sn=True, need_wgrad=need_wgrad)
        wv.layers = Layers(self.ops, self.store, need_wgrad)
        return wv

    def forward(self, sketch, img, wv, save=True):
        """sketch, img NHWC [N,H,W,3] -> (patch logits [N,H/32,W/32,1], class logits [N,1,1,25], ctx)."""
        ops, p, L = self.ops, "discriminator", wv.layers
        L.begin(img.shape[0], img.device)
        h = torch.cat([sketch, img], dim=3)
        bctx = []
        for k in range(1, 5):                                                                             # :866-877
            h, c = L.block_fwd("en", "%s/layer_%d" % (p, k), [h], "lrelu")
            bctx.append(c)
        top, c5 = L.block_fwd("en", p + "/layer_5", [h], "lrelu")                                        # :880-882
        disc, cd = L.conv_fwd("c4", p + "/layer_5", [top])
        pooled = ops.spatial_mean_fwd(h)                                                                  # :885
        w2, bf = wv.get(self.FC)
        logits = ops.conv_fwd([(pooled, False)], w2.view(1, 1, *w2.shape), bf)
        ctx = dict(bctx=bctx, c5=c5, cd=cd, h=h, pooled=pooled, wv=wv) if save else None
        return disc, logits, ctx

    def backward(self, g_disc, g_logits, ctx, need_x_grad):
        ops, p = self.ops, "discriminator"
        wv, h = ctx["wv"], ctx["h"]
        L = wv.layers
        g_top = L.conv_bwd("c4", p + "/layer_5", g_disc, ctx["cd"])[0]
        g_h = L.block_bwd(g_top, ctx["c5"])[0]
        if g_logits is not None:
            w2, _ = wv.get(self.FC)
            if wv.need_wgrad:
                gw, db = wv.grads(self.FC)
                ops.conv_wgrad([(ctx["pooled"], False)], g_logits, gw.view(1, 1, *gw.shape), db)
            g_pool = ops.conv_dgrad(g_logits, w2.view(1, 1, *w2.shape), 0, w2.shape[0])
            ops.add_(g_h, ops.spatial_mean_bwd(g_pool, h.shape[1], h.shape[2]))
        for k in (4, 3, 2):
            g_h = L.block_bwd(g_h, ctx["bctx"][k - 1])[0]
        g_x = L.block_bwd(g_h, ctx["bctx"][0], need_x=need_x_grad)[0]
        L.finish_backward()
        return g_x[..., 3:].contiguous() if need_x_grad else None
