"""`--block_type Pix2Pix`: U-Net generator and PatchGAN discriminator, forward and hand-written backward.

Reference: models_collection.generate_pix2pix (:444-538) = image_encoder_pix2pix (:408-441) + encode_feat_with_text
(:150-248) + noise FC (:479-492) + five conv2d_transpose layers with skip connections (:505-529);
discriminate_pix2pix (:789-841).  Filters are 4x4 without bias; normalisation is plain batch norm (:36-46).

The 4x4 layers run in PHASE FORM on the stride-1 SAME kernels of the MRU path, so no new contraction kernel is involved:
  * stride-2 convolution with pad 1 (nchw_conv, :380-391)   = 3x3 SAME convolution over space_to_depth(x) [N,h,w,4C];
  * conv2d_transpose, stride 2, SAME (nchw_deconv, :394-405) = depth_to_space of a 3x3 SAME convolution to 4*Co channels;
  * stride-1 convolution with pad 1 (H -> H-1)              = a 5x5 SAME convolution (first filter row / column zero)
    cropped by one row and column.
`ops.phase_weights` scatters the 4x4 filter into the odd-size one (4/9 resp. 16/25 of its taps are non-zero: the price is
2.25x resp. 1.56x the multiply-adds of a dedicated kernel), `ops.phase_wgrad` gathers its gradient back.
"""
from __future__ import annotations

import torch

from . import blocks, text_fusion
from .ops_base import ACT_MIU, ACT_NONE, ACT_TANH


# ------------------------------------------------------------------------------------------------
# layers
# ------------------------------------------------------------------------------------------------
class _Filters:
    """Expanded filters of one forward/backward pass (built once per filter) and their gradient buffers."""

    def __init__(self, store, ops, need_wgrad=True):
        self.store, self.ops, self.need_wgrad = store, ops, need_wgrad
        self.w, self.dw = {}, {}

    def get(self, name, mode):
        if name not in self.w:
            self.w[name] = (self.ops.phase_weights(self.store.p[name], mode), mode)
        return self.w[name][0]

    def grad(self, name):
        if name not in self.dw:
            self.dw[name] = self.ops.zeros_f32(self.w[name][0].shape)
        return self.dw[name]

    def finish_backward(self):
        for name, dw in self.dw.items():
            self.ops.phase_wgrad(dw, self.store.g[name], self.w[name][1])
        self.dw = {}


def conv_s2_fwd(ops, fl, name, x):
    """nchw_conv(stride 2).  x NHWC [N,2h,2w,C] -> ([N,h,w,Co], ctx)."""
    xs = ops.space_to_depth(x)
    return ops.conv_fwd([(xs, False)], fl.get(name, "conv"), None), xs


def conv_s2_bwd(ops, fl, name, gy, xs, need_x):
    w = fl.get(name, "conv")
    if fl.need_wgrad:
        ops.conv_wgrad([(xs, False)], gy, fl.grad(name), None)
    if not need_x:
        return None
    return ops.depth_to_space(ops.conv_dgrad(gy, w, 0, w.shape[2]))


def conv_s1_fwd(ops, fl, name, x):
    """nchw_conv(stride 1): [N,H,W,C] -> [N,H-1,W-1,Co]."""
    y = ops.conv_fwd([(x, False)], fl.get(name, "k5"), None)
    return ops.copy_rect(y, x.shape[1] - 1, x.shape[2] - 1)


def conv_s1_bwd(ops, fl, name, gy, x, need_x=True):
    w = fl.get(name, "k5")
    gp = ops.copy_rect(gy, x.shape[1], x.shape[2])              # the cropped row / column received no gradient
    if fl.need_wgrad:
        ops.conv_wgrad([(x, False)], gp, fl.grad(name), None)
    return ops.conv_dgrad(gp, w, 0, w.shape[2]) if need_x else None


def deconv_fwd(ops, fl, name, srcs, act=ACT_NONE):
    """nchw_deconv of the channel concat of `srcs` (never materialised).  Returns (y [N,2h,2w,Co], y3 = its phase form)."""
    y3 = ops.conv_fwd([(s, False) for s in srcs], fl.get(name, "deconv"), None, act=act)
    return ops.depth_to_space(y3), y3


def deconv_bwd(ops, fl, name, g3, srcs, need):
    """g3: gradient in phase form [N,h,w,4Co].  Returns the gradients of the sources flagged in `need`."""
    w = fl.get(name, "deconv")
    if fl.need_wgrad:
        ops.conv_wgrad([(s, False) for s in srcs], g3, fl.grad(name), None)
    out, off = [], 0
    for s, nd in zip(srcs, need):
        c = s.shape[3]
        out.append(ops.conv_dgrad(g3, w, off, c) if nd else None)
        off += c
    return out


def bn_fwd(ops, store, scope, x, labels0):
    """Plain batch norm (models_collection.py:36-46) = the conditional kernel with a one-row table."""
    mean, rstd = ops.chan_stats(x)
    y = ops.cbn_act_fwd(x, mean, rstd, store.p[scope + "/scale"].view(1, -1), store.p[scope + "/offset"].view(1, -1), labels0,
                        ACT_NONE)
    return y, (x, mean, rstd)


def bn_bwd(ops, store, scope, gy, c, labels0):
    x, mean, rstd = c
    return ops.cbn_act_bwd(gy, x, mean, rstd, store.p[scope + "/scale"].view(1, -1), store.p[scope + "/offset"].view(1, -1),
                           labels0, store.g[scope + "/scale"].view(1, -1), store.g[scope + "/offset"].view(1, -1), ACT_NONE)


def _slopes(store):
    """lrelu(x, 0.2) = max(0.2 x, x) and relu as the PReLU kernel with a constant slope."""
    return (torch.full((1,), 0.2, dtype=store.dtype, device=store.device), torch.zeros(1, dtype=store.dtype, device=store.device))


# ------------------------------------------------------------------------------------------------
# generator
# ------------------------------------------------------------------------------------------------
class Pix2PixGenerator:
    def __init__(self, ops, store, size=64, lstm_hybrid=True):
        self.ops, self.store, self.size, self.lstm_hybrid = ops, store, size, lstm_hybrid
        self.a_lrelu, self.a_relu = _slopes(store)

    def forward(self, sketch_nchw, text_ids_host, labels, noise, save=True):
        """Same boundary as generator.Generator.forward; `labels` is unused (plain batch norm), as in the reference."""
        ops, st, p = self.ops, self.store, "generator"
        fl = _Filters(st, ops)
        N = sketch_nchw.shape[0]
        zeros = torch.zeros(N, dtype=torch.int32, device=sketch_nchw.device)
        s0 = ops.nchw_to_nhwc(sketch_nchw)
        e1, xs1 = conv_s2_fwd(ops, fl, p + "/encoder_1/conv/filter", s0)                     # :423-425
        enc, ectx = [e1], [dict(xs=xs1)]
        for k in range(2, 6):                                                                 # :434-439
            r = ops.prelu_fwd(enc[-1], self.a_lrelu)
            c, xs = conv_s2_fwd(ops, fl, p + "/encoder_%d/conv/filter" % k, r)
            y, bc = bn_fwd(ops, st, p + "/encoder_%d" % k, c, zeros)
            enc.append(y)
            ectx.append(dict(xs=xs, bn=bc))
        tctx = None
        if self.lstm_hybrid:
            feat, tctx = text_fusion.text_fusion_fwd(ops, st, enc[4], text_ids_host, save)   # :474-477
        else:
            feat = enc[4]
        nc, nh, nw = enc[4].shape[3] // 8, enc[4].shape[1], enc[4].shape[2]                  # :479-492
        wfc, bfc = st.p[p + "/fully_connected/weights"], st.p[p + "/fully_connected/biases"]
        nz = noise.view(N, 1, 1, noise.shape[1])
        fc = ops.conv_fwd([(nz, False)], wfc.view(1, 1, *wfc.shape), bfc, act=ACT_MIU, out_dtype=torch.float32)
        nzf = ops.nchw_to_nhwc(fc.view(N, nc, nh, nw))
        dctx = []
        d = None
        for i in range(5):                                                                    # :505-529
            k = 5 - i
            srcs = [feat, nzf] if i == 0 else [d, enc[k - 1]]
            rs = [ops.prelu_fwd(s, self.a_relu) for s in srcs]
            last = k == 1
            y, y3 = deconv_fwd(ops, fl, p + "/decoder_%d/deconv/filter" % k, rs, act=ACT_TANH if last else ACT_NONE)
            c = dict(srcs=srcs, rs=rs)
            if last:
                c["y3"] = y3
                d = y
            else:
                d, c["bn"] = bn_fwd(ops, st, p + "/decoder_%d" % k, y, zeros)
            dctx.append(c)
        ctx = None
        if save:
            ctx = dict(fl=fl, ectx=ectx, enc=enc, tctx=tctx, nz=nz, fc=fc, dctx=dctx, zeros=zeros, N=N, nc=nc, nh=nh, nw=nw)
        return d, ctx

    def backward(self, g_out, ctx):
        """g_out: dL/d(image) NHWC.  Accumulates all generator weight gradients into store.grad."""
        ops, st, p = self.ops, self.store, "generator"
        fl, enc, dctx, zeros = ctx["fl"], ctx["enc"], ctx["dctx"], ctx["zeros"]
        g_enc = [None] * 5          # gradient reaching each encoder output through its relu'd skip connection
        g_feat = g_nz = None
        g = g_out
        for i in reversed(range(5)):
            k = 5 - i
            c = dctx[i]
            if k == 1:
                g3 = ops.act_bwd(ops.space_to_depth(g), c["y3"], ACT_TANH)
            else:
                g3 = ops.space_to_depth(bn_bwd(ops, st, p + "/decoder_%d" % k, g, c["bn"], zeros))
            g_rs = deconv_bwd(ops, fl, p + "/decoder_%d/deconv/filter" % k, g3, c["rs"], [True, True])
            g_a = ops.prelu_bwd(g_rs[0], c["srcs"][0], self.a_relu, None)
            g_b = ops.prelu_bwd(g_rs[1], c["srcs"][1], self.a_relu, None)
            if i == 0:
                g_feat, g_nz = g_a, g_b
            else:
                g, g_enc[k - 1] = g_a, g_b
        # noise FC
        N, nc, nh, nw = ctx["N"], ctx["nc"], ctx["nh"], ctx["nw"]
        g_fc = ops.nhwc_to_nchw(g_nz, out_dtype=torch.float32).view(N, 1, 1, nc * nh * nw)
        g_fc = ops.act_bwd(g_fc, ctx["fc"], ACT_MIU)
        wfc = st.p[p + "/fully_connected/weights"]
        ops.conv_wgrad([(ctx["nz"], False)], g_fc, st.g[p + "/fully_connected/weights"].view(1, 1, *wfc.shape),
                       st.g[p + "/fully_connected/biases"])
        # text fusion, then down the encoder
        g_cur = text_fusion.text_fusion_bwd(ops, st, g_feat, ctx["tctx"]) if self.lstm_hybrid else g_feat
        for k in (5, 4, 3, 2):
            if g_enc[k - 1] is not None:
                ops.add_(g_cur, g_enc[k - 1])
            e = ctx["ectx"][k - 1]
            g_c = bn_bwd(ops, st, p + "/encoder_%d" % k, g_cur, e["bn"], zeros)
            g_r = conv_s2_bwd(ops, fl, p + "/encoder_%d/conv/filter" % k, g_c, e["xs"], need_x=True)
            g_cur = ops.prelu_bwd(g_r, enc[k - 2], self.a_lrelu, None)
        ops.add_(g_cur, g_enc[0])
        conv_s2_bwd(ops, fl, p + "/encoder_1/conv/filter", g_cur, ctx["ectx"][0]["xs"], need_x=False)
        fl.finish_backward()


# ------------------------------------------------------------------------------------------------
# discriminator
# ------------------------------------------------------------------------------------------------
class Pix2PixDiscriminator:
    """PatchGAN over (sketch, image) pairs with the ACGAN class head; spectral norm on the head only."""

    FC = "discriminator/fully_connected"

    def __init__(self, ops, store, size=64):
        self.ops, self.store, self.size = ops, store, size
        self.a_lrelu, _ = _slopes(store)

    def new_weight_view(self, need_wgrad=True):
        wv = blocks.WeightView(self.store, self.ops, sn=True, need_wgrad=need_wgrad)
        wv.filters = _Filters(self.store, self.ops, need_wgrad)
        return wv

    def forward(self, sketch, img, wv, save=True):
        """sketch, img NHWC [N,H,W,3] -> (patch logits [N,H/8-2,W/8-2,1], class logits [N,1,1,25], ctx)."""
        ops, st, p, fl = self.ops, self.store, "discriminator", wv.filters
        N = img.shape[0]
        zeros = torch.zeros(N, dtype=torch.int32, device=img.device)
        x = torch.cat([sketch, img], dim=3)                                                   # :811
        c1, xs1 = conv_s2_fwd(ops, fl, p + "/layer_1/conv/filter", x)                         # :814-817
        h = ops.prelu_fwd(c1, self.a_lrelu)
        lctx = [dict(xs=xs1, pre=c1)]
        for k in (2, 3, 4):                                                                   # :822-829
            name = p + "/layer_%d/conv/filter" % k
            if k == 4:
                c, xs = conv_s1_fwd(ops, fl, name, h), h
            else:
                c, xs = conv_s2_fwd(ops, fl, name, h)
            y, bc = bn_fwd(ops, st, p + "/layer_%d" % k, c, zeros)
            h = ops.prelu_fwd(y, self.a_lrelu)
            lctx.append(dict(xs=xs, bn=bc, pre=y))
        disc = conv_s1_fwd(ops, fl, p + "/layer_5/conv/filter", h)                            # :832-833
        pooled = ops.spatial_mean_fwd(h)                                                      # :836
        w2, bf = wv.get(self.FC)
        logits = ops.conv_fwd([(pooled, False)], w2.view(1, 1, *w2.shape), bf)                # :837
        ctx = dict(lctx=lctx, h=h, pooled=pooled, wv=wv, zeros=zeros) if save else None
        return disc, logits, ctx

    def backward(self, g_disc, g_logits, ctx, need_x_grad):
        """Returns dL/d(img) NHWC [N,H,W,3] if need_x_grad else None; accumulates weight gradients when wv.need_wgrad."""
        ops, st, p = self.ops, self.store, "discriminator"
        wv, lctx, h, zeros = ctx["wv"], ctx["lctx"], ctx["h"], ctx["zeros"]
        fl, nw = wv.filters, wv.need_wgrad
        g_h = conv_s1_bwd(ops, fl, p + "/layer_5/conv/filter", g_disc, h)
        if g_logits is not None:
            w2, _ = wv.get(self.FC)
            if nw:
                gw, db = wv.grads(self.FC)
                ops.conv_wgrad([(ctx["pooled"], False)], g_logits, gw.view(1, 1, *gw.shape), db)
            g_pool = ops.conv_dgrad(g_logits, w2.view(1, 1, *w2.shape), 0, w2.shape[0])
            ops.add_(g_h, ops.spatial_mean_bwd(g_pool, h.shape[1], h.shape[2]))
        for k in (4, 3, 2):
            c = lctx[k - 1]
            g_y = ops.prelu_bwd(g_h, c["pre"], self.a_lrelu, None)
            g_c = bn_bwd(ops, st, p + "/layer_%d" % k, g_y, c["bn"], zeros) if nw else self._bn_bwd_x_only(g_y, c, k, zeros)
            name = p + "/layer_%d/conv/filter" % k
            g_h = conv_s1_bwd(ops, fl, name, g_c, c["xs"]) if k == 4 else conv_s2_bwd(ops, fl, name, g_c, c["xs"], True)
        g_c1 = ops.prelu_bwd(g_h, lctx[0]["pre"], self.a_lrelu, None)
        g_x = conv_s2_bwd(ops, fl, p + "/layer_1/conv/filter", g_c1, lctx[0]["xs"], need_x_grad)
        if nw:
            fl.finish_backward()
        return g_x[..., 3:].contiguous() if need_x_grad else None

    def _bn_bwd_x_only(self, g_y, c, k, zeros):
        """G step: the discriminator's table gradients are not wanted -- accumulate them into scratch."""
        ops, st, scope = self.ops, self.store, "discriminator/layer_%d" % k
        x, mean, rstd = c["bn"]
        scale, offset = st.p[scope + "/scale"].view(1, -1), st.p[scope + "/offset"].view(1, -1)
        return ops.cbn_act_bwd(g_y, x, mean, rstd, scale, offset, zeros, ops.zeros_f32(scale.shape), ops.zeros_f32(offset.shape),
                               ACT_NONE)
