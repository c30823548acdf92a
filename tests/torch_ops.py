"""TEST INFRASTRUCTURE: a plain-torch implementation of the operator contract (ops_base.OpsBase).

It exists for two reasons only:
  1. on a CPU-only box it lets the tests check the hand-derived backward passes of the host code
     (blocks.py / generator.py / discriminator.py / text_fusion.py / trainer.py) against torch autograd on the
     oracle, before any GPU time is spent;
  2. on the GPU box every CUDA op is unit-tested against the op of the same name here.
The product never imports this module and has no CPU path.
"""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F

from sketchyscenecolorization_b200.ops_base import ACT_LRELU, ACT_MIU, ACT_NONE, ACT_RELU, ACT_TANH, OpsBase


def _same_pad(size, k, stride):
    out = -(-size // stride)
    total = max((out - 1) * stride + k - size, 0)
    return total // 2, total - total // 2


def _up(x):   # NHWC nearest x2
    return x.repeat_interleave(2, dim=1).repeat_interleave(2, dim=2)


def _sum2(x):  # NHWC 2x2 sum
    return x[:, ::2, ::2] + x[:, 1::2, ::2] + x[:, ::2, 1::2] + x[:, 1::2, 1::2]


def _act(y, act):
    if act == ACT_LRELU:
        return torch.where(y > 0, y, 0.2 * y)
    if act == ACT_TANH:
        return torch.tanh(y)
    if act == ACT_MIU:
        return (y + torch.sqrt(0.09 + y * y)) / 2
    if act == ACT_RELU:
        return torch.relu(y)
    return y


class TorchOps(OpsBase):
    def __init__(self, dtype=torch.float64, device="cpu"):
        self.act_dtype = dtype
        self.cdt = dtype if dtype == torch.float64 else torch.float32   # compute dtype
        self.device = torch.device(device)

    def _c(self, t):
        return t.to(self.cdt)

    def _od(self, out_dtype):
        """fp32 requests map to the compute dtype so that an fp64 run stays fp64 end to end."""
        if out_dtype is None:
            return self.act_dtype
        return self.cdt if out_dtype == torch.float32 else out_dtype

    # ---- conv family
    def _cat(self, srcs):
        return torch.cat([_up(self._c(e[0])) if e[1] else self._c(e[0]) for e in srcs], dim=-1)

    def conv_fwd(self, srcs, w, b, *, stride=1, act=ACT_NONE, out_dtype=None, out=None, acc=False):
        if out is not None:
            y = self.conv_fwd(srcs, w, b, stride=stride, act=act, out_dtype=out.dtype)
            return out.add_(y) if acc else out.copy_(y)
        x = self._cat(srcs).permute(0, 3, 1, 2)
        k = w.shape[0]
        pt, pb = _same_pad(x.shape[2], k, stride)
        pl, pr = _same_pad(x.shape[3], k, stride)
        y = F.conv2d(F.pad(x, (pl, pr, pt, pb)), self._c(w).permute(3, 2, 0, 1), stride=stride)
        if b is not None:
            y = y + self._c(b).view(1, -1, 1, 1)
        return _act(y, act).permute(0, 2, 3, 1).contiguous().to(self._od(out_dtype))

    def conv_dgrad(self, gy, w, c_off, c_len, *, ups=False, out=None, acc=False, out_dtype=None, gy_patch=None):
        k = w.shape[0]
        ws = self._c(w)[:, :, c_off:c_off + c_len, :]
        g = F.conv_transpose2d(self._c(gy).permute(0, 3, 1, 2), ws.permute(3, 2, 0, 1), padding=k // 2)
        g = g.permute(0, 2, 3, 1)
        if ups:
            g = _sum2(g)
        if out is not None:
            if acc:
                out += g.to(out.dtype)
            else:
                out.copy_(g)
            return out
        return g.contiguous().to(self._od(out_dtype))

    def conv_wgrad(self, srcs, gy, dw, db, *, stride=1, gy_patch=None):
        x = self._cat(srcs).permute(0, 3, 1, 2)
        k = dw.shape[0]
        pt, pb = _same_pad(x.shape[2], k, stride)
        pl, pr = _same_pad(x.shape[3], k, stride)
        xp = F.pad(x, (pl, pr, pt, pb)).detach().requires_grad_(False)
        g = self._c(gy).permute(0, 3, 1, 2)
        wshape = (dw.shape[3], dw.shape[2], k, k)
        gw = torch.nn.grad.conv2d_weight(xp, wshape, g, stride=stride)
        dw += gw.permute(2, 3, 1, 0).to(dw.dtype)
        if db is not None:
            db += g.sum(dim=(0, 2, 3)).to(db.dtype)

    # ---- norms / activations
    def chan_stats(self, x):
        xc = self._c(x)
        mean = xc.mean(dim=(0, 1, 2))
        var = ((xc - mean) ** 2).mean(dim=(0, 1, 2))
        return mean, torch.rsqrt(var + 1e-5)

    def cbn_act_fwd(self, x, mean, rstd, scale, offset, labels, act=ACT_MIU):
        xh = (self._c(x) - mean) * rstd
        lab = labels.long()
        y = xh * self._c(scale)[lab][:, None, None, :] + self._c(offset)[lab][:, None, None, :]
        return _act(y, act).to(self.act_dtype)

    def cbn_act_bwd(self, gy, x, mean, rstd, scale, offset, labels, dscale, doffset, act=ACT_MIU, dbias=None):
        lab = labels.long()
        xh = (self._c(x) - mean) * rstd
        gam = self._c(scale)[lab][:, None, None, :]
        y = xh * gam + self._c(offset)[lab][:, None, None, :]
        g = self._c(gy)
        if act == ACT_MIU:
            g = g * 0.5 * (1 + y / torch.sqrt(0.09 + y * y))
        s1 = g.sum(dim=(1, 2))
        s2 = (g * xh).sum(dim=(1, 2))
        doffset.index_add_(0, lab, s1.to(doffset.dtype))
        dscale.index_add_(0, lab, s2.to(dscale.dtype))
        gxh = g * gam
        M = x.shape[0] * x.shape[1] * x.shape[2]
        m1 = gxh.sum(dim=(0, 1, 2)) / M
        m2 = (gxh * xh).sum(dim=(0, 1, 2)) / M
        gx = rstd * (gxh - m1 - xh * m2)
        if dbias is not None:
            dbias += gx.reshape(-1, gx.shape[-1]).sum(0).to(dbias.dtype).reshape(dbias.shape)
        return gx.to(self.act_dtype)

    def prelu_fwd(self, x, a):
        xc = self._c(x)
        return torch.where(a * xc >= xc, a * xc, xc).to(self.act_dtype)

    def prelu_bwd(self, gy, x, a, da, dbias=None, acc_into=None):
        if acc_into is not None:
            return acc_into.copy_(acc_into + self.prelu_bwd(gy, x, a, da))
        xc, g = self._c(x), self._c(gy)
        m = a * xc >= xc
        if da is not None:
            da += (g * xc * m).sum().to(da.dtype)
        gx = torch.where(m, a * g, g)
        if dbias is not None:
            dbias += gx.reshape(-1, gx.shape[-1]).sum(0).to(dbias.dtype).reshape(dbias.shape)
        return gx.to(self.act_dtype)

    def colsum_(self, x, out):
        out += self._c(x).reshape(-1, x.shape[-1]).sum(0).to(out.dtype).reshape(out.shape)
        return out

    def minmax_fwd(self, x):
        xc = self._c(x)
        mn = xc.amin(dim=(1, 2))
        mx = xc.amax(dim=(1, 2))
        gate = (xc - mn[:, None, None, :]) / (mx - mn)[:, None, None, :]
        return gate.to(self.act_dtype), mn, mx

    def minmax_bwd(self, ggate, x, mn, mx, dbias=None):
        xc, g = self._c(x), self._c(ggate)
        mnb, mxb = mn[:, None, None, :], mx[:, None, None, :]
        d = mxb - mnb
        gx = g / d
        g_mx = -(g * (xc - mnb)).sum(dim=(1, 2), keepdim=True) / (d * d)
        g_mn = (g * (xc - mxb)).sum(dim=(1, 2), keepdim=True) / (d * d)
        is_mx = (xc == mxb).to(xc.dtype)
        is_mn = (xc == mnb).to(xc.dtype)
        gx = gx + is_mx * g_mx / is_mx.sum(dim=(1, 2), keepdim=True) + is_mn * g_mn / is_mn.sum(dim=(1, 2), keepdim=True)
        gx = gx * torch.where(xc > 0, torch.ones_like(xc), torch.full_like(xc, 0.2))
        if dbias is not None:
            dbias += gx.reshape(-1, gx.shape[-1]).sum(0).to(dbias.dtype).reshape(dbias.shape)
        return gx.to(self.act_dtype)

    def act_bwd(self, gy, y, act):
        yc, g = self._c(y), self._c(gy)
        if act == ACT_TANH:
            r = g * (1 - yc * yc)
        elif act == ACT_MIU:
            xx = yc - 0.0225 / yc
            r = g * 0.5 * (1 + xx / torch.sqrt(0.09 + xx * xx))
        else:
            raise ValueError(act)
        return r.to(gy.dtype)

    # ---- gating
    def gate_fma_fwd(self, ht, rg, im):
        return (self._c(ht) + self._c(rg) * self._c(im)).to(self.act_dtype)

    def gate_fma_bwd(self, g, rg, im):
        return (self._c(g) * self._c(im)).to(self.act_dtype), (self._c(g) * self._c(rg)).to(self.act_dtype)

    def gate_prelu_fwd(self, ht, rg, im, a):
        return self.prelu_fwd(self.gate_fma_fwd(ht, rg, im), a)

    def gate_prelu_bwd(self, gp, ht, rg, im, a, da, g_ht=None, acc=False):
        gx = self.prelu_bwd(gp, self.gate_fma_fwd(ht, rg, im), a, da)
        if g_ht is not None:
            g_ht.copy_(g_ht + gx if acc else gx)
        return self.gate_fma_bwd(gx, rg, im)

    def mul_up_fwd(self, rg, ht_low):
        return (self._c(rg) * _up(self._c(ht_low))).to(self.act_dtype)

    def mul_up_bwd(self, g, rg, ht_low):
        gc = self._c(g)
        return (gc * _up(self._c(ht_low))).to(self.act_dtype), _sum2(gc * self._c(rg)).to(self.act_dtype)

    def blend_fwd(self, sk_low, h2, zg):
        z = self._c(zg)
        return (_up(self._c(sk_low)) * (1 - z) + self._c(h2) * z).to(self.act_dtype)

    def blend_bwd(self, g, sk_low, h2, zg):
        gc, z = self._c(g), self._c(zg)
        return (_sum2(gc * (1 - z)).to(self.act_dtype), (gc * z).to(self.act_dtype),
                (gc * (self._c(h2) - _up(self._c(sk_low)))).to(self.act_dtype))

    def addpool_fwd(self, a, b):
        return (_sum2(self._c(a) + self._c(b)) / 4).to(self.act_dtype)

    def unpool_bwd(self, g):
        return (_up(self._c(g)) / 4).to(self.act_dtype)

    def meanpool_fwd(self, x):
        return (_sum2(self._c(x)) / 4).to(x.dtype)

    def upsample_fwd(self, x):
        return _up(x).contiguous()

    def zeros_f32(self, shape):
        return torch.zeros(shape, dtype=self.cdt, device=self.device)

    def add_(self, dst, src):
        dst += src.to(dst.dtype)
        return dst

    def spatial_mean_fwd(self, x):
        return self._c(x).mean(dim=(1, 2), keepdim=True).to(self.act_dtype)

    def spatial_mean_bwd(self, g, H, W):
        return (self._c(g) / (H * W)).expand(g.shape[0], H, W, g.shape[3]).contiguous().to(self.act_dtype)

    # ---- layout
    def nchw_to_nhwc(self, x, out_dtype=None):
        return x.permute(0, 2, 3, 1).contiguous().to(self._od(out_dtype))

    def nhwc_to_nchw(self, x, out_dtype=None):
        return x.permute(0, 3, 1, 2).contiguous().to(self._od(out_dtype))

    def cast(self, x, dtype):
        if dtype == torch.float32:
            dtype = self.cdt
        return x.to(dtype)

    # ---- phase form of the 4x4 convolutions
    def space_to_depth(self, x):
        N, H, W, C = x.shape
        return x.reshape(N, H // 2, 2, W // 2, 2, C).permute(0, 1, 3, 2, 4, 5).reshape(N, H // 2, W // 2, 4 * C).contiguous()

    def depth_to_space(self, x):
        N, h, w, C4 = x.shape
        C = C4 // 4
        return x.reshape(N, h, w, 2, 2, C).permute(0, 1, 3, 2, 4, 5).reshape(N, 2 * h, 2 * w, C).contiguous()

    @staticmethod
    def _phase_index(mode):
        """ky (0..3) -> (row of the expanded filter, phase) per axis."""
        if mode == "conv":        # input row 2y + ky - 1 = 2 (y + dy) + py
            return [((ky - 1) // 2 + 1, (ky - 1) % 2) for ky in range(4)]
        if mode == "deconv":      # output row 2y + py takes input row y + dy through tap ky = py + 1 - 2 dy
            return [((((ky + 1) % 2) + 1 - ky) // 2 + 1, (ky + 1) % 2) for ky in range(4)]
        return [(ky + 1, 0) for ky in range(4)]

    def phase_weights(self, f, mode):
        idx = self._phase_index(mode)
        A, B = f.shape[2], f.shape[3]
        if mode == "conv":
            w = f.new_zeros(3, 3, 4, A, B)
        elif mode == "deconv":
            w = f.new_zeros(3, 3, B, 4, A)
        else:
            w = f.new_zeros(5, 5, A, B)
        for ky in range(4):
            for kx in range(4):
                (ry, py), (rx, px) = idx[ky], idx[kx]
                if mode == "conv":
                    w[ry, rx, py * 2 + px] = f[ky, kx]
                elif mode == "deconv":
                    w[ry, rx, :, py * 2 + px] = f[ky, kx].t()
                else:
                    w[ry, rx] = f[ky, kx]
        if mode == "conv":
            return w.reshape(3, 3, 4 * A, B)
        if mode == "deconv":
            return w.reshape(3, 3, B, 4 * A)
        return w

    def phase_wgrad(self, dw, df, mode):
        idx = self._phase_index(mode)
        A, B = df.shape[2], df.shape[3]
        if mode == "conv":
            dw = dw.reshape(3, 3, 4, A, B)
        elif mode == "deconv":
            dw = dw.reshape(3, 3, B, 4, A)
        for ky in range(4):
            for kx in range(4):
                (ry, py), (rx, px) = idx[ky], idx[kx]
                if mode == "conv":
                    df[ky, kx] += dw[ry, rx, py * 2 + px].to(df.dtype)
                elif mode == "deconv":
                    df[ky, kx] += dw[ry, rx, :, py * 2 + px].t().to(df.dtype)
                else:
                    df[ky, kx] += dw[ry, rx].to(df.dtype)

    def tanh_fwd(self, x):
        return torch.tanh(self._c(x)).to(self.act_dtype)

    def copy_rect(self, x, H, W):
        N, h, w, C = x.shape
        out = x.new_zeros(N, H, W, C)
        out[:, :min(h, H), :min(w, W)] = x[:, :min(h, H), :min(w, W)]
        return out

    # ---- instance-matching model
    def affine_act(self, x, scale, shift, res=None, rscale=None, rshift=None, relu=False):
        y = self._c(x) * self._c(scale) + self._c(shift)
        if res is not None:
            y = y + (self._c(res) if rscale is None else self._c(res) * self._c(rscale) + self._c(rshift))
        return (torch.relu(y) if relu else y).to(x.dtype)

    def pad_cast_rows(self, x, cp, dtype):
        y = torch.zeros(x.shape[0], cp, dtype=self._od(dtype), device=x.device)
        y[:, :x.shape[1]] = x
        return y

    def maxpool3x3s2(self, x):
        xn = self._c(x).permute(0, 3, 1, 2)
        pt, pb = _same_pad(xn.shape[2], 3, 2)
        pl, pr = _same_pad(xn.shape[3], 3, 2)
        y = F.max_pool2d(F.pad(xn, (pl, pr, pt, pb), value=float("-inf")), 3, 2)
        return y.permute(0, 2, 3, 1).contiguous().to(x.dtype)

    def space_to_batch(self, x, r):
        N, H, W, C = x.shape
        return x.view(N, H // r, r, W // r, r, C).permute(2, 4, 0, 1, 3, 5).reshape(r * r * N, H // r, W // r, C).contiguous()

    def batch_to_space(self, x, r):
        NB, h, w, C = x.shape
        N = NB // (r * r)
        return x.view(r, r, N, h, w, C).permute(2, 3, 0, 4, 1, 5).reshape(N, h * r, w * r, C).contiguous()

    def resize_bilinear_sigmoid(self, x, H, W):
        """TF-1 resize_bilinear, align_corners False: source = destination * (h / H), no half-pixel shift."""
        N, h, w, C = x.shape
        xc = self._c(x)
        fy = torch.arange(H, dtype=self.cdt, device=x.device) * (h / H)
        fx = torch.arange(W, dtype=self.cdt, device=x.device) * (w / W)
        y0, x0 = fy.floor().long(), fx.floor().long()
        y1, x1 = (y0 + 1).clamp(max=h - 1), (x0 + 1).clamp(max=w - 1)
        ly, lx = (fy - y0).view(1, H, 1, 1), (fx - x0).view(1, 1, W, 1)
        top = xc[:, y0][:, :, x0] + (xc[:, y0][:, :, x1] - xc[:, y0][:, :, x0]) * lx
        bot = xc[:, y1][:, :, x0] + (xc[:, y1][:, :, x1] - xc[:, y1][:, :, x0]) * lx
        up = top + (bot - top) * ly
        return up, torch.sigmoid(up)

    # ---- real-data input: the CPU oracle's restatement of get_paired_input
    def paired_input(self, cartoon, sketch, out_hw, seed=0, dequantize=True):
        from oracle import input_oracle
        im, sk = input_oracle.paired_input(cartoon.cpu().numpy(), sketch.cpu().numpy(), tuple(out_hw), seed, dequantize)
        return torch.from_numpy(im), torch.from_numpy(sk)

    # ---- text fusion
    @staticmethod
    def _into(out, val):
        if out is None:
            return val
        out.copy_(val)
        return out

    def l2norm_rows_fwd(self, x, out=None):
        inv = torch.rsqrt(torch.clamp((x * x).sum(1), min=1e-12))
        return self._into(out, x * inv[:, None]), inv

    def l2norm_rows_bwd(self, gy, y, inv):
        # y = x*inv (inv treated as a function of x unless clamped; clamp never active on real data)
        return (gy - y * (gy * y).sum(1, keepdim=True)) * inv[:, None]

    def embedding_fwd(self, table, ids, t, out=None):
        return self._into(out, table[ids[:, t].long()].to(self.cdt))

    def embedding_bwd(self, g, ids, t, dtable):
        dtable.index_add_(0, ids[:, t].long(), g.to(dtable.dtype))

    def embedding_all_fwd(self, table, ids):
        return table[ids.long().t()].to(self.cdt).contiguous()

    def embedding_all_bwd(self, g, ids, dtable):
        dtable.index_add_(0, ids.long().t().reshape(-1), g.reshape(-1, g.shape[-1]).to(dtable.dtype))

    def lstm_seq_supported(self, N, D):
        return True

    def lstm_seq_fwd(self, gx, kh, ids):
        T, N, D4 = gx.shape
        D = D4 // 4
        h_all, c_all = gx.new_zeros((T + 1, N, D)), gx.new_zeros((T + 1, N, D))
        pre_all = gx.new_zeros((T, N, D4))
        for t in range(T):
            pre = gx[t] + h_all[t] @ kh
            i, j, f, o = pre.chunk(4, dim=1)
            c = c_all[t] * torch.sigmoid(f + 1.0) + torch.sigmoid(i) * torch.tanh(j)
            h = torch.tanh(c) * torch.sigmoid(o)
            m = (ids[:, t] != 0)[:, None]
            pre_all[t], c_all[t + 1], h_all[t + 1] = pre, torch.where(m, c, c_all[t]), torch.where(m, h, h_all[t])
        return h_all, c_all, pre_all

    def lstm_seq_bwd(self, g_hext, pre_all, c_all, kh, ids):
        T, N, D = g_hext.shape
        g_pre_all = torch.zeros_like(pre_all)
        g_h, g_c = torch.zeros_like(g_hext[0]), torch.zeros_like(g_hext[0])
        for t in range(T - 1, -1, -1):
            g_pre, g_c, g_pass = self.lstm_cell_bwd(g_c, g_hext[t] + g_h, pre_all[t], c_all[t], c_all[t + 1], ids, t, 1)
            g_pre_all[t] = g_pre
            g_h = g_pre @ kh.t() + g_pass
        return g_pre_all

    def lstm_cell_fwd(self, gates, gates2, grow, c_prev, h_prev, ids, t, P, out_h=None, save_pre=True):
        pre = gates.clone()
        if gates2 is not None:
            pre = pre + gates2
        if grow is not None:
            pre = pre + grow.repeat_interleave(P, dim=0)
        i, j, f, o = pre.chunk(4, dim=1)
        c = c_prev * torch.sigmoid(f + 1.0) + torch.sigmoid(i) * torch.tanh(j)
        h = torch.tanh(c) * torch.sigmoid(o)
        m = (ids[:, t] != 0).repeat_interleave(P)[:, None]
        return torch.where(m, c, c_prev), self._into(out_h, torch.where(m, h, h_prev)), (pre if save_pre else None)

    def lstm_cell_bwd(self, gc, gh, pre, c_prev, c, ids, t, P, out_gpre=None):
        m = (ids[:, t] != 0).repeat_interleave(P)[:, None]
        i, j, f, o = pre.chunk(4, dim=1)
        si, sf, so, tj = torch.sigmoid(i), torch.sigmoid(f + 1.0), torch.sigmoid(o), torch.tanh(j)
        cn = c_prev * sf + si * tj          # recompute (c holds c_prev on masked rows)
        tc = torch.tanh(cn)
        g_o = gh * tc * so * (1 - so)
        g_cn = gc + gh * so * (1 - tc * tc)
        g_f = g_cn * c_prev * sf * (1 - sf)
        g_i = g_cn * tj * si * (1 - si)
        g_j = g_cn * si * (1 - tj * tj)
        g_pre = torch.cat([g_i, g_j, g_f, g_o], 1)
        z = torch.zeros_like(gc)
        return (self._into(out_gpre, torch.where(m, g_pre, torch.zeros_like(g_pre))), torch.where(m, g_cn * sf, gc),
                torch.where(m, z, gh))

    def rows_group_sum(self, x, P, out=None):
        return self._into(out, x.view(-1, P, x.shape[1]).sum(1))

    def atanh_relu_fwd(self, h):
        return torch.relu(0.5 * (torch.log(1.001 + h) - torch.log(1.001 - h)))

    def atanh_relu_bwd(self, gy, h):
        y = 0.5 * (torch.log(1.001 + h) - torch.log(1.001 - h))
        return torch.where(y > 0, gy * 0.5 * (1 / (1.001 + h) + 1 / (1.001 - h)), torch.zeros_like(gy))

    # ---- spectral norm
    def sn_fwd(self, w2d, u):
        eps = 1e-12
        a = u @ w2d.t()
        na = a.norm()
        v = a / (na + eps)
        b = v @ w2d
        nb = b.norm()
        u_new = b / (nb + eps)
        sigma = (b @ u_new.t())[0, 0]
        return w2d / sigma, dict(v=v, u=u.clone(), u_new=u_new, b=b, a=a, na=na, nb=nb, sigma=sigma)

    def sn_bwd(self, gwbar, w2d, ctx, dw):
        eps = 1e-12
        v, u, b, a, na, nb, sigma = ctx["v"], ctx["u"], ctx["b"], ctx["a"], ctx["na"], ctx["nb"], ctx["sigma"]
        gsig = -(gwbar * w2d).sum() / (sigma * sigma)
        # sigma = nb^2/(nb+eps);  d sigma/d b = (nb + 2 eps)/(nb+eps)^2 * b
        gb = gsig * (nb + 2 * eps) / (nb + eps) ** 2 * b                 # [1,C]
        gv = gb @ w2d.t()                                               # [1,K]
        ga = gv / (na + eps) - a * ((gv * a).sum() / ((na + eps) ** 2 * na))
        dw += gwbar / sigma + v.t() @ gb + ga.t() @ u

    # ---- losses
    def softplus_mean(self, d, sign):
        dc = self._c(d)
        loss = F.softplus(sign * dc).mean()
        g = sign * torch.sigmoid(sign * dc) / dc.numel()
        return loss, g.to(d.dtype)

    def ce_loss(self, logits, labels, focal, weight):
        lg = self._c(logits).view(logits.shape[0], -1)
        lab = labels.long()
        N = lg.shape[0]
        p = torch.softmax(lg, 1)
        pt = p.gather(1, lab[:, None])[:, 0]
        ce = -torch.log(pt)
        onehot = F.one_hot(lab, lg.shape[1]).to(lg.dtype)
        if focal:
            loss = ((1 - pt) ** 2 * ce).mean()
            # d/dlogits of (1-pt)^2 * ce : dpt/dz = pt*(onehot - p); dce/dz = p - onehot
            coef = (-2 * (1 - pt) * ce)[:, None] * (pt[:, None] * (onehot - p)) + ((1 - pt) ** 2)[:, None] * (p - onehot)
        else:
            loss = ce.mean()
            coef = p - onehot
        return weight * loss, (weight * coef / N).view(logits.shape).to(logits.dtype)

    def smooth_l1(self, target, gen, weight):
        d = self._c(target) - self._c(gen)
        ab = d.abs()
        loss = torch.where(ab < 1.0, 0.5 * ab * ab, ab - 0.5).mean()
        g = -torch.clamp(d, -1.0, 1.0) * (weight / d.numel())
        return weight * loss, g.to(gen.dtype)

    def reg_loss(self, store):
        tot = torch.zeros((), dtype=self.cdt)
        for s in store.specs:
            if s.trainable and s.reg > 0:
                tot = tot + s.reg * 0.5 * (store.p[s.name].to(self.cdt) ** 2).sum()
        return tot

    # ---- optimiser
    def add_reg_grad(self, store):
        for s in store.specs:
            if s.trainable and s.reg > 0:
                store.g[s.name] += s.reg * store.p[s.name]

    def adam_step(self, store, lr, add_reg_grad=True, lr_dev=None):
        if add_reg_grad:
            self.add_reg_grad(store)
        store.adam_t += 1
        g = store.grad
        store.adam_v.mul_(0.9).add_(0.1 * g * g)
        lr_t = lr * math.sqrt(1 - 0.9 ** store.adam_t)
        store.flat -= lr_t * g / (store.adam_v.sqrt() + 1e-8)

    def optimizer_step(self, store, kind, lr, add_reg_grad=True, lr_dev=None):
        if add_reg_grad:
            self.add_reg_grad(store)
        if lr_dev is not None:
            lr = float(lr_dev)
        g = store.grad
        if kind == 'rmsprop':
            store.adam_v.mul_(0.9).add_(0.1 * g * g)
            store.flat -= lr * g / (store.adam_v + 1e-10).sqrt()
        elif kind == 'adadelta':
            store.adam_v.mul_(0.95).add_(0.05 * g * g)
            upd = (store.opt_s2 + 1e-8).sqrt() * (store.adam_v + 1e-8).rsqrt() * g
            store.opt_s2.mul_(0.95).add_(0.05 * upd * upd)
            store.flat -= lr * upd
        elif kind == 'adagrad':
            store.adam_v.add_(g * g)
            store.flat -= lr * g / store.adam_v.sqrt()
        else:
            raise ValueError(kind)
