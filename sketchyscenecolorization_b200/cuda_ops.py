"""CudaOps: the operator contract (ops_base.OpsBase) on hand-written sm_100a kernels behind the C-ABI of
include/fgcolor.h.  PyTorch tensors are used only as device-buffer containers (allocation + data_ptr);
every arithmetic op is a call into libfgcolor.so on torch's current CUDA stream.

There is no CPU implementation: constructing CudaOps without a CUDA device or without the built library
raises.
"""
from __future__ import annotations

import ctypes as C
import math
import os

import torch

from . import _lib
from ._lib import FgcSrc, check
from .ops_base import ACT_LRELU, ACT_MIU, ACT_NONE, ACT_TANH, OpsBase

_DT = {torch.float32: 0, torch.bfloat16: 1}


def _same_pad(size, k, stride):
    out = -(-size // stride)
    total = max((out - 1) * stride + k - size, 0)
    return out, total // 2


class CudaOps(OpsBase):
    supports_cuda_graphs = True     # nothing allocates or synchronises inside the library; TrainSession replays graphs on it

    def __init__(self, device="cuda:0", act_dtype=torch.float32, conv_terms=2):
        """act_dtype float32: fp32 storage, tensor-core products in the split mode -- conv_terms = 2: bf16x3 (x = x1 + x2,
        three products per K step: the parity mode of the MRU / Pix2Pix networks), conv_terms = 3: six products
        (x = x1 + x2 + x3; forward convolutions without an epilogue activation -- every convolution of the Residual and
        background generators -- get the three missing products x2 w2 + x1 w3 + x3 w1 as accumulating bf16 passes: ~110
        batch-normalised layers amplify the bf16x3 residue 2^-16 to 4e-3 .. 1e-2 on the picture).
        act_dtype bfloat16: bf16 storage, single-pass bf16 products (training throughput mode)."""
        if conv_terms not in (2, 3):
            raise ValueError("conv_terms must be 2 (bf16x3) or 3 (six products)")
        self.conv_terms = conv_terms
        self.fold_narrow_dgrad = os.environ.get("FGC_FOLD_DGRAD", "1") != "0"
        self.fuse_gate_prelu = os.environ.get("FGC_GATE_PRELU", "1") != "0"
        self.phase_dgrad = os.environ.get("FGC_PHASE_DGRAD", "0") == "1"       # measured neutral (profiles/r2ae_*): opt-in
        if not torch.cuda.is_available():
            raise RuntimeError("CudaOps needs a CUDA device (sm_100a); there is no CPU path in this package")
        if act_dtype not in _DT:
            raise ValueError("act_dtype must be float32 (bf16x3 tensor-core mode) or bfloat16 (single-pass bf16)")
        self.lib = _lib.load()
        self.device = torch.device(device)
        self.act_dtype = act_dtype
        torch.cuda.set_device(self.device)
        # constants of conv_dgrad_pooled_gy, on the device before any step is captured into a CUDA graph
        self._phase_mix = {d: torch.tensor(self._PHASE_MIX[d], dtype=torch.float32, device=self.device) for d in (0, 1)}

    # ---------------- helpers ----------------
    @staticmethod
    def _s():
        return torch.cuda.current_stream().cuda_stream

    @staticmethod
    def _p(t):
        if t is None:
            return None
        assert t.is_contiguous(), "libfgcolor needs contiguous buffers"
        return t.data_ptr()

    @staticmethod
    def _dt(t):
        return _DT[t.dtype]

    def _empty(self, shape, dtype):
        return torch.empty(shape, dtype=dtype, device=self.device)

    def _f32(self, t):
        assert t.dtype == torch.float32 and t.is_contiguous()
        return t.data_ptr()

    def run_aside(self, fn):
        """On a side stream (three, round robin) forked from and later joined to the current one; inside a captured step this
        is a parallel branch of the CUDA graph.  Buffers allocated by fn belong to the side stream's pool: they are recycled
        only by later side work, which starts with a wait on the current stream, i.e. after every reader enqueued so far.
        FGC_SIDE_STREAMS=0 runs everything inline."""
        if os.environ.get("FGC_SIDE_STREAMS", "1") == "0":
            return fn(), (lambda: None)
        main = torch.cuda.current_stream()
        pool = self.__dict__.setdefault("_side_streams", [])
        if not pool:
            pool.extend(torch.cuda.Stream(device=self.device) for _ in range(3))
            self._side_next = 0
        side = pool[self._side_next % len(pool)]
        self._side_next += 1
        side.wait_stream(main)
        with torch.cuda.stream(side):
            r = fn()
        return r, (lambda: main.wait_stream(side))

    def launch_count(self):
        return int(self.lib.fgc_launch_count())

    # ---------------- convolution family ----------------
    def _srcs(self, srcs):
        dt = srcs[0][0].dtype
        arr = (FgcSrc * len(srcs))()
        H = W = None
        for i, e in enumerate(srcs):
            t, ups = e[0], e[1]
            patch = e[2] if len(e) > 2 else None
            assert t.dtype == dt, "all conv sources must share a dtype"
            assert t.dim() == 4
            arr[i].ptr = self._p(t)
            arr[i].C = t.shape[3]
            arr[i].ups = 1 if ups else 0
            arr[i].patch = None if patch is None else self._p(patch)
            h, w = (t.shape[1] * 2, t.shape[2] * 2) if ups else (t.shape[1], t.shape[2])
            assert H is None or (H, W) == (h, w), "conv sources disagree on the spatial size"
            H, W = h, w
        return arr, srcs[0][0].shape[0], H, W, _DT[dt]

    def _ws(self, cs, k, nout, dt):
        arr = (C.c_int * len(cs))(*cs)
        n = self.lib.fgc_conv2d_ws_bytes(arr, len(cs), k, nout, dt)
        return self._empty((n,), torch.uint8)

    def _conv_fwd_column_folded(self, src, w, b, act, out_dtype):
        """A k x k, stride-1 layer with very few outputs over one wide bf16 source (the generator's 7x7, 64 -> 3 head,
        models_collection.py:372-374).  On the tensor path its natural form is k*k narrow (N = 16) instructions per K step,
        each paying a full 128-pixel operand fetch: 1.22 ms for 44 GFLOP.  Folded: z = the k x 1 vertical convolution with
        k*Cout outputs (filter column kw in channels kw*Cout..: one N = 32 instruction per filter row; the filter is handed
        over as a k x k one whose only non-zero column is the centre, flag 2 lets the halo kernel skip the rest), then
        y[.., w, co] = act(b + sum_kw z[.., w + kw - pad, kw*Cout + co]) -- a small streaming pass."""
        x = src[0]
        N, H, W, cin = x.shape
        k, cout = w.shape[0], w.shape[3]
        w2 = torch.zeros((k, k, cin, k * cout), dtype=torch.float32, device=self.device)
        # pure re-indexing: w2[kh, centre, c, kw*Cout + co] = w[kh, kw, c, co]
        w2[:, (k - 1) // 2].copy_(w.permute(0, 2, 1, 3).reshape(k, cin, k * cout))
        arr, N, H, W, dt = self._srcs([src])
        pad = (k - 1) // 2
        z = self._empty((N, H, W, k * cout), torch.float32)
        ws = self._ws([cin], k, k * cout, dt)
        check(self.lib.fgc_conv2d_fwd_acc(arr, 1, dt, N, H, W, self._f32(w2), k, cin, k * cout, None, 1, pad, pad, H, W, ACT_NONE, 2,
                                          self._p(z), self._dt(z), self._p(ws), self._s()), "conv2d_fwd (column folded)")
        y = self._empty((N, H, W, cout), out_dtype or self.act_dtype)
        check(self.lib.fgc_tapsum_w(self._f32(z), N, H, W, k, cout, k * cout, None if b is None else self._f32(b.reshape(-1)), act,
                                    self._p(y), self._dt(y), self._s()), "tapsum_w")
        return y

    def conv_fwd(self, srcs, w, b, *, stride=1, act=ACT_NONE, out_dtype=None, out=None, acc=False):
        if out is not None:
            return self._conv_fwd_into(srcs, w, b, stride, act, out, acc)
        if (len(srcs) == 1 and stride == 1 and w.shape[0] >= 5 and w.shape[0] * w.shape[3] <= 32 and srcs[0][0].dtype == torch.bfloat16
                and srcs[0][0].shape[3] >= 64 and not srcs[0][1] and srcs[0][0].shape[2] % 8 == 0):
            return self._conv_fwd_column_folded(srcs[0], w, b, act, out_dtype)
        arr, N, H, W, dt = self._srcs(srcs)
        k, cin, cout = w.shape[0], w.shape[2], w.shape[3]
        OH, pt = _same_pad(H, k, stride)
        OW, pl = _same_pad(W, k, stride)
        y = self._empty((N, OH, OW, cout), out_dtype or self.act_dtype)
        ws = self._ws([e[0].shape[3] for e in srcs], k, cout, dt)
        check(self.lib.fgc_conv2d_fwd(arr, len(srcs), dt, N, H, W, self._f32(w), k, cin, cout,
                                      None if b is None else self._f32(b.reshape(-1)), stride, pt, pl, OH, OW, act,
                                      self._p(y), self._dt(y), self._p(ws), self._s()), "conv2d_fwd")
        if self.conv_terms == 3 and srcs[0][0].dtype == torch.float32 and act == ACT_NONE and y.dtype == torch.float32:
            self._six_product_passes(srcs, w, k, cin, cout, stride, pt, pl, OH, OW, y)
        return y

    def _conv_fwd_into(self, srcs, w, b, stride, act, out, acc):
        """out (+)= act(conv + b): fgc_conv2d_fwd_acc, flag bit 0."""
        arr, N, H, W, dt = self._srcs(srcs)
        k, cin, cout = w.shape[0], w.shape[2], w.shape[3]
        OH, pt = _same_pad(H, k, stride)
        OW, pl = _same_pad(W, k, stride)
        assert tuple(out.shape) == (N, OH, OW, cout) and out.is_contiguous()
        ws = self._ws([e[0].shape[3] for e in srcs], k, cout, dt)
        check(self.lib.fgc_conv2d_fwd_acc(arr, len(srcs), dt, N, H, W, self._f32(w), k, cin, cout,
                                          None if b is None else self._f32(b.reshape(-1)), stride, pt, pl, OH, OW, act,
                                          1 if acc else 0, self._p(out), self._dt(out), self._p(ws), self._s()), "conv2d_fwd_acc")
        if self.conv_terms == 3 and srcs[0][0].dtype == torch.float32 and act == ACT_NONE and out.dtype == torch.float32:
            self._six_product_passes(srcs, w, k, cin, cout, stride, pt, pl, OH, OW, out)
        return out

    def _split(self, x, level, bf16):
        out = self._empty(x.shape, torch.bfloat16 if bf16 else torch.float32)
        check(self.lib.fgc_split_term(self._f32(x), x.numel(), level, self._p(out) if bf16 else None, None if bf16 else self._p(out),
                                      self._s()), "split_term")
        return out

    def _six_product_passes(self, srcs, w, k, cin, cout, stride, pt, pl, OH, OW, y):
        """y (= x1 w1 + x1 w2 + x2 w1 from the bf16x3 pass) += x2 w2 + x1 w3 + x3 w1: three single-pass bf16 convolutions
        that accumulate into the fp32 result.  A single-pass convolution rounds its fp32 filter argument to bf16, so handing
        it the filter's residual after one / two terms multiplies by w2 / w3."""
        xs = [e[0] for e in srcs]
        ups = [e[1] for e in srcs]
        for lx, lw in ((1, 1), (0, 2), (2, 0)):
            terms = [(self._split(x, lx, True), u) for x, u in zip(xs, ups)]
            wt = w if lw == 0 else self._split(w, lw, False)
            arr, N, H, W, dt = self._srcs(terms)
            ws = self._ws([t[0].shape[3] for t in terms], k, cout, dt)
            check(self.lib.fgc_conv2d_fwd_acc(arr, len(terms), dt, N, H, W, self._f32(wt), k, cin, cout, None, stride, pt, pl, OH, OW,
                                              ACT_NONE, 1, self._p(y), self._dt(y), self._p(ws), self._s()), "conv2d_fwd_acc")

    def small_patch(self, x, k, ups=False, mirror=False):
        """bf16 training mode only: flattened (tap, channel) copy of a narrow source for the TMA-fed conv kernels."""
        N, h, w, Cc = x.shape
        H, W = (2 * h, 2 * w) if ups else (h, w)
        if x.dtype != torch.bfloat16 or Cc >= 64 or k < 3 or W % 8 != 0:
            return None
        cp = 8 * ((k * k * Cc + 7) // 8)
        out = self._empty((N, H, W, cp), torch.bfloat16)
        check(self.lib.fgc_im2col_small(self._p(x), self._dt(x), N, H, W, Cc, 1 if ups else 0, k, 1 if mirror else 0, self._p(out),
                                        self._s()),
              "im2col_small")
        return out

    def conv_dgrad(self, gy, w, c_off, c_len, *, ups=False, out=None, acc=False, out_dtype=None, gy_patch=None):
        N, H, W, cout = gy.shape
        k, cin = w.shape[0], w.shape[2]
        assert w.shape[3] == cout
        if (k == 3 and k * c_len <= 32 and not ups and gy_patch is None and gy.dtype == torch.bfloat16 and cout >= 64 and W % 8 == 0
                and self.fold_narrow_dgrad):
            # gradient towards a NARROW source (the 8-channel stem features under a 128-wide layer) = a stride-1 convolution of
            # gy with the flipped, transposed filter to c_len outputs: k*k N = 16 instructions per K step on the tensor path,
            # k N = 32 ones in the column-folded form of the 7x7 head
            wt = w[:, :, c_off:c_off + c_len, :].flip(0, 1).permute(0, 1, 3, 2).contiguous()
            g = self._conv_fwd_column_folded((gy, False), wt, None, ACT_NONE, out_dtype if out is None else out.dtype)
            if out is None:
                return g
            return self.add_(out, g) if acc else out.copy_(g)
        oshape = (N, H // 2, W // 2, c_len) if ups else (N, H, W, c_len)
        if out is None:
            out = self._empty(oshape, out_dtype or self.act_dtype)
            acc = False
        assert tuple(out.shape) == oshape
        scratch = self._empty((N, H, W, c_len), torch.float32) if ups else None
        ws = self._ws([cout], k, c_len, self._dt(gy))
        check(self.lib.fgc_conv2d_dgrad(self._p(gy), self._dt(gy), N, H, W, self._f32(w), k, cin, cout, c_off, c_len,
                                        1 if ups else 0, 1 if acc else 0, self._p(out), self._dt(out),
                                        self._p(scratch), self._p(ws), self._p(gy_patch), self._s()), "conv2d_dgrad")
        return out

    # per dimension: filter row tau of the phase filter collects the forward taps t with A[d][tau][t] = 1 (see conv_dgrad_pooled_gy)
    _PHASE_MIX = {0: ((0, 0, 1), (1, 1, 0), (0, 0, 0)), 1: ((0, 0, 0), (0, 1, 1), (1, 0, 0))}
    _PHASE_MASK = {0: 0b011, 1: 0b110}

    def conv_dgrad_pooled_gy(self, g_low, w, c_off, c_len):
        """Per output phase (dy, dx) of the full-resolution gradient, the source pixel 2q + d - (t - 1) of forward tap t lies
        in the low-resolution cell q + o: d = 0 -> t = 0, 1 in o = 0 and t = 2 in o = -1; d = 1 -> t = 0 in o = +1 and t = 1, 2
        in o = 0.  So each phase is a 2x2-tap convolution of g_low with filters pre-summed over the taps that share a cell
        (x 1/4 for the mean): four fgc_conv2d_fwd_phase launches that write their phase of the result directly."""
        N, h, wd, cout = g_low.shape
        k = w.shape[0]
        if not self.phase_dgrad or k != 3 or g_low.dtype != torch.bfloat16 or cout < 64 or wd % 8 != 0:
            return super().conv_dgrad_pooled_gy(g_low, w, c_off, c_len)
        wt = (w[:, :, c_off:c_off + c_len, :].permute(0, 1, 3, 2) * 0.25).contiguous()          # [ty, tx, co, ci]
        out = self._empty((N, 2 * h, 2 * wd, c_len), self.act_dtype)
        arr, _, _, _, dt = self._srcs([(g_low, False)])
        ws = self._ws([cout], 3, c_len, dt)
        mix = self._phase_mix
        for dy in (0, 1):
            for dx in (0, 1):
                f = torch.einsum('ab,cd,bdxy->acxy', mix[dy], mix[dx], wt).contiguous()
                rc = self.lib.fgc_conv2d_fwd_phase(arr, 1, dt, N, h, wd, self._f32(f), 3, cout, c_len, self._PHASE_MASK[dx],
                                                   self._PHASE_MASK[dy], dy, dx, self._p(out), self._dt(out), self._p(ws), self._s())
                if rc == -3 and dy == 0 and dx == 0:          # FGC_EUNSUPPORTED: not a halo-kernel layer
                    return super().conv_dgrad_pooled_gy(g_low, w, c_off, c_len)
                check(rc, "conv2d_fwd_phase")
        return out

    def conv_wgrad(self, srcs, gy, dw, db, *, stride=1, gy_patch=None):
        arr, N, H, W, dt = self._srcs(srcs)
        k, cin, cout = dw.shape[0], dw.shape[2], dw.shape[3]
        OH, pt = _same_pad(H, k, stride)
        OW, pl = _same_pad(W, k, stride)
        assert tuple(gy.shape) == (N, OH, OW, cout)
        check(self.lib.fgc_conv2d_wgrad(arr, len(srcs), dt, N, H, W, self._p(gy), self._dt(gy), k, cin, cout, stride, pt, pl,
                                        OH, OW, self._f32(dw), None if db is None else self._f32(db.reshape(-1)), self._p(gy_patch),
                                        self._s()),
              "conv2d_wgrad")

    # ---------------- normalisation / activations ----------------
    def chan_stats(self, x):
        Cc = x.shape[-1]
        M = x.numel() // Cc
        acc = self._empty((2 * Cc,), torch.float64)
        stats = self._empty((2 * Cc,), torch.float32)
        check(self.lib.fgc_chan_stats(self._p(x), self._dt(x), M, Cc, self._p(acc), self._p(stats), self._s()), "chan_stats")
        return stats[:Cc], stats[Cc:]

    @staticmethod
    def _stats(mean, rstd):
        Cc = mean.numel()
        if mean.data_ptr() + 4 * Cc == rstd.data_ptr():
            return mean.data_ptr()
        return None

    def _stats_ptr(self, mean, rstd, keep):
        p = self._stats(mean, rstd)
        if p is None:
            t = torch.cat([mean.reshape(-1), rstd.reshape(-1)]).contiguous()
            keep.append(t)
            p = t.data_ptr()
        return p

    def cbn_act_fwd(self, x, mean, rstd, scale, offset, labels, act=ACT_MIU):
        N, H, W, Cc = x.shape
        y = self._empty(x.shape, x.dtype)
        keep = []
        check(self.lib.fgc_cbn_act_fwd(self._p(x), self._dt(x), N, H * W, Cc, self._stats_ptr(mean, rstd, keep), self._f32(scale),
                                       self._f32(offset), self._p(labels), act, self._p(y), self._s()), "cbn_act_fwd")
        return y

    def cbn_act_bwd(self, gy, x, mean, rstd, scale, offset, labels, dscale, doffset, act=ACT_MIU, dbias=None):
        N, H, W, Cc = x.shape
        assert gy.dtype == x.dtype
        gx = self._empty(x.shape, x.dtype)
        scratch = self._empty((2 * N * Cc + 2 * Cc,), torch.float32)
        keep = []
        check(self.lib.fgc_cbn_act_bwd(self._p(gy), self._p(x), self._dt(x), N, H * W, Cc, self._stats_ptr(mean, rstd, keep),
                                       self._f32(scale), self._f32(offset), self._p(labels), act, self._f32(dscale),
                                       self._f32(doffset), self._p(gx), self._p(scratch),
                                       None if dbias is None else self._f32(dbias.reshape(-1)), self._s()), "cbn_act_bwd")
        return gx

    def prelu_fwd(self, x, a):
        y = self._empty(x.shape, x.dtype)
        check(self.lib.fgc_prelu_fwd(self._p(x), self._dt(x), x.numel(), self._f32(a), self._p(y), self._s()), "prelu_fwd")
        return y

    def prelu_bwd(self, gy, x, a, da, dbias=None, acc_into=None):
        assert gy.dtype == x.dtype
        if acc_into is not None:
            assert dbias is None and acc_into.shape == x.shape and acc_into.dtype == x.dtype
            check(self.lib.fgc_prelu_bwd_acc(self._p(gy), self._p(x), self._dt(x), x.numel(), self._f32(a),
                                             None if da is None else self._f32(da), self._p(acc_into), self._s()), "prelu_bwd_acc")
            return acc_into
        gx = self._empty(x.shape, x.dtype)
        check(self.lib.fgc_prelu_bwd(self._p(gy), self._p(x), self._dt(x), x.numel(), x.shape[-1], self._f32(a),
                                     None if da is None else self._f32(da),
                                     None if dbias is None else self._f32(dbias.reshape(-1)), self._p(gx), self._s()), "prelu_bwd")
        return gx

    def colsum_(self, x, out):
        """out[C] += sum over all leading dims of x[..., C]"""
        Cc = x.shape[-1]
        check(self.lib.fgc_colsum(self._p(x), self._dt(x), x.numel() // Cc, Cc, self._f32(out.reshape(-1)), self._s()), "colsum")
        return out

    # min-max gates: statistics per (sample, channel), so the two passes (reduce, apply) CAN run over a few samples at a time,
    # the apply pass then reading what the reduce pass has just pulled through the 126 MB L2.  Measured (profiles/r2v_*): it
    # loses -- 98.0 ms per iteration unchunked, 98.7 / 99.9 / 101.4 ms at 80 / 40 / 20 MB chunks: the short launches cost more
    # than the L2 hits return.  Off by default (FGC_MINMAX_CHUNK_MB=0); kept as an option with its test.
    _MM_CHUNK_BYTES = int(os.environ.get("FGC_MINMAX_CHUNK_MB", "0")) << 20

    def _mm_chunks(self, x):
        N = x.shape[0]
        per = x[0].numel() * x.element_size()
        if self._MM_CHUNK_BYTES <= 0 or N * per <= 2 * self._MM_CHUNK_BYTES:
            return [(0, N)]
        b = max(1, self._MM_CHUNK_BYTES // per)
        return [(i, min(N, i + b)) for i in range(0, N, b)]

    def minmax_fwd(self, x):
        N, H, W, Cc = x.shape
        gate = self._empty(x.shape, x.dtype)
        mn = self._empty((N, Cc), torch.float32)
        mx = self._empty((N, Cc), torch.float32)
        scratch = self._empty((2 * N * Cc,), torch.int32)
        for a, b in self._mm_chunks(x):
            check(self.lib.fgc_minmax_fwd(self._p(x[a:b]), self._dt(x), b - a, H * W, Cc, self._p(gate[a:b]), self._p(mn[a:b]),
                                          self._p(mx[a:b]), self._p(scratch[2 * a * Cc:2 * b * Cc]), self._s()), "minmax_fwd")
        return gate, mn, mx

    def minmax_bwd(self, ggate, x, mn, mx, dbias=None):
        N, H, W, Cc = x.shape
        assert ggate.dtype == x.dtype
        gpre = self._empty(x.shape, x.dtype)
        scratch = self._empty((4 * N * Cc,), torch.float32)
        for a, b in self._mm_chunks(x):
            check(self.lib.fgc_minmax_bwd(self._p(ggate[a:b]), self._p(x[a:b]), self._dt(x), b - a, H * W, Cc, self._f32(mn[a:b]),
                                          self._f32(mx[a:b]), self._p(gpre[a:b]), self._p(scratch[4 * a * Cc:4 * b * Cc]),
                                          None if dbias is None else self._f32(dbias.reshape(-1)), self._s()), "minmax_bwd")
        return gpre

    def act_bwd(self, gy, y, act):
        assert gy.dtype == y.dtype
        gx = self._empty(y.shape, y.dtype)
        check(self.lib.fgc_act_bwd(self._p(gy), self._p(y), self._dt(y), y.numel(), act, self._p(gx), self._s()), "act_bwd")
        return gx

    # ---------------- gating ----------------
    def gate_fma_fwd(self, ht, rg, im):
        out = self._empty(ht.shape, ht.dtype)
        check(self.lib.fgc_gate_fma_fwd(self._p(ht), self._p(rg), self._p(im), self._dt(ht), ht.numel(), self._p(out), self._s()),
              "gate_fma_fwd")
        return out

    def gate_fma_bwd(self, g, rg, im):
        g_rg = self._empty(g.shape, g.dtype)
        g_im = self._empty(g.shape, g.dtype)
        check(self.lib.fgc_gate_fma_bwd(self._p(g), self._p(rg), self._p(im), self._dt(g), g.numel(), self._p(g_rg), self._p(g_im),
                                        self._s()), "gate_fma_bwd")
        return g_rg, g_im

    def gate_prelu_fwd(self, ht, rg, im, a):
        out = torch.empty_like(ht)
        check(self.lib.fgc_gate_prelu_fwd(self._p(ht), self._p(rg), self._p(im), self._dt(ht), ht.numel(), self._f32(a), self._p(out),
                                          self._s()), "gate_prelu_fwd")
        return out

    def gate_prelu_bwd(self, gp, ht, rg, im, a, da, g_ht=None, acc=False):
        g_rg, g_im = torch.empty_like(rg), torch.empty_like(im)
        check(self.lib.fgc_gate_prelu_bwd(self._p(gp), self._p(ht), self._p(rg), self._p(im), self._dt(ht), ht.numel(), self._f32(a),
                                          None if da is None else self._f32(da), self._p(g_ht), 1 if acc else 0, self._p(g_rg),
                                          self._p(g_im), self._s()), "gate_prelu_bwd")
        return g_rg, g_im

    def mul_up_fwd(self, rg, ht_low):
        N, h, w, Cc = ht_low.shape
        out = self._empty(rg.shape, rg.dtype)
        check(self.lib.fgc_mul_up_fwd(self._p(rg), self._p(ht_low), self._dt(rg), N, h, w, Cc, self._p(out), self._s()), "mul_up_fwd")
        return out

    def mul_up_bwd(self, g, rg, ht_low):
        N, h, w, Cc = ht_low.shape
        g_rg = self._empty(g.shape, g.dtype)
        g_ht = self._empty(ht_low.shape, g.dtype)
        check(self.lib.fgc_mul_up_bwd(self._p(g), self._p(rg), self._p(ht_low), self._dt(g), N, h, w, Cc, self._p(g_rg),
                                      self._p(g_ht), self._s()), "mul_up_bwd")
        return g_rg, g_ht

    def blend_fwd(self, sk_low, h2, zg):
        N, h, w, Cc = sk_low.shape
        out = self._empty(h2.shape, h2.dtype)
        check(self.lib.fgc_blend_fwd(self._p(sk_low), self._p(h2), self._p(zg), self._dt(h2), N, h, w, Cc, self._p(out), self._s()),
              "blend_fwd")
        return out

    def blend_bwd(self, g, sk_low, h2, zg):
        N, h, w, Cc = sk_low.shape
        g_sk = self._empty(sk_low.shape, g.dtype)
        g_h2 = self._empty(g.shape, g.dtype)
        g_zg = self._empty(g.shape, g.dtype)
        check(self.lib.fgc_blend_bwd(self._p(g), self._p(sk_low), self._p(h2), self._p(zg), self._dt(g), N, h, w, Cc,
                                     self._p(g_sk), self._p(g_h2), self._p(g_zg), self._s()), "blend_bwd")
        return g_sk, g_h2, g_zg

    def addpool_fwd(self, a, b):
        N, H, W, Cc = a.shape
        out = self._empty((N, H // 2, W // 2, Cc), a.dtype)
        check(self.lib.fgc_addpool_fwd(self._p(a), self._p(b), self._dt(a), N, H // 2, W // 2, Cc, self._p(out), self._s()),
              "addpool_fwd")
        return out

    def meanpool_fwd(self, x):
        return self.addpool_fwd(x, None)

    def unpool_bwd(self, g):
        N, h, w, Cc = g.shape
        out = self._empty((N, 2 * h, 2 * w, Cc), g.dtype)
        check(self.lib.fgc_unpool_bwd(self._p(g), self._dt(g), N, h, w, Cc, self._p(out), self._s()), "unpool_bwd")
        return out

    def upsample_fwd(self, x):
        N, h, w, Cc = x.shape
        out = self._empty((N, 2 * h, 2 * w, Cc), x.dtype)
        check(self.lib.fgc_upsample2x(self._p(x), self._dt(x), N, h, w, Cc, self._p(out), self._s()), "upsample2x")
        return out

    def zeros_f32(self, shape):
        return torch.zeros(shape, dtype=torch.float32, device=self.device)

    def add_(self, dst, src):
        assert dst.numel() == src.numel()
        check(self.lib.fgc_axpy(self._p(dst), self._p(src), self._dt(dst), self._dt(src), dst.numel(), 1.0, self._s()), "axpy")
        return dst

    def spatial_mean_fwd(self, x):
        N, H, W, Cc = x.shape
        out = self._empty((N, 1, 1, Cc), x.dtype)
        check(self.lib.fgc_spatial_mean_fwd(self._p(x), self._dt(x), N, H * W, Cc, self._p(out), self._s()), "spatial_mean_fwd")
        return out

    def spatial_mean_bwd(self, g, H, W):
        N, Cc = g.shape[0], g.shape[3]
        out = self._empty((N, H, W, Cc), g.dtype)
        check(self.lib.fgc_spatial_mean_bwd(self._p(g), self._dt(g), N, H * W, Cc, self._p(out), self._s()), "spatial_mean_bwd")
        return out

    # ---------------- layout ----------------
    def nchw_to_nhwc(self, x, out_dtype=None):
        N, Cc, H, W = x.shape
        y = self._empty((N, H, W, Cc), out_dtype or self.act_dtype)
        check(self.lib.fgc_nchw_to_nhwc(self._p(x), self._dt(x), N, Cc, H * W, self._p(y), self._dt(y), self._s()), "nchw_to_nhwc")
        return y

    def nhwc_to_nchw(self, x, out_dtype=None):
        N, H, W, Cc = x.shape
        y = self._empty((N, Cc, H, W), out_dtype or self.act_dtype)
        check(self.lib.fgc_nhwc_to_nchw(self._p(x), self._dt(x), N, Cc, H * W, self._p(y), self._dt(y), self._s()), "nhwc_to_nchw")
        return y

    def cast(self, x, dtype):
        if x.dtype == dtype:
            return x
        y = self._empty(x.shape, dtype)
        check(self.lib.fgc_cast(self._p(x), self._dt(x), self._p(y), self._dt(y), x.numel(), self._s()), "cast")
        return y

    # ---------------- phase form of the 4x4 convolutions ----------------
    _PHASE_MODE = {"conv": 0, "deconv": 1, "k5": 2}

    def space_to_depth(self, x):
        N, H, W, Cc = x.shape
        assert H % 2 == 0 and W % 2 == 0
        y = self._empty((N, H // 2, W // 2, 4 * Cc), x.dtype)
        check(self.lib.fgc_space_to_depth(self._p(x), self._dt(x), N, H // 2, W // 2, Cc, self._p(y), self._s()), "space_to_depth")
        return y

    def depth_to_space(self, x):
        N, h, w, C4 = x.shape
        assert C4 % 4 == 0
        y = self._empty((N, 2 * h, 2 * w, C4 // 4), x.dtype)
        check(self.lib.fgc_depth_to_space(self._p(x), self._dt(x), N, h, w, C4 // 4, self._p(y), self._s()), "depth_to_space")
        return y

    def tanh_fwd(self, x):
        y = self._empty(x.shape, x.dtype)
        check(self.lib.fgc_tanh_fwd(self._p(x), self._dt(x), x.numel(), self._p(y), self._s()), "tanh_fwd")
        return y

    def copy_rect(self, x, H, W):
        N, h, w, Cc = x.shape
        y = self._empty((N, H, W, Cc), x.dtype)
        check(self.lib.fgc_copy_rect(self._p(x), self._dt(x), N, h, w, Cc, self._p(y), H, W, self._s()), "copy_rect")
        return y

    @staticmethod
    def _phase_shape(f, mode):
        A, B = f.shape[2], f.shape[3]
        return {"conv": (3, 3, 4 * A, B), "deconv": (3, 3, B, 4 * A), "k5": (5, 5, A, B)}[mode]

    def phase_weights(self, f, mode):
        assert f.dtype == torch.float32 and tuple(f.shape[:2]) == (4, 4)
        w = self._empty(self._phase_shape(f, mode), torch.float32)
        check(self.lib.fgc_phase_weights(self._f32(f), f.shape[2], f.shape[3], self._PHASE_MODE[mode], self._p(w), self._s()),
              "phase_weights")
        return w

    def phase_wgrad(self, dw, df, mode):
        assert tuple(dw.shape) == self._phase_shape(df, mode)
        check(self.lib.fgc_phase_wgrad(self._f32(dw), df.shape[2], df.shape[3], self._PHASE_MODE[mode], self._f32(df), self._s()),
              "phase_wgrad")

    # ---------------- real-data input ----------------
    def paired_input(self, cartoon, sketch, out_hw, seed=0, dequantize=True):
        N, R = cartoon.shape[0], cartoon.shape[1]
        assert cartoon.dtype == torch.uint8 and cartoon.shape == (N, R, R, 3) and sketch.shape == cartoon.shape
        assert sketch.dtype in (torch.uint8, torch.float32)
        images, sketches = self._empty((N, 3) + tuple(out_hw), torch.float32), self._empty((N, 3) + tuple(out_hw), torch.float32)
        scratch = self._empty((2 * N,), torch.int32)
        check(self.lib.fgc_paired_input(self._p(cartoon), self._p(sketch), 0 if sketch.dtype == torch.uint8 else 1, N, R,
                                        out_hw[0], out_hw[1], int(seed) & 0xFFFFFFFFFFFFFFFF, 1 if dequantize else 0,
                                        self._p(images), self._p(sketches), self._p(scratch), self._s()), "paired_input")
        return images, sketches

    # ---------------- instance-matching model ----------------
    def affine_act(self, x, scale, shift, res=None, rscale=None, rshift=None, relu=False):
        Cc = x.shape[-1]
        y = torch.empty_like(x)
        check(self.lib.fgc_affine_act(self._p(x), self._dt(x), x.numel() // Cc, Cc, self._f32(scale), self._f32(shift),
                                      self._p(res), None if rscale is None else self._f32(rscale),
                                      None if rshift is None else self._f32(rshift), 1 if relu else 0, self._p(y), self._s()),
              "affine_act")
        return y

    def pad_cast_rows(self, x, cp, dtype):
        R, Cc = x.shape
        y = self._empty((R, cp), dtype)
        check(self.lib.fgc_pad_cast_rows(self._p(x), self._dt(x), R, Cc, cp, self._p(y), self._dt(y), self._s()), "pad_cast_rows")
        return y

    def maxpool3x3s2(self, x):
        N, H, W, Cc = x.shape
        y = self._empty((N, (H + 1) // 2, (W + 1) // 2, Cc), x.dtype)
        check(self.lib.fgc_maxpool3x3s2(self._p(x), self._dt(x), N, H, W, Cc, self._p(y), self._s()), "maxpool3x3s2")
        return y

    def space_to_batch(self, x, r):
        N, H, W, Cc = x.shape
        y = self._empty((N * r * r, H // r, W // r, Cc), x.dtype)
        check(self.lib.fgc_space_to_batch(self._p(x), self._dt(x), N, H, W, Cc, r, self._p(y), self._s()), "space_to_batch")
        return y

    def batch_to_space(self, x, r):
        NB, h, w, Cc = x.shape
        N = NB // (r * r)
        y = self._empty((N, h * r, w * r, Cc), x.dtype)
        check(self.lib.fgc_batch_to_space(self._p(x), self._dt(x), N, h, w, Cc, r, self._p(y), self._s()), "batch_to_space")
        return y

    def resize_bilinear_sigmoid(self, x, H, W):
        N, h, w, Cc = x.shape
        up, sg = self._empty((N, H, W, Cc), torch.float32), self._empty((N, H, W, Cc), torch.float32)
        check(self.lib.fgc_resize_bilinear(self._f32(x), N, h, w, Cc, H, W, self._p(up), self._p(sg), self._s()), "resize_bilinear")
        return up, sg

    # ---------------- text fusion ----------------
    def _dst(self, out, shape, dtype=torch.float32):
        if out is None:
            return self._empty(shape, dtype)
        assert tuple(out.shape) == tuple(shape) and out.dtype == dtype and out.is_contiguous()
        return out

    def l2norm_rows_fwd(self, x, out=None):
        R, D = x.shape
        y = self._dst(out, (R, D))
        inv = self._empty((R,), torch.float32)
        check(self.lib.fgc_l2norm_rows_fwd(self._f32(x), R, D, self._p(y), self._p(inv), self._s()), "l2norm_rows_fwd")
        return y, inv

    def l2norm_rows_bwd(self, gy, y, inv):
        R, D = y.shape
        gx = self._empty((R, D), torch.float32)
        check(self.lib.fgc_l2norm_rows_bwd(self._f32(gy), self._f32(y), self._f32(inv), R, D, self._p(gx), self._s()),
              "l2norm_rows_bwd")
        return gx

    def embedding_fwd(self, table, ids, t, out=None):
        N, T = ids.shape
        D = table.shape[1]
        out = self._dst(out, (N, D))
        check(self.lib.fgc_embedding_fwd(self._f32(table), self._p(ids), N, T, t, D, self._p(out), self._s()), "embedding_fwd")
        return out

    def embedding_bwd(self, g, ids, t, dtable):
        N, T = ids.shape
        D = dtable.shape[1]
        check(self.lib.fgc_embedding_bwd(self._f32(g), self._p(ids), N, T, t, D, self._f32(dtable), self._s()), "embedding_bwd")

    def embedding_all_fwd(self, table, ids):
        N, T = ids.shape
        D = table.shape[1]
        out = self._empty((T, N, D), torch.float32)
        check(self.lib.fgc_embedding_all_fwd(self._f32(table), self._p(ids), N, T, D, self._p(out), self._s()), "embedding_all_fwd")
        return out

    def embedding_all_bwd(self, g, ids, dtable):
        N, T = ids.shape
        D = dtable.shape[1]
        assert tuple(g.shape) == (T, N, D)
        check(self.lib.fgc_embedding_all_bwd(self._f32(g), self._p(ids), N, T, D, self._f32(dtable), self._s()), "embedding_all_bwd")

    def lstm_seq_supported(self, N, D):
        return D % 4 == 0 and 16 <= D <= 512 and N <= 256

    def lstm_seq_fwd(self, gx, kh, ids):
        T, N, D4 = gx.shape
        D = D4 // 4
        assert tuple(kh.shape) == (D, D4) and tuple(ids.shape) == (N, T)
        h_all, c_all = self._empty((T + 1, N, D), torch.float32), self._empty((T + 1, N, D), torch.float32)
        pre_all = self._empty((T, N, D4), torch.float32)
        bar = self._empty((16,), torch.int32)
        check(self.lib.fgc_lstm_seq_fwd(self._f32(gx), self._f32(kh), self._p(ids), T, N, D, self._p(h_all), self._p(c_all),
                                        self._p(pre_all), self._p(bar), self._s()), "lstm_seq_fwd")
        return h_all, c_all, pre_all

    def lstm_seq_bwd(self, g_hext, pre_all, c_all, kh, ids):
        T, N, D = g_hext.shape
        g_pre_all = self._empty((T, N, 4 * D), torch.float32)
        bar = self._empty((16,), torch.int32)
        check(self.lib.fgc_lstm_seq_bwd(self._f32(g_hext), self._f32(pre_all), self._f32(c_all), self._f32(kh), self._p(ids), T, N, D,
                                        self._p(g_pre_all), self._p(bar), self._s()), "lstm_seq_bwd")
        return g_pre_all

    def lstm_cell_fwd(self, gates, gates2, grow, c_prev, h_prev, ids, t, P, out_h=None, save_pre=True):
        R, D = c_prev.shape
        N, T = ids.shape
        assert R == N * P
        c = self._empty((R, D), torch.float32)
        h = self._dst(out_h, (R, D))
        pre = self._empty((R, 4 * D), torch.float32) if save_pre else None
        check(self.lib.fgc_lstm_cell_fwd(self._f32(gates), None if gates2 is None else self._f32(gates2),
                                         None if grow is None else self._f32(grow), self._f32(c_prev), self._f32(h_prev),
                                         self._p(ids), T, t, N, P, D, self._p(c), self._p(h), self._p(pre), self._s()),
              "lstm_cell_fwd")
        return c, h, pre

    def lstm_cell_bwd(self, gc, gh, pre, c_prev, c, ids, t, P, out_gpre=None):
        R, D = c_prev.shape
        N, T = ids.shape
        g_pre = self._dst(out_gpre, (R, 4 * D))
        g_c_prev = self._empty((R, D), torch.float32)
        g_h_pass = self._empty((R, D), torch.float32)
        check(self.lib.fgc_lstm_cell_bwd(self._f32(gc), self._f32(gh), self._f32(pre), self._f32(c_prev), self._p(ids), T, t, N, P,
                                         D, self._p(g_pre), self._p(g_c_prev), self._p(g_h_pass), self._s()), "lstm_cell_bwd")
        return g_pre, g_c_prev, g_h_pass

    def rows_group_sum(self, x, P, out=None):
        R, Cc = x.shape
        out = self._dst(out, (R // P, Cc))
        check(self.lib.fgc_rows_group_sum(self._f32(x), R // P, P, Cc, self._p(out), self._s()), "rows_group_sum")
        return out

    def atanh_relu_fwd(self, h):
        y = self._empty(h.shape, torch.float32)
        check(self.lib.fgc_atanh_relu_fwd(self._f32(h), h.numel(), self._p(y), self._s()), "atanh_relu_fwd")
        return y

    def atanh_relu_bwd(self, gy, h):
        gx = self._empty(h.shape, torch.float32)
        check(self.lib.fgc_atanh_relu_bwd(self._f32(gy), self._f32(h), h.numel(), self._p(gx), self._s()), "atanh_relu_bwd")
        return gx

    # ---------------- spectral norm ----------------
    def sn_fwd(self, w2d, u):
        K, Cc = w2d.shape
        wbar = self._empty((K, Cc), torch.float32)
        work = self._empty((K + 2 * Cc + 8,), torch.float32)
        uc = u.reshape(-1).clone()
        check(self.lib.fgc_sn_fwd(self._f32(w2d), self._f32(uc), K, Cc, self._p(wbar), self._p(work), self._s()), "sn_fwd")
        return wbar, dict(work=work, u=uc, u_new=work[K + Cc:K + 2 * Cc].view(1, Cc), sigma=work[K + 2 * Cc + 3])

    def sn_bwd(self, gwbar, w2d, ctx, dw):
        K, Cc = w2d.shape
        gv = self._empty((K,), torch.float32)
        check(self.lib.fgc_sn_bwd(self._f32(gwbar), self._f32(w2d), self._f32(ctx["u"]), K, Cc, self._p(ctx["work"]), self._p(gv),
                                  self._f32(dw), self._s()), "sn_bwd")

    # ---------------- losses ----------------
    def softplus_mean(self, d, sign):
        buf = self.zeros_f32((2,))
        gd = self._empty(d.shape, d.dtype)
        check(self.lib.fgc_softplus_mean(self._p(d), self._dt(d), d.numel(), float(sign), self._p(buf), 1, self._p(gd), self._s()),
              "softplus_mean")
        return buf[1], gd

    def ce_loss(self, logits, labels, focal, weight):
        N = logits.shape[0]
        Cc = logits.numel() // N
        buf = self.zeros_f32((2,))
        g = self._empty(logits.shape, logits.dtype)
        check(self.lib.fgc_ce_loss(self._p(logits), self._dt(logits), self._p(labels), N, Cc, 1 if focal else 0, float(weight),
                                   self._p(buf), 1, self._p(g), self._s()), "ce_loss")
        return buf[1], g

    def smooth_l1(self, target, gen, weight):
        assert target.dtype == gen.dtype
        buf = self.zeros_f32((2,))
        g = self._empty(gen.shape, gen.dtype)
        check(self.lib.fgc_smooth_l1(self._p(target), self._p(gen), self._dt(gen), gen.numel(), float(weight), self._p(buf), 1,
                                     self._p(g), self._s()), "smooth_l1")
        return buf[1], g

    def reg_loss(self, store):
        buf = self.zeros_f32((2,))
        check(self.lib.fgc_reg_loss(self._f32(store.flat), self._p(store.chunk_start), self._p(store.chunk_len),
                                    self._p(store.chunk_reg), store.chunk_start.numel(), self._p(buf), 1, self._s()), "reg_loss")
        return buf[1]

    # ---------------- optimiser ----------------
    def adam_step(self, store, lr, add_reg_grad=True, lr_dev=None):
        """lr_dev: optional fp32 device scalar holding lr*sqrt(1-0.9^t); then `lr` and store.adam_t are left to the caller
        (CUDA-graph replay must not bake the step size into the kernel arguments)."""
        lr_t = 0.0
        if lr_dev is None:
            store.adam_t += 1
            lr_t = lr * math.sqrt(1.0 - 0.9 ** store.adam_t)
        check(self.lib.fgc_adam_step(self._f32(store.flat), self._f32(store.grad), self._f32(store.adam_v),
                                     self._p(store.chunk_start), self._p(store.chunk_len), self._p(store.chunk_reg),
                                     store.chunk_start.numel(), float(lr_t), None if lr_dev is None else self._f32(lr_dev),
                                     0.9, 1e-8, 1 if add_reg_grad else 0, self._s()), "adam_step")

    _OPT_KIND = {'rmsprop': 1, 'adadelta': 2, 'adagrad': 3}

    def optimizer_step(self, store, kind, lr, add_reg_grad=True, lr_dev=None):
        s2 = store.opt_s2 if kind == 'adadelta' else None
        check(self.lib.fgc_opt_step(self._f32(store.flat), self._f32(store.grad), self._f32(store.adam_v),
                                    None if s2 is None else self._f32(s2), self._p(store.chunk_start), self._p(store.chunk_len),
                                    self._p(store.chunk_reg), store.chunk_start.numel(), self._OPT_KIND[kind], float(lr),
                                    None if lr_dev is None else self._f32(lr_dev), 1 if add_reg_grad else 0, self._s()), "opt_step")


def enable_op_timing(ops):
    """Debug aid: wrap every operator of a CudaOps instance with CUDA events (synchronising after each call) and
    collect per-operator totals in ops.op_times = {name: [calls, ms]}.  Used by scripts/op_breakdown.py only."""
    import functools
    ops.op_times = {}
    ops.op_detail = {}

    def wrap(name, fn):
        @functools.wraps(fn)
        def inner(*a, **k):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            r = fn(*a, **k)
            e1.record()
            e1.synchronize()
            key = name
            ms = e0.elapsed_time(e1)
            if name in ("conv_fwd", "conv_dgrad", "conv_wgrad"):
                t0 = a[0][0][0] if name != "conv_dgrad" else a[0]
                key = name + ("/f32" if t0.dtype == torch.float32 else "")
                if name == "conv_dgrad":
                    sig = "%s gy%s k%d cin%d[%d:+%d]%s" % (key, tuple(a[0].shape), a[1].shape[0], a[1].shape[2], a[2], a[3],
                                                          " ups" if k.get("ups") else "")
                else:
                    wshape = a[1].shape if name == "conv_fwd" else a[2].shape
                    sig = "%s N%d %dx%d srcs[%s] k%d cout%d" % (key, t0.shape[0], t0.shape[1] * (2 if a[0][0][1] else 1),
                                                               t0.shape[2] * (2 if a[0][0][1] else 1),
                                                               ",".join("%d%s%s" % (e[0].shape[3], "u" if e[1] else "", "p" if len(e) > 2 and e[2] is not None else "")
                                                                        for e in a[0]),
                                                               wshape[0], wshape[3])
                d = ops.op_detail.setdefault(sig, [0, 0.0])
                d[0] += 1
                d[1] += ms
            t = ops.op_times.setdefault(key, [0, 0.0])
            t[0] += 1
            t[1] += ms
            return r
        return inner

    for name in dir(OpsBase):
        if name.startswith("_"):
            continue
        fn = getattr(ops, name, None)
        if callable(fn) and name in CudaOps.__dict__:
            setattr(ops, name, wrap(name, fn))
    return ops
