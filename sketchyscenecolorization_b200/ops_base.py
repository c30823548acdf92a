"""Operator contract of the fg-colorization hot path.

The host-side network code (generator.py / discriminator.py / text_fusion.py / losses.py) is written
against this interface only.  The product implementation is `cuda_ops.CudaOps` (hand-written sm_100a
kernels behind the C-ABI of include/fgcolor.h); it raises at construction when the shared library or
a CUDA device is missing -- there is NO CPU implementation in this package.

Conventions
  * activations are NHWC, contiguous, dtype `self.act_dtype` (fp32, or bf16 in the training mode);
    2-D row tensors [R, C] are the NHWC special case H=W=1.
  * weights / biases / tables / statistics / gradients of weights are always fp32; conv weights keep the
    reference HWIO layout [k,k,Cin,Cout] (mru.py:118), a 2-D matrix [K,N] is HWIO with k=1.
  * `srcs` of a conv is a list of (tensor, ups) pairs -- or (tensor, ups, patch) triples, see `small_patch` --
    concatenated along channels in list order
    (tf.concat, mru.py:403,552,572); ups=True reads the source through a nearest-neighbour x2 upsample
    (mru.upsample, mru.py:22-28) without materialising it.
  * ops that take `acc=True` add into the given output instead of overwriting it.
"""
from __future__ import annotations

ACT_NONE, ACT_LRELU, ACT_TANH, ACT_MIU, ACT_RELU = 0, 1, 2, 3, 4     # ACT_RELU: convolution epilogues only


class OpsBase:
    act_dtype = None

    # ---------------- scheduling ----------------
    def run_aside(self, fn):
        """Run fn() so that its work MAY overlap with what the caller enqueues next (work that depends on nothing the caller is
        about to produce: spectral normalisation of the weights, the caption's word LSTM, weight gradients off the critical
        path).  Returns (fn(), join): call join() before anything consumes the results.  Default: inline, join is a no-op."""
        return fn(), (lambda: None)

    # ---------------- convolution family (mru.conv2d, mru.py:95-140) ----------------
    def conv_fwd(self, srcs, w, b, *, stride=1, act=ACT_NONE, out_dtype=None, out=None, acc=False):
        """y = act(conv2d_SAME(concat(srcs), w) + b);  w HWIO fp32, b fp32 [Cout] or None.  out= writes into the given tensor,
        with acc=True adds to it."""
        raise NotImplementedError

    def conv_dgrad(self, gy, w, c_off, c_len, *, ups=False, out=None, acc=False, out_dtype=None, gy_patch=None):
        """Gradient of a stride-1 conv w.r.t. input channels [c_off, c_off+c_len) of the concatenated input.
        ups=True: the source was read through the x2 upsample, the result is the 2x2-summed low-res gradient."""
        raise NotImplementedError

    def conv_dgrad_pooled_gy(self, g_low, w, c_off, c_len):
        """conv_dgrad(up2(g_low) / 4, w, c_off, c_len): the input gradient of a 3x3 stride-1 layer that is followed by the 2x2
        mean pool (mru.py:437-457), from the LOW-resolution output gradient; result [N, 2h, 2w, c_len].  Implementations may
        evaluate it per output phase as 2x2-tap convolutions of g_low (4 taps instead of 9, nothing upsampled in memory)."""
        return self.conv_dgrad(self.unpool_bwd(g_low), w, c_off, c_len)

    def conv_wgrad(self, srcs, gy, dw, db, *, stride=1, gy_patch=None):
        """dw += d/dw, db += sum_{n,h,w} gy  (always accumulating; dw HWIO fp32 view, db fp32 [Cout] or None)."""
        raise NotImplementedError

    # ---------------- normalisation / activations ----------------
    def chan_stats(self, x):
        """(mean[C], rstd[C]) over N,H,W; biased variance, eps 1e-5 (models_collection.batchnorm, :26,34)."""
        raise NotImplementedError

    def cbn_act_fwd(self, x, mean, rstd, scale, offset, labels, act=ACT_MIU):
        """act(scale[l_n,c] * (x-mean)*rstd + offset[l_n,c]); act in {ACT_NONE, ACT_MIU}."""
        raise NotImplementedError

    def cbn_act_bwd(self, gy, x, mean, rstd, scale, offset, labels, dscale, doffset, act=ACT_MIU, dbias=None):
        """returns gx; accumulates into dscale/doffset [25,C].  dbias (here, in prelu_bwd and in minmax_bwd): optional fp32
        [C] that receives += the column sums of the returned gradient = the bias gradient of the convolution in front."""
        raise NotImplementedError

    def colsum_(self, x, out):
        """out[C] += column sums of x[..., C]"""
        raise NotImplementedError

    def prelu_fwd(self, x, a):
        """max(a*x, x), a = fp32 scalar tensor (models_collection.prelu, :56-60)."""
        raise NotImplementedError

    def prelu_bwd(self, gy, x, a, da, dbias=None, acc_into=None):
        """returns gx; accumulates da (None => skip).  acc_into: optional tensor the gradient is ADDED to (and returned)
        instead of a new one (not together with dbias)."""
        raise NotImplementedError

    def minmax_fwd(self, x):
        """(gate, mn[N,C], mx[N,C]) with gate = (x-mn)/(mx-mn) per (n,c) over H,W, no epsilon (mru.py:415-416)."""
        raise NotImplementedError

    def minmax_bwd(self, ggate, x, mn, mx, dbias=None):
        """gradient w.r.t. the PRE-lrelu conv output: min-max backward (arg-min/arg-max routing, ties split
        evenly) times lrelu'(x) with leak 0.2; x is the post-lrelu tensor given to minmax_fwd."""
        raise NotImplementedError

    def act_bwd(self, gy, y, act):
        """gradient through an activation fused into a conv epilogue, from the OUTPUT y:
        ACT_TANH: gy*(1-y^2);  ACT_MIU: gy*0.5*(1 + x/sqrt(0.09+x^2)) with x = y - 0.0225/y."""
        raise NotImplementedError

    # ---------------- gating (mru.py:426,453,572,589) ----------------
    def gate_fma_fwd(self, ht, rg, im):
        """ht + rg*im"""
        raise NotImplementedError

    def gate_fma_bwd(self, g, rg, im):
        """(g*im, g*rg)"""
        raise NotImplementedError

    def gate_prelu_fwd(self, ht, rg, im, a):
        """prelu(ht + rg*im, a) in one pass (the discriminator's cell); the sum is not kept."""
        raise NotImplementedError

    def gate_prelu_bwd(self, gp, ht, rg, im, a, da, g_ht=None, acc=False):
        """Backward of gate_prelu_fwd from its operands: returns (g_rg, g_im); accumulates da (None => skip); g_ht (optional
        tensor) receives the gradient towards ht -- written, or added when acc."""
        raise NotImplementedError

    def mul_up_fwd(self, rg, ht_low):
        """rg * up2(ht_low)"""
        raise NotImplementedError

    def mul_up_bwd(self, g, rg, ht_low):
        """(g * up2(ht_low), sum2x2(g * rg))"""
        raise NotImplementedError

    def blend_fwd(self, sk_low, h2, zg):
        """up2(sk_low)*(1-zg) + h2*zg"""
        raise NotImplementedError

    def blend_bwd(self, g, sk_low, h2, zg):
        """(sum2x2(g*(1-zg)), g*zg, g*(h2-up2(sk_low)))"""
        raise NotImplementedError

    def addpool_fwd(self, a, b):
        """mean_pool2x2(a + b)  (mru.py:453,457)"""
        raise NotImplementedError

    def unpool_bwd(self, g):
        """up2(g)/4"""
        raise NotImplementedError

    def meanpool_fwd(self, x):
        raise NotImplementedError

    def small_patch(self, x, k, ups=False, mirror=False):
        """Optional accelerator for a NARROW conv source (C < 64): a pre-flattened (tap, channel) copy of x for kernel size k
        that TMA-fed kernels fetch instead of gathering element-wise; pass it as the third element of the source tuple.
        Returns None where it would not help (implementations are free to ignore it: the value of the conv is unchanged).
        The same works for a NARROW gradient tensor under a large filter (the 7x7, 64 -> 3 head): conv_dgrad(gy_patch=) takes
        the mirror=True patch of gy, conv_wgrad(gy_patch=) the plain one."""
        return None

    def upsample_fwd(self, x):
        """up2(x), materialised (only the weight-gradient path wants it in memory)"""
        raise NotImplementedError

    def zeros_f32(self, shape):
        raise NotImplementedError

    def add_(self, dst, src):
        """dst += src (same shape); returns dst"""
        raise NotImplementedError

    def spatial_mean_fwd(self, x):
        """[N,H,W,C] -> fp32-or-act [N,1,1,C] mean over H,W (models_collection.py:783)"""
        raise NotImplementedError

    def spatial_mean_bwd(self, g, H, W):
        """[N,1,1,C] -> [N,H,W,C] broadcast g/(H*W)"""
        raise NotImplementedError

    # ---------------- layout ----------------
    def nchw_to_nhwc(self, x, out_dtype=None):
        raise NotImplementedError

    def nhwc_to_nchw(self, x, out_dtype=None):
        raise NotImplementedError

    def cast(self, x, dtype):
        raise NotImplementedError

    # ---------------- 4x4 convolutions of the Pix2Pix / Residual variants in phase form (models_collection.py:380-405) ------
    def space_to_depth(self, x):
        """[N,2h,2w,C] -> [N,h,w,4C], channel (py*2+px)*C + c = pixel (2y+py, 2x+px) (tf.space_to_depth order)."""
        raise NotImplementedError

    def depth_to_space(self, x):
        """Inverse of space_to_depth: [N,h,w,4C] -> [N,2h,2w,C]."""
        raise NotImplementedError

    def phase_weights(self, f, mode):
        """4x4 filter -> the odd-size filter of the equivalent stride-1 SAME convolution (fp32, zero where no tap lands):
        mode 'conv'   f [4,4,C,Co] (stride 2, pad 1)              -> [3,3,4C,Co] over space_to_depth(x);
        mode 'deconv' f [4,4,Co,C] (conv2d_transpose, stride 2)   -> [3,3,C,4Co], output through depth_to_space;
        mode 'k5'     f [4,4,C,Co] (stride 1, pad 1: H -> H-1)    -> [5,5,C,Co], output cropped by one row / column."""
        raise NotImplementedError

    def phase_wgrad(self, dw, df, mode):
        """df += the entries of dw (gradient of the expanded filter) at the positions phase_weights writes."""
        raise NotImplementedError

    def tanh_fwd(self, x):
        """tanh as a pass of its own (after a batch norm: generate_residual, models_collection.py:665); its gradient is
        act_bwd(g, y, ACT_TANH)."""
        raise NotImplementedError

    def copy_rect(self, x, H, W):
        """[N,h,w,C] -> [N,H,W,C]: the overlapping top-left rectangle is copied, the rest is zero (crop or zero-pad)."""
        raise NotImplementedError

    # ---------------- instance-matching model (Instance_Matching/deeplab_model.py, RMI_model.py) ----------------
    def affine_act(self, x, scale, shift, res=None, rscale=None, rshift=None, relu=False):
        """act(x*scale[c] + shift[c] + r) with r = 0 | res | res*rscale[c] + rshift[c]: stored-moment batch norm
        (deeplab_model._batch_norm, :213-233), the residual sum and relu of _bottleneck_residual (:237-264)."""
        raise NotImplementedError

    def pad_cast_rows(self, x, cp, dtype):
        """[R, C] -> [R, cp] (cp >= C, zero columns appended) in `dtype`."""
        raise NotImplementedError

    def maxpool3x3s2(self, x):
        """tf.nn.max_pool 3x3, stride 2, SAME (deeplab_model.py:72)."""
        raise NotImplementedError

    def space_to_batch(self, x, r):
        """[N,H,W,C] -> [r*r*N, H/r, W/r, C], batch (py*r + px)*N + n = pixel phase (py, px) of sample n: a SAME convolution
        of the result is tf.nn.atrous_conv2d(x, rate=r) in the same form (deeplab_model.py:289-291)."""
        raise NotImplementedError

    def batch_to_space(self, x, r):
        raise NotImplementedError

    def resize_bilinear_sigmoid(self, x, H, W):
        """(tf.image.resize_bilinear(x, [H, W]), its sigmoid) of fp32 NHWC x (RMI_model.py:150-151)."""
        raise NotImplementedError

    # ---------------- real-data input (input_pipeline.get_paired_input, :72-126) ----------------
    def paired_input(self, cartoon, sketch, out_hw, seed=0, dequantize=True):
        """cartoon uint8 [N,R,R,3], sketch uint8 | fp32 (0..255 distance map) [N,R,R,3] on the op device ->
        (images, sketches) fp32 NCHW [N,3,H,W] in [-1,1]: image picked at the integer resize factor (TF-1 BILINEAR),
        min-max normalised per picture, + U[0,1/256) noise, sketch block mean (AREA)."""
        raise NotImplementedError

    # ---------------- text fusion (models_collection.encode_feat_with_text, :150-248) ----------------
    def l2norm_rows_fwd(self, x, out=None):
        """rows [R,D] fp32: (x * rsqrt(max(sum x^2,1e-12)), inv[R]).  `out`: optional destination of the normalised rows
        (a slot of a per-time-step stack, so that the weight gradients can run once over all steps)."""
        raise NotImplementedError

    def l2norm_rows_bwd(self, gy, y, inv):
        raise NotImplementedError

    def embedding_fwd(self, table, ids, t, out=None):
        """table[ids[:, t]] -> [N,D];  ids int32 [N,T] on the device; `out`: optional destination"""
        raise NotImplementedError

    def embedding_bwd(self, g, ids, t, dtable):
        """dtable[ids[n,t]] += g[n]"""
        raise NotImplementedError

    def embedding_all_fwd(self, table, ids):
        """table[ids[n, t]] for every step: [T, N, D]"""
        raise NotImplementedError

    def embedding_all_bwd(self, g, ids, dtable):
        """dtable[ids[n, t]] += g[t, n] over all steps"""
        raise NotImplementedError

    def lstm_seq_supported(self, N, D):
        """Whether lstm_seq_fwd / lstm_seq_bwd take this batch / hidden size (else the caller steps the cell itself)."""
        return False

    def lstm_seq_fwd(self, gx, kh, ids):
        """A whole BasicLSTMCell recurrence from the zero state (the word LSTM, models_collection.py:173-213): gx [T,N,4D] is the
        input half of the gate pre-activations (x_t @ kernel[0:Din] + bias) of every step, kh [D,4D] the recurrent rows of the
        kernel; <pad> steps (ids[n,t] == 0) keep the state.  Returns (h_all [T+1,N,D], c_all [T+1,N,D], pre_all [T,N,4D]);
        slot 0 of the state stacks is the zero initial state, slot t+1 the state after step t."""
        raise NotImplementedError

    def lstm_seq_bwd(self, g_hext, pre_all, c_all, kh, ids):
        """BPTT through lstm_seq_fwd: g_hext [T,N,D] is the gradient reaching h(t) from outside the recurrence; returns the
        gate gradients of every step [T,N,4D] (the operand of the weight / input gradients)."""
        raise NotImplementedError

    def lstm_cell_fwd(self, gates, gates2, grow, c_prev, h_prev, ids, t, P, out_h=None, save_pre=True):
        """BasicLSTMCell pointwise part.  pre = gates [+ gates2] [+ grow[r // P]]  ([R,4D], order i,j,f,o);
        c = c_prev*sig(f+1) + sig(i)*tanh(j); h = tanh(c)*sig(o); rows whose sample token ids[r//P, t] == 0 (<pad>) keep
        (c_prev, h_prev).  Returns (c, h, pre); `out_h`: optional destination of h; save_pre=False (inference) returns pre = None."""
        raise NotImplementedError

    def lstm_cell_bwd(self, gc, gh, pre, c_prev, c, ids, t, P, out_gpre=None):
        """returns (g_pre [R,4D], g_c_prev [R,D], g_h_pass [R,D]) where g_h_pass = gh on masked rows else 0;
        `out_gpre`: optional destination of g_pre."""
        raise NotImplementedError

    def rows_group_sum(self, x, P, out=None):
        """[N*P, C] -> [N, C] sum over each group of P consecutive rows; `out`: optional destination"""
        raise NotImplementedError

    def atanh_relu_fwd(self, h):
        """relu(0.5*(log(1.001+h) - log(1.001-h)))  (models_collection.py:239-241)"""
        raise NotImplementedError

    def atanh_relu_bwd(self, gy, h):
        raise NotImplementedError

    # ---------------- spectral norm (sn.py:12-52) ----------------
    def sn_fwd(self, w2d, u):
        """w2d [K,C] fp32, u [1,C] -> (w_bar [K,C], ctx) with ctx holding v, u_new, sigma and norms."""
        raise NotImplementedError

    def sn_bwd(self, gwbar, w2d, ctx, dw):
        """dw += dL/dW given gwbar = dL/dW_bar (differentiates through sigma and the power iteration)."""
        raise NotImplementedError

    # ---------------- losses (graph_single.py:340-353,388-402,552-555) ----------------
    def softplus_mean(self, d, sign):
        """(mean(softplus(sign*d)) as fp32 0-d tensor, gradient w.r.t. d in d's dtype)"""
        raise NotImplementedError

    def ce_loss(self, logits, labels, focal, weight):
        """weight * mean_n( (1-p_t)^2 if focal else 1) * CE_n ) and its gradient w.r.t. logits [N,1,1,C]."""
        raise NotImplementedError

    def smooth_l1(self, target, gen, weight):
        """weight*mean(smooth_l1(target-gen)) and gradient w.r.t. gen."""
        raise NotImplementedError

    def reg_loss(self, store):
        """sum_i reg_i * sum(w_i^2)/2 over a ParamStore (fp32 0-d tensor)."""
        raise NotImplementedError

    # ---------------- optimiser (graph_single.py:584-593, tf.train.AdamOptimizer(beta1=0, beta2=0.9)) ----
    def adam_step(self, store, lr, add_reg_grad=True, lr_dev=None):
        """g += reg*w; v = .9v+.1g^2; w -= lr*sqrt(1-.9^t)*g/(sqrt(v)+1e-8); increments store.adam_t first."""
        raise NotImplementedError

    def optimizer_step(self, store, kind, lr, add_reg_grad=True, lr_dev=None):
        """The reference's other optimisers (graph_single.get_optimizer, :584-593) as one fused pass over the flat buffers;
        kind 'rmsprop' (decay 0.9, momentum 0, eps 1e-10; store.adam_v is the rms slot, initialised to 1), 'adadelta'
        (rho 0.95, eps 1e-8; slots store.adam_v and store.opt_s2, zeros), 'adagrad' (store.adam_v = accumulator, initialised
        to 0.1).  lr_dev: optional device scalar overriding lr (CUDA-graph replay)."""
        raise NotImplementedError
