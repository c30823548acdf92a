/* fgcolor.h -- C-ABI of libfgcolor.so: hand-written sm_100a kernels for the
 * SketchySceneColorization foreground-instance-colorization hot path.
 *
 * The reference (TensorFlow-1, pure Python) has no FFI of its own; the seam is the set of TF ops its layer
 * library calls.  Each entry point below replaces one such call site (cited as file:line, paths relative to
 * Foreground_Instance_Colorization/obj_lib/).  The Python host code binds these with ctypes
 * (sketchyscenecolorization_b200/_lib.py); INTEGRATION.md shows the stub a reference maintainer would add.
 *
 * Conventions
 *   - plain pointers + sizes only; all pointers are DEVICE pointers owned by the caller and must stay valid
 *     until `stream` reaches the end of the call; nothing is allocated, nothing synchronises.
 *   - every call returns 0 on success or a negative FGC_E* code; fgc_last_error() gives the message
 *     (thread-local).
 *   - activations are NHWC; `dtype` 0 = fp32, 1 = bf16 (the storage type of the tensor behind a void*).
 *     Weights, biases, tables, statistics and weight gradients are always fp32; conv weights are HWIO.
 */
#ifndef FGCOLOR_H_
#define FGCOLOR_H_
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void* fgc_stream;     /* cudaStream_t */

enum { FGC_OK = 0, FGC_EINVAL = -1, FGC_ECUDA = -2, FGC_EUNSUPPORTED = -3 };
enum { FGC_F32 = 0, FGC_BF16 = 1 };
enum { FGC_ACT_NONE = 0, FGC_ACT_LRELU = 1, FGC_ACT_TANH = 2, FGC_ACT_MIU = 3, FGC_ACT_RELU = 4 /* convolution epilogues only */ };

const char* fgc_last_error(void);
int fgc_version(void);
/* number of kernel launches issued through this library by the calling process (bench.py `gpu_launches`) */
long long fgc_launch_count(void);
/* host-side CRC-32C (Castagnoli) of data[0:n) continuing from crc (0 = fresh), unmasked: the checksum TensorFlow's
 * tensor-bundle snapshot files carry (tf.train.Saver V2; main_procedure.py:141,235-237) */
unsigned int fgc_crc32c(const void* data, size_t n, unsigned int crc);

/* ---------------------------------------------------------------------------------------------------------
 * Convolution family.  Replaces tf.nn.conv2d(NCHW, SAME)+bias(+activation) in mru.conv2d (mru.py:95-140),
 * tf.matmul in mru.fully_connected (mru.py:75) and BasicLSTMCell (models_collection.py:184-187), the
 * tf.concat of conv inputs (mru.py:403,552,572; models_collection.py:323-353) and mru.upsample
 * (mru.py:22-28) in front of a conv, plus the gradients TF autodiff derives for them.
 * -------------------------------------------------------------------------------------------------------*/
typedef struct {
  const void* ptr;   /* NHWC [N, H/(ups?2:1), W/(ups?2:1), C] */
  int C;
  int ups;           /* 1: read through nearest-neighbour x2 upsample */
  const void* patch; /* optional (NULL = none), narrow sources only: bf16 [N,H,W,8*ceil(k*k*C/8)] written by
                      * fgc_im2col_small for this conv's k -- lets the TMA-fed kernels fetch the source's flattened
                      * (tap, channel) slabs as ordinary 64-channel blocks instead of gathering them element-wise */
} fgc_src;
/* out[n,h,w,q] = x[n, h+kh-pad, w+kw-pad, c] for q = (kh*k+kw)*C + c < k*k*C, else 0 (stride 1, SAME); out is bf16 with
 * 8*ceil(k*k*C/8) channels (the TMA boxes of the consumers are 64 channels wide; the rest is out-of-bounds zero fill).  ups = 1 reads x through the x2 nearest-neighbour upsample. */
int fgc_im2col_small(const void* x, int dtype, int N, int H, int W, int C, int ups, int k, int mirror, void* out,
                     fgc_stream stream);
/* mirror = 1 flattens the mirrored taps, out[.., q] = x[n, h-(kh-pad), w-(kw-pad), c]: the patch tensor of a NARROW gradient
 * tensor for fgc_conv2d_dgrad (gy_patch), whose taps run mirrored; fgc_conv2d_wgrad's gy_patch takes mirror = 0. */

/* Bytes of device workspace `ws` a forward (n_out = Cout, sources = the conv inputs) or input-gradient
 * (n_out = c_len, one source of Cout channels) call needs for its packed bf16 weight tiles. */
size_t fgc_conv2d_ws_bytes(const int* src_C, int nsrc, int k, int n_out, int src_dtype);
/* 0 = tcgen05 tensor-core path (default), 1 = CUDA-core checker (also: env FGC_CONV_IMPL=simple) */
int fgc_set_conv_impl(int impl);
/* tuning / A-B switches of the tensor-core path (also: env FGC_HALO, FGC_SMALL): halo = 1 routes stride-1 SAME layers
 * with wide bf16 sources through the halo-reuse kernels (tensor-map TMA), 0 through the per-tap gather kernels
 * (2: also 1x1 layers; 3: force the 64 x 8 pixel tile variant wherever it fits -- test coverage on small problems);
 * small = 1 routes the narrow stem-level layers (all sources < 64 channels, <= 8 outputs) through the CUDA-core
 * direct kernels.  A negative value leaves that switch unchanged. */
int fgc_set_conv_flags(int halo, int small);
/* launches so far per convolution kernel family: out[0] halo-reuse fwd/dgrad, out[1] per-tap gather fwd/dgrad,
 * out[2] direct narrow fwd/dgrad, out[3] direct narrow wgrad, out[4] per-tap tensor-core wgrad, out[5] halo-reuse wgrad
 * (tests assert the routing) */
int fgc_debug_conv_counts(long long out[6]);
/* debug aid: per-role event trace (role, event, tile, clock64) of CTA 0 of the implicit-GEMM kernel; buf holds
 * 4 + 4*capacity int64 on the device, buf[0] is the event count.  NULL disables. */
int fgc_debug_set_trace(long long* buf, int capacity);
/* measurement aid: on = 1 makes the forward / input-gradient calls skip their weight-packing launch -- the caller guarantees
 * that `ws` still holds the tiles an identical earlier call packed there.  bench.py times the dominant convolution kernel
 * alone with it (the roofline figure is per kernel launch); the product path never sets it. */
int fgc_debug_keep_packed(int on);

/* y[N,OH,OW,Cout] = act(conv(concat_c(srcs), w) + bias); SAME padding (pad_t/pad_l = TF's top/left pad).
 * w: fp32 HWIO [k,k,Cin_total,Cout]; bias fp32 [Cout] or NULL.  fp32 sources run the bf16x3 split-accumulate
 * tensor-core mode, bf16 sources single-pass bf16; accumulation is fp32 in TMEM either way. */
int fgc_conv2d_fwd(const fgc_src* srcs, int nsrc, int src_dtype, int N, int H, int W,
                   const float* w, int k, int Cin_total, int Cout, const float* bias,
                   int stride, int pad_t, int pad_l, int OH, int OW, int act,
                   void* y, int y_dtype, void* ws, fgc_stream stream);
/* The same with flags in `accumulate`: bit 0 = add into y (y += act(conv + bias): the extra passes of the six-product mode);
 * bit 1 = only the centre filter column of w is non-zero (see fgc_tapsum_w). */
int fgc_conv2d_fwd_acc(const fgc_src* srcs, int nsrc, int src_dtype, int N, int H, int W,
                       const float* w, int k, int Cin_total, int Cout, const float* bias,
                       int stride, int pad_t, int pad_l, int OH, int OW, int act, int accumulate,
                       void* y, int y_dtype, void* ws, fgc_stream stream);
/* One PHASE of a convolution whose result lives on a 2x finer grid: y [N, 2H, 2W, Cout] receives, at its pixels
 * (2h + dy, 2w + dx), the stride-1 SAME convolution of the (bf16) sources with w (fp32 HWIO [k,k,Cin_total,Cout]); only the
 * filter columns / rows whose bit is set in kw_mask / kh_mask carry weights (the rest of w is zero by contract and is
 * skipped: no operand fetch, no MMA).  Use: the input gradient of a 3x3 layer whose OUTPUT gradient is the x2 nearest-
 * neighbour upsample of a low-resolution tensor (the cell's Conv_2 under mean_pool, mru.py:437-457) is, per output phase, a
 * 2x2-tap convolution of the low-resolution gradient with pre-summed filters -- 4 taps instead of 9 and no full-resolution
 * gradient in memory.  Halo-reuse kernel only: FGC_EUNSUPPORTED (nothing launched) when the layer does not qualify. */
int fgc_conv2d_fwd_phase(const fgc_src* srcs, int nsrc, int src_dtype, int N, int H, int W, const float* w, int k, int Cin_total,
                         int Cout, int kw_mask, int kh_mask, int dy, int dx, void* y, int y_dtype, void* ws, fgc_stream stream);
/* Second half of a column-folded k x k convolution with few outputs (models_collection.py:372-374, the 7x7 64 -> 3 head):
 * z [N,H,W,Cz] fp32 holds, in channel kw*Cout + co, the k x 1 VERTICAL convolution with filter column kw (computed by
 * fgc_conv2d_fwd_acc with flag 2 on a filter whose other columns are zero); y[n,h,w,co] = act(bias[co] +
 * sum_kw z[n,h,w + kw - (k-1)/2, kw*Cout + co]) with zero padding in w.  On the tensor path a 3-output 7x7 layer is 49
 * N = 16 instructions per K step; folded it is 7 N = 32 ones. */
int fgc_tapsum_w(const float* z, int N, int H, int W, int k, int Cout, int Cz, const float* bias, int act, void* y, int y_dtype,
                 fgc_stream stream);
/* Term `level` (0, 1, 2) of the three-way bf16 split of an fp32 tensor: r = x minus its first `level` bf16 terms;
 * out_bf16 (optional) = bf16(r), out_f32 (optional) = r.  With bf16x3 (fp32 sources, the parity mode of mru.conv2d) these
 * give the six-product mode: x1w1 + x1w2 + x2w1 (one bf16x3 pass) + x2w2 + x1w3 + x3w1 (three accumulating bf16 passes). */
int fgc_split_term(const float* x, long long n, int level, void* out_bf16, float* out_f32, fgc_stream stream);

/* gx[N,H,W,c_len] (=|+=) d/d(input channels [c_off,c_off+c_len)) of a stride-1 SAME conv given gy[N,H,W,Cout].
 * ups=1: gx is the 2x2-summed low-res gradient [N,H/2,W/2,c_len]; `scratch` must then hold N*H*W*c_len floats. */
int fgc_conv2d_dgrad(const void* gy, int gy_dtype, int N, int H, int W, const float* w, int k, int Cin_total,
                     int Cout, int c_off, int c_len, int ups, int accumulate, void* gx, int gx_dtype,
                     void* scratch, void* ws, const void* gy_patch /*NULL ok: mirrored patch tensor of a narrow gy*/,
                     fgc_stream stream);

/* dw[k,k,Cin_total,Cout] += sum_m x[m+tap,ci]*gy[m,co];  db[Cout] += sum_m gy (db may be NULL). */
int fgc_conv2d_wgrad(const fgc_src* srcs, int nsrc, int src_dtype, int N, int H, int W,
                     const void* gy, int gy_dtype, int k, int Cin_total, int Cout,
                     int stride, int pad_t, int pad_l, int OH, int OW,
                     float* dw, float* db, const void* gy_patch /*NULL ok: patch tensor of a narrow gy under a large filter
                     (the 7x7, 64 -> 3 head): the kernel then reduces over gy's taps instead of gathering x 49 times*/,
                     fgc_stream stream);

/* ---------------------------------------------------------------------------------------------------------
 * Conditional batch-norm with batch statistics (models_collection.batchnorm, :22-34) fused with miu_relu
 * (:63-65); PReLU (:56-60); spatial min-max gate normalisation (mru.py:415-416,560-561,568-569).
 * -------------------------------------------------------------------------------------------------------*/
/* stats[0:C]=mean, stats[C:2C]=rstd over M=N*H*W rows; acc: 2*C doubles of scratch (zeroed by the call). */
int fgc_chan_stats(const void* x, int dtype, long long M, int C, double* acc, float* stats, fgc_stream stream);
int fgc_cbn_act_fwd(const void* x, int dtype, int N, int HW, int C, const float* stats, const float* scale,
                    const float* offset, const int32_t* labels, int act, void* y, fgc_stream stream);
/* scratch: 2*N*C + 2*C floats.  dscale/doffset [n_labels,C] accumulate.
 * dbias (NULL ok, here and in fgc_prelu_bwd / fgc_minmax_bwd): fp32 [C], += column sums of the gradient written -- the
 * bias gradient of the convolution whose output this normalisation / activation consumed (mru.conv2d adds the bias before
 * the normaliser, mru.py:125-140), so that no separate pass over the gradient tensor is needed for it. */
int fgc_cbn_act_bwd(const void* gy, const void* x, int dtype, int N, int HW, int C, const float* stats,
                    const float* scale, const float* offset, const int32_t* labels, int act,
                    float* dscale, float* doffset, void* gx, float* scratch, float* dbias, fgc_stream stream);
int fgc_prelu_fwd(const void* x, int dtype, long long n, const float* a, void* y, fgc_stream stream);
/* C: channel count of the NHWC tensor (only used when dbias != NULL) */
int fgc_prelu_bwd(const void* gy, const void* x, int dtype, long long n, int C, const float* a, float* da /*NULL ok*/,
                  float* dbias /*NULL ok*/, void* gx, fgc_stream stream);
/* the same without the bias-gradient output, ADDED into gx (the hidden-state gradient of a cell already holds its other
 * branches: one pass less than prelu_bwd followed by an add) */
int fgc_prelu_bwd_acc(const void* gy, const void* x, int dtype, long long n, const float* a, float* da /*NULL ok*/, void* gx,
                      fgc_stream stream);
/* out[C] += column sums of x[M,C] (bias gradients that no fused kernel produces) */
int fgc_colsum(const void* x, int dtype, long long M, int C, float* out, fgc_stream stream);
/* gate=(x-mn)/(mx-mn) per (n,c); mn,mx fp32 [N,C]; scratch: 2*N*C uint32. */
int fgc_minmax_fwd(const void* x, int dtype, int N, int HW, int C, void* gate, float* mn, float* mx,
                   uint32_t* scratch, fgc_stream stream);
/* gradient w.r.t. the pre-lrelu conv output (leak 0.2); scratch: 4*N*C floats. */
int fgc_minmax_bwd(const void* ggate, const void* x, int dtype, int N, int HW, int C, const float* mn,
                   const float* mx, void* gpre, float* scratch, float* dbias /*NULL ok*/, fgc_stream stream);
/* y = tanh(x) as a pass of its own (after a batch norm: generate_residual, models_collection.py:665); gradient: fgc_act_bwd */
int fgc_tanh_fwd(const void* x, int dtype, long long n, void* y, fgc_stream stream);
/* gradient through an activation fused in a conv epilogue, from its output y (tanh / miu_relu). */
int fgc_act_bwd(const void* gy, const void* y, int dtype, long long n, int act, void* gx, fgc_stream stream);

/* ---------------------------------------------------------------------------------------------------------
 * Gating / resampling: ht + rg*img (mru.py:426), rg*up(ht) (:572), ht*(1-zg)+h*zg (:589),
 * mean_pool(ht_orig+h_new) (:453,457 + :15-19), mean_pool pyramids (models_collection.py:84-86,268-272,697-699),
 * reduce_mean over H,W (:783).
 * -------------------------------------------------------------------------------------------------------*/
int fgc_gate_fma_fwd(const void* ht, const void* rg, const void* im, int dtype, long long n, void* out, fgc_stream s);
int fgc_gate_fma_bwd(const void* g, const void* rg, const void* im, int dtype, long long n, void* g_rg, void* g_im, fgc_stream s);
/* the same gate followed by PReLU in one pass (the discriminator's cell: norm_activ = prelu, mru.py:426-430,
 * models_collection.py:56-60): out = prelu(ht + rg*im, *a).  Backward from the three operands (the sum is recomputed):
 * g_x = gp * prelu'(x); *da += sum gp*x over the leaky side (da may be NULL); g_rg = g_x*im; g_im = g_x*rg;
 * g_ht (may be NULL) = g_x, or += g_x when acc_ht. */
int fgc_gate_prelu_fwd(const void* ht, const void* rg, const void* im, int dtype, long long n, const float* a, void* out,
                       fgc_stream s);
int fgc_gate_prelu_bwd(const void* gp, const void* ht, const void* rg, const void* im, int dtype, long long n, const float* a,
                       float* da, void* g_ht, int acc_ht, void* g_rg, void* g_im, fgc_stream s);
/* full-res tensors are [N,2h,2w,C], low-res [N,h,w,C] */
int fgc_mul_up_fwd(const void* rg, const void* ht_low, int dtype, int N, int h, int w, int C, void* out, fgc_stream s);
int fgc_mul_up_bwd(const void* g, const void* rg, const void* ht_low, int dtype, int N, int h, int w, int C,
                   void* g_rg, void* g_ht_low, fgc_stream s);
int fgc_blend_fwd(const void* sk_low, const void* h2, const void* zg, int dtype, int N, int h, int w, int C, void* out, fgc_stream s);
int fgc_blend_bwd(const void* g, const void* sk_low, const void* h2, const void* zg, int dtype, int N, int h, int w, int C,
                  void* g_sk_low, void* g_h2, void* g_zg, fgc_stream s);
/* out[N,h,w,C] = mean2x2(a + b); b may be NULL (plain mean_pool) */
int fgc_addpool_fwd(const void* a, const void* b, int dtype, int N, int h, int w, int C, void* out, fgc_stream s);
/* out[N,2h,2w,C] = up2(g)/4 */
int fgc_unpool_bwd(const void* g, int dtype, int N, int h, int w, int C, void* out, fgc_stream s);
/* out[N,2h,2w,C] = nearest-neighbour x2 upsample of x (mru.upsample, mru.py:22-28), materialised: only the weight-gradient
 * kernel wants it in memory (its TMA tiles cannot replicate pixels); forward convs read the low-res tensor directly */
int fgc_upsample2x(const void* x, int dtype, int N, int h, int w, int C, void* out, fgc_stream s);
/* out[N,h,w,C] (=|+=) sum over each 2x2 block of g[N,2h,2w,C]  (gradient of mru.upsample, mru.py:22-28) */
int fgc_sum2x2(const void* g, int g_dtype, int N, int h, int w, int C, void* out, int out_dtype, int accumulate, fgc_stream s);
int fgc_axpy(void* dst, const void* src, int dst_dtype, int src_dtype, long long n, float alpha, fgc_stream s); /* dst += alpha*src */
int fgc_spatial_mean_fwd(const void* x, int dtype, int N, int HW, int C, void* out, fgc_stream s);
int fgc_spatial_mean_bwd(const void* g, int dtype, int N, int HW, int C, void* out, fgc_stream s);
int fgc_nchw_to_nhwc(const void* x, int x_dtype, int N, int C, int HW, void* y, int y_dtype, fgc_stream s);
int fgc_nhwc_to_nchw(const void* x, int x_dtype, int N, int C, int HW, void* y, int y_dtype, fgc_stream s);
int fgc_cast(const void* x, int x_dtype, void* y, int y_dtype, long long n, fgc_stream s);

/* ---------------------------------------------------------------------------------------------------------
 * Caption encoder pieces (models_collection.encode_feat_with_text, :150-248): tf.nn.l2_normalize (:202,216),
 * embedding_lookup (:182), BasicLSTMCell gates (:213,226) with the tf.cond pad skip (:235), the
 * 0.5*(log(1.001+h)-log(1.001-h)) -> relu output transform (:239-241).  All fp32.
 * -------------------------------------------------------------------------------------------------------*/
int fgc_l2norm_rows_fwd(const float* x, int R, int D, float* y, float* inv, fgc_stream s);
int fgc_l2norm_rows_bwd(const float* gy, const float* y, const float* inv, int R, int D, float* gx, fgc_stream s);
int fgc_embedding_fwd(const float* table, const int32_t* ids, int N, int T, int t, int D, float* out, fgc_stream s);
int fgc_embedding_bwd(const float* g, const int32_t* ids, int N, int T, int t, int D, float* dtable, fgc_stream s);
/* models_collection.py:182,211 for every time step at once: out[t, n, :] = table[ids[n, t], :]  ([T, N, D]) and its scatter-add */
int fgc_embedding_all_fwd(const float* table, const int32_t* ids, int N, int T, int D, float* out, fgc_stream stream);
int fgc_embedding_all_bwd(const float* g, const int32_t* ids, int N, int T, int D, float* dtable, fgc_stream stream);
/* models_collection.py:173-213, the word LSTM (BasicLSTMCell :184, dynamic over T tokens, <pad> steps skipped :235) as ONE
 * persistent launch.  gx [T, N, 4D]: the input half of the gate pre-activations, x_t @ kernel[0:Din] + bias, for every step;
 * kh [D, 4D]: the recurrent rows kernel[Din:Din+D] (row-major, gate order i, j, f, o); ids [N, T] (0 = <pad>: state kept).
 * Writes h_all / c_all [T+1, N, D] (slot 0 = the zero initial state, slot t+1 = state after step t) and pre_all [T, N, 4D].
 * barrier: 4 bytes of device scratch.  D: multiple of 4, 16..512; N <= 256.  The backward call runs BPTT over the same
 * sequence: g_hext [T, N, D] = gradient arriving at h(t) from outside the recurrence -> g_pre_all [T, N, 4D]. */
int fgc_lstm_seq_fwd(const float* gx, const float* kh, const int32_t* ids, int T, int N, int D, float* h_all, float* c_all,
                     float* pre_all, unsigned int* barrier, fgc_stream stream);
int fgc_lstm_seq_bwd(const float* g_hext, const float* pre_all, const float* c_all, const float* kh, const int32_t* ids, int T,
                     int N, int D, float* g_pre_all, unsigned int* barrier, fgc_stream stream);
/* pre = gates (+gates2) (+grow[r/P]); R = N*P rows of 4*D; gates2/grow may be NULL; pre may be NULL (inference: not kept) */
int fgc_lstm_cell_fwd(const float* gates, const float* gates2, const float* grow, const float* c_prev,
                      const float* h_prev, const int32_t* ids, int T, int t, int N, int P, int D,
                      float* c, float* h, float* pre, fgc_stream s);
int fgc_lstm_cell_bwd(const float* gc, const float* gh, const float* pre, const float* c_prev,
                      const int32_t* ids, int T, int t, int N, int P, int D,
                      float* g_pre, float* g_c_prev, float* g_h_pass, fgc_stream s);
int fgc_rows_group_sum(const float* x, int N, int P, int C, float* out, fgc_stream s);
int fgc_atanh_relu_fwd(const float* h, long long n, float* y, fgc_stream s);
int fgc_atanh_relu_bwd(const float* gy, const float* h, long long n, float* gx, fgc_stream s);

/* ---------------------------------------------------------------------------------------------------------
 * Spectral normalisation (sn.spectral_normed_weight, sn.py:12-52; one power iteration, gradient through
 * sigma and through the iteration).  w [K,C]; u [C]; work: K + 2*C + 8 floats laid out as
 * a[K] | b[C] | u_new[C] | scal[8] = {na^2 acc, na, nb, sigma, dot acc, cb, s acc, -}.
 * -------------------------------------------------------------------------------------------------------*/
int fgc_sn_fwd(const float* w, const float* u, int K, int C, float* wbar, float* work, fgc_stream s);
int fgc_sn_bwd(const float* gwbar, const float* w, const float* u, int K, int C, float* work /*from fwd*/,
               float* gv /*K floats scratch*/, float* dw, fgc_stream s);

/* ---------------------------------------------------------------------------------------------------------
 * Losses (graph_single.get_losses, :317-581) and optimiser (graph_single.py:139-142,588).
 * Every loss call adds weight*loss to lossbuf[slot] and to lossbuf[0]; gradients are w.r.t. the logits/images.
 * -------------------------------------------------------------------------------------------------------*/
int fgc_softplus_mean(const void* d, int dtype, long long n, float sign, float* lossbuf, int slot, void* gd, fgc_stream s); /* :401-402 */
int fgc_ce_loss(const void* logits, int dtype, const int32_t* labels, int N, int C, int focal, float weight,
                float* lossbuf, int slot, void* glogits, fgc_stream s);                                                     /* :343-352 */
int fgc_smooth_l1(const void* target, const void* gen, int dtype, long long n, float weight, float* lossbuf, int slot,
                  void* ggen, fgc_stream s);                                                                                /* :552-555 */
/* chunk table over a flat parameter buffer: chunk i covers [start[i], start[i]+len[i]) with l2 scale reg[i] */
int fgc_reg_loss(const float* flat, const long long* start, const int32_t* len, const float* reg, int nchunks,
                 float* lossbuf, int slot, fgc_stream s);                                                                   /* :570-576 */
/* g += reg*w (if add_reg); v = b2 v + (1-b2) g^2; w -= lr_t * g / (sqrt(v)+eps), lr_t = lr*sqrt(1-b2^t) (beta1 = 0).
 * lr_t_dev (device pointer, may be NULL) overrides lr_t: lets a captured CUDA graph be replayed with a new step size. */
int fgc_adam_step(float* flat, float* grad, float* v, const long long* start, const int32_t* len, const float* reg,
                  int nchunks, float lr_t, const float* lr_t_dev, float beta2, float eps, int add_reg, fgc_stream s);

/* The reference's other optimisers (graph_single.get_optimizer, :584-593) over the same chunk table: kind 1 =
 * RMSPropOptimizer(decay 0.9, momentum 0, epsilon 1e-10), s1 = rms (the caller initialises it to 1 as TF does); kind 2 =
 * AdadeltaOptimizer (rho 0.95, epsilon 1e-8), s1 = accum, s2 = accum_update; kind 3 = AdagradOptimizer, s1 = accumulator
 * (initial value 0.1).  g += reg*w first (if add_reg); lr_dev (device scalar, may be NULL) overrides lr. */
int fgc_opt_step(float* flat, float* grad, float* s1, float* s2 /*kind 2 only*/, const long long* start, const int32_t* len,
                 const float* reg, int nchunks, int kind, float lr, const float* lr_dev, int add_reg, fgc_stream s);

/* ---------------------------------------------------------------------------------------------------------
 * Layout passes that put the 4x4 layers of the Pix2Pix / Residual variants (models_collection.nchw_conv, :380-391;
 * nchw_deconv = tf.nn.conv2d_transpose, :394-405) on the stride-1 SAME convolutions above ("phase form"):
 *   4x4 stride 2 pad 1        = 3x3 SAME over space_to_depth(x);
 *   conv2d_transpose 4x4 s2   = depth_to_space of a 3x3 SAME convolution to 4*Co channels;
 *   4x4 stride 1 pad 1        = 5x5 SAME (first filter row / column zero), cropped by one row and column.
 * Depth channel order is tf.space_to_depth's: (py*2+px)*C + c  <->  pixel (2y+py, 2x+px).
 * -------------------------------------------------------------------------------------------------------*/
int fgc_space_to_depth(const void* x /*[N,2h,2w,C]*/, int dtype, int N, int h, int w, int C, void* out /*[N,h,w,4C]*/, fgc_stream s);
int fgc_depth_to_space(const void* x /*[N,h,w,4C]*/, int dtype, int N, int h, int w, int C, void* out /*[N,2h,2w,C]*/, fgc_stream s);
/* out[N,H,W,C] = the top-left min(h,H) x min(w,W) rectangle of x[N,h,w,C], zero elsewhere (crop, or zero-pad a gradient) */
int fgc_copy_rect(const void* x, int dtype, int N, int h, int w, int C, void* out, int H, int W, fgc_stream s);
/* f[4,4,A,B] fp32 -> w: mode 0 (conv, A = Cin, B = Cout) [3,3,4A,B]; mode 1 (transposed conv, TF filter layout
 * [kh,kw,out,in]: A = Cout, B = Cin) [3,3,B,4A]; mode 2 (stride-1 conv) [5,5,A,B].  Taps that do not land are zero. */
int fgc_phase_weights(const float* f, int A, int B, int mode, float* w, fgc_stream s);
/* df[4,4,A,B] += the entries of dw at the positions fgc_phase_weights writes (adjoint of the scatter) */
int fgc_phase_wgrad(const float* dw, int A, int B, int mode, float* df, fgc_stream s);

/* ---------------------------------------------------------------------------------------------------------
 * Real-data input path (input_pipeline.get_paired_input, :72-126; the queues of :131-181 batch its outputs):
 * raw record payloads -> the tensors the graph is fed.  cartoon: uint8 [N,R,R,3] (`cartoon_data`, R = 384);
 * sketch: uint8 (sketch_dtype 0, `sketch_data`) or fp32 (sketch_dtype 1: the 0..255 distance map the reference computes
 * on the host with scipy through tf.py_func, :90-100) [N,R,R,3].  Per sample: image = BILINEAR resize (TF-1 legacy
 * kernel; at the integer factors R/OH, R/OW a pixel pick), (image - min) / (max - min + 1) over the whole resized picture,
 * + U[0, 1/256) dequantisation noise when `dequantize` (counter based: value i of splitmix64(seed), i = NCHW output
 * index, top 24 bits * 2^-32), * 2 - 1; sketch = AREA resize (block mean) / 255 * 2 - 1.  Outputs fp32 NCHW
 * [N,3,OH,OW].  scratch: 2*N uint32.  FGC_EUNSUPPORTED when R is not a multiple of OH and OW.
 * -------------------------------------------------------------------------------------------------------*/
int fgc_paired_input(const uint8_t* cartoon, const void* sketch, int sketch_dtype, int N, int R, int OH, int OW,
                     unsigned long long seed, int dequantize, float* images, float* sketches, uint32_t* scratch,
                     fgc_stream stream);

/* ---------------------------------------------------------------------------------------------------------
 * Streaming operators of the instance-matching model (BASELINE.json configs[4]; reference Instance_Matching/): the ResNet-101
 * trunk of deeplab_model.py and the output head of RMI_model.py.  Its contractions are the convolutions above (1x1 / 3x3 /
 * 7x7, stride 1 / 2); tf.nn.atrous_conv2d (:289-291) runs -- as inside TensorFlow -- as a plain SAME convolution between
 * space_to_batch and batch_to_space, and because every other operator of a residual group is per pixel the whole dilated
 * group stays in the batch form.
 * -------------------------------------------------------------------------------------------------------*/
/* y = act(x*scale[c] + shift[c] + r), r = 0 (res NULL) | res (rscale NULL) | res*rscale[c] + rshift[c]; relu: 0 / 1.
 * deeplab_model._batch_norm with stored moments (:213-233) folded to scale = gamma*rsqrt(variance/factor + 0.001),
 * shift = beta - mean/factor*scale; _relu (:299-301); the residual sum of _bottleneck_residual (:256-262).  x: [M, C]. */
int fgc_affine_act(const void* x, int dtype, long long M, int C, const float* scale, const float* shift, const void* res,
                   const float* rscale, const float* rshift, int relu, void* y, fgc_stream s);
/* tf.nn.max_pool(x, [1,3,3,1], [1,2,2,1], 'SAME') (deeplab_model.py:72): y [N, ceil(H/2), ceil(W/2), C] */
int fgc_maxpool3x3s2(const void* x, int dtype, int N, int H, int W, int C, void* y, fgc_stream s);
/* y[(py*r + px)*N + n, h, w, :] = x[n, h*r + py, w*r + px, :]: [N,H,W,C] -> [r*r*N, H/r, W/r, C], and its inverse (h, w =
 * the batch-form size, N = the ORIGINAL batch) */
int fgc_space_to_batch(const void* x, int dtype, int N, int H, int W, int C, int r, void* y, fgc_stream s);
int fgc_batch_to_space(const void* x, int dtype, int N, int h, int w, int C, int r, void* y, fgc_stream s);
/* y[r, c] = c < C ? x[r, c] : 0 for c < Cp, converted to y_dtype: a row operand whose width is not a multiple of 8 bf16
 * elements (the 500-wide state of the multimodal LSTM) padded to one the 16-byte operand fetches of the tensor path take */
int fgc_pad_cast_rows(const void* x, int x_dtype, long long R, int C, int Cp, void* y, int y_dtype, fgc_stream s);
/* tf.image.resize_bilinear(x, [H, W]) (align_corners False, TF-1: source = destination * h/H) of fp32 x [N,h,w,C] into up
 * [N,H,W,C]; sigm (optional) = sigmoid(up) (RMI_model.py:150-151) */
int fgc_resize_bilinear(const float* x, int N, int h, int w, int C, int H, int W, float* up, float* sigm, fgc_stream s);

#ifdef __cplusplus
}
#endif
#endif  /* FGCOLOR_H_ */
