// C-ABI entry points of the convolution family (see include/fgcolor.h).
// FGC_CONV_IMPL=simple selects the CUDA-core checker in conv_simple.cu; the default is the tcgen05 path.
#include <cstdlib>
#include <cstring>

#include "conv_geom.cuh"

namespace fgc {
int conv_fwd_simple(const ConvGeom& g, int src_dtype, const float* w, int Cin_total, int Cout, const float* bias, int act, int accumulate,
                    void* y, int y_dtype, cudaStream_t s);
int conv_dgrad_simple(const void* gy, int gy_dtype, int N, int H, int W, const float* w, int k, int Cin_total, int Cout,
                      int c_off, int c_len, int ups, int accumulate, void* gx, int gx_dtype, cudaStream_t s);
int conv_wgrad_simple(const ConvGeom& g, int src_dtype, const void* gy, int Cin_total, int Cout, float* dw, cudaStream_t s);
int colsum_launch(const void* x, int dtype, long long M, int C, float* out, cudaStream_t s);
int conv_igemm_run(ConvGeom& g, int src_dtype, const float* w, long long tap_stride, long long k_stride, long long n_stride,
                   long long base, int nout, const float* bias, int act, int accumulate, void* y, int y_dtype, void* ws,
                   cudaStream_t s, int pool2 = 0);
int conv_wgrad_run(ConvGeom& g, int src_dtype, const void* gy, int Cin_total, int Cout, float* dw, cudaStream_t s,
                   const void* gy_patch = nullptr);
size_t conv_ws_bytes(const ConvGeom& g, int nout, int x3);
extern int g_halo_mode, g_small_mode, g_keep_packed;
extern long long g_conv_counts[6];
extern int g_center_col;
extern int g_phase_kw_mask, g_phase_kh_mask, g_phase_dy, g_phase_dx;
extern long long* g_trace;
extern int g_trace_cap;

static int g_impl = -1;   // 0 = tcgen05, 1 = simple
static int conv_impl() {
  if (g_impl < 0) {
    const char* e = getenv("FGC_CONV_IMPL");
    g_impl = (e && !strcmp(e, "simple")) ? 1 : 0;
  }
  return g_impl;
}

static int build_geom(ConvGeom& g, const fgc_src* srcs, int nsrc, int N, int H, int W, int k, int stride, int pad_t, int pad_l,
                      int OH, int OW, int sign) {
  memset(&g, 0, sizeof(g));
  FGC_REQUIRE(nsrc >= 1 && nsrc <= kMaxSrc, "conv: %d sources (max %d)", nsrc, kMaxSrc);
  g.nsrc = nsrc;
  for (int i = 0; i < nsrc; i++) {
    g.src[i] = srcs[i].ptr;
    g.C[i] = srcs[i].C;
    g.ups[i] = srcs[i].ups;
    g.patch[i] = srcs[i].patch;
    FGC_REQUIRE(srcs[i].C > 0 && srcs[i].ptr, "conv: bad source %d", i);
    if (srcs[i].ups) FGC_REQUIRE(H % 2 == 0 && W % 2 == 0, "conv: upsampled source needs even H, W");
  }
  g.N = N; g.H = H; g.W = W; g.OH = OH; g.OW = OW;
  g.k = k; g.stride = stride; g.pad_t = pad_t; g.pad_l = pad_l; g.sign = sign;
  return FGC_OK;
}
}  // namespace fgc

using namespace fgc;

extern "C" {

// debug: event trace of CTA 0 of the implicit-GEMM kernel; buf = [4 + 4*capacity] int64 on the device (buf[0] = count)
int fgc_debug_set_trace(long long* buf, int capacity) {
  g_trace = buf;
  g_trace_cap = capacity;
  return FGC_OK;
}

int fgc_set_conv_flags(int halo, int small) {
  if (halo >= 0) g_halo_mode = halo;
  if (small >= 0) g_small_mode = small;
  return FGC_OK;
}

int fgc_debug_keep_packed(int on) {
  g_keep_packed = on ? 1 : 0;
  return FGC_OK;
}

int fgc_debug_conv_counts(long long out[6]) {
  for (int i = 0; i < 6; i++) out[i] = g_conv_counts[i];
  return FGC_OK;
}

int fgc_set_conv_impl(int impl) {
  g_impl = impl ? 1 : 0;
  return FGC_OK;
}

size_t fgc_conv2d_ws_bytes(const int* src_C, int nsrc, int k, int n_out, int src_dtype) {
  ConvGeom g;
  memset(&g, 0, sizeof(g));
  g.nsrc = nsrc;
  for (int i = 0; i < nsrc && i < kMaxSrc; i++) g.C[i] = src_C[i];
  g.k = k;
  g.N = g.OH = g.OW = g.H = g.W = 1;
  finish_geom(g);
  return conv_ws_bytes(g, n_out, src_dtype == FGC_F32) + 256;
}

int fgc_conv2d_fwd(const fgc_src* srcs, int nsrc, int src_dtype, int N, int H, int W,
                   const float* w, int k, int Cin_total, int Cout, const float* bias,
                   int stride, int pad_t, int pad_l, int OH, int OW, int act,
                   void* y, int y_dtype, void* ws, fgc_stream stream) {
  return fgc_conv2d_fwd_acc(srcs, nsrc, src_dtype, N, H, W, w, k, Cin_total, Cout, bias, stride, pad_t, pad_l, OH, OW, act, 0, y,
                            y_dtype, ws, stream);
}

int fgc_conv2d_fwd_acc(const fgc_src* srcs, int nsrc, int src_dtype, int N, int H, int W,
                       const float* w, int k, int Cin_total, int Cout, const float* bias,
                       int stride, int pad_t, int pad_l, int OH, int OW, int act, int accumulate,
                       void* y, int y_dtype, void* ws, fgc_stream stream) {
  ConvGeom g;
  int e = build_geom(g, srcs, nsrc, N, H, W, k, stride, pad_t, pad_l, OH, OW, 1);
  if (e) return e;
  int cin = finish_geom(g);
  FGC_REQUIRE(cin == Cin_total, "conv_fwd: sources have %d channels, weights expect %d", cin, Cin_total);
  cudaStream_t s = as_stream(stream);
  if (conv_impl() == 1) {
    e = conv_fwd_simple(g, src_dtype, w, Cin_total, Cout, bias, act, accumulate & 1, y, y_dtype, s);
    if (e) return e;
    FGC_LAUNCH_CHECK("conv_fwd_simple");
    return FGC_OK;
  }
  // accumulate: bit 0 = add into y; bit 1 = only the centre filter COLUMN of w is non-zero (the column-folded 7x7 head,
  // fgc_tapsum_w): kernels that walk the filter by columns skip the others, the rest multiply the zeros
  g_center_col = (accumulate & 2) ? 1 : 0;
  e = conv_igemm_run(g, src_dtype, w, (long long)Cin_total * Cout, Cout, 1, 0, Cout, bias, act, accumulate & 1, y, y_dtype, ws, s);
  g_center_col = 0;
  return e;
}

int fgc_conv2d_fwd_phase(const fgc_src* srcs, int nsrc, int src_dtype, int N, int H, int W, const float* w, int k, int Cin_total,
                         int Cout, int kw_mask, int kh_mask, int dy, int dx, void* y, int y_dtype, void* ws, fgc_stream stream) {
  FGC_REQUIRE(k % 2 == 1 && k <= 15 && (dy == 0 || dy == 1) && (dx == 0 || dx == 1), "conv_fwd_phase: bad arguments");
  FGC_REQUIRE((kw_mask & ((1 << k) - 1)) && (kh_mask & ((1 << k) - 1)), "conv_fwd_phase: empty tap mask");
  if (conv_impl() == 1 || src_dtype != FGC_BF16) return FGC_EUNSUPPORTED;
  ConvGeom g;
  int e = build_geom(g, srcs, nsrc, N, H, W, k, 1, (k - 1) / 2, (k - 1) / 2, H, W, 1);
  if (e) return e;
  int cin = finish_geom(g);
  FGC_REQUIRE(cin == Cin_total, "conv_fwd_phase: sources have %d channels, weights expect %d", cin, Cin_total);
  g_phase_kw_mask = kw_mask; g_phase_kh_mask = kh_mask; g_phase_dy = dy; g_phase_dx = dx;
  e = conv_igemm_run(g, src_dtype, w, (long long)Cin_total * Cout, Cout, 1, 0, Cout, nullptr, FGC_ACT_NONE, 0, y, y_dtype, ws,
                     as_stream(stream), 2);
  g_phase_kw_mask = g_phase_kh_mask = g_phase_dy = g_phase_dx = 0;
  return e == 1 ? FGC_EUNSUPPORTED : e;          // 1 = kNotTaken: the halo-reuse kernel does not take this layer
}

int fgc_conv2d_dgrad(const void* gy, int gy_dtype, int N, int H, int W, const float* w, int k, int Cin_total,
                     int Cout, int c_off, int c_len, int ups, int accumulate, void* gx, int gx_dtype,
                     void* scratch, void* ws, const void* gy_patch, fgc_stream stream) {
  FGC_REQUIRE(k % 2 == 1, "conv_dgrad: odd kernel sizes only (got %d)", k);
  FGC_REQUIRE(c_off >= 0 && c_len > 0 && c_off + c_len <= Cin_total, "conv_dgrad: bad channel slice");
  cudaStream_t s = as_stream(stream);
  if (conv_impl() == 1) {
    int e = conv_dgrad_simple(gy, gy_dtype, N, H, W, w, k, Cin_total, Cout, c_off, c_len, ups, accumulate, gx, gx_dtype, s);
    if (e) return e;
    FGC_LAUNCH_CHECK("conv_dgrad_simple");
    return FGC_OK;
  }
  // a stride-1 SAME conv over gy with the taps mirrored and the weight matrix transposed
  fgc_src src{gy, Cout, 0, gy_patch};
  ConvGeom g;
  int pad = (k - 1) / 2;
  int e = build_geom(g, &src, 1, N, H, W, k, 1, pad, pad, H, W, -1);
  if (e) return e;
  finish_geom(g);
  if (!ups)
    return conv_igemm_run(g, gy_dtype, w, (long long)Cin_total * Cout, 1, Cout, (long long)c_off * Cout, c_len, nullptr,
                          FGC_ACT_NONE, accumulate, gx, gx_dtype, ws, s);
  // gradient of an x2-upsampled source = 2x2 sums of the full-resolution gradient: folded into the epilogue of the
  // halo-reuse kernel (warp shuffles) where that kernel takes the layer, else via a full-resolution fp32 scratch
  e = conv_igemm_run(g, gy_dtype, w, (long long)Cin_total * Cout, 1, Cout, (long long)c_off * Cout, c_len, nullptr, FGC_ACT_NONE,
                     accumulate, gx, gx_dtype, ws, s, 1);
  if (e != 1) return e;
  FGC_REQUIRE(scratch != nullptr, "conv_dgrad: ups needs a full-resolution fp32 scratch");
  e = conv_igemm_run(g, gy_dtype, w, (long long)Cin_total * Cout, 1, Cout, (long long)c_off * Cout, c_len, nullptr, FGC_ACT_NONE, 0,
                     scratch, FGC_F32, ws, s);
  if (e) return e;
  return fgc_sum2x2(scratch, FGC_F32, N, H / 2, W / 2, c_len, gx, gx_dtype, accumulate, stream);
}

int fgc_colsum(const void* x, int dtype, long long M, int C, float* out, fgc_stream stream) {
  FGC_REQUIRE(M > 0 && C > 0 && out, "colsum: bad arguments");
  int e = colsum_launch(x, dtype, M, C, out, as_stream(stream));
  if (e) return e;
  FGC_LAUNCH_CHECK("colsum");
  return FGC_OK;
}

int fgc_conv2d_wgrad(const fgc_src* srcs, int nsrc, int src_dtype, int N, int H, int W,
                     const void* gy, int gy_dtype, int k, int Cin_total, int Cout,
                     int stride, int pad_t, int pad_l, int OH, int OW,
                     float* dw, float* db, const void* gy_patch, fgc_stream stream) {
  FGC_REQUIRE(src_dtype == gy_dtype, "conv_wgrad: sources and gy must share a dtype");
  ConvGeom g;
  int e = build_geom(g, srcs, nsrc, N, H, W, k, stride, pad_t, pad_l, OH, OW, 1);
  if (e) return e;
  int cin = finish_geom(g);
  FGC_REQUIRE(cin == Cin_total, "conv_wgrad: sources have %d channels, weights expect %d", cin, Cin_total);
  cudaStream_t s = as_stream(stream);
  if (db) {
    e = colsum_launch(gy, gy_dtype, g.M, Cout, db, s);
    if (e) return e;
  }
  if (conv_impl() == 1) {
    e = conv_wgrad_simple(g, src_dtype, gy, Cin_total, Cout, dw, s);
    if (e) return e;
    FGC_LAUNCH_CHECK("conv_wgrad_simple");
    return FGC_OK;
  }
  return conv_wgrad_run(g, src_dtype, gy, Cin_total, Cout, dw, s, gy_patch);
}

}  // extern "C"
