// Streaming kernels of the instance-matching model's vision trunk and output head (sm_100a); BASELINE.json configs[4].
//
// Reference call sites (Instance_Matching/):
//   deeplab_model.py  _batch_norm :186-235 (stored moments: an affine map per channel), _relu :299-301, the residual sum of
//                     _bottleneck_residual :237-264, tf.nn.max_pool 3x3 / stride 2 / SAME :72, tf.nn.atrous_conv2d :289-291
//   RMI_model.py      tf.image.resize_bilinear + tf.sigmoid :150-151
//
// All HBM-bound: one thread owns V consecutive channels of one pixel (128-bit accesses along C), grid-stride loops over
// grids capped at 16 blocks per SM.  tf.nn.atrous_conv2d with rate r is -- also inside TensorFlow -- a plain SAME
// convolution between space_to_batch and batch_to_space; every other operator of a residual group is per pixel, so a whole
// dilated group runs in the batch form and the two permutations below are paid once per group.
#include "common.cuh"

namespace fgc {
int ew_grid(long long work, int threads);

#define FGC_DISPATCH_V(vw, V, ...)                                                    \
  do {                                                                                \
    if ((vw) >= 8) { constexpr int V = 8; __VA_ARGS__; }                              \
    else if ((vw) >= 4) { constexpr int V = 4; __VA_ARGS__; }                         \
    else { constexpr int V = 1; __VA_ARGS__; }                                        \
  } while (0)
#define FGC_DISPATCH_TV(dtype, vw, T, V, ...)                                         \
  do {                                                                                \
    if ((dtype) == FGC_F32) { using T = float; FGC_DISPATCH_V(vw, V, __VA_ARGS__); }  \
    else if ((dtype) == FGC_BF16) { using T = __nv_bfloat16; FGC_DISPATCH_V(vw, V, __VA_ARGS__); } \
    else {                                                                            \
      fgc::set_error("bad dtype %d", (int)(dtype));                                   \
      return FGC_EINVAL;                                                              \
    }                                                                                 \
  } while (0)

// V per-channel parameters starting at a multiple of V (16-byte loads: the per-element scalar loads made this pass L1-bound)
template <int V> __device__ __forceinline__ void ldp(const float* __restrict__ p, float (&v)[kMaxV]) {
  if constexpr (V == 8) {
    const float4 a = __ldg(reinterpret_cast<const float4*>(p)), b = __ldg(reinterpret_cast<const float4*>(p) + 1);
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
  } else if constexpr (V == 4) {
    const float4 a = __ldg(reinterpret_cast<const float4*>(p));
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w;
  } else {
    v[0] = __ldg(p);
  }
}

// y = act(x*scale[c] + shift[c] + (res ? res*rscale[c] + rshift[c] : 0)); rscale == NULL: the residual is added as it is
template <typename T, int V>
__global__ void affine_act_kernel(const T* __restrict__ x, long long nvec, int C, const float* __restrict__ scale,
                                  const float* __restrict__ shift, const T* __restrict__ res, const float* __restrict__ rscale,
                                  const float* __restrict__ rshift, int relu, T* __restrict__ y) {
  const int CV = C / V;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < nvec; i += (long long)gridDim.x * blockDim.x) {
    const int c0 = (int)(i % CV) * V;
    float a[kMaxV], r[kMaxV], sc[kMaxV], sh[kMaxV], rs[kMaxV], rt[kMaxV];
    ldv<T, V>(x + i * V, a);
    ldp<V>(scale + c0, sc);
    ldp<V>(shift + c0, sh);
    if (res) ldv<T, V>(res + i * V, r);
    if (rscale) { ldp<V>(rscale + c0, rs); ldp<V>(rshift + c0, rt); }
#pragma unroll
    for (int k = 0; k < V; k++) {
      float v = fmaf(a[k], sc[k], sh[k]);
      if (res) v += rscale ? fmaf(r[k], rs[k], rt[k]) : r[k];
      a[k] = relu ? fmaxf(v, 0.f) : v;
    }
    stv<T, V>(y + i * V, a);
  }
}

// tf.nn.max_pool 3x3, stride 2, SAME: OH = ceil(H/2), the odd pad pixel at the bottom / right, padding never wins
template <typename T, int V>
__global__ void maxpool3x3s2_kernel(const T* __restrict__ x, int H, int W, int C, int OH, int OW, int pt, int pl, long long nvec,
                                    T* __restrict__ y) {
  const int CV = C / V;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < nvec; i += (long long)gridDim.x * blockDim.x) {
    const int cv = (int)(i % CV);
    long long p = i / CV;
    const int ow = (int)(p % OW);
    p /= OW;
    const int oh = (int)(p % OH);
    const long long n = p / OH;
    float m[kMaxV], a[kMaxV];
#pragma unroll
    for (int k = 0; k < V; k++) m[k] = -INFINITY;
    for (int kh = 0; kh < 3; kh++) {
      const int ih = oh * 2 + kh - pt;
      if (ih < 0 || ih >= H) continue;
      for (int kw = 0; kw < 3; kw++) {
        const int iw = ow * 2 + kw - pl;
        if (iw < 0 || iw >= W) continue;
        ldv<T, V>(x + ((n * H + ih) * (long long)W + iw) * C + (long long)cv * V, a);
#pragma unroll
        for (int k = 0; k < V; k++) m[k] = fmaxf(m[k], a[k]);
      }
    }
    stv<T, V>(y + i * V, m);
  }
}

// TO_BATCH: dst[(py*r + px)*N + n, h, w, :] = src[n, h*r + py, w*r + px, :]  (tf.space_to_batch, zero paddings); else inverse
template <typename T, int V, bool TO_BATCH>
__global__ void space_batch_kernel(const T* __restrict__ src, int N, int h, int w, int C, int r, long long nvec, T* __restrict__ dst) {
  const int CV = C / V;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < nvec; i += (long long)gridDim.x * blockDim.x) {
    const int cv = (int)(i % CV);
    long long p = i / CV;                       // pixel index of the batch-form tensor [r*r*N, h, w]
    const int x = (int)(p % w);
    p /= w;
    const int y = (int)(p % h);
    p /= h;
    const int n = (int)(p % N);
    const int ph = (int)(p / N);
    const int py = ph / r, px = ph % r;
    const long long sp = (((long long)n * h * r + (long long)y * r + py) * ((long long)w * r) + (long long)x * r + px) * C + (long long)cv * V;
    float a[kMaxV];
    if (TO_BATCH) {
      ldv<T, V>(src + sp, a);
      stv<T, V>(dst + i * V, a);
    } else {
      ldv<T, V>(src + i * V, a);
      stv<T, V>(dst + sp, a);
    }
  }
}

template <typename TI, typename TO>
__global__ void pad_cast_rows_kernel(const TI* __restrict__ x, long long total, int C, int Cp, TO* __restrict__ y) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / Cp;
    const int c = (int)(i - r * Cp);
    st1<TO>(y + i, c < C ? ld1<TI>(x + r * C + c) : 0.f);
  }
}
// fp32 -> bf16, C and Cp multiples of 4: four columns per thread (16-byte load, 8-byte store)
__global__ void pad_cast_rows_f32_bf16_vec4_kernel(const float* __restrict__ x, long long total4, int C4, int Cp4, uint2* __restrict__ y) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total4; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / Cp4;
    const int c = (int)(i - r * Cp4);
    uint2 o = make_uint2(0u, 0u);
    if (c < C4) {
      const float4 v = __ldcs(reinterpret_cast<const float4*>(x) + r * C4 + c);
      o = make_uint2(bf16x2_bits(v.x, v.y), bf16x2_bits(v.z, v.w));
    }
    y[i] = o;
  }
}

// tf.image.resize_bilinear, align_corners=False (TF-1 legacy: in = out * h/H, no half-pixel shift), fp32, and its sigmoid
__global__ void resize_bilinear_kernel(const float* __restrict__ x, int h, int w, int C, int H, int W, long long total,
                                       float* __restrict__ up, float* __restrict__ sigm) {
  const float sy = (float)h / (float)H, sx = (float)w / (float)W;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    long long p = i / C;
    const int ox = (int)(p % W);
    p /= W;
    const int oy = (int)(p % H);
    const long long n = p / H;
    const float fy = oy * sy, fx = ox * sx;
    const int y0 = (int)floorf(fy), x0 = (int)floorf(fx);
    const int y1 = y0 + 1 < h ? y0 + 1 : h - 1, x1 = x0 + 1 < w ? x0 + 1 : w - 1;
    const float ly = fy - y0, lx = fx - x0;
    const float* b = x + n * (long long)h * w * C + c;
    const float tl = __ldg(b + ((long long)y0 * w + x0) * C), tr = __ldg(b + ((long long)y0 * w + x1) * C);
    const float bl = __ldg(b + ((long long)y1 * w + x0) * C), br = __ldg(b + ((long long)y1 * w + x1) * C);
    const float top = tl + (tr - tl) * lx, bot = bl + (br - bl) * lx;
    const float v = top + (bot - top) * ly;
    up[i] = v;
    if (sigm) sigm[i] = 1.f / (1.f + expf(-v));
  }
}
}  // namespace fgc

using namespace fgc;

extern "C" {

int fgc_affine_act(const void* x, int dtype, long long M, int C, const float* scale, const float* shift, const void* res,
                   const float* rscale, const float* rshift, int relu, void* y, fgc_stream stream) {
  FGC_REQUIRE(x && y && scale && shift && M > 0 && C > 0 && (!rscale || (res && rshift)), "affine_act: bad arguments");
  cudaStream_t s = as_stream(stream);
  int vec = vmin(vec_width(x, C, dtype), vec_width(y, C, dtype));
  if (res) vec = vmin(vec, vec_width(res, C, dtype));
  if (!aligned16(scale) || !aligned16(shift) || (rscale && (!aligned16(rscale) || !aligned16(rshift)))) vec = 1;
  FGC_DISPATCH_TV(dtype, vec, T, V, {
    long long nv = M * C / V;
    affine_act_kernel<T, V><<<ew_grid(nv, 256), 256, 0, s>>>((const T*)x, nv, C, scale, shift, (const T*)res, rscale, rshift, relu, (T*)y);
  });
  count_launch();
  FGC_LAUNCH_CHECK("affine_act");
  return FGC_OK;
}

int fgc_maxpool3x3s2(const void* x, int dtype, int N, int H, int W, int C, void* y, fgc_stream stream) {
  FGC_REQUIRE(x && y && N > 0 && H > 0 && W > 0 && C > 0, "maxpool3x3s2: bad arguments");
  cudaStream_t s = as_stream(stream);
  const int OH = (H + 1) / 2, OW = (W + 1) / 2;
  const int th = (OH - 1) * 2 + 3 - H, tw = (OW - 1) * 2 + 3 - W;
  const int pt = (th > 0 ? th : 0) / 2, pl = (tw > 0 ? tw : 0) / 2;
  int vec = vmin(vec_width(x, C, dtype), vec_width(y, C, dtype));
  FGC_DISPATCH_TV(dtype, vec, T, V, {
    long long nv = (long long)N * OH * OW * C / V;
    maxpool3x3s2_kernel<T, V><<<ew_grid(nv, 256), 256, 0, s>>>((const T*)x, H, W, C, OH, OW, pt, pl, nv, (T*)y);
  });
  count_launch();
  FGC_LAUNCH_CHECK("maxpool3x3s2");
  return FGC_OK;
}

static int space_batch(const void* x, int dtype, int N, int h, int w, int C, int r, void* y, bool to_batch, cudaStream_t s) {
  int vec = vmin(vec_width(x, C, dtype), vec_width(y, C, dtype));
  FGC_DISPATCH_TV(dtype, vec, T, V, {
    long long nv = (long long)N * r * r * h * w * C / V;
    if (to_batch) space_batch_kernel<T, V, true><<<ew_grid(nv, 256), 256, 0, s>>>((const T*)x, N, h, w, C, r, nv, (T*)y);
    else space_batch_kernel<T, V, false><<<ew_grid(nv, 256), 256, 0, s>>>((const T*)x, N, h, w, C, r, nv, (T*)y);
  });
  count_launch();
  return FGC_OK;
}
int fgc_space_to_batch(const void* x, int dtype, int N, int H, int W, int C, int r, void* y, fgc_stream stream) {
  FGC_REQUIRE(x && y && N > 0 && C > 0 && r > 0 && H > 0 && W > 0 && H % r == 0 && W % r == 0, "space_to_batch: bad arguments");
  int e = space_batch(x, dtype, N, H / r, W / r, C, r, y, true, as_stream(stream));
  if (e) return e;
  FGC_LAUNCH_CHECK("space_to_batch");
  return FGC_OK;
}
int fgc_batch_to_space(const void* x, int dtype, int N, int h, int w, int C, int r, void* y, fgc_stream stream) {
  FGC_REQUIRE(x && y && N > 0 && C > 0 && r > 0 && h > 0 && w > 0, "batch_to_space: bad arguments");
  int e = space_batch(x, dtype, N, h, w, C, r, y, false, as_stream(stream));
  if (e) return e;
  FGC_LAUNCH_CHECK("batch_to_space");
  return FGC_OK;
}

int fgc_pad_cast_rows(const void* x, int x_dtype, long long R, int C, int Cp, void* y, int y_dtype, fgc_stream stream) {
  FGC_REQUIRE(x && y && R > 0 && C > 0 && Cp >= C, "pad_cast_rows: bad arguments");
  cudaStream_t s = as_stream(stream);
  const long long total = R * Cp;
  const int grid = ew_grid(total, 256);
  if (x_dtype == FGC_F32 && y_dtype == FGC_BF16 && C % 4 == 0 && Cp % 4 == 0 && aligned16(x) && aligned8(y)) {
    const long long total4 = total / 4;
    pad_cast_rows_f32_bf16_vec4_kernel<<<ew_grid(total4, 256), 256, 0, s>>>((const float*)x, total4, C / 4, Cp / 4, (uint2*)y);
  } else if (x_dtype == FGC_F32 && y_dtype == FGC_F32) pad_cast_rows_kernel<float, float><<<grid, 256, 0, s>>>((const float*)x, total, C, Cp, (float*)y);
  else if (x_dtype == FGC_F32 && y_dtype == FGC_BF16) pad_cast_rows_kernel<float, __nv_bfloat16><<<grid, 256, 0, s>>>((const float*)x, total, C, Cp, (__nv_bfloat16*)y);
  else if (x_dtype == FGC_BF16 && y_dtype == FGC_F32) pad_cast_rows_kernel<__nv_bfloat16, float><<<grid, 256, 0, s>>>((const __nv_bfloat16*)x, total, C, Cp, (float*)y);
  else if (x_dtype == FGC_BF16 && y_dtype == FGC_BF16) pad_cast_rows_kernel<__nv_bfloat16, __nv_bfloat16><<<grid, 256, 0, s>>>((const __nv_bfloat16*)x, total, C, Cp, (__nv_bfloat16*)y);
  else { set_error("pad_cast_rows: bad dtypes"); return FGC_EINVAL; }
  count_launch();
  FGC_LAUNCH_CHECK("pad_cast_rows");
  return FGC_OK;
}

int fgc_resize_bilinear(const float* x, int N, int h, int w, int C, int H, int W, float* up, float* sigm, fgc_stream stream) {
  FGC_REQUIRE(x && up && N > 0 && h > 0 && w > 0 && C > 0 && H > 0 && W > 0, "resize_bilinear: bad arguments");
  long long total = (long long)N * H * W * C;
  resize_bilinear_kernel<<<ew_grid(total, 256), 256, 0, as_stream(stream)>>>(x, h, w, C, H, W, total, up, sigm);
  count_launch();
  FGC_LAUNCH_CHECK("resize_bilinear");
  return FGC_OK;
}

}  // extern "C"
