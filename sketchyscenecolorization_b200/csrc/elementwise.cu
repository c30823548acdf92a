// Normalisation, activation, gating and resampling kernels of libfgcolor (sm_100a).
//
// All of these are HBM-bound streaming passes over NHWC activations: one thread owns V=4 consecutive
// channels (128-bit fp32 / 64-bit bf16 accesses, coalesced along C), grids are sized to a multiple of the
// SM count and grid-stride over the tensor.  Reductions over H*W or N*H*W are two-stage: per-thread fp32
// partials over a short run of rows, then one atomic per (thread, channel) into a small global table.
//
// Reference call sites (Foreground_Instance_Colorization/obj_lib/):
//   models_collection.batchnorm :22-34, prelu :56-60, miu_relu :63-65; mru.lrelu :10-12,
//   min-max gates mru.py:415-416,560-561,568-569; gating mru.py:426,453,572,589; mean_pool mru.py:15-19;
//   upsample mru.py:22-28.
#include "common.cuh"

namespace fgc {

static int g_num_sms = 0;
int num_sms() {
  if (!g_num_sms) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev);
    if (g_num_sms <= 0) g_num_sms = 148;
  }
  return g_num_sms;
}
// grid for a grid-stride loop over `work` thread-items with `threads` per block
int ew_grid(long long work, int threads) {
  long long b = (work + threads - 1) / threads;
  long long cap = (long long)num_sms() * 16;
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return (int)b;
}

// order-preserving float <-> uint32 map for atomicMin/atomicMax on floats
__device__ __forceinline__ uint32_t f2ord(float f) {
  uint32_t u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float ord2f(uint32_t u) {
  return __uint_as_float((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u);
}

// `vw` = vector width chosen with vec_width(): 8, 4 or 1 elements per thread and access
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

// ======================================================================================================
// channel statistics (tf.nn.moments over N,H,W; biased variance)
// ======================================================================================================
template <typename T, int V>
__global__ void __launch_bounds__(256, 4) chan_stats_kernel(const T* __restrict__ x, long long M, int C, int rows_per_block, double* acc) {
  extern __shared__ float sh_part[];          // [lanes][2*C] per-thread partial sums (no shared-memory atomics: the fp64
                                              // CAS loops of the first version, 16 row lanes deep per address, were its tail)
  const int CV = C / V;
  const int lanes = blockDim.x / CV;          // row lanes per block
  const int v = threadIdx.x % CV, rl = threadIdx.x / CV;
  long long r0 = (long long)blockIdx.x * rows_per_block;
  long long r1 = r0 + rows_per_block;
  if (r1 > M) r1 = M;
  if (rl < lanes) {
    // fp32 partials per thread (a thread sees rows_per_block / lanes rows: a few hundred), fp64 from the block level on
    float s[V], ss[V];
#pragma unroll
    for (int i = 0; i < V; i++) s[i] = ss[i] = 0.f;
    // four row loads per group, the next group's loads issued before this group's sums (walk_rows, common.cuh)
    walk_rows<T, V, 4>(x + r0 * C + v * V, C, rl, (int)(r1 - r0), lanes, [&](int, const float (&a)[kMaxV]) {
#pragma unroll
      for (int i = 0; i < V; i++) { s[i] += a[i]; ss[i] = fmaf(a[i], a[i], ss[i]); }
    });
    float* mine = sh_part + (size_t)rl * 2 * C + v * V;
#pragma unroll
    for (int i = 0; i < V; i++) { mine[i] = s[i]; mine[C + i] = ss[i]; }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) {
    double t = 0.0;
    for (int l = 0; l < lanes; l++) t += (double)sh_part[(size_t)l * 2 * C + i];
    atomicAdd(&acc[i], t);
  }
}
__global__ void chan_stats_finalize(const double* acc, long long M, int C, float* stats) {
  int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  double mean = acc[c] / (double)M;
  double var = acc[C + c] / (double)M - mean * mean;
  if (var < 0) var = 0;
  stats[c] = (float)mean;
  stats[C + c] = (float)(1.0 / sqrt(var + 1e-5));
}

// ======================================================================================================
// conditional BN apply (+ miu_relu)
// ======================================================================================================
// The "apply" passes below share one thread layout: grid (row blocks, N); a thread owns V fixed channels of image n and
// walks rows (pixels), so the per-(n,c) parameters (statistics, class-table rows, min/max) are loaded ONCE per thread
// instead of once per element -- the first versions were bound by L1/TEX parameter traffic (ncu: 87-96% L1, < 2 TB/s).
template <typename T, int V>
__global__ void __launch_bounds__(256, 4) cbn_act_fwd_kernel(const T* __restrict__ x, int HW, int C, int rows_per_block, const float* __restrict__ stats,
                                   const float* __restrict__ scale, const float* __restrict__ offset,
                                   const int32_t* __restrict__ labels, int act, T* __restrict__ y) {
  const int CV = C / V;
  const int lanes = blockDim.x / CV;
  const int v = threadIdx.x % CV, rl = threadIdx.x / CV;
  if (rl >= lanes) return;
  const int n = blockIdx.y;
  const int l = labels[n];
  float A[V], B[V];              // (x - mean) * rstd * scale + offset = x * A + B
#pragma unroll
  for (int k = 0; k < V; k++) {
    int c = v * V + k;
    A[k] = stats[C + c] * scale[l * C + c];
    B[k] = fmaf(-stats[c], A[k], offset[l * C + c]);
  }
  const int r0 = blockIdx.x * rows_per_block, r1 = min(r0 + rows_per_block, HW);
  const long long base = (long long)n * HW * C + v * V;
  const bool miu = act == FGC_ACT_MIU;
  walk_rows<T, V, 4>(x + base, C, r0 + rl, r1, lanes, [&](int r, const float (&a)[kMaxV]) {
    float o[kMaxV];
#pragma unroll
    for (int k = 0; k < V; k++) {
      float t = fmaf(a[k], A[k], B[k]);
      o[k] = miu ? miu_relu_fast(t) : t;
    }
    stv<T, V>(y + base + (long long)r * C, o);
  });
}

// per-(n,c) sums of g and g*xhat where g = gy*act'(y)    -> sums[0][n][c], sums[1][n][c]
template <typename T, int V>
__global__ void __launch_bounds__(256, 3) cbn_bwd_reduce_kernel(const T* __restrict__ gy, const T* __restrict__ x, int HW, int C, int rows_per_block,
                                      const float* __restrict__ stats, const float* __restrict__ scale,
                                      const float* __restrict__ offset, const int32_t* __restrict__ labels, int act,
                                      int N, float* sums) {
  extern __shared__ float sh_red[];              // [lanes][2][C] per-thread partial sums (deposited, then added per column:
                                                 // fp32 atomicAdd on shared memory is a CAS loop, `lanes` deep per address)
  const int CV = C / V;
  const int lanes = blockDim.x / CV;
  const int v = threadIdx.x % CV, rl = threadIdx.x / CV;
  const int n = blockIdx.y;
  const int l = labels[n];
  int r0 = blockIdx.x * rows_per_block, r1 = rl < lanes ? min(r0 + rows_per_block, HW) : 0;
  float rs[V], nm[V], A[V], B[V], s1[V], s2[V];      // xhat = x * rs + nm;  scale * xhat + offset = x * A + B
#pragma unroll
  for (int k = 0; k < V; k++) {
    int c = v * V + k;
    rs[k] = stats[C + c]; nm[k] = -stats[c] * rs[k];
    A[k] = rs[k] * scale[l * C + c]; B[k] = fmaf(-stats[c], A[k], offset[l * C + c]);
    s1[k] = s2[k] = 0.f;
  }
  const bool miu = act == FGC_ACT_MIU;
  // two independent row loads per tensor in flight per thread; four (profiles/r1q) cost registers / resident CTAs and ran
  // 15-20% slower on the 604 MB tensor
  {
    const long long base = (long long)n * HW * C + v * V;
    walk_rows2<T, V, 2>(x + base, gy + base, C, r0 + rl, r1, lanes, [&](int, const float (&a)[kMaxV], const float (&g)[kMaxV]) {
#pragma unroll
      for (int k = 0; k < V; k++) {
        float gg = g[k];
        if (miu) gg *= miu_relu_grad_fast(fmaf(a[k], A[k], B[k]));
        s1[k] += gg; s2[k] = fmaf(gg, fmaf(a[k], rs[k], nm[k]), s2[k]);
      }
    });
  }
  if (rl < lanes) {
    float* mine = sh_red + (size_t)rl * 2 * C + v * V;
#pragma unroll
    for (int k = 0; k < V; k++) { mine[k] = s1[k]; mine[C + k] = s2[k]; }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) {       // one global atomic per table entry and block
    float t = 0.f;
    for (int q2 = 0; q2 < lanes; q2++) t += sh_red[(size_t)q2 * 2 * C + i];
    int q = i / C, c = i - q * C;
    atomicAdd(&sums[((long long)q * N + n) * C + c], t);
  }
}
// one thread per channel: table gradients and batch means.  The per-sample loads are independent (unrolled, all in flight
// together) and the class-table updates are fire-and-forget red.add -- the first version chained a load-add-store per
// sample and spent 48 us per call on L2 latency alone.
__global__ void cbn_bwd_finalize_kernel(const float* __restrict__ sums, int N, int C, long long M, const float* __restrict__ scale,
                                        const int32_t* __restrict__ labels, float* dscale, float* doffset, float* m12) {
  // a warp per channel: its lanes take the samples (the one-thread-per-channel form walked the 64 samples in sequence and
  // was pure latency: 33 us per call); class-table updates are fire-and-forget red.add
  const int c = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (c >= C) return;
  double m1 = 0, m2 = 0;
  for (int n = lane; n < N; n += 32) {
    const int l = __ldg(labels + n);
    const float s1 = __ldg(sums + (long long)n * C + c), s2 = __ldg(sums + ((long long)N + n) * C + c);
    const float g = __ldg(scale + l * C + c);
    atomicAdd(doffset + l * C + c, s1);
    atomicAdd(dscale + l * C + c, s2);
    m1 += (double)g * s1; m2 += (double)g * s2;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) { m1 += __shfl_xor_sync(0xffffffffu, m1, o); m2 += __shfl_xor_sync(0xffffffffu, m2, o); }
  if (lane == 0) {
    m12[c] = (float)(m1 / (double)M);
    m12[C + c] = (float)(m2 / (double)M);
  }
}
template <typename T, int V>
__global__ void __launch_bounds__(256, 3) cbn_bwd_apply_kernel(const T* __restrict__ gy, const T* __restrict__ x, int HW, int C, int rows_per_block,
                                     const float* __restrict__ stats, const float* __restrict__ scale,
                                     const float* __restrict__ offset, const int32_t* __restrict__ labels, int act,
                                     const float* __restrict__ m12, T* __restrict__ gx, float* dbias) {
  extern __shared__ float sh_db[];               // [lanes][C] per-thread column sums of gx (only when dbias != NULL)
  const int CV = C / V;
  const int lanes = blockDim.x / CV;
  const int v = threadIdx.x % CV, rl = threadIdx.x / CV;
  const bool on = rl < lanes;
  const int n = blockIdx.y;
  const int l = labels[n];
  // gx = rstd * (g * scale - m1 - xhat * m2) = g * G + x * X + K  with  G = rstd * scale, X = -rstd^2 * m2,
  // K = rstd * (mean * rstd * m2 - m1);   act'(scale * xhat + offset) is evaluated at x * A + B
  float A[V], B[V], G[V], X[V], K[V], bs[V];
#pragma unroll
  for (int k = 0; k < V; k++) {
    int c = v * V + k;
    const float mean = stats[c], rstd = stats[C + c], ga = scale[l * C + c];
    G[k] = rstd * ga;
    A[k] = G[k]; B[k] = fmaf(-mean, G[k], offset[l * C + c]);      // (A aliases G: one register)
    X[k] = -rstd * rstd * m12[C + c];
    K[k] = rstd * (mean * rstd * m12[C + c] - m12[c]);
    bs[k] = 0.f;
  }
  const bool miu = act == FGC_ACT_MIU;
  const int r0 = blockIdx.x * rows_per_block, r1 = on ? min(r0 + rows_per_block, HW) : 0;
  const long long base = (long long)n * HW * C + v * V;
  walk_rows2<T, V, 2>(x + base, gy + base, C, r0 + rl, r1, lanes, [&](int r, const float (&a)[kMaxV], const float (&g)[kMaxV]) {
    float o[kMaxV];
#pragma unroll
    for (int k = 0; k < V; k++) {
      float gg = g[k];
      if (miu) gg *= miu_relu_grad_fast(fmaf(a[k], A[k], B[k]));
      o[k] = fmaf(gg, G[k], fmaf(a[k], X[k], K[k]));
      bs[k] += o[k];
    }
    stv<T, V>(gx + base + (long long)r * C, o);
  });
  if (dbias) {                                   // bias gradient of the convolution in front: column sums of the result
    if (on) {
      float* mine = sh_db + (size_t)rl * C + v * V;
#pragma unroll
      for (int k = 0; k < V; k++) mine[k] = bs[k];
    }
    __syncthreads();
    for (int i = threadIdx.x; i < C; i += blockDim.x) {
      float t = 0.f;
      for (int q2 = 0; q2 < lanes; q2++) t += sh_db[(size_t)q2 * C + i];
      atomicAdd(dbias + i, t);
    }
  }
}

// ======================================================================================================
// PReLU
// ======================================================================================================
template <typename T, int V>
__global__ void prelu_fwd_kernel(const T* __restrict__ x, long long nvec, const float* __restrict__ ap, T* __restrict__ y) {
  const float a = *ap;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < nvec; i += (long long)gridDim.x * blockDim.x) {
    float v[kMaxV], o[kMaxV];
    ldv<T, V>(x + i * V, v);
#pragma unroll
    for (int k = 0; k < V; k++) o[k] = (a * v[k] >= v[k]) ? a * v[k] : v[k];
    stv<T, V>(y + i * V, o);
  }
}
// tanh as a pass of its own (after a batch norm: generate_residual, models_collection.py:665)
template <typename T, int V>
__global__ void tanh_fwd_kernel(const T* __restrict__ x, long long nvec, T* __restrict__ y) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < nvec; i += (long long)gridDim.x * blockDim.x) {
    float v[kMaxV], o[kMaxV];
    ldv<T, V>(x + i * V, v);
#pragma unroll
    for (int k = 0; k < V; k++) o[k] = tanhf(v[k]);
    stv<T, V>(y + i * V, o);
  }
}
template <typename T, int V>
__global__ void prelu_bwd_kernel(const T* __restrict__ gy, const T* __restrict__ x, long long nvec,
                                 const float* __restrict__ ap, float* da, T* __restrict__ gx, int accumulate) {
  __shared__ float red[32];
  const float a = *ap;
  float acc = 0.f;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < nvec; i += (long long)gridDim.x * blockDim.x) {
    float v[kMaxV], g[kMaxV], o[kMaxV];
    ldv<T, V>(x + i * V, v);
    ldv<T, V>(gy + i * V, g);
#pragma unroll
    for (int k = 0; k < V; k++) {
      bool m = a * v[k] >= v[k];
      o[k] = m ? a * g[k] : g[k];
      if (m) acc += g[k] * v[k];
    }
    if (accumulate) {                            // gx += : the gradient joins what other branches have left there
      float old[kMaxV];
      ldv<T, V>(gx + i * V, old);
#pragma unroll
      for (int k = 0; k < V; k++) o[k] += old[k];
    }
    stv<T, V>(gx + i * V, o);
  }
  if (da) {
    float t = block_sum(acc, red);
    if (threadIdx.x == 0) atomicAdd(da, t);
  }
}

// row-walking variant that also accumulates the column sums of gx (bias gradient of the convolution in front)
template <typename T, int V>
__global__ void prelu_bwd_rows_kernel(const T* __restrict__ gy, const T* __restrict__ x, long long M, int C, int rows_per_block,
                                      const float* __restrict__ ap, float* da, T* __restrict__ gx, float* dbias) {
  extern __shared__ float sh_db[];               // [lanes][C] per-thread column sums
  __shared__ float red[32];
  const int CV = C / V;
  const int lanes = blockDim.x / CV;
  const int v = threadIdx.x % CV, rl = threadIdx.x / CV;
  const bool on = rl < lanes;
  const float a = *ap;
  float acc = 0.f, bs[V];
#pragma unroll
  for (int k = 0; k < V; k++) bs[k] = 0.f;
  const long long r0 = (long long)blockIdx.x * rows_per_block;
  const long long r1 = on ? (r0 + rows_per_block < M ? r0 + rows_per_block : M) : 0;
  {
    const long long base = r0 * C + v * V;
    walk_rows2<T, V, 2>(x + base, gy + base, C, rl, (int)(r1 - r0), lanes, [&](int r, const float (&xv)[kMaxV], const float (&g)[kMaxV]) {
      float o[kMaxV];
#pragma unroll
      for (int k = 0; k < V; k++) {
        bool m = a * xv[k] >= xv[k];
        o[k] = m ? a * g[k] : g[k];
        if (m) acc += g[k] * xv[k];
        bs[k] += o[k];
      }
      stv<T, V>(gx + base + (long long)r * C, o);
    });
  }
  if (on) {
    float* mine = sh_db + (size_t)rl * C + v * V;
#pragma unroll
    for (int k = 0; k < V; k++) mine[k] = bs[k];
  }
  if (da) {
    float t = block_sum(acc, red);
    if (threadIdx.x == 0) atomicAdd(da, t);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < C; i += blockDim.x) {
    float t = 0.f;
    for (int q2 = 0; q2 < lanes; q2++) t += sh_db[(size_t)q2 * C + i];
    atomicAdd(dbias + i, t);
  }
}

// ======================================================================================================
// min-max gate normalisation
// ======================================================================================================
template <typename T, int V>
__global__ void __launch_bounds__(256, 4) minmax_reduce_kernel(const T* __restrict__ x, int HW, int C, int rows_per_block, uint32_t* mn_ord,
                                     uint32_t* mx_ord) {
  extern __shared__ uint32_t sh_mm[];            // [2][C] block-level min / max (order-preserving encoding)
  const int CV = C / V;
  const int lanes = blockDim.x / CV;
  const int v = threadIdx.x % CV, rl = threadIdx.x / CV;
  const int n = blockIdx.y;
  for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) sh_mm[i] = i < C ? 0xFFFFFFFFu : 0u;
  __syncthreads();
  int r0 = blockIdx.x * rows_per_block, r1 = rl < lanes ? min(r0 + rows_per_block, HW) : 0;
  float lo[V], hi[V];
#pragma unroll
  for (int k = 0; k < V; k++) { lo[k] = INFINITY; hi[k] = -INFINITY; }
  walk_rows<T, V, 4>(x + (long long)n * HW * C + v * V, C, r0 + rl, r1, lanes, [&](int, const float (&a)[kMaxV]) {
#pragma unroll
    for (int k = 0; k < V; k++) { lo[k] = fminf(lo[k], a[k]); hi[k] = fmaxf(hi[k], a[k]); }
  });
  // few channel vectors per row (CV a power of two below 32, e.g. the 8-channel gates of unit 1: ONE vector): lanes CV apart hold
  // the same channels and are combined with shuffles, so a warp sends CV * V atomics instead of 32 * V to the same few words
  bool deposit = rl < lanes;
  if (CV < 32 && (CV & (CV - 1)) == 0) {         // then lanes * CV == blockDim.x: every thread of every warp takes part
#pragma unroll
    for (int k = 0; k < V; k++)
      for (int o = 16; o >= CV; o >>= 1) {
        lo[k] = fminf(lo[k], __shfl_xor_sync(0xffffffffu, lo[k], o));
        hi[k] = fmaxf(hi[k], __shfl_xor_sync(0xffffffffu, hi[k], o));
      }
    deposit = (int)(threadIdx.x & 31) < CV;
  }
  if (deposit) {
#pragma unroll
    for (int k = 0; k < V; k++) {
      atomicMin(&sh_mm[v * V + k], f2ord(lo[k]));
      atomicMax(&sh_mm[C + v * V + k], f2ord(hi[k]));
    }
  }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    atomicMin(&mn_ord[(long long)n * C + c], sh_mm[c]);
    atomicMax(&mx_ord[(long long)n * C + c], sh_mm[C + c]);
  }
}
template <typename T, int V>
__global__ void __launch_bounds__(256, 4) minmax_apply_kernel(const T* __restrict__ x, int HW, int C, int rows_per_block, const uint32_t* __restrict__ mn_ord,
                                    const uint32_t* __restrict__ mx_ord, T* __restrict__ gate, float* mn, float* mx) {
  const int CV = C / V;
  const int lanes = blockDim.x / CV;
  const int v = threadIdx.x % CV, rl = threadIdx.x / CV;
  if (rl >= lanes) return;
  const int n = blockIdx.y;
  float lo[V], inv[V];           // (x - mn) / (mx - mn) as (x - mn) * inv: one IEEE division per (n, c) instead of one per element
#pragma unroll
  for (int k = 0; k < V; k++) {
    long long q = (long long)n * C + v * V + k;
    lo[k] = ord2f(mn_ord[q]);
    const float hi = ord2f(mx_ord[q]);
    inv[k] = 1.0f / (hi - lo[k]);
    if (blockIdx.x == 0 && rl == 0) { mn[q] = lo[k]; mx[q] = hi; }
  }
  const int r0 = blockIdx.x * rows_per_block, r1 = min(r0 + rows_per_block, HW);
  const long long base = (long long)n * HW * C + v * V;
  walk_rows<T, V, 4>(x + base, C, r0 + rl, r1, lanes, [&](int r, const float (&a)[kMaxV]) {
    float o[kMaxV];
#pragma unroll
    for (int k = 0; k < V; k++) o[k] = (a[k] - lo[k]) * inv[k];
    stv<T, V>(gate + base + (long long)r * C, o);
  });
}
// sums[0] = sum g*(x-mn), sums[1] = sum g, sums[2] = #(x==mx), sums[3] = #(x==mn)   each [N,C]
// (the two min-max backward kernels carry 48 per-(n,c) registers: four rows per tensor in flight at 2 blocks per SM measured
// 0.540 ms for the pair on the 604 MB tensor against 0.632 with two rows at 3 blocks per SM; the cBN pair prefers the latter)
template <typename T, int V>
__global__ void __launch_bounds__(256, 2) minmax_bwd_reduce_kernel(const T* __restrict__ gg, const T* __restrict__ x, int HW, int C, int rows_per_block,
                                         const float* __restrict__ mn, const float* __restrict__ mx, int N, float* sums) {
  extern __shared__ float sh_red[];              // [lanes][4][C] per-thread partial sums, added per column below
  const int CV = C / V;
  const int lanes = blockDim.x / CV;
  const int v = threadIdx.x % CV, rl = threadIdx.x / CV;
  const int n = blockIdx.y;
  int r0 = blockIdx.x * rows_per_block, r1 = rl < lanes ? min(r0 + rows_per_block, HW) : 0;
  float lo[V], hi[V], s0[V], s1[V], c0[V], c1[V];
#pragma unroll
  for (int k = 0; k < V; k++) {
    lo[k] = mn[(long long)n * C + v * V + k]; hi[k] = mx[(long long)n * C + v * V + k];
    s0[k] = s1[k] = c0[k] = c1[k] = 0.f;
  }
  {
    const long long base = (long long)n * HW * C + v * V;
    walk_rows2<T, V, 4>(x + base, gg + base, C, r0 + rl, r1, lanes, [&](int, const float (&a)[kMaxV], const float (&g)[kMaxV]) {
#pragma unroll
      for (int k = 0; k < V; k++) {
        s0[k] += g[k] * (a[k] - lo[k]); s1[k] += g[k];
        c0[k] += (a[k] == hi[k]) ? 1.f : 0.f;
        c1[k] += (a[k] == lo[k]) ? 1.f : 0.f;
      }
    });
  }
  if (rl < lanes) {
    float* mine = sh_red + (size_t)rl * 4 * C + v * V;
#pragma unroll
    for (int k = 0; k < V; k++) { mine[k] = s0[k]; mine[C + k] = s1[k]; mine[2 * C + k] = c0[k]; mine[3 * C + k] = c1[k]; }
  }
  __syncthreads();
  long long NC = (long long)N * C;
  for (int i = threadIdx.x; i < 4 * C; i += blockDim.x) {
    float val = 0.f;
    for (int q2 = 0; q2 < lanes; q2++) val += sh_red[(size_t)q2 * 4 * C + i];
    int q = i / C, c = i - q * C;
    if (val != 0.f) atomicAdd(&sums[q * NC + (long long)n * C + c], val);
  }
}
template <typename T, int V>
__global__ void __launch_bounds__(256, 2) minmax_bwd_apply_kernel(const T* __restrict__ gg, const T* __restrict__ x, int HW, int C, int rows_per_block,
                                        const float* __restrict__ mn, const float* __restrict__ mx, int N,
                                        const float* __restrict__ sums, T* __restrict__ gpre, float* dbias) {
  extern __shared__ float sh_db[];               // [lanes][C] per-thread column sums of gpre (only when dbias != NULL)
  const int CV = C / V;
  const int lanes = blockDim.x / CV;
  const int v = threadIdx.x % CV, rl = threadIdx.x / CV;
  const bool on = rl < lanes;
  const int n = blockIdx.y;
  const long long NC = (long long)N * C;
  float lo[V], hi[V], invd[V], a_mx[V], a_mn[V], bs[V];   // per-(n,c): min, max, 1/range, shares of the arg-max / arg-min pixels
#pragma unroll
  for (int k = 0; k < V; k++) {
    long long q = (long long)n * C + v * V + k;
    lo[k] = mn[q]; hi[k] = mx[q];
    const float d = hi[k] - lo[k];
    invd[k] = 1.0f / d;
    float A = sums[q], S = sums[NC + q];
    a_mx[k] = (-A / (d * d)) / sums[2 * NC + q];
    a_mn[k] = ((A - d * S) / (d * d)) / sums[3 * NC + q];
    bs[k] = 0.f;
  }
  const int r0 = blockIdx.x * rows_per_block, r1 = on ? min(r0 + rows_per_block, HW) : 0;
  const long long base = (long long)n * HW * C + v * V;
  walk_rows2<T, V, 4>(x + base, gg + base, C, r0 + rl, r1, lanes, [&](int r, const float (&a)[kMaxV], const float (&g)[kMaxV]) {
    float o[kMaxV];
#pragma unroll
    for (int k = 0; k < V; k++) {
      float rr = g[k] * invd[k];
      if (a[k] == hi[k]) rr += a_mx[k];
      if (a[k] == lo[k]) rr += a_mn[k];
      o[k] = rr * (a[k] > 0.f ? 1.f : 0.2f);
      bs[k] += o[k];
    }
    stv<T, V>(gpre + base + (long long)r * C, o);
  });
  if (dbias) {                                   // bias gradient of the convolution in front: column sums of the result
    if (on) {
      float* mine = sh_db + (size_t)rl * C + v * V;
#pragma unroll
      for (int k = 0; k < V; k++) mine[k] = bs[k];
    }
    __syncthreads();
    for (int i = threadIdx.x; i < C; i += blockDim.x) {
      float t = 0.f;
      for (int q2 = 0; q2 < lanes; q2++) t += sh_db[(size_t)q2 * C + i];
      atomicAdd(dbias + i, t);
    }
  }
}

// ======================================================================================================
// activation backward from the output
// ======================================================================================================
template <typename T, int V>
__global__ void act_bwd_kernel(const T* __restrict__ gy, const T* __restrict__ y, long long nvec, int act, T* __restrict__ gx) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < nvec; i += (long long)gridDim.x * blockDim.x) {
    float g[kMaxV], v[kMaxV], o[kMaxV];
    ldv<T, V>(gy + i * V, g);
    ldv<T, V>(y + i * V, v);
#pragma unroll
    for (int k = 0; k < V; k++) {
      if (act == FGC_ACT_TANH) o[k] = g[k] * (1.f - v[k] * v[k]);
      else {  // miu_relu: x = y - 0.0225/y
        float xx = v[k] - 0.0225f / v[k];
        o[k] = g[k] * miu_relu_grad(xx);
      }
    }
    stv<T, V>(gx + i * V, o);
  }
}

// ======================================================================================================
// gating
// ======================================================================================================
template <typename T, int V>
__global__ void gate_fma_fwd_kernel(const T* __restrict__ ht, const T* __restrict__ rg, const T* __restrict__ im, long long nvec,
                                    T* __restrict__ out) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < nvec; i += (long long)gridDim.x * blockDim.x) {
    float a[kMaxV], b[kMaxV], c[kMaxV], o[kMaxV];
    ldv<T, V>(ht + i * V, a); ldv<T, V>(rg + i * V, b); ldv<T, V>(im + i * V, c);
#pragma unroll
    for (int k = 0; k < V; k++) o[k] = a[k] + b[k] * c[k];
    stv<T, V>(out + i * V, o);
  }
}
template <typename T, int V>
__global__ void gate_fma_bwd_kernel(const T* __restrict__ g, const T* __restrict__ rg, const T* __restrict__ im, long long nvec,
                                    T* __restrict__ g_rg, T* __restrict__ g_im) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < nvec; i += (long long)gridDim.x * blockDim.x) {
    float a[kMaxV], b[kMaxV], c[kMaxV], o1[kMaxV], o2[kMaxV];
    ldv<T, V>(g + i * V, a); ldv<T, V>(rg + i * V, b); ldv<T, V>(im + i * V, c);
#pragma unroll
    for (int k = 0; k < V; k++) { o1[k] = a[k] * c[k]; o2[k] = a[k] * b[k]; }
    stv<T, V>(g_rg + i * V, o1); stv<T, V>(g_im + i * V, o2);
  }
}

// The discriminator's cell runs PReLU directly on ht + rg*im (mru.py:426-430 with prelu as the normaliser): one pass instead
// of two, and the sum is never stored -- the backward pass recomputes it from the three operands it keeps anyway.
template <typename T, int V>
__global__ void gate_prelu_fwd_kernel(const T* __restrict__ ht, const T* __restrict__ rg, const T* __restrict__ im, long long nvec,
                                      const float* __restrict__ ap, T* __restrict__ out) {
  const float a = *ap;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < nvec; i += (long long)gridDim.x * blockDim.x) {
    float h[kMaxV], b[kMaxV], c[kMaxV], o[kMaxV];
    ldv<T, V>(ht + i * V, h); ldv<T, V>(rg + i * V, b); ldv<T, V>(im + i * V, c);
#pragma unroll
    for (int k = 0; k < V; k++) {
      const float x = fmaf(b[k], c[k], h[k]);
      o[k] = (a * x >= x) ? a * x : x;
    }
    stv<T, V>(out + i * V, o);
  }
}
// g_x = g_p * prelu'(x), x = ht + rg*im recomputed;  da += sum g_p * x over the leaky side;  g_rg = g_x*im, g_im = g_x*rg,
// g_ht (+)= g_x  (acc_ht: add into the gradient the skip branch has already written there; g_ht may be NULL)
template <typename T, int V>
__global__ void gate_prelu_bwd_kernel(const T* __restrict__ gp, const T* __restrict__ ht, const T* __restrict__ rg,
                                      const T* __restrict__ im, long long nvec, const float* __restrict__ ap, float* da,
                                      T* __restrict__ g_ht, int acc_ht, T* __restrict__ g_rg, T* __restrict__ g_im) {
  __shared__ float red[32];
  const float a = *ap;
  float acc = 0.f;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < nvec; i += (long long)gridDim.x * blockDim.x) {
    float g[kMaxV], h[kMaxV], b[kMaxV], c[kMaxV], gx[kMaxV], o1[kMaxV], o2[kMaxV];
    ldv<T, V>(gp + i * V, g); ldv<T, V>(ht + i * V, h); ldv<T, V>(rg + i * V, b); ldv<T, V>(im + i * V, c);
#pragma unroll
    for (int k = 0; k < V; k++) {
      const float x = fmaf(b[k], c[k], h[k]);
      const bool m = a * x >= x;
      gx[k] = m ? a * g[k] : g[k];
      if (m) acc += g[k] * x;
      o1[k] = gx[k] * c[k];
      o2[k] = gx[k] * b[k];
    }
    stv<T, V>(g_rg + i * V, o1); stv<T, V>(g_im + i * V, o2);
    if (g_ht) {
      if (acc_ht) {
        float old[kMaxV];
        ldv<T, V>(g_ht + i * V, old);
#pragma unroll
        for (int k = 0; k < V; k++) gx[k] += old[k];
      }
      stv<T, V>(g_ht + i * V, gx);
    }
  }
  if (da) {
    float t = block_sum(acc, red);
    if (threadIdx.x == 0) atomicAdd(da, t);
  }
}

// index helpers for low-res thread -> 4 full-res positions.  low-res tensor [N,h,w,C], full [N,2h,2w,C]
struct UpIdx {
  long long lo;      // element offset in the low-res tensor
  long long hi[4];   // element offsets in the full-res tensor
};
template <int V>
__device__ __forceinline__ UpIdx up_index(long long i, int h, int w, int C) {
  const int CV = C / V;
  int cv = (int)(i % CV);
  long long p = i / CV;
  int x = (int)(p % w);
  long long q = p / w;
  int y = (int)(q % h);
  long long n = q / h;
  UpIdx r;
  r.lo = i * V;
  long long base = ((n * (2 * h) + 2 * y) * (2LL * w) + 2 * x) * C + cv * V;
  r.hi[0] = base; r.hi[1] = base + C; r.hi[2] = base + 2LL * w * C; r.hi[3] = base + 2LL * w * C + C;
  return r;
}

template <typename T, int V>
__global__ void mul_up_fwd_kernel(const T* __restrict__ rg, const T* __restrict__ ht, long long nlow, int h, int w, int C,
                                  T* __restrict__ out) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < nlow; i += (long long)gridDim.x * blockDim.x) {
    UpIdx u = up_index<V>(i, h, w, C);
    float a[kMaxV];
    ldv<T, V>(ht + u.lo, a);
#pragma unroll
    for (int j = 0; j < 4; j++) {
      float b[kMaxV], o[kMaxV];
      ldv<T, V>(rg + u.hi[j], b);
#pragma unroll
      for (int k = 0; k < V; k++) o[k] = a[k] * b[k];
      stv<T, V>(out + u.hi[j], o);
    }
  }
}
template <typename T, int V>
__global__ void mul_up_bwd_kernel(const T* __restrict__ g, const T* __restrict__ rg, const T* __restrict__ ht, long long nlow,
                                  int h, int w, int C, T* __restrict__ g_rg, T* __restrict__ g_ht) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < nlow; i += (long long)gridDim.x * blockDim.x) {
    UpIdx u = up_index<V>(i, h, w, C);
    float a[kMaxV], acc[kMaxV] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    ldv<T, V>(ht + u.lo, a);
#pragma unroll
    for (int j = 0; j < 4; j++) {
      float gg[kMaxV], b[kMaxV], o[kMaxV];
      ldv<T, V>(g + u.hi[j], gg);
      ldv<T, V>(rg + u.hi[j], b);
#pragma unroll
      for (int k = 0; k < V; k++) { o[k] = gg[k] * a[k]; acc[k] += gg[k] * b[k]; }
      stv<T, V>(g_rg + u.hi[j], o);
    }
    stv<T, V>(g_ht + u.lo, acc);
  }
}
template <typename T, int V>
__global__ void blend_fwd_kernel(const T* __restrict__ sk, const T* __restrict__ h2, const T* __restrict__ zg, long long nlow,
                                 int h, int w, int C, T* __restrict__ out) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < nlow; i += (long long)gridDim.x * blockDim.x) {
    UpIdx u = up_index<V>(i, h, w, C);
    float s[kMaxV];
    ldv<T, V>(sk + u.lo, s);
#pragma unroll
    for (int j = 0; j < 4; j++) {
      float a[kMaxV], z[kMaxV], o[kMaxV];
      ldv<T, V>(h2 + u.hi[j], a);
      ldv<T, V>(zg + u.hi[j], z);
#pragma unroll
      for (int k = 0; k < V; k++) o[k] = s[k] * (1.f - z[k]) + a[k] * z[k];
      stv<T, V>(out + u.hi[j], o);
    }
  }
}
template <typename T, int V>
__global__ void blend_bwd_kernel(const T* __restrict__ g, const T* __restrict__ sk, const T* __restrict__ h2,
                                 const T* __restrict__ zg, long long nlow, int h, int w, int C, T* __restrict__ g_sk,
                                 T* __restrict__ g_h2, T* __restrict__ g_zg) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < nlow; i += (long long)gridDim.x * blockDim.x) {
    UpIdx u = up_index<V>(i, h, w, C);
    float s[kMaxV], acc[kMaxV] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    ldv<T, V>(sk + u.lo, s);
#pragma unroll
    for (int j = 0; j < 4; j++) {
      float gg[kMaxV], a[kMaxV], z[kMaxV], o1[kMaxV], o2[kMaxV];
      ldv<T, V>(g + u.hi[j], gg);
      ldv<T, V>(h2 + u.hi[j], a);
      ldv<T, V>(zg + u.hi[j], z);
#pragma unroll
      for (int k = 0; k < V; k++) {
        acc[k] += gg[k] * (1.f - z[k]);
        o1[k] = gg[k] * z[k];
        o2[k] = gg[k] * (a[k] - s[k]);
      }
      stv<T, V>(g_h2 + u.hi[j], o1);
      stv<T, V>(g_zg + u.hi[j], o2);
    }
    stv<T, V>(g_sk + u.lo, acc);
  }
}
template <typename T, int V>
__global__ void addpool_fwd_kernel(const T* __restrict__ a, const T* __restrict__ b, long long nlow, int h, int w, int C,
                                   T* __restrict__ out) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < nlow; i += (long long)gridDim.x * blockDim.x) {
    UpIdx u = up_index<V>(i, h, w, C);
    float acc[kMaxV] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int j = 0; j < 4; j++) {
      float p[kMaxV];
      ldv<T, V>(a + u.hi[j], p);
#pragma unroll
      for (int k = 0; k < V; k++) acc[k] += p[k];
      if (b) {
        ldv<T, V>(b + u.hi[j], p);
#pragma unroll
        for (int k = 0; k < V; k++) acc[k] += p[k];
      }
    }
#pragma unroll
    for (int k = 0; k < V; k++) acc[k] *= 0.25f;
    stv<T, V>(out + u.lo, acc);
  }
}
template <typename T, int V>
__global__ void unpool_bwd_kernel(const T* __restrict__ g, long long nlow, int h, int w, int C, float scale, T* __restrict__ out) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < nlow; i += (long long)gridDim.x * blockDim.x) {
    UpIdx u = up_index<V>(i, h, w, C);
    float a[kMaxV];
    ldv<T, V>(g + u.lo, a);
#pragma unroll
    for (int k = 0; k < V; k++) a[k] *= scale;
#pragma unroll
    for (int j = 0; j < 4; j++) stv<T, V>(out + u.hi[j], a);
  }
}
// out[N,h,w,C] (+)= sum2x2(g[N,2h,2w,C])   (used by the upsampled-source dgrad)
template <typename TI, typename TO>
__global__ void sum2x2_kernel(const TI* __restrict__ g, long long nlow, int h, int w, int C, TO* __restrict__ out, int acc) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < nlow; i += (long long)gridDim.x * blockDim.x) {
    UpIdx u = up_index<1>(i, h, w, C);
    float s = ld1<TI>(g + u.hi[0]) + ld1<TI>(g + u.hi[1]) + ld1<TI>(g + u.hi[2]) + ld1<TI>(g + u.hi[3]);
    if (acc) s += ld1<TO>(out + u.lo);
    st1<TO>(out + u.lo, s);
  }
}

template <typename TD, typename TS>
__global__ void axpy_kernel(TD* __restrict__ dst, const TS* __restrict__ src, long long n, float alpha) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    st1<TD>(dst + i, ld1<TD>(dst + i) + alpha * ld1<TS>(src + i));
}
template <typename T, int V>
__global__ void axpy_vec_kernel(T* __restrict__ dst, const T* __restrict__ src, long long nvec, float alpha) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < nvec; i += (long long)gridDim.x * blockDim.x) {
    float a[kMaxV], b[kMaxV];
    ldv<T, V>(dst + i * V, a);
    ldv<T, V>(src + i * V, b);
#pragma unroll
    for (int k = 0; k < V; k++) a[k] += alpha * b[k];
    stv<T, V>(dst + i * V, a);
  }
}

template <typename T>
__global__ void spatial_mean_fwd_kernel(const T* __restrict__ x, int HW, int C, T* __restrict__ out) {
  int n = blockIdx.y;
  int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  float s = 0.f;
  for (int r = 0; r < HW; r++) s += ld1<T>(x + ((long long)n * HW + r) * C + c);
  st1<T>(out + (long long)n * C + c, s / (float)HW);
}
template <typename T>
__global__ void spatial_mean_bwd_kernel(const T* __restrict__ g, long long total, int HW, int C, T* __restrict__ out) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    int c = (int)(i % C);
    long long n = i / ((long long)HW * C);
    st1<T>(out + i, ld1<T>(g + n * C + c) / (float)HW);
  }
}

template <typename TI, typename TO>
__global__ void nchw_to_nhwc_kernel(const TI* __restrict__ x, long long total, int C, int HW, TO* __restrict__ y) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    int c = (int)(i % C);
    long long p = i / C;
    int s = (int)(p % HW);
    long long n = p / HW;
    st1<TO>(y + i, ld1<TI>(x + (n * C + c) * HW + s));
  }
}
template <typename TI, typename TO>
__global__ void nhwc_to_nchw_kernel(const TI* __restrict__ x, long long total, int C, int HW, TO* __restrict__ y) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    int s = (int)(i % HW);
    long long p = i / HW;
    int c = (int)(p % C);
    long long n = p / C;
    st1<TO>(y + i, ld1<TI>(x + (n * HW + s) * C + c));
  }
}
template <typename TI, typename TO>
__global__ void cast_kernel(const TI* __restrict__ x, TO* __restrict__ y, long long n) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    st1<TO>(y + i, ld1<TI>(x + i));
}

// ======================================================================================================
// layout passes of the phase-form 4x4 convolutions (Pix2Pix / Residual variants, models_collection.py:380-405)
// ======================================================================================================
// TO_DEPTH: dst[n,y,x,(py*2+px)*C + c] = src[n,2y+py,2x+px,c] (tf.space_to_depth order); else the inverse copy
template <typename T, int V, bool TO_DEPTH>
__device__ __forceinline__ void depth_space_body(const T* __restrict__ src, long long nlow, int h, int w, int C, T* __restrict__ dst) {
  const int CV = C / V;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < nlow; i += (long long)gridDim.x * blockDim.x) {
    UpIdx u = up_index<V>(i, h, w, C);                      // hi[py*2+px]: the four full-resolution pixels of low-res pixel i / CV
    const long long d = (i / CV) * 4LL * C + (i % CV) * V;  // this channel vector inside phase 0 of the depth tensor
    float a[kMaxV];
#pragma unroll
    for (int j = 0; j < 4; j++) {
      if (TO_DEPTH) {
        ldv<T, V>(src + u.hi[j], a);
        stv<T, V>(dst + d + (long long)j * C, a);
      } else {
        ldv<T, V>(src + d + (long long)j * C, a);
        stv<T, V>(dst + u.hi[j], a);
      }
    }
  }
}
template <typename T, int V>
__global__ void space_to_depth_kernel(const T* __restrict__ src, long long nlow, int h, int w, int C, T* __restrict__ dst) {
  depth_space_body<T, V, true>(src, nlow, h, w, C, dst);
}
template <typename T, int V>
__global__ void depth_to_space_kernel(const T* __restrict__ src, long long nlow, int h, int w, int C, T* __restrict__ dst) {
  depth_space_body<T, V, false>(src, nlow, h, w, C, dst);
}
// dst[N,H,W,C]: top-left min(h,H) x min(w,W) rectangle of src[N,h,w,C], zero elsewhere
template <typename T, int V>
__global__ void copy_rect_kernel(const T* __restrict__ src, int h, int w, int C, T* __restrict__ dst, long long nvec, int H, int W) {
  const int CV = C / V;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < nvec; i += (long long)gridDim.x * blockDim.x) {
    int cv = (int)(i % CV);
    long long p = i / CV;
    int x = (int)(p % W);
    long long q = p / W;
    int y = (int)(q % H);
    long long n = q / H;
    float a[kMaxV];
#pragma unroll
    for (int k = 0; k < V; k++) a[k] = 0.f;
    if (y < h && x < w) ldv<T, V>(src + ((n * h + y) * (long long)w + x) * C + (long long)cv * V, a);
    stv<T, V>(dst + i * V, a);
  }
}
// position of 4x4 filter element (ky, kx, a, b) of f[4,4,A,B] inside the expanded odd-size filter (see fgc_phase_weights)
__device__ __forceinline__ long long phase_dest(int e, int A, int B, int mode) {
  int b = e % B;
  int t = e / B;
  int a = t % A;
  t /= A;
  int kx = t & 3, ky = t >> 2;
  if (mode == 2) return (((long long)(ky + 1) * 5 + (kx + 1)) * A + a) * B + b;
  const int py = (ky & 1) ^ 1, px = (kx & 1) ^ 1;                          // ky: 0 1 2 3 -> phase 1 0 1 0 (both modes)
  if (mode == 0) {                                                         // rows 0 1 1 2: input row 2y+ky-1 = 2(y+dy)+py
    const int ry = (ky + 1) >> 1, rx = (kx + 1) >> 1;
    return (((long long)(ry * 3 + rx)) * (4 * A) + (py * 2 + px) * A + a) * B + b;
  }
  const int ry = 2 - ((ky + 1) >> 1), rx = 2 - ((kx + 1) >> 1);            // rows 2 1 1 0: tap ky = py + 1 - 2 dy
  return (((long long)(ry * 3 + rx)) * B + b) * (4 * A) + (py * 2 + px) * A + a;
}
__global__ void phase_weights_kernel(const float* __restrict__ f, int n, int A, int B, int mode, float* __restrict__ w) {
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < n; e += gridDim.x * blockDim.x) w[phase_dest(e, A, B, mode)] = __ldg(f + e);
}
__global__ void phase_wgrad_kernel(const float* __restrict__ dw, int n, int A, int B, int mode, float* __restrict__ df) {
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < n; e += gridDim.x * blockDim.x) df[e] += __ldg(dw + phase_dest(e, A, B, mode));
}

// choose block / rows-per-block for the per-(n,c) row reductions
struct RowRed { int threads, V, lanes, rows_per_block, nblk; };
static RowRed rowred_plan(int C, int vec, long long rows, long long other_blocks) {
  RowRed p;
  p.V = vec >= 8 ? 8 : (vec >= 4 ? 4 : 1);
  int CV = C / p.V;
  p.threads = 256;
  if (CV > 256) p.threads = ((CV + 31) / 32) * 32;
  p.lanes = p.threads / CV;
  static int waves = 0;                                             // env FGC_ROWRED_WAVES: tuning knob (blocks per SM overall)
  if (!waves) { const char* e = getenv("FGC_ROWRED_WAVES"); waves = e ? atoi(e) : 24; if (waves < 1) waves = 24; }
  long long target = (long long)num_sms() * waves;                  // blocks wanted overall (several waves: short tail)
  long long want = target / (other_blocks > 0 ? other_blocks : 1);
  if (want < 1) want = 1;
  long long rpb = (rows + want - 1) / want;
  long long min_rpb = (long long)p.lanes * 8;
  if (rpb < min_rpb) rpb = min_rpb;
  if (rpb > rows) rpb = rows;
  p.rows_per_block = (int)rpb;
  p.nblk = (int)((rows + rpb - 1) / rpb);
  return p;
}


// ---- terms of the three-way bf16 split of an fp32 tensor (the six-product mode of the convolutions):
// x = x1 + x2 + x3 + O(2^-24 x) with x1 = bf16(x), x2 = bf16(x - x1), x3 = bf16(x - x1 - x2)
__global__ void split_term_kernel(const float* __restrict__ x, long long n, int level, __nv_bfloat16* __restrict__ out_bf16,
                                  float* __restrict__ out_f32) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    float r = x[i];
    for (int l = 0; l < level; l++) r -= __bfloat162float(__float2bfloat16_rn(r));      // residual after `level` terms
    if (out_bf16) out_bf16[i] = __float2bfloat16_rn(r);
    if (out_f32) out_f32[i] = r;
  }
}

// ---- second half of a column-folded k x k convolution with few outputs (the 7x7, 64 -> 3 head): z[n,h,w, kw*Cout + co] holds
// the k x 1 (vertical) convolution with filter column kw; y[n,h,w,co] = act(bias[co] + sum_kw z[n,h,w + kw - pad, kw*Cout + co])
template <typename TO>
__global__ void tapsum_w_kernel(const float* __restrict__ z, long long npix, int W, int k, int Cout, int Cz, const float* __restrict__ bias,
                                int act, TO* __restrict__ y) {
  const int pad = (k - 1) / 2;
  const long long total = npix * Cout;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int co = (int)(i % Cout);
    const long long p = i / Cout;
    const int w = (int)(p % W);
    float acc = bias ? bias[co] : 0.f;
    for (int kw = 0; kw < k; kw++) {
      const int ww = w + kw - pad;
      if (ww >= 0 && ww < W) acc += __ldg(z + (p + kw - pad) * Cz + kw * Cout + co);
    }
    if (act == FGC_ACT_TANH) acc = tanhf(acc);
    else if (act == FGC_ACT_LRELU) acc = acc > 0.f ? acc : 0.2f * acc;
    else if (act == FGC_ACT_MIU) acc = miu_relu(acc);
    else if (act == FGC_ACT_RELU) acc = fmaxf(acc, 0.f);
    st1<TO>(y + i, acc);
  }
}
// one thread per PIXEL with its CO outputs in registers (CO = 8: the folded input gradients towards the 8-channel stem features,
// two 16-byte loads per filter column and one 16-byte store; CO = 3: the 7x7 head).  The per-element form above splits a 64-bit
// index twice per output element (~100 instructions for 3 loads): 234 us for a 528 MB pass.
template <typename TO, int CO>
__global__ void __launch_bounds__(256) tapsum_w_pix_kernel(const float* __restrict__ z, unsigned npix, int W, int k, int Cz,
                                                          const float* __restrict__ bias, int act, TO* __restrict__ y) {
  const int pad = (k - 1) / 2;
  float b[CO];
#pragma unroll
  for (int q = 0; q < CO; q++) b[q] = bias ? __ldg(bias + q) : 0.f;
  for (unsigned p = blockIdx.x * blockDim.x + threadIdx.x; p < npix; p += gridDim.x * blockDim.x) {
    const int w = (int)(p % (unsigned)W);
    float acc[CO];
#pragma unroll
    for (int q = 0; q < CO; q++) acc[q] = b[q];
    for (int kw = 0; kw < k; kw++) {
      const int ww = w + kw - pad;
      if (ww < 0 || ww >= W) continue;
      const float* zp = z + ((long long)p + kw - pad) * Cz + kw * CO;
      if (CO % 4 == 0) {
#pragma unroll
        for (int q = 0; q < CO; q += 4) {
          const float4 t = __ldg(reinterpret_cast<const float4*>(zp + q));
          acc[q] += t.x; acc[q + 1] += t.y; acc[q + 2] += t.z; acc[q + 3] += t.w;
        }
      } else {
#pragma unroll
        for (int q = 0; q < CO; q++) acc[q] += __ldg(zp + q);
      }
    }
#pragma unroll
    for (int q = 0; q < CO; q++) {
      float v = acc[q];
      if (act == FGC_ACT_TANH) v = tanhf(v);
      else if (act == FGC_ACT_LRELU) v = v > 0.f ? v : 0.2f * v;
      else if (act == FGC_ACT_MIU) v = miu_relu(v);
      else if (act == FGC_ACT_RELU) v = fmaxf(v, 0.f);
      acc[q] = v;
    }
    if (CO == 8 && sizeof(TO) == 2) {
      *reinterpret_cast<uint4*>(y + (size_t)p * CO) = make_uint4(bf16x2_bits(acc[0], acc[1]), bf16x2_bits(acc[2], acc[3]),
                                                                 bf16x2_bits(acc[4], acc[5]), bf16x2_bits(acc[6], acc[7]));
    } else {
#pragma unroll
      for (int q = 0; q < CO; q++) st1<TO>(y + (size_t)p * CO + q, acc[q]);
    }
  }
}
}  // namespace fgc

using namespace fgc;

extern "C" {

int fgc_chan_stats(const void* x, int dtype, long long M, int C, double* acc, float* stats, fgc_stream stream) {
  FGC_REQUIRE(M > 0 && C > 0 && C <= 1024, "chan_stats: bad shape M=%lld C=%d", M, C);
  cudaStream_t s = as_stream(stream);
  cudaMemsetAsync(acc, 0, sizeof(double) * 2 * C, s);
  int vec = vec_width(x, C, dtype);
  RowRed p = rowred_plan(C, vec, M, 6);      // one wave of 4 blocks per SM: every block ends in 2*C fp64 atomics
  FGC_DISPATCH_TV(dtype, vec, T, V,
                  (chan_stats_kernel<T, V><<<p.nblk, p.threads, (size_t)p.lanes * 2 * C * sizeof(float), s>>>((const T*)x, M, C, p.rows_per_block, acc)));
  chan_stats_finalize<<<cdiv(C, 128), 128, 0, s>>>(acc, M, C, stats);
  count_launch(2);
  FGC_LAUNCH_CHECK("chan_stats");
  return FGC_OK;
}

int fgc_cbn_act_fwd(const void* x, int dtype, int N, int HW, int C, const float* stats, const float* scale,
                    const float* offset, const int32_t* labels, int act, void* y, fgc_stream stream) {
  FGC_REQUIRE(act == FGC_ACT_NONE || act == FGC_ACT_MIU, "cbn_act_fwd: act %d", act);
  cudaStream_t s = as_stream(stream);
  int vec = vmin(vec_width(x, C, dtype), vec_width(y, C, dtype));
  long long n = (long long)N * HW * C;
  (void)n;
  RowRed p = rowred_plan(C, vec, HW, N);
  FGC_DISPATCH_TV(dtype, vec, T, V, {
    cbn_act_fwd_kernel<T, V><<<dim3(p.nblk, N), p.threads, 0, s>>>((const T*)x, HW, C, p.rows_per_block, stats, scale, offset, labels,
                                                                  act, (T*)y);
  });
  count_launch();
  FGC_LAUNCH_CHECK("cbn_act_fwd");
  return FGC_OK;
}

int fgc_cbn_act_bwd(const void* gy, const void* x, int dtype, int N, int HW, int C, const float* stats,
                    const float* scale, const float* offset, const int32_t* labels, int act,
                    float* dscale, float* doffset, void* gx, float* scratch, float* dbias, fgc_stream stream) {
  cudaStream_t s = as_stream(stream);
  int vec = vmin(vmin(vec_width(x, C, dtype), vec_width(gy, C, dtype)), vec_width(gx, C, dtype));
  float* sums = scratch;                       // [2,N,C]
  float* m12 = scratch + 2LL * N * C;          // [2,C]
  cudaMemsetAsync(sums, 0, sizeof(float) * 2 * N * C, s);
  const int rvec = vec;
  RowRed p = rowred_plan(C, rvec, HW, N);
  long long n = (long long)N * HW * C;
  FGC_DISPATCH_TV(dtype, rvec, T, V, {
    cbn_bwd_reduce_kernel<T, V><<<dim3(p.nblk, N), p.threads, (size_t)p.lanes * 2 * C * sizeof(float), s>>>((const T*)gy, (const T*)x, HW, C,
                                                                                          p.rows_per_block, stats, scale, offset,
                                                                                          labels, act, N, sums);
  });
  cbn_bwd_finalize_kernel<<<cdiv(C, 8), 256, 0, s>>>(sums, N, C, (long long)N * HW, scale, labels, dscale, doffset, m12);
  (void)n;
  RowRed pa = rowred_plan(C, vec, HW, N);
  FGC_DISPATCH_TV(dtype, vec, T, V, {
    cbn_bwd_apply_kernel<T, V><<<dim3(pa.nblk, N), pa.threads, dbias ? (size_t)pa.lanes * C * sizeof(float) : 0, s>>>(
        (const T*)gy, (const T*)x, HW, C, pa.rows_per_block, stats, scale, offset, labels, act, m12, (T*)gx, dbias);
  });
  count_launch(3);
  FGC_LAUNCH_CHECK("cbn_act_bwd");
  return FGC_OK;
}

int fgc_prelu_fwd(const void* x, int dtype, long long n, const float* a, void* y, fgc_stream stream) {
  cudaStream_t s = as_stream(stream);
  int vec = vmin(vec_width(x, n, dtype), vec_width(y, n, dtype));
  FGC_DISPATCH_TV(dtype, vec, T, V, {
    long long nvec = n / V;
    prelu_fwd_kernel<T, V><<<ew_grid(nvec, 256), 256, 0, s>>>((const T*)x, nvec, a, (T*)y);
  });
  count_launch();
  FGC_LAUNCH_CHECK("prelu_fwd");
  return FGC_OK;
}
int fgc_tanh_fwd(const void* x, int dtype, long long n, void* y, fgc_stream stream) {
  FGC_REQUIRE(n > 0 && x && y, "tanh_fwd: bad arguments");
  cudaStream_t s = as_stream(stream);
  int vec = vmin(vec_width(x, n, dtype), vec_width(y, n, dtype));
  FGC_DISPATCH_TV(dtype, vec, T, V, {
    long long nvec = n / V;
    tanh_fwd_kernel<T, V><<<ew_grid(nvec, 256), 256, 0, s>>>((const T*)x, nvec, (T*)y);
  });
  count_launch();
  FGC_LAUNCH_CHECK("tanh_fwd");
  return FGC_OK;
}
static int prelu_bwd_run(const void* gy, const void* x, int dtype, long long n, int C, const float* a, float* da, float* dbias,
                         void* gx, int accumulate, fgc_stream stream);
int fgc_prelu_bwd(const void* gy, const void* x, int dtype, long long n, int C, const float* a, float* da, float* dbias,
                  void* gx, fgc_stream stream) {
  return prelu_bwd_run(gy, x, dtype, n, C, a, da, dbias, gx, 0, stream);
}
int fgc_prelu_bwd_acc(const void* gy, const void* x, int dtype, long long n, const float* a, float* da, void* gx, fgc_stream stream) {
  return prelu_bwd_run(gy, x, dtype, n, 0, a, da, nullptr, gx, 1, stream);
}
static int prelu_bwd_run(const void* gy, const void* x, int dtype, long long n, int C, const float* a, float* da, float* dbias,
                         void* gx, int accumulate, fgc_stream stream) {
  cudaStream_t s = as_stream(stream);
  if (dbias) {
    FGC_REQUIRE(C > 0 && n % C == 0 && C <= 1024, "prelu_bwd: bad channel count %d for the bias-gradient variant", C);
    int vec = vmin(vmin(vec_width(x, C, dtype), vec_width(gy, C, dtype)), vec_width(gx, C, dtype));
    const long long M = n / C;
    RowRed p = rowred_plan(C, vec, M, 1);
    FGC_DISPATCH_TV(dtype, vec, T, V, {
      prelu_bwd_rows_kernel<T, V><<<p.nblk, p.threads, (size_t)p.lanes * C * sizeof(float), s>>>((const T*)gy, (const T*)x, M, C, p.rows_per_block, a, da,
                                                                             (T*)gx, dbias);
    });
    count_launch();
    FGC_LAUNCH_CHECK("prelu_bwd");
    return FGC_OK;
  }
  int vec = vmin(vmin(vec_width(x, n, dtype), vec_width(gy, n, dtype)), vec_width(gx, n, dtype));
  FGC_DISPATCH_TV(dtype, vec, T, V, {
    long long nvec = n / V;
    prelu_bwd_kernel<T, V><<<ew_grid(nvec, 256), 256, 0, s>>>((const T*)gy, (const T*)x, nvec, a, da, (T*)gx, accumulate);
  });
  count_launch();
  FGC_LAUNCH_CHECK("prelu_bwd");
  return FGC_OK;
}

int fgc_minmax_fwd(const void* x, int dtype, int N, int HW, int C, void* gate, float* mn, float* mx,
                   uint32_t* scratch, fgc_stream stream) {
  cudaStream_t s = as_stream(stream);
  int vec = vmin(vec_width(x, C, dtype), vec_width(gate, C, dtype));
  uint32_t* mn_ord = scratch;
  uint32_t* mx_ord = scratch + (long long)N * C;
  cudaMemsetAsync(mn_ord, 0xFF, sizeof(uint32_t) * N * C, s);
  cudaMemsetAsync(mx_ord, 0x00, sizeof(uint32_t) * N * C, s);
  const int rvec = vec;
  RowRed p = rowred_plan(C, rvec, HW, N);
  long long n = (long long)N * HW * C;
  FGC_DISPATCH_TV(dtype, rvec, T, V, {
    minmax_reduce_kernel<T, V><<<dim3(p.nblk, N), p.threads, 2 * C * sizeof(uint32_t), s>>>((const T*)x, HW, C, p.rows_per_block,
                                                                                            mn_ord, mx_ord);
  });
  (void)n;
  RowRed pa = rowred_plan(C, vec, HW, N);
  FGC_DISPATCH_TV(dtype, vec, T, V, {
    minmax_apply_kernel<T, V><<<dim3(pa.nblk, N), pa.threads, 0, s>>>((const T*)x, HW, C, pa.rows_per_block, mn_ord, mx_ord, (T*)gate,
                                                                     mn, mx);
  });
  count_launch(2);
  FGC_LAUNCH_CHECK("minmax_fwd");
  return FGC_OK;
}
int fgc_minmax_bwd(const void* ggate, const void* x, int dtype, int N, int HW, int C, const float* mn,
                   const float* mx, void* gpre, float* scratch, float* dbias, fgc_stream stream) {
  cudaStream_t s = as_stream(stream);
  int vec = vmin(vmin(vec_width(x, C, dtype), vec_width(ggate, C, dtype)), vec_width(gpre, C, dtype));
  cudaMemsetAsync(scratch, 0, sizeof(float) * 4 * N * C, s);
  const int rvec = vec;
  RowRed p = rowred_plan(C, rvec, HW, N);
  long long n = (long long)N * HW * C;
  FGC_DISPATCH_TV(dtype, rvec, T, V, {
    minmax_bwd_reduce_kernel<T, V><<<dim3(p.nblk, N), p.threads, (size_t)p.lanes * 4 * C * sizeof(float), s>>>((const T*)ggate, (const T*)x, HW, C,
                                                                                             p.rows_per_block, mn, mx, N, scratch);
  });
  (void)n;
  RowRed pa = rowred_plan(C, vec, HW, N);
  FGC_DISPATCH_TV(dtype, vec, T, V, {
    minmax_bwd_apply_kernel<T, V><<<dim3(pa.nblk, N), pa.threads, dbias ? (size_t)pa.lanes * C * sizeof(float) : 0, s>>>(
        (const T*)ggate, (const T*)x, HW, C, pa.rows_per_block, mn, mx, N, scratch, (T*)gpre, dbias);
  });
  count_launch(2);
  FGC_LAUNCH_CHECK("minmax_bwd");
  return FGC_OK;
}

int fgc_act_bwd(const void* gy, const void* y, int dtype, long long n, int act, void* gx, fgc_stream stream) {
  FGC_REQUIRE(act == FGC_ACT_TANH || act == FGC_ACT_MIU, "act_bwd: act %d", act);
  cudaStream_t s = as_stream(stream);
  int vec = vmin(vmin(vec_width(gy, n, dtype), vec_width(y, n, dtype)), vec_width(gx, n, dtype));
  FGC_DISPATCH_TV(dtype, vec, T, V, {
    long long nvec = n / V;
    act_bwd_kernel<T, V><<<ew_grid(nvec, 256), 256, 0, s>>>((const T*)gy, (const T*)y, nvec, act, (T*)gx);
  });
  count_launch();
  FGC_LAUNCH_CHECK("act_bwd");
  return FGC_OK;
}

int fgc_gate_fma_fwd(const void* ht, const void* rg, const void* im, int dtype, long long n, void* out, fgc_stream stream) {
  cudaStream_t s = as_stream(stream);
  int vec = vmin(vmin(vmin(vec_width(ht, n, dtype), vec_width(rg, n, dtype)), vec_width(im, n, dtype)), vec_width(out, n, dtype));
  FGC_DISPATCH_TV(dtype, vec, T, V, {
    long long nvec = n / V;
    gate_fma_fwd_kernel<T, V><<<ew_grid(nvec, 256), 256, 0, s>>>((const T*)ht, (const T*)rg, (const T*)im, nvec, (T*)out);
  });
  count_launch();
  FGC_LAUNCH_CHECK("gate_fma_fwd");
  return FGC_OK;
}
int fgc_gate_fma_bwd(const void* g, const void* rg, const void* im, int dtype, long long n, void* g_rg, void* g_im,
                     fgc_stream stream) {
  cudaStream_t s = as_stream(stream);
  int vec = vmin(vmin(vmin(vmin(vec_width(g, n, dtype), vec_width(rg, n, dtype)), vec_width(im, n, dtype)), vec_width(g_rg, n, dtype)), vec_width(g_im, n, dtype));
  FGC_DISPATCH_TV(dtype, vec, T, V, {
    long long nvec = n / V;
    gate_fma_bwd_kernel<T, V><<<ew_grid(nvec, 256), 256, 0, s>>>((const T*)g, (const T*)rg, (const T*)im, nvec, (T*)g_rg, (T*)g_im);
  });
  count_launch();
  FGC_LAUNCH_CHECK("gate_fma_bwd");
  return FGC_OK;
}

int fgc_gate_prelu_fwd(const void* ht, const void* rg, const void* im, int dtype, long long n, const float* a, void* out,
                       fgc_stream stream) {
  FGC_REQUIRE(ht && rg && im && a && out && n > 0, "gate_prelu_fwd: bad arguments");
  cudaStream_t s = as_stream(stream);
  int vec = vmin(vmin(vmin(vec_width(ht, n, dtype), vec_width(rg, n, dtype)), vec_width(im, n, dtype)), vec_width(out, n, dtype));
  FGC_DISPATCH_TV(dtype, vec, T, V, {
    long long nvec = n / V;
    gate_prelu_fwd_kernel<T, V><<<ew_grid(nvec, 256), 256, 0, s>>>((const T*)ht, (const T*)rg, (const T*)im, nvec, a, (T*)out);
  });
  count_launch();
  FGC_LAUNCH_CHECK("gate_prelu_fwd");
  return FGC_OK;
}
int fgc_gate_prelu_bwd(const void* gp, const void* ht, const void* rg, const void* im, int dtype, long long n, const float* a,
                       float* da, void* g_ht, int acc_ht, void* g_rg, void* g_im, fgc_stream stream) {
  FGC_REQUIRE(gp && ht && rg && im && a && g_rg && g_im && n > 0, "gate_prelu_bwd: bad arguments");
  cudaStream_t s = as_stream(stream);
  int vec = vmin(vmin(vmin(vec_width(gp, n, dtype), vec_width(ht, n, dtype)), vec_width(rg, n, dtype)), vec_width(im, n, dtype));
  vec = vmin(vmin(vec, vec_width(g_rg, n, dtype)), vec_width(g_im, n, dtype));
  if (g_ht) vec = vmin(vec, vec_width(g_ht, n, dtype));
  FGC_DISPATCH_TV(dtype, vec, T, V, {
    long long nvec = n / V;
    gate_prelu_bwd_kernel<T, V><<<ew_grid(nvec, 256), 256, 0, s>>>((const T*)gp, (const T*)ht, (const T*)rg, (const T*)im, nvec, a, da,
                                                                 (T*)g_ht, acc_ht, (T*)g_rg, (T*)g_im);
  });
  count_launch();
  FGC_LAUNCH_CHECK("gate_prelu_bwd");
  return FGC_OK;
}

#define FGC_UP_LAUNCH(name, kern, ...)                                                          \
  do {                                                                                          \
    cudaStream_t s = as_stream(stream);                                                         \
    long long nlow = (long long)N * h * w * C;                                                  \
    FGC_DISPATCH_TV(dtype, vec, T, V, {                                                         \
      long long nv = nlow / V;                                                                  \
      kern<T, V><<<ew_grid(nv, 256), 256, 0, s>>>(__VA_ARGS__);                                 \
    });                                                                                         \
    count_launch();                                                                             \
    FGC_LAUNCH_CHECK(name);                                                                     \
    return FGC_OK;                                                                              \
  } while (0)

int fgc_mul_up_fwd(const void* rg, const void* ht_low, int dtype, int N, int h, int w, int C, void* out, fgc_stream stream) {
  int vec = vmin(vmin(vec_width(rg, C, dtype), vec_width(ht_low, C, dtype)), vec_width(out, C, dtype));
  FGC_UP_LAUNCH("mul_up_fwd", mul_up_fwd_kernel, (const T*)rg, (const T*)ht_low, nv, h, w, C, (T*)out);
}
int fgc_mul_up_bwd(const void* g, const void* rg, const void* ht_low, int dtype, int N, int h, int w, int C,
                   void* g_rg, void* g_ht_low, fgc_stream stream) {
  int vec = vmin(vmin(vmin(vmin(vec_width(g, C, dtype), vec_width(rg, C, dtype)), vec_width(ht_low, C, dtype)), vec_width(g_rg, C, dtype)), vec_width(g_ht_low, C, dtype));
  FGC_UP_LAUNCH("mul_up_bwd", mul_up_bwd_kernel, (const T*)g, (const T*)rg, (const T*)ht_low, nv, h, w, C, (T*)g_rg, (T*)g_ht_low);
}
int fgc_blend_fwd(const void* sk_low, const void* h2, const void* zg, int dtype, int N, int h, int w, int C, void* out,
                  fgc_stream stream) {
  int vec = vmin(vmin(vmin(vec_width(sk_low, C, dtype), vec_width(h2, C, dtype)), vec_width(zg, C, dtype)), vec_width(out, C, dtype));
  FGC_UP_LAUNCH("blend_fwd", blend_fwd_kernel, (const T*)sk_low, (const T*)h2, (const T*)zg, nv, h, w, C, (T*)out);
}
int fgc_blend_bwd(const void* g, const void* sk_low, const void* h2, const void* zg, int dtype, int N, int h, int w, int C,
                  void* g_sk_low, void* g_h2, void* g_zg, fgc_stream stream) {
  int vec = vmin(vmin(vmin(vmin(vmin(vmin(vec_width(g, C, dtype), vec_width(sk_low, C, dtype)), vec_width(h2, C, dtype)), vec_width(zg, C, dtype)), vec_width(g_sk_low, C, dtype)), vec_width(g_h2, C, dtype)), vec_width(g_zg, C, dtype));
  FGC_UP_LAUNCH("blend_bwd", blend_bwd_kernel, (const T*)g, (const T*)sk_low, (const T*)h2, (const T*)zg, nv, h, w, C,
                (T*)g_sk_low, (T*)g_h2, (T*)g_zg);
}
int fgc_addpool_fwd(const void* a, const void* b, int dtype, int N, int h, int w, int C, void* out, fgc_stream stream) {
  int vec = vmin(vmin(vec_width(a, C, dtype), b ? vec_width(b, C, dtype) : 8), vec_width(out, C, dtype));
  FGC_UP_LAUNCH("addpool_fwd", addpool_fwd_kernel, (const T*)a, (const T*)b, nv, h, w, C, (T*)out);
}
int fgc_unpool_bwd(const void* g, int dtype, int N, int h, int w, int C, void* out, fgc_stream stream) {
  int vec = vmin(vec_width(g, C, dtype), vec_width(out, C, dtype));
  FGC_UP_LAUNCH("unpool_bwd", unpool_bwd_kernel, (const T*)g, nv, h, w, C, 0.25f, (T*)out);
}
int fgc_upsample2x(const void* x, int dtype, int N, int h, int w, int C, void* out, fgc_stream stream) {
  int vec = vmin(vec_width(x, C, dtype), vec_width(out, C, dtype));
  const void* g = x;
  FGC_UP_LAUNCH("upsample2x", unpool_bwd_kernel, (const T*)g, nv, h, w, C, 1.0f, (T*)out);
}

int fgc_sum2x2(const void* g, int g_dtype, int N, int h, int w, int C, void* out, int out_dtype, int accumulate,
               fgc_stream stream) {
  cudaStream_t s = as_stream(stream);
  long long nlow = (long long)N * h * w * C;
  int grid = ew_grid(nlow, 256);
  if (g_dtype == FGC_F32 && out_dtype == FGC_F32)
    sum2x2_kernel<float, float><<<grid, 256, 0, s>>>((const float*)g, nlow, h, w, C, (float*)out, accumulate);
  else if (g_dtype == FGC_F32 && out_dtype == FGC_BF16)
    sum2x2_kernel<float, __nv_bfloat16><<<grid, 256, 0, s>>>((const float*)g, nlow, h, w, C, (__nv_bfloat16*)out, accumulate);
  else if (g_dtype == FGC_BF16 && out_dtype == FGC_BF16)
    sum2x2_kernel<__nv_bfloat16, __nv_bfloat16><<<grid, 256, 0, s>>>((const __nv_bfloat16*)g, nlow, h, w, C, (__nv_bfloat16*)out, accumulate);
  else if (g_dtype == FGC_BF16 && out_dtype == FGC_F32)
    sum2x2_kernel<__nv_bfloat16, float><<<grid, 256, 0, s>>>((const __nv_bfloat16*)g, nlow, h, w, C, (float*)out, accumulate);
  else { set_error("sum2x2: bad dtypes"); return FGC_EINVAL; }
  count_launch();
  FGC_LAUNCH_CHECK("sum2x2");
  return FGC_OK;
}

int fgc_axpy(void* dst, const void* src, int dst_dtype, int src_dtype, long long n, float alpha, fgc_stream stream) {
  cudaStream_t s = as_stream(stream);
  int vw = dst_dtype == src_dtype ? vmin(vec_width(dst, n, dst_dtype), vec_width(src, n, src_dtype)) : 1;
  if (vw >= 8) {
    if (dst_dtype == FGC_F32) axpy_vec_kernel<float, 8><<<ew_grid(n / 8, 256), 256, 0, s>>>((float*)dst, (const float*)src, n / 8, alpha);
    else axpy_vec_kernel<__nv_bfloat16, 8><<<ew_grid(n / 8, 256), 256, 0, s>>>((__nv_bfloat16*)dst, (const __nv_bfloat16*)src, n / 8, alpha);
  } else if (vw >= 4) {
    if (dst_dtype == FGC_F32) axpy_vec_kernel<float, 4><<<ew_grid(n / 4, 256), 256, 0, s>>>((float*)dst, (const float*)src, n / 4, alpha);
    else axpy_vec_kernel<__nv_bfloat16, 4><<<ew_grid(n / 4, 256), 256, 0, s>>>((__nv_bfloat16*)dst, (const __nv_bfloat16*)src, n / 4, alpha);
  } else {
    int grid = ew_grid(n, 256);
    if (dst_dtype == FGC_F32 && src_dtype == FGC_F32) axpy_kernel<float, float><<<grid, 256, 0, s>>>((float*)dst, (const float*)src, n, alpha);
    else if (dst_dtype == FGC_F32 && src_dtype == FGC_BF16) axpy_kernel<float, __nv_bfloat16><<<grid, 256, 0, s>>>((float*)dst, (const __nv_bfloat16*)src, n, alpha);
    else if (dst_dtype == FGC_BF16 && src_dtype == FGC_F32) axpy_kernel<__nv_bfloat16, float><<<grid, 256, 0, s>>>((__nv_bfloat16*)dst, (const float*)src, n, alpha);
    else if (dst_dtype == FGC_BF16 && src_dtype == FGC_BF16) axpy_kernel<__nv_bfloat16, __nv_bfloat16><<<grid, 256, 0, s>>>((__nv_bfloat16*)dst, (const __nv_bfloat16*)src, n, alpha);
    else { set_error("axpy: bad dtypes"); return FGC_EINVAL; }
  }
  count_launch();
  FGC_LAUNCH_CHECK("axpy");
  return FGC_OK;
}

int fgc_spatial_mean_fwd(const void* x, int dtype, int N, int HW, int C, void* out, fgc_stream stream) {
  cudaStream_t s = as_stream(stream);
  FGC_DISPATCH_DTYPE(dtype, T, (spatial_mean_fwd_kernel<T><<<dim3(cdiv(C, 128), N), 128, 0, s>>>((const T*)x, HW, C, (T*)out)));
  count_launch();
  FGC_LAUNCH_CHECK("spatial_mean_fwd");
  return FGC_OK;
}
int fgc_spatial_mean_bwd(const void* g, int dtype, int N, int HW, int C, void* out, fgc_stream stream) {
  cudaStream_t s = as_stream(stream);
  long long total = (long long)N * HW * C;
  FGC_DISPATCH_DTYPE(dtype, T, (spatial_mean_bwd_kernel<T><<<ew_grid(total, 256), 256, 0, s>>>((const T*)g, total, HW, C, (T*)out)));
  count_launch();
  FGC_LAUNCH_CHECK("spatial_mean_bwd");
  return FGC_OK;
}

#define FGC_DISPATCH_2(dt_in, dt_out, TI, TO, ...)                                                   \
  do {                                                                                               \
    if ((dt_in) == FGC_F32 && (dt_out) == FGC_F32) { using TI = float; using TO = float; __VA_ARGS__; }                 \
    else if ((dt_in) == FGC_F32 && (dt_out) == FGC_BF16) { using TI = float; using TO = __nv_bfloat16; __VA_ARGS__; }   \
    else if ((dt_in) == FGC_BF16 && (dt_out) == FGC_F32) { using TI = __nv_bfloat16; using TO = float; __VA_ARGS__; }   \
    else if ((dt_in) == FGC_BF16 && (dt_out) == FGC_BF16) { using TI = __nv_bfloat16; using TO = __nv_bfloat16; __VA_ARGS__; } \
    else { fgc::set_error("bad dtypes %d %d", (int)(dt_in), (int)(dt_out)); return FGC_EINVAL; }    \
  } while (0)

int fgc_nchw_to_nhwc(const void* x, int x_dtype, int N, int C, int HW, void* y, int y_dtype, fgc_stream stream) {
  cudaStream_t s = as_stream(stream);
  long long total = (long long)N * C * HW;
  FGC_DISPATCH_2(x_dtype, y_dtype, TI, TO,
                 (nchw_to_nhwc_kernel<TI, TO><<<ew_grid(total, 256), 256, 0, s>>>((const TI*)x, total, C, HW, (TO*)y)));
  count_launch();
  FGC_LAUNCH_CHECK("nchw_to_nhwc");
  return FGC_OK;
}
int fgc_nhwc_to_nchw(const void* x, int x_dtype, int N, int C, int HW, void* y, int y_dtype, fgc_stream stream) {
  cudaStream_t s = as_stream(stream);
  long long total = (long long)N * C * HW;
  FGC_DISPATCH_2(x_dtype, y_dtype, TI, TO,
                 (nhwc_to_nchw_kernel<TI, TO><<<ew_grid(total, 256), 256, 0, s>>>((const TI*)x, total, C, HW, (TO*)y)));
  count_launch();
  FGC_LAUNCH_CHECK("nhwc_to_nchw");
  return FGC_OK;
}
int fgc_cast(const void* x, int x_dtype, void* y, int y_dtype, long long n, fgc_stream stream) {
  cudaStream_t s = as_stream(stream);
  FGC_DISPATCH_2(x_dtype, y_dtype, TI, TO, (cast_kernel<TI, TO><<<ew_grid(n, 256), 256, 0, s>>>((const TI*)x, (TO*)y, n)));
  count_launch();
  FGC_LAUNCH_CHECK("cast");
  return FGC_OK;
}

int fgc_split_term(const float* x, long long n, int level, void* out_bf16, float* out_f32, fgc_stream stream) {
  FGC_REQUIRE(x && n > 0 && level >= 0 && level <= 2 && (out_bf16 || out_f32), "split_term: bad arguments");
  split_term_kernel<<<ew_grid(n, 256), 256, 0, as_stream(stream)>>>(x, n, level, (__nv_bfloat16*)out_bf16, out_f32);
  count_launch();
  FGC_LAUNCH_CHECK("split_term");
  return FGC_OK;
}

int fgc_tapsum_w(const float* z, int N, int H, int W, int k, int Cout, int Cz, const float* bias, int act, void* y, int y_dtype,
                 fgc_stream stream) {
  FGC_REQUIRE(z && y && k % 2 == 1 && Cz >= k * Cout, "tapsum_w: bad arguments");
  const long long npix = (long long)N * H * W;
  cudaStream_t s = as_stream(stream);
  const bool al16 = ((reinterpret_cast<uintptr_t>(z) | reinterpret_cast<uintptr_t>(y)) & 15) == 0 && (Cz & 3) == 0;
  const unsigned np = (unsigned)npix;
  if (npix < (1LL << 31) && Cout == 8 && al16) {
    if (y_dtype == FGC_F32) tapsum_w_pix_kernel<float, 8><<<ew_grid(npix, 256), 256, 0, s>>>(z, np, W, k, Cz, bias, act, (float*)y);
    else tapsum_w_pix_kernel<__nv_bfloat16, 8><<<ew_grid(npix, 256), 256, 0, s>>>(z, np, W, k, Cz, bias, act, (__nv_bfloat16*)y);
  } else if (npix < (1LL << 31) && Cout == 3) {
    if (y_dtype == FGC_F32) tapsum_w_pix_kernel<float, 3><<<ew_grid(npix, 256), 256, 0, s>>>(z, np, W, k, Cz, bias, act, (float*)y);
    else tapsum_w_pix_kernel<__nv_bfloat16, 3><<<ew_grid(npix, 256), 256, 0, s>>>(z, np, W, k, Cz, bias, act, (__nv_bfloat16*)y);
  } else if (y_dtype == FGC_F32) tapsum_w_kernel<float><<<ew_grid(npix * Cout, 256), 256, 0, s>>>(z, npix, W, k, Cout, Cz, bias, act, (float*)y);
  else tapsum_w_kernel<__nv_bfloat16><<<ew_grid(npix * Cout, 256), 256, 0, s>>>(z, npix, W, k, Cout, Cz, bias, act, (__nv_bfloat16*)y);
  count_launch();
  FGC_LAUNCH_CHECK("tapsum_w");
  return FGC_OK;
}

/* ---- phase-form layout passes ---- */
int fgc_space_to_depth(const void* x, int dtype, int N, int h, int w, int C, void* out, fgc_stream stream) {
  FGC_REQUIRE(N > 0 && h > 0 && w > 0 && C > 0 && x && out, "space_to_depth: bad arguments");
  int vec = vmin(vec_width(x, C, dtype), vec_width(out, C, dtype));
  FGC_UP_LAUNCH("space_to_depth", space_to_depth_kernel, (const T*)x, nv, h, w, C, (T*)out);
}
int fgc_depth_to_space(const void* x, int dtype, int N, int h, int w, int C, void* out, fgc_stream stream) {
  FGC_REQUIRE(N > 0 && h > 0 && w > 0 && C > 0 && x && out, "depth_to_space: bad arguments");
  int vec = vmin(vec_width(x, C, dtype), vec_width(out, C, dtype));
  FGC_UP_LAUNCH("depth_to_space", depth_to_space_kernel, (const T*)x, nv, h, w, C, (T*)out);
}
int fgc_copy_rect(const void* x, int dtype, int N, int h, int w, int C, void* out, int H, int W, fgc_stream stream) {
  FGC_REQUIRE(N > 0 && h > 0 && w > 0 && H > 0 && W > 0 && C > 0 && x && out, "copy_rect: bad arguments");
  cudaStream_t s = as_stream(stream);
  int vec = vmin(vec_width(x, C, dtype), vec_width(out, C, dtype));
  long long n = (long long)N * H * W * C;
  FGC_DISPATCH_TV(dtype, vec, T, V, {
    long long nv = n / V;
    copy_rect_kernel<T, V><<<ew_grid(nv, 256), 256, 0, s>>>((const T*)x, h, w, C, (T*)out, nv, H, W);
  });
  count_launch();
  FGC_LAUNCH_CHECK("copy_rect");
  return FGC_OK;
}
static long long phase_numel(int A, int B, int mode) { return (mode == 2 ? 25LL : 36LL) * A * B; }
int fgc_phase_weights(const float* f, int A, int B, int mode, float* w, fgc_stream stream) {
  FGC_REQUIRE(A > 0 && B > 0 && mode >= 0 && mode <= 2 && f && w && 16LL * A * B < (1LL << 31), "phase_weights: bad arguments");
  cudaStream_t s = as_stream(stream);
  if (cudaMemsetAsync(w, 0, sizeof(float) * phase_numel(A, B, mode), s) != cudaSuccess) return check_launch("phase_weights memset");
  int n = 16 * A * B;
  phase_weights_kernel<<<ew_grid(n, 256), 256, 0, s>>>(f, n, A, B, mode, w);
  count_launch();
  FGC_LAUNCH_CHECK("phase_weights");
  return FGC_OK;
}
int fgc_phase_wgrad(const float* dw, int A, int B, int mode, float* df, fgc_stream stream) {
  FGC_REQUIRE(A > 0 && B > 0 && mode >= 0 && mode <= 2 && dw && df && 16LL * A * B < (1LL << 31), "phase_wgrad: bad arguments");
  cudaStream_t s = as_stream(stream);
  int n = 16 * A * B;
  phase_wgrad_kernel<<<ew_grid(n, 256), 256, 0, s>>>(dw, n, A, B, mode, df);
  count_launch();
  FGC_LAUNCH_CHECK("phase_wgrad");
  return FGC_OK;
}

}  // extern "C"
