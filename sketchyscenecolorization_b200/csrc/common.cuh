// Shared helpers for libfgcolor kernels (sm_100a).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/fgcolor.h"

namespace fgc {

void set_error(const char* fmt, ...);
void count_launch(int n = 1);
int check_launch(const char* what);

#define FGC_REQUIRE(cond, ...)          \
  do {                                  \
    if (!(cond)) {                      \
      fgc::set_error(__VA_ARGS__);      \
      return FGC_EINVAL;                \
    }                                   \
  } while (0)

#define FGC_LAUNCH_CHECK(what)                   \
  do {                                           \
    int _e = fgc::check_launch(what);            \
    if (_e) return _e;                           \
  } while (0)

static inline cudaStream_t as_stream(fgc_stream s) { return reinterpret_cast<cudaStream_t>(s); }
static inline int cdiv(long long a, long long b) { return (int)((a + b - 1) / b); }
static inline size_t dsize(int dtype) { return dtype == FGC_BF16 ? 2 : 4; }

// ---- storage-type generic scalar / vector access -------------------------------------------------
template <typename T> __device__ __forceinline__ float ld1(const T* p);
template <> __device__ __forceinline__ float ld1<float>(const float* p) { return __ldg(p); }
template <> __device__ __forceinline__ float ld1<__nv_bfloat16>(const __nv_bfloat16* p) {
  return __bfloat162float(*p);
}
template <typename T> __device__ __forceinline__ void st1(T* p, float v);
template <> __device__ __forceinline__ void st1<float>(float* p, float v) { *p = v; }
template <> __device__ __forceinline__ void st1<__nv_bfloat16>(__nv_bfloat16* p, float v) {
  *p = __float2bfloat16_rn(v);
}

// V consecutive elements (V = 8, 4 or 1; pointer aligned to V elements), widened to fp32 in v[0..V)
constexpr int kMaxV = 8;
template <typename T, int V> __device__ __forceinline__ void ldv(const T* p, float (&v)[kMaxV]);
template <> __device__ __forceinline__ void ldv<float, 1>(const float* p, float (&v)[kMaxV]) { v[0] = __ldg(p); }
template <> __device__ __forceinline__ void ldv<float, 4>(const float* p, float (&v)[kMaxV]) {
  float4 t = __ldg(reinterpret_cast<const float4*>(p));
  v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
}
template <> __device__ __forceinline__ void ldv<float, 8>(const float* p, float (&v)[kMaxV]) {
  float4 t = __ldg(reinterpret_cast<const float4*>(p)), u = __ldg(reinterpret_cast<const float4*>(p) + 1);
  v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w; v[4] = u.x; v[5] = u.y; v[6] = u.z; v[7] = u.w;
}
template <> __device__ __forceinline__ void ldv<__nv_bfloat16, 1>(const __nv_bfloat16* p, float (&v)[kMaxV]) {
  v[0] = __bfloat162float(*p);
}
template <> __device__ __forceinline__ void ldv<__nv_bfloat16, 4>(const __nv_bfloat16* p, float (&v)[kMaxV]) {
  uint2 t = __ldg(reinterpret_cast<const uint2*>(p));
  float2 fa = __bfloat1622float2(*reinterpret_cast<__nv_bfloat162*>(&t.x)), fb = __bfloat1622float2(*reinterpret_cast<__nv_bfloat162*>(&t.y));
  v[0] = fa.x; v[1] = fa.y; v[2] = fb.x; v[3] = fb.y;
}
template <> __device__ __forceinline__ void ldv<__nv_bfloat16, 8>(const __nv_bfloat16* p, float (&v)[kMaxV]) {
  uint4 t = __ldg(reinterpret_cast<const uint4*>(p));
  float2 f0 = __bfloat1622float2(*reinterpret_cast<__nv_bfloat162*>(&t.x)), f1 = __bfloat1622float2(*reinterpret_cast<__nv_bfloat162*>(&t.y));
  float2 f2 = __bfloat1622float2(*reinterpret_cast<__nv_bfloat162*>(&t.z)), f3 = __bfloat1622float2(*reinterpret_cast<__nv_bfloat162*>(&t.w));
  v[0] = f0.x; v[1] = f0.y; v[2] = f1.x; v[3] = f1.y; v[4] = f2.x; v[5] = f2.y; v[6] = f3.x; v[7] = f3.y;
}
// ---- raw vector loads + row walker ------------------------------------------------------------------
// ldv widens right behind the load; in an unrolled loop ptxas then keeps "load, unpack, load, unpack" order and every thread
// has ONE load in flight (chan_stats_kernel: LDGs 27-32 instructions apart, 3.7 TB/s where a bare read loop of the same shape
// reaches 6.5, scripts/microbench/readbw.cu).  RawV keeps the loaded bytes as they are; walk_rows issues the U loads of a group
// of rows back to back, then their arithmetic.  PREFETCH = true also issues the NEXT group's loads before the arithmetic of the
// current one: measured slower everywhere (scripts/microbench/rowwalk.cu, profiles/r2ai_rowwalk_sweep.log: the second set of
// raw registers spills under the occupancy caps these kernels need) -- kept for the record, off.  Measured best on the 604 MB
// tensor: one tensor U = 4 at 4 blocks of 256 per SM (6.6 TB/s read-only, 6.3-6.5 read + write), two tensors U = 2 at 3 blocks
// per SM (6.4 TB/s), in a single launch of ~24 blocks per SM.
template <typename T, int V> struct RawV;
template <> struct RawV<__nv_bfloat16, 8> { uint4 q; };
template <> struct RawV<__nv_bfloat16, 4> { uint2 q; };
template <> struct RawV<__nv_bfloat16, 1> { unsigned short q; };
template <> struct RawV<float, 8> { float4 q, q1; };
template <> struct RawV<float, 4> { float4 q; };
template <> struct RawV<float, 1> { float q; };
__device__ __forceinline__ void ldraw(const __nv_bfloat16* p, RawV<__nv_bfloat16, 8>& r) { r.q = __ldg(reinterpret_cast<const uint4*>(p)); }
__device__ __forceinline__ void ldraw(const __nv_bfloat16* p, RawV<__nv_bfloat16, 4>& r) { r.q = __ldg(reinterpret_cast<const uint2*>(p)); }
__device__ __forceinline__ void ldraw(const __nv_bfloat16* p, RawV<__nv_bfloat16, 1>& r) { r.q = __ldg(reinterpret_cast<const unsigned short*>(p)); }
__device__ __forceinline__ void ldraw(const float* p, RawV<float, 8>& r) {
  r.q = __ldg(reinterpret_cast<const float4*>(p)); r.q1 = __ldg(reinterpret_cast<const float4*>(p) + 1);
}
__device__ __forceinline__ void ldraw(const float* p, RawV<float, 4>& r) { r.q = __ldg(reinterpret_cast<const float4*>(p)); }
__device__ __forceinline__ void ldraw(const float* p, RawV<float, 1>& r) { r.q = __ldg(p); }
__device__ __forceinline__ float bf16lo(uint32_t w) { return __uint_as_float(w << 16); }
__device__ __forceinline__ float bf16hi(uint32_t w) { return __uint_as_float(w & 0xFFFF0000u); }
__device__ __forceinline__ void widen(const RawV<__nv_bfloat16, 8>& r, float (&v)[kMaxV]) {
  v[0] = bf16lo(r.q.x); v[1] = bf16hi(r.q.x); v[2] = bf16lo(r.q.y); v[3] = bf16hi(r.q.y);
  v[4] = bf16lo(r.q.z); v[5] = bf16hi(r.q.z); v[6] = bf16lo(r.q.w); v[7] = bf16hi(r.q.w);
}
__device__ __forceinline__ void widen(const RawV<__nv_bfloat16, 4>& r, float (&v)[kMaxV]) {
  v[0] = bf16lo(r.q.x); v[1] = bf16hi(r.q.x); v[2] = bf16lo(r.q.y); v[3] = bf16hi(r.q.y);
}
__device__ __forceinline__ void widen(const RawV<__nv_bfloat16, 1>& r, float (&v)[kMaxV]) { v[0] = __uint_as_float((uint32_t)r.q << 16); }
__device__ __forceinline__ void widen(const RawV<float, 8>& r, float (&v)[kMaxV]) {
  v[0] = r.q.x; v[1] = r.q.y; v[2] = r.q.z; v[3] = r.q.w; v[4] = r.q1.x; v[5] = r.q1.y; v[6] = r.q1.z; v[7] = r.q1.w;
}
__device__ __forceinline__ void widen(const RawV<float, 4>& r, float (&v)[kMaxV]) { v[0] = r.q.x; v[1] = r.q.y; v[2] = r.q.z; v[3] = r.q.w; }
__device__ __forceinline__ void widen(const RawV<float, 1>& r, float (&v)[kMaxV]) { v[0] = r.q; }

// rows rb, rb + lanes, ... < r1 of one tensor; p points at this thread's V elements of row 0, rows are C elements apart;
// f(row, values) per row, in row order.
template <typename T, int V, int U, bool PREFETCH = false, typename F>
__device__ __forceinline__ void walk_rows(const T* __restrict__ p, int C, int rb, int r1, int lanes, F&& f) {
  const int nrows = rb < r1 ? (r1 - rb + lanes - 1) / lanes : 0;
  const int nfull = nrows / U;
  if (!PREFETCH) {                               // U loads, then their arithmetic (fewer live registers)
    for (int g = 0; g < nfull; g++, rb += U * lanes) {
      RawV<T, V> t[U];
#pragma unroll
      for (int u = 0; u < U; u++) ldraw(p + (long long)(rb + u * lanes) * C, t[u]);
#pragma unroll
      for (int u = 0; u < U; u++) {
        float a[kMaxV];
        widen(t[u], a);
        f(rb + u * lanes, a);
      }
    }
    for (; rb < r1; rb += lanes) {
      RawV<T, V> t;
      ldraw(p + (long long)rb * C, t);
      float a[kMaxV];
      widen(t, a);
      f(rb, a);
    }
    return;
  }
  RawV<T, V> cur[U], nxt[U];
  if (nfull > 0) {
#pragma unroll
    for (int u = 0; u < U; u++) ldraw(p + (long long)(rb + u * lanes) * C, cur[u]);
  }
  for (int g = 0; g < nfull; g++) {
    const bool more = g + 1 < nfull;
    if (more) {
#pragma unroll
      for (int u = 0; u < U; u++) ldraw(p + (long long)(rb + (U + u) * lanes) * C, nxt[u]);
    }
#pragma unroll
    for (int u = 0; u < U; u++) {
      float a[kMaxV];
      widen(cur[u], a);
      f(rb + u * lanes, a);
    }
    if (more) {
#pragma unroll
      for (int u = 0; u < U; u++) cur[u] = nxt[u];
    }
    rb += U * lanes;
  }
  for (; rb < r1; rb += lanes) {
    RawV<T, V> t;
    ldraw(p + (long long)rb * C, t);
    float a[kMaxV];
    widen(t, a);
    f(rb, a);
  }
}
// the same over two tensors read at the same offsets: f(row, values of p, values of q)
template <typename T, int V, int U, bool PREFETCH = false, typename F>
__device__ __forceinline__ void walk_rows2(const T* __restrict__ p, const T* __restrict__ q, int C, int rb, int r1, int lanes, F&& f) {
  const int nrows = rb < r1 ? (r1 - rb + lanes - 1) / lanes : 0;
  const int nfull = nrows / U;
  if (!PREFETCH) {
    for (int g = 0; g < nfull; g++, rb += U * lanes) {
      RawV<T, V> tp[U], tq[U];
#pragma unroll
      for (int u = 0; u < U; u++) {
        const long long o = (long long)(rb + u * lanes) * C;
        ldraw(p + o, tp[u]); ldraw(q + o, tq[u]);
      }
#pragma unroll
      for (int u = 0; u < U; u++) {
        float a[kMaxV], b[kMaxV];
        widen(tp[u], a); widen(tq[u], b);
        f(rb + u * lanes, a, b);
      }
    }
    for (; rb < r1; rb += lanes) {
      RawV<T, V> tp, tq;
      const long long o = (long long)rb * C;
      ldraw(p + o, tp); ldraw(q + o, tq);
      float a[kMaxV], b[kMaxV];
      widen(tp, a); widen(tq, b);
      f(rb, a, b);
    }
    return;
  }
  RawV<T, V> cp[U], cq[U], np[U], nq[U];
  if (nfull > 0) {
#pragma unroll
    for (int u = 0; u < U; u++) {
      const long long o = (long long)(rb + u * lanes) * C;
      ldraw(p + o, cp[u]); ldraw(q + o, cq[u]);
    }
  }
  for (int g = 0; g < nfull; g++) {
    const bool more = g + 1 < nfull;
    if (more) {
#pragma unroll
      for (int u = 0; u < U; u++) {
        const long long o = (long long)(rb + (U + u) * lanes) * C;
        ldraw(p + o, np[u]); ldraw(q + o, nq[u]);
      }
    }
#pragma unroll
    for (int u = 0; u < U; u++) {
      float a[kMaxV], b[kMaxV];
      widen(cp[u], a); widen(cq[u], b);
      f(rb + u * lanes, a, b);
    }
    if (more) {
#pragma unroll
      for (int u = 0; u < U; u++) { cp[u] = np[u]; cq[u] = nq[u]; }
    }
    rb += U * lanes;
  }
  for (; rb < r1; rb += lanes) {
    RawV<T, V> tp, tq;
    const long long o = (long long)rb * C;
    ldraw(p + o, tp); ldraw(q + o, tq);
    float a[kMaxV], b[kMaxV];
    widen(tp, a); widen(tq, b);
    f(rb, a, b);
  }
}

__device__ __forceinline__ uint32_t bf16x2_bits(float a, float b) {
  __nv_bfloat162 t = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&t);
}
template <typename T, int V> __device__ __forceinline__ void stv(T* p, const float (&v)[kMaxV]);
template <> __device__ __forceinline__ void stv<float, 1>(float* p, const float (&v)[kMaxV]) { *p = v[0]; }
template <> __device__ __forceinline__ void stv<float, 4>(float* p, const float (&v)[kMaxV]) {
  *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
}
template <> __device__ __forceinline__ void stv<float, 8>(float* p, const float (&v)[kMaxV]) {
  reinterpret_cast<float4*>(p)[0] = make_float4(v[0], v[1], v[2], v[3]);
  reinterpret_cast<float4*>(p)[1] = make_float4(v[4], v[5], v[6], v[7]);
}
template <> __device__ __forceinline__ void stv<__nv_bfloat16, 1>(__nv_bfloat16* p, const float (&v)[kMaxV]) {
  *p = __float2bfloat16_rn(v[0]);
}
template <> __device__ __forceinline__ void stv<__nv_bfloat16, 4>(__nv_bfloat16* p, const float (&v)[kMaxV]) {
  *reinterpret_cast<uint2*>(p) = make_uint2(bf16x2_bits(v[0], v[1]), bf16x2_bits(v[2], v[3]));
}
template <> __device__ __forceinline__ void stv<__nv_bfloat16, 8>(__nv_bfloat16* p, const float (&v)[kMaxV]) {
  *reinterpret_cast<uint4*>(p) = make_uint4(bf16x2_bits(v[0], v[1]), bf16x2_bits(v[2], v[3]), bf16x2_bits(v[4], v[5]), bf16x2_bits(v[6], v[7]));
}

__device__ __forceinline__ float miu_relu(float x) { return 0.5f * (x + sqrtf(0.09f + x * x)); }
__device__ __forceinline__ float miu_relu_grad(float x) { return 0.5f * (1.f + x * rsqrtf(0.09f + x * x)); }
__device__ __forceinline__ float sigmoidf_(float x) { return 1.f / (1.f + __expf(-x)); }
// single-MUFU forms for the streaming passes (relative error 2^-23 / 2^-22.9: far inside every tolerance of the path; the
// IEEE sqrtf / division sequences cost 8-12 issue slots per element, and those passes are issue-bound next to HBM)
__device__ __forceinline__ float sqrt_approx(float x) { float y; asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float rsqrt_approx(float x) { float y; asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float miu_relu_fast(float x) { return 0.5f * (x + sqrt_approx(fmaf(x, x, 0.09f))); }
__device__ __forceinline__ float miu_relu_grad_fast(float x) { return fmaf(0.5f * x, rsqrt_approx(fmaf(x, x, 0.09f)), 0.5f); }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// block-wide sum; result valid in thread 0.  `red` must hold >= 32 floats.
__device__ __forceinline__ float block_sum(float v, float* red) {
  v = warp_sum(v);
  int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  __syncthreads();
  if (lane == 0) red[wid] = v;
  __syncthreads();
  if (wid == 0) {
    int nw = (blockDim.x + 31) >> 5;
    v = lane < nw ? red[lane] : 0.f;
    v = warp_sum(v);
  }
  return v;
}

// dispatch on storage dtype
#define FGC_DISPATCH_DTYPE(dtype, T, ...)                               \
  do {                                                                  \
    if ((dtype) == FGC_F32) { using T = float; __VA_ARGS__; }           \
    else if ((dtype) == FGC_BF16) { using T = __nv_bfloat16; __VA_ARGS__; } \
    else { fgc::set_error("bad dtype %d", (int)(dtype)); return FGC_EINVAL; } \
  } while (0)

static inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }
static inline bool aligned8(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 7) == 0; }
// widest vector access (8, 4 or 1 elements) usable on a tensor whose channel count is C
static inline int vec_width(const void* p, long long C, int dtype) {
  if (C % 8 == 0 && aligned16(p) && (dtype == FGC_BF16 || (reinterpret_cast<uintptr_t>(p) & 31) == 0)) return 8;
  if (C % 4 == 0 && (dtype == FGC_BF16 ? aligned8(p) : aligned16(p))) return 4;
  return 1;
}
static inline int vmin(int a, int b) { return a < b ? a : b; }

}  // namespace fgc
