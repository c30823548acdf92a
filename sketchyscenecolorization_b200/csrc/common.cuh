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
