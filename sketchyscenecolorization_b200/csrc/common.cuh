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

// 4 consecutive elements (pointer must be aligned to 4 elements)
template <typename T> __device__ __forceinline__ void ld4(const T* p, float (&v)[4]);
template <> __device__ __forceinline__ void ld4<float>(const float* p, float (&v)[4]) {
  float4 t = __ldg(reinterpret_cast<const float4*>(p));
  v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
}
template <> __device__ __forceinline__ void ld4<__nv_bfloat16>(const __nv_bfloat16* p, float (&v)[4]) {
  uint2 t = __ldg(reinterpret_cast<const uint2*>(p));
  __nv_bfloat162 a = *reinterpret_cast<__nv_bfloat162*>(&t.x);
  __nv_bfloat162 b = *reinterpret_cast<__nv_bfloat162*>(&t.y);
  float2 fa = __bfloat1622float2(a), fb = __bfloat1622float2(b);
  v[0] = fa.x; v[1] = fa.y; v[2] = fb.x; v[3] = fb.y;
}
template <typename T> __device__ __forceinline__ void st4(T* p, const float (&v)[4]);
template <> __device__ __forceinline__ void st4<float>(float* p, const float (&v)[4]) {
  *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
}
template <> __device__ __forceinline__ void st4<__nv_bfloat16>(__nv_bfloat16* p, const float (&v)[4]) {
  __nv_bfloat162 a = __floats2bfloat162_rn(v[0], v[1]);
  __nv_bfloat162 b = __floats2bfloat162_rn(v[2], v[3]);
  uint2 t;
  t.x = *reinterpret_cast<uint32_t*>(&a);
  t.y = *reinterpret_cast<uint32_t*>(&b);
  *reinterpret_cast<uint2*>(p) = t;
}

// V-wide access (V = 4 or 1)
template <typename T, int V> __device__ __forceinline__ void ldv(const T* p, float (&v)[4]) {
  if constexpr (V == 4) ld4<T>(p, v);
  else { v[0] = ld1<T>(p); v[1] = v[2] = v[3] = 0.f; }
}
template <typename T, int V> __device__ __forceinline__ void stv(T* p, const float (&v)[4]) {
  if constexpr (V == 4) st4<T>(p, v);
  else st1<T>(p, v[0]);
}

__device__ __forceinline__ float miu_relu(float x) { return 0.5f * (x + sqrtf(0.09f + x * x)); }
__device__ __forceinline__ float miu_relu_grad(float x) { return 0.5f * (1.f + x * rsqrtf(0.09f + x * x)); }
__device__ __forceinline__ float sigmoidf_(float x) { return 1.f / (1.f + __expf(-x)); }

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
// can we use 4-wide vector access on a tensor whose channel count is C?
static inline bool vec4_ok(const void* p, long long C, int dtype) {
  if (C % 4) return false;
  return dtype == FGC_BF16 ? aligned8(p) : aligned16(p);
}

}  // namespace fgc
