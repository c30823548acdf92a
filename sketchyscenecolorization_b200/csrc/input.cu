// Real-data input path: the raw records of the training / validation queues -> the normalised tensors the graph is fed.
// Replaces the per-sample TF ops of obj_lib/input_pipeline.get_paired_input (:72-126): decode_raw + cast, resize_images
// (image BILINEAR, sketch AREA; TF-1 legacy kernels: align_corners = False, no half-pixel centres), whole-image min-max
// normalisation, dequantisation noise, the [-1,1] map and the NHWC -> NCHW transpose -- one reduction pass and one apply
// pass over a whole batch.  HBM-bound byte work: per sample 2 x 442 KB of uint8 in, 2 x 442 KB of fp32 out at 192 x 192.
#include "common.cuh"

namespace fgc {

// splitmix64: the i-th output of the generator seeded with `seed` (tests restate it in numpy and match the noise exactly)
__device__ __forceinline__ uint32_t noise24(unsigned long long seed, unsigned long long i) {
  unsigned long long z = seed + (i + 1ull) * 0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  z ^= z >> 31;
  return (uint32_t)(z >> 40);                                  // 24 bits
}

// tf.image.resize_images(BILINEAR) with the TF-1 legacy kernel (src = dst * in / out, no half-pixel centres) at an INTEGER
// factor lands exactly on source pixel (oy * fy, ox * fx): the interpolation weights are zero and the resize is a pick.
// (The AREA resize of the sketch is only defined here for integer factors, so the call requires them.)
__device__ __forceinline__ float pick_px(const uint8_t* __restrict__ img, int R, int fy, int fx, int oy, int ox, int c) {
  return (float)__ldg(img + ((size_t)(oy * fy) * R + (size_t)ox * fx) * 3 + c);
}

// non-negative floats order like their bit patterns; the minimum is kept as the maximum of the complement so that one
// memset(0) initialises both slots
__device__ __forceinline__ uint32_t enc_min(float v) { return ~__float_as_uint(v); }
__device__ __forceinline__ float dec_min(uint32_t u) { return __uint_as_float(~u); }

// pass 1: scratch[2n] = ~bits(min), scratch[2n+1] = bits(max) over the RESIZED picture of sample n (tf.reduce_min / max
// of the whole [H,W,3] tensor, :109).  grid (blocks per sample, N)
__global__ void __launch_bounds__(256) paired_minmax_kernel(const uint8_t* __restrict__ cartoon, int R, int OH, int OW,
                                                            int fy, int fx, uint32_t* __restrict__ scratch) {
  const int n = blockIdx.y;
  const uint8_t* img = cartoon + (size_t)n * R * R * 3;
  float mn = 3.0e38f, mx = 0.f;
  for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < OH * OW; p += gridDim.x * blockDim.x) {
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      float v = pick_px(img, R, fy, fx, p / OW, p % OW, c);
      mn = fminf(mn, v);
      mx = fmaxf(mx, v);
    }
  }
#pragma unroll
  for (int o = 16; o; o >>= 1) {
    mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, o));
    mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  }
  __shared__ float smn[8], smx[8];
  if ((threadIdx.x & 31) == 0) { smn[threadIdx.x >> 5] = mn; smx[threadIdx.x >> 5] = mx; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < (int)(blockDim.x >> 5); ++w) { mn = fminf(mn, smn[w]); mx = fmaxf(mx, smx[w]); }
    if (mn <= mx) {                                            // a block with no pixel leaves the slots alone
      atomicMax(scratch + 2 * n, enc_min(mn));
      atomicMax(scratch + 2 * n + 1, __float_as_uint(mx));
    }
  }
}

// tf.image.resize_images(AREA) at an integer factor: the block mean
template <typename TS>
__device__ __forceinline__ float area_px(const TS* __restrict__ sk, int R, int fy, int fx, int oy, int ox, int c) {
  float acc = 0.f;
  for (int dy = 0; dy < fy; ++dy) {
    const TS* row = sk + ((size_t)(oy * fy + dy) * R + (size_t)ox * fx) * 3 + c;
    for (int dx = 0; dx < fx; ++dx) acc = __fadd_rn(acc, (float)__ldg(row + dx * 3));
  }
  return __fdiv_rn(acc, (float)(fy * fx));
}

// pass 2: one thread per output pixel, three channels; consecutive threads write consecutive floats of each NCHW plane
template <typename TS>
__global__ void __launch_bounds__(256) paired_apply_kernel(const uint8_t* __restrict__ cartoon, const TS* __restrict__ sketch,
                                                           int R, int OH, int OW, int fy, int fx,
                                                           const uint32_t* __restrict__ scratch, unsigned long long seed,
                                                           int dequantize, float* __restrict__ images,
                                                           float* __restrict__ sketches) {
  const int n = blockIdx.y;
  const uint8_t* img = cartoon + (size_t)n * R * R * 3;
  const TS* sk = sketch + (size_t)n * R * R * 3;
  const float mn = dec_min(__ldg(scratch + 2 * n)), mx = __uint_as_float(__ldg(scratch + 2 * n + 1));
  const float den = __fadd_rn(__fsub_rn(mx, mn), 1.f);         // max - min + 1 (:109)
  const size_t plane = (size_t)OH * OW;
  for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < OH * OW; p += gridDim.x * blockDim.x) {
    const int oy = p / OW, ox = p % OW;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const size_t o = ((size_t)n * 3 + c) * plane + p;
      float q = __fdiv_rn(__fsub_rn(pick_px(img, R, fy, fx, oy, ox, c), mn), den);
      if (dequantize) q = __fadd_rn(q, (float)noise24(seed, o) * (1.f / 4294967296.f));   // U[0, 1/256), :110
      images[o] = __fmaf_rn(q, 2.f, -1.f);
      float s = __fdiv_rn(area_px<TS>(sk, R, fy, fx, oy, ox, c), 255.f);         // :111
      sketches[o] = __fmaf_rn(s, 2.f, -1.f);
    }
  }
}

// ---- the production shape, 384 -> 192 with uint8 sketches: factor 2, rows 8-byte aligned -----------------------------------
// One thread per FOUR output pixels: 24 contiguous source bytes per row as three 8-byte loads (one cartoon row, two sketch
// rows), one 16-byte store per channel plane and tensor.  Every value a pixel can take is tabulated once per CTA in shared
// memory -- the normalised picture value of each of the 256 byte values, the final sketch value of each of the 1021 possible
// 2x2 sums -- with exactly the roundings of the generic kernel, so the per-element IEEE divisions (which kept the first
// version issue-bound, ncu r1s: 78% issue slots, 3.8 TB/s) become table reads.
__global__ void __launch_bounds__(256) paired_minmax2_kernel(const uint8_t* __restrict__ cartoon, int R, int OH, int OW,
                                                             uint32_t* __restrict__ scratch) {
  const int n = blockIdx.y;
  const uint8_t* img = cartoon + (size_t)n * R * R * 3;
  const int QW = OW / 4;
  unsigned mn = 255u, mx = 0u;
  for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < OH * QW; t += gridDim.x * blockDim.x) {
    const int oy = t / QW, ox = (t % QW) * 4;
    union { uint2 v[3]; uint8_t b[24]; } a;
    const uint2* pa = reinterpret_cast<const uint2*>(img + ((size_t)(2 * oy) * R + 2 * ox) * 3);
#pragma unroll
    for (int i = 0; i < 3; ++i) a.v[i] = __ldg(pa + i);
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        unsigned v = a.b[6 * j + c];
        mn = min(mn, v);
        mx = max(mx, v);
      }
  }
#pragma unroll
  for (int o = 16; o; o >>= 1) {
    mn = min(mn, __shfl_xor_sync(0xffffffffu, mn, o));
    mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  }
  __shared__ unsigned smn[8], smx[8];
  if ((threadIdx.x & 31) == 0) { smn[threadIdx.x >> 5] = mn; smx[threadIdx.x >> 5] = mx; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < (int)(blockDim.x >> 5); ++w) { mn = min(mn, smn[w]); mx = max(mx, smx[w]); }
    if (mn <= mx) {
      atomicMax(scratch + 2 * n, enc_min((float)mn));
      atomicMax(scratch + 2 * n + 1, __float_as_uint((float)mx));
    }
  }
}

__global__ void __launch_bounds__(256) paired_apply2_kernel(const uint8_t* __restrict__ cartoon, const uint8_t* __restrict__ sketch,
                                                            int R, int OH, int OW, const uint32_t* __restrict__ scratch,
                                                            unsigned long long seed, int dequantize, float* __restrict__ images,
                                                            float* __restrict__ sketches) {
  __shared__ float t_img[256];      // (v - min) / (max - min + 1), before the noise
  __shared__ float t_sk[1024];      // sum of a 2x2 block (0..1020) -> mean / 255 * 2 - 1
  const int n = blockIdx.y;
  const uint8_t* img = cartoon + (size_t)n * R * R * 3;
  const uint8_t* sk = sketch + (size_t)n * R * R * 3;
  const float mn = dec_min(__ldg(scratch + 2 * n)), mx = __uint_as_float(__ldg(scratch + 2 * n + 1));
  const float den = __fadd_rn(__fsub_rn(mx, mn), 1.f);
  for (int v = threadIdx.x; v < 256; v += blockDim.x) t_img[v] = __fdiv_rn(__fsub_rn((float)v, mn), den);
  for (int v = threadIdx.x; v < 1024; v += blockDim.x)
    t_sk[v] = __fmaf_rn(__fdiv_rn(__fmul_rn((float)v, 0.25f), 255.f), 2.f, -1.f);
  __syncthreads();
  const size_t plane = (size_t)OH * OW;
  const int QW = OW / 4;
  for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < OH * QW; t += gridDim.x * blockDim.x) {
    const int oy = t / QW, ox = (t % QW) * 4;
    union { uint2 v[3]; uint8_t b[24]; } a, s0, s1;
    const uint2* pa = reinterpret_cast<const uint2*>(img + ((size_t)(2 * oy) * R + 2 * ox) * 3);
    const uint2* p0 = reinterpret_cast<const uint2*>(sk + ((size_t)(2 * oy) * R + 2 * ox) * 3);
    const uint2* p1 = reinterpret_cast<const uint2*>(sk + ((size_t)(2 * oy + 1) * R + 2 * ox) * 3);
#pragma unroll
    for (int i = 0; i < 3; ++i) { a.v[i] = __ldg(pa + i); s0.v[i] = __ldg(p0 + i); s1.v[i] = __ldg(p1 + i); }
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const size_t o = ((size_t)n * 3 + c) * plane + (size_t)oy * OW + ox;
      float im[4], sq[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float q = t_img[a.b[6 * j + c]];
        if (dequantize) q = __fadd_rn(q, (float)noise24(seed, o + j) * (1.f / 4294967296.f));
        im[j] = __fmaf_rn(q, 2.f, -1.f);
        sq[j] = t_sk[(unsigned)s0.b[6 * j + c] + s0.b[6 * j + 3 + c] + s1.b[6 * j + c] + s1.b[6 * j + 3 + c]];
      }
      *reinterpret_cast<float4*>(images + o) = make_float4(im[0], im[1], im[2], im[3]);
      *reinterpret_cast<float4*>(sketches + o) = make_float4(sq[0], sq[1], sq[2], sq[3]);
    }
  }
}

}  // namespace fgc

using namespace fgc;

extern "C" int fgc_paired_input(const uint8_t* cartoon, const void* sketch, int sketch_dtype, int N, int R, int OH, int OW,
                                unsigned long long seed, int dequantize, float* images, float* sketches, uint32_t* scratch,
                                fgc_stream stream) {
  FGC_REQUIRE(N > 0 && R > 0 && OH > 0 && OW > 0 && OH <= R && OW <= R, "paired_input: bad sizes N=%d R=%d out=%dx%d", N, R, OH, OW);
  FGC_REQUIRE(sketch_dtype == 0 || sketch_dtype == 1, "paired_input: sketch_dtype must be 0 (uint8) or 1 (fp32)");
  if (R % OH || R % OW) {
    set_error("paired_input: AREA resize at a non-integer factor (%d -> %dx%d) is not implemented", R, OH, OW);
    return FGC_EUNSUPPORTED;
  }
  cudaStream_t s = as_stream(stream);
  if (cudaMemsetAsync(scratch, 0, sizeof(uint32_t) * 2 * (size_t)N, s) != cudaSuccess) return check_launch("paired_input memset");
  const int fy = R / OH, fx = R / OW;
  const int px = OH * OW;
  const bool fast = sketch_dtype == 0 && fy == 2 && fx == 2 && OW % 4 == 0 && R % 8 == 0 && (reinterpret_cast<uintptr_t>(cartoon) & 7) == 0 &&
                    (reinterpret_cast<uintptr_t>(sketch) & 7) == 0 && (reinterpret_cast<uintptr_t>(images) & 15) == 0 &&
                    (reinterpret_cast<uintptr_t>(sketches) & 15) == 0;
  if (fast) {        // ~3 loop trips per thread amortise the tables; 12 x N CTAs cover the 148 SMs from N = 13 up
    dim3 grid(max(1, min(cdiv(px / 4, 256), N >= 32 ? 12 : 36)), N);
    paired_minmax2_kernel<<<grid, 256, 0, s>>>(cartoon, R, OH, OW, scratch);
  } else {
    dim3 grid1(max(1, min(cdiv(px, 256 * 4), 64)), N);
    paired_minmax_kernel<<<grid1, 256, 0, s>>>(cartoon, R, OH, OW, fy, fx, scratch);
  }
  count_launch();
  FGC_LAUNCH_CHECK("paired_minmax");
  if (fast) {
    dim3 grid(max(1, min(cdiv(px / 4, 256), N >= 32 ? 12 : 36)), N);
    paired_apply2_kernel<<<grid, 256, 0, s>>>(cartoon, (const uint8_t*)sketch, R, OH, OW, scratch, seed, dequantize, images, sketches);
  } else {
    dim3 grid(max(1, cdiv(px, 256)), N);
    if (sketch_dtype == 0)
      paired_apply_kernel<uint8_t><<<grid, 256, 0, s>>>(cartoon, (const uint8_t*)sketch, R, OH, OW, fy, fx, scratch, seed,
                                                        dequantize, images, sketches);
    else
      paired_apply_kernel<float><<<grid, 256, 0, s>>>(cartoon, (const float*)sketch, R, OH, OW, fy, fx, scratch, seed,
                                                      dequantize, images, sketches);
  }
  count_launch();
  FGC_LAUNCH_CHECK("paired_apply");
  return FGC_OK;
}
