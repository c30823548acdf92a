// CUDA-core direct convolutions for the narrow layers of the MRU networks (sm_100a).
//
// The stem-level layers -- 7x7 3 -> 8 stems (models_collection.py:93,710), the unit-1 gates 11 -> 8 and image
// branches 3 -> 8 (mru.py:408-424) and the input gradients that end in 8 or 3 channels -- carry < 0.5% of the FLOPs
// but run at 192x192: on the tensor path they pay a 64-wide K slab and a 16-wide N tile for 3..11 real channels and
// a scalar gather per element.  They are HBM-bound streaming layers, so they get direct kernels:
//   conv_small_fwd_kernel   : one thread per output pixel, all (<= 8) output channels in registers; the input tile
//                             (with halo) and the weights sit in shared memory as fp32.  Serves forward and (with
//                             mirrored taps / transposed weight addressing) input-gradient calls.
//   conv_small_wgrad_kernel : one thread per row of dW[k*k*Cin, Cout<=8], persistent CTAs walk 16x16 pixel tiles and
//                             keep the partial dW in registers; one red.add per entry and CTA at the end.
// Eligibility: every source "small" (C < 64 or C % 8 != 0, conv_geom.cuh), <= 8 output channels.
#include "conv_geom.cuh"

namespace fgc {
int num_sms();
extern long long g_conv_counts[6];

constexpr int kTS = 16;        // output tile edge

struct SmallArgs {
  ConvGeom g;
  const float* w;            // value(tap, c, n) = w[tap*tap_stride + c*k_stride + n*n_stride + base]
  long long tap_stride, k_stride, n_stride, base;
  int nout;
  const float* bias;
  int act, accumulate;
  void* y;
  int y_dtype;
  int ctot;
  int tile_h, tile_w;        // input tile extent: (kTS-1)*stride + k
  int tiles_x, tiles_y;
  int dbg;                   // timing experiments (env FGC_SF_DBG): 1 skip the products, 2 skip the tile staging, 4 skip the weight staging
};

__device__ __forceinline__ float small_act(float v, int act) {
  switch (act) {
    case FGC_ACT_LRELU: return v > 0.f ? v : 0.2f * v;
    case FGC_ACT_TANH: return tanhf(v);
    case FGC_ACT_MIU: return miu_relu(v);
    case FGC_ACT_RELU: return fmaxf(v, 0.f);
    default: return v;
  }
}

// non-coherent scalar load widened to fp32 (the staging loops below issue several before the first dependent shared-memory store)
template <typename T> __device__ __forceinline__ float ldnc(const T* p);
template <> __device__ __forceinline__ float ldnc<float>(const float* p) { return __ldg(p); }
template <> __device__ __forceinline__ float ldnc<__nv_bfloat16>(const __nv_bfloat16* p) {
  return __uint_as_float((uint32_t)__ldg(reinterpret_cast<const unsigned short*>(p)) << 16);
}

// input tile of every source -> shared memory, channel-planar fp32 [ctot][tile_h][tile_w], zero outside the image
template <typename T>
__device__ __forceinline__ void load_x_tile(const ConvGeom& g, int n, int ih0, int iw0, int tile_h, int tile_w, float* xt) {
  const int plane = tile_h * tile_w;
  for (int s = 0; s < g.nsrc; s++) {
    const T* src = reinterpret_cast<const T*>(g.src[s]);
    const int C = g.C[s], ups = g.ups[s];
    const int Hs = ups ? (g.H >> 1) : g.H, Ws = ups ? (g.W >> 1) : g.W;
    float* dst = xt + g.cbase[s] * plane;
    if (!ups && C == 8 && sizeof(T) == 2 && (reinterpret_cast<uintptr_t>(src) & 15) == 0) {
      // 8 bf16 channels = one 16-byte load per pixel of the tile (the stem features next to the 3-channel picture in the unit-1
      // gates: 72 % of that tile's elements); two pixels per thread and step in flight
      const long long img = (long long)n * g.H;
      const int B = (int)blockDim.x;
      for (int pos = threadIdx.x; pos < plane; pos += 2 * B) {
        uint4 raw[2];
        int po[2];
#pragma unroll
        for (int u = 0; u < 2; u++) {
          raw[u] = make_uint4(0, 0, 0, 0);
          po[u] = pos + u * B;
          if (po[u] < plane) {
            const int r = po[u] / tile_w, col = po[u] - r * tile_w;
            const int ih = ih0 + r, iw = iw0 + col;
            if ((unsigned)ih < (unsigned)g.H && (unsigned)iw < (unsigned)g.W)
              raw[u] = __ldg(reinterpret_cast<const uint4*>(src + ((img + ih) * g.W + iw) * 8));
          }
        }
#pragma unroll
        for (int u = 0; u < 2; u++) {
          if (po[u] >= plane) continue;
          float* d = dst + po[u];
          d[0] = __uint_as_float(raw[u].x << 16); d[plane] = __uint_as_float(raw[u].x & 0xFFFF0000u);
          d[2 * plane] = __uint_as_float(raw[u].y << 16); d[3 * plane] = __uint_as_float(raw[u].y & 0xFFFF0000u);
          d[4 * plane] = __uint_as_float(raw[u].z << 16); d[5 * plane] = __uint_as_float(raw[u].z & 0xFFFF0000u);
          d[6 * plane] = __uint_as_float(raw[u].w << 16); d[7 * plane] = __uint_as_float(raw[u].w & 0xFFFF0000u);
        }
      }
      continue;
    }
    if (!ups) {
      // the flattened (row, column, channel) index of the tile advances by blockDim.x per step: its three parts are carried
      // along instead of divided out per element (the generic loop below costs ~120 instructions per element -- three integer
      // divisions and 64-bit indexing -- and was a quarter to a half of these kernels' time on the 3 / 8 / 11-channel layers)
      const int dq = (int)blockDim.x / C, dr = (int)blockDim.x - dq * C;
      const int pc = (int)threadIdx.x / C;
      int c = (int)threadIdx.x - pc * C, r = pc / tile_w, col = pc - r * tile_w;
      const long long img = (long long)n * g.H;
      const int total = plane * C, B = (int)blockDim.x;
      // eight elements per thread and step: the loads first, then the shared-memory stores -- one load in flight per thread made
      // the staging of a tile pure latency (7 dependent round trips per 16 x 16 tile of an 11-channel source: the weight-gradient
      // kernel took 700 us whatever the filter size)
      constexpr int SU = 8;
      for (int idx = threadIdx.x; idx < total; idx += SU * B) {
        float v[SU];
        int so[SU];
#pragma unroll
        for (int u = 0; u < SU; u++) {
          v[u] = 0.f;
          so[u] = -1;
          if (idx + u * B < total) {
            const int ih = ih0 + r, iw = iw0 + col;
            so[u] = c * plane + r * tile_w + col;
            if ((unsigned)ih < (unsigned)g.H && (unsigned)iw < (unsigned)g.W) v[u] = ldnc<T>(src + ((img + ih) * g.W + iw) * C + c);
            c += dr; col += dq;
            if (c >= C) { c -= C; col++; }
            while (col >= tile_w) { col -= tile_w; r++; }
          }
        }
#pragma unroll
        for (int u = 0; u < SU; u++)
          if (so[u] >= 0) dst[so[u]] = v[u];
      }
      continue;
    }
    for (int idx = threadIdx.x; idx < plane * C; idx += blockDim.x) {
      const int c = idx % C, pos = idx / C;
      const int r = pos / tile_w, col = pos - r * tile_w;
      int ih = ih0 + r, iw = iw0 + col;
      float v = 0.f;
      if ((unsigned)ih < (unsigned)g.H && (unsigned)iw < (unsigned)g.W) {
        if (ups) { ih >>= 1; iw >>= 1; }
        v = ld1<T>(src + (((long long)n * Hs + ih) * Ws + iw) * C + c);
      }
      dst[c * plane + pos] = v;
    }
  }
}

constexpr int kTW = 64;        // forward tile: 16 rows x 64 columns, 4 pixels (16 columns apart) per thread

template <typename T>
__global__ void __launch_bounds__(256) conv_small_fwd_kernel(const __grid_constant__ SmallArgs a) {
  extern __shared__ __align__(16) float sm_small[];
  const ConvGeom& g = a.g;
  const int k = g.k, ctot = a.ctot, KK = k * k * ctot;
  float* wsm = sm_small;                     // [KK][8]
  float* xt = sm_small + KK * 8;             // [ctot][tile_h][tile_w]
  const int tid = threadIdx.x;
  int bt = blockIdx.x;
  const int tx = bt % a.tiles_x; bt /= a.tiles_x;
  const int ty = bt % a.tiles_y;
  const int n = bt / a.tiles_y;
  for (int i = (a.dbg & 4) ? KK * 8 : tid; i < KK * 8; i += blockDim.x) {
    const int q = i >> 3, co = i & 7;
    const int tap = q / ctot, c = q - tap * ctot;
    wsm[i] = co < a.nout ? __ldg(a.w + tap * a.tap_stride + c * a.k_stride + co * a.n_stride + a.base) : 0.f;
  }
  // input row of output row oh under filter row kh:  oh*stride + sign*(kh - pad_t)
  const int org_h = g.sign > 0 ? -g.pad_t : g.pad_t - (k - 1);
  const int org_w = g.sign > 0 ? -g.pad_l : g.pad_l - (k - 1);
  if (!(a.dbg & 2)) load_x_tile<T>(g, n, ty * kTS * g.stride + org_h, tx * kTW * g.stride + org_w, a.tile_h, a.tile_w, xt);
  __syncthreads();
  // each weight vector read from shared memory feeds 4 pixels: the kernel is bound by shared-memory return bandwidth,
  // not by the FMA pipe, so the register blocking is what sets its speed
  const int ly = tid >> 4, lx = tid & 15;
  const int plane = a.tile_h * a.tile_w;
  float acc[4][8];
#pragma unroll
  for (int p = 0; p < 4; p++)
#pragma unroll
    for (int q = 0; q < 8; q++) acc[p][q] = 0.f;
  for (int kh = (a.dbg & 1) ? k : 0; kh < k; kh++) {
    const int khl = g.sign > 0 ? kh : k - 1 - kh;
    for (int kw = 0; kw < k; kw++) {
      const int kwl = g.sign > 0 ? kw : k - 1 - kw;
      const float* xp = xt + (ly * g.stride + khl) * a.tile_w + lx * g.stride + kwl;
      const float4* wp = reinterpret_cast<const float4*>(wsm + (size_t)(kh * k + kw) * ctot * 8);
#pragma unroll 2
      for (int c = 0; c < ctot; c++) {
        const float4 w0 = wp[2 * c], w1 = wp[2 * c + 1];
#pragma unroll
        for (int p = 0; p < 4; p++) {
          const float x = xp[c * plane + p * 16 * g.stride];
          acc[p][0] += x * w0.x; acc[p][1] += x * w0.y; acc[p][2] += x * w0.z; acc[p][3] += x * w0.w;
          acc[p][4] += x * w1.x; acc[p][5] += x * w1.y; acc[p][6] += x * w1.z; acc[p][7] += x * w1.w;
        }
      }
    }
  }
  const int oh = ty * kTS + ly;
  if (oh >= g.OH) return;
  float bias[8];
#pragma unroll
  for (int q = 0; q < 8; q++) bias[q] = (a.bias && q < a.nout) ? __ldg(a.bias + q) : 0.f;
#pragma unroll
  for (int p = 0; p < 4; p++) {
    const int ow = tx * kTW + lx + 16 * p;
    if (ow >= g.OW) continue;
    const long long m = ((long long)n * g.OH + oh) * g.OW + ow;
    float v[8];
#pragma unroll
    for (int q = 0; q < 8; q++) v[q] = small_act(acc[p][q] + bias[q], a.act);
    if (a.y_dtype == FGC_F32) {
      float* yp = reinterpret_cast<float*>(a.y) + m * a.nout;
#pragma unroll
      for (int q = 0; q < 8; q++)
        if (q < a.nout) yp[q] = a.accumulate ? yp[q] + v[q] : v[q];
    } else {
      __nv_bfloat16* yp = reinterpret_cast<__nv_bfloat16*>(a.y) + m * a.nout;
      if (a.nout == 8 && (reinterpret_cast<uintptr_t>(a.y) & 15) == 0) {
        if (a.accumulate) {
          uint4 pv = *reinterpret_cast<const uint4*>(yp);
          const __nv_bfloat162* pp = reinterpret_cast<const __nv_bfloat162*>(&pv);
#pragma unroll
          for (int e = 0; e < 4; e++) {
            float2 f = __bfloat1622float2(pp[e]);
            v[2 * e] += f.x; v[2 * e + 1] += f.y;
          }
        }
        *reinterpret_cast<uint4*>(yp) = make_uint4(bf16x2_bits(v[0], v[1]), bf16x2_bits(v[2], v[3]), bf16x2_bits(v[4], v[5]),
                                                   bf16x2_bits(v[6], v[7]));
      } else {
#pragma unroll
        for (int q = 0; q < 8; q++)
          if (q < a.nout) yp[q] = __float2bfloat16_rn(a.accumulate ? __bfloat162float(yp[q]) + v[q] : v[q]);
      }
    }
  }
}

struct SmallWgradArgs {
  ConvGeom g;
  const void* gy;            // [N, OH, OW, Cout]
  int Cout;
  int ctot;
  float* dw;                 // HWIO fp32 [k*k*ctot][Cout], accumulated
  int tile_h, tile_w;
  int tiles_x, tiles_y, ntiles;
  int KKP, S;                // threads per pixel slice (multiple of 32), pixel slices per CTA
  int dbg;                   // timing experiments (env FGC_SWG_DBG): 1 skip the final atomics, 2 skip the products, 4 / 8 skip the x / gy staging
};

// NKK rows of dW per thread (kk = lane-group index + j*KKP): every gy vector read from shared memory feeds NKK rows
template <typename T, int NKK>
__global__ void __launch_bounds__(512) conv_small_wgrad_kernel(const __grid_constant__ SmallWgradArgs a) {
  extern __shared__ __align__(16) float sm_small[];
  const ConvGeom& g = a.g;
  const int k = g.k, ctot = a.ctot, KK = k * k * ctot;
  const int plane = a.tile_h * a.tile_w;
  float* xt = sm_small;                          // [ctot][tile_h][tile_w]
  float* gyt = sm_small + ((ctot * plane + 3) & ~3);   // [256][8], 16-byte aligned rows
  const int tid = threadIdx.x;
  const int kt = tid % a.KKP, slice = tid / a.KKP;
  int xoff[NKK];
  bool act_[NKK];
#pragma unroll
  for (int j = 0; j < NKK; j++) {
    const int kk = kt + j * a.KKP;
    act_[j] = kk < KK;
    const int kc = act_[j] ? kk : 0;
    const int tap = kc / ctot, c = kc - tap * ctot;
    const int kh = tap / k, kw = tap - kh * k;
    xoff[j] = c * plane + kh * a.tile_w + kw;
  }
  float acc[NKK][8];
#pragma unroll
  for (int j = 0; j < NKK; j++)
#pragma unroll
    for (int q = 0; q < 8; q++) acc[j][q] = 0.f;
  const T* gy = reinterpret_cast<const T*>(a.gy);
  for (int t = blockIdx.x; t < a.ntiles; t += gridDim.x) {
    int bt = t;
    const int tx = bt % a.tiles_x; bt /= a.tiles_x;
    const int ty = bt % a.tiles_y;
    const int n = bt / a.tiles_y;
    __syncthreads();                             // previous tile fully consumed
    const bool gy_vec = sizeof(T) == 2 && a.Cout == 8 && (reinterpret_cast<uintptr_t>(gy) & 15) == 0 && blockDim.x >= 256;
    uint4 raw = make_uint4(0, 0, 0, 0);          // 8 bf16 outputs per pixel = one 16-byte load per thread, issued before the x tile's
    if (gy_vec && tid < 256 && !(a.dbg & 8)) {   // loads so that a tile costs ONE exposed memory round trip
      const int oh = ty * kTS + (tid >> 4), ow = tx * kTS + (tid & 15);
      if (oh < g.OH && ow < g.OW) raw = __ldg(reinterpret_cast<const uint4*>(gy + (((long long)n * g.OH + oh) * g.OW + ow) * 8));
    }
    if (!(a.dbg & 4)) load_x_tile<T>(g, n, ty * kTS * g.stride - g.pad_t, tx * kTS * g.stride - g.pad_l, a.tile_h, a.tile_w, xt);
    if (a.dbg & 8) {
    } else if (gy_vec) {
      if (tid < 256) {
        const int p = tid;
        float4 lo4, hi4;
        lo4.x = __uint_as_float(raw.x << 16); lo4.y = __uint_as_float(raw.x & 0xFFFF0000u);
        lo4.z = __uint_as_float(raw.y << 16); lo4.w = __uint_as_float(raw.y & 0xFFFF0000u);
        hi4.x = __uint_as_float(raw.z << 16); hi4.y = __uint_as_float(raw.z & 0xFFFF0000u);
        hi4.z = __uint_as_float(raw.w << 16); hi4.w = __uint_as_float(raw.w & 0xFFFF0000u);
        *reinterpret_cast<float4*>(gyt + p * 8) = lo4;
        *reinterpret_cast<float4*>(gyt + p * 8 + 4) = hi4;
      }
    } else {
#pragma unroll 4
      for (int i = tid; i < 256 * 8; i += blockDim.x) {
        const int p = i >> 3, co = i & 7;
        const int oh = ty * kTS + (p >> 4), ow = tx * kTS + (p & 15);
        float v = 0.f;
        if (co < a.Cout && oh < g.OH && ow < g.OW) v = ldnc<T>(gy + (((long long)n * g.OH + oh) * g.OW + ow) * a.Cout + co);
        gyt[i] = v;
      }
    }
    __syncthreads();
    // a slice takes half rows of the tile (8 consecutive pixels: 32 units over S slices), unrolled with constant offsets --
    // per pixel two 16-byte gy reads, one x read and the 8 FMAs, no index arithmetic (the per-pixel form spent 7 of its 40
    // instructions per pixel on it)
    const int st = g.stride;
    for (int u = (a.dbg & 2) ? 32 : slice; u < 32; u += a.S) {
      const int py = u >> 1, px0 = (u & 1) << 3;
      const float* gp = gyt + (py * 16 + px0) * 8;
      const int pbase = (py * a.tile_w + px0) * st;
      if (st == 1) {
#pragma unroll
        for (int q = 0; q < 8; q++) {
          const float4 g0 = *reinterpret_cast<const float4*>(gp + q * 8), g1 = *reinterpret_cast<const float4*>(gp + q * 8 + 4);
#pragma unroll
          for (int j = 0; j < NKK; j++) {
            const float x = xt[xoff[j] + pbase + q];
            acc[j][0] += x * g0.x; acc[j][1] += x * g0.y; acc[j][2] += x * g0.z; acc[j][3] += x * g0.w;
            acc[j][4] += x * g1.x; acc[j][5] += x * g1.y; acc[j][6] += x * g1.z; acc[j][7] += x * g1.w;
          }
        }
      } else {
#pragma unroll 2
        for (int q = 0; q < 8; q++) {
          const float4 g0 = *reinterpret_cast<const float4*>(gp + q * 8), g1 = *reinterpret_cast<const float4*>(gp + q * 8 + 4);
#pragma unroll
          for (int j = 0; j < NKK; j++) {
            const float x = xt[xoff[j] + pbase + q * st];
            acc[j][0] += x * g0.x; acc[j][1] += x * g0.y; acc[j][2] += x * g0.z; acc[j][3] += x * g0.w;
            acc[j][4] += x * g1.x; acc[j][5] += x * g1.y; acc[j][6] += x * g1.z; acc[j][7] += x * g1.w;
          }
        }
      }
    }
  }
  // the S pixel slices of the CTA hold partial sums of the same rows: combined in shared memory, then ONE atomic per element of dW
  // and CTA (a 3 x 3, 3-channel filter is 27 rows under 16 slices: 444 CTAs x 512 threads x 8 atomics on 216 addresses cost
  // 480 of the kernel's 690 us)
  __syncthreads();                               // the last tile's operands are dead: the buffer is reused
  float* part = sm_small;                        // [NKK][S][KKP][8]
#pragma unroll
  for (int j = 0; j < NKK; j++) {
    float* mine = part + ((size_t)(j * a.S + slice) * a.KKP + kt) * 8;
    *reinterpret_cast<float4*>(mine) = make_float4(acc[j][0], acc[j][1], acc[j][2], acc[j][3]);
    *reinterpret_cast<float4*>(mine + 4) = make_float4(acc[j][4], acc[j][5], acc[j][6], acc[j][7]);
  }
  __syncthreads();
  if (a.dbg & 1) return;
  for (int o = tid; o < NKK * a.KKP * 8; o += blockDim.x) {
    const int j = o / (a.KKP * 8), rem = o - j * a.KKP * 8;
    const int kk = (rem >> 3) + j * a.KKP, q = rem & 7;
    if (kk >= KK || q >= a.Cout) continue;
    float t = 0.f;
    for (int sl = 0; sl < a.S; sl++) t += part[((size_t)(j * a.S + sl) * a.KKP) * 8 + rem];
    atomicAdd(a.dw + (long long)kk * a.Cout + q, t);
  }
}

// Patch tensor of a narrow source: out[n,h,w,q] = x[n, h+kh-pad, w+kw-pad, c], q = (kh*k+kw)*C + c, zero beyond k*k*C and
// outside the image (stride 1, SAME).  One thread writes one 16-byte chunk (8 consecutive q).
template <typename T>
__global__ void im2col_small_kernel(const T* __restrict__ x, long long nchunks, int H, int W, int C, int ups, int k, int CP,
                                    int sign, __nv_bfloat16* __restrict__ out) {
  // per flattened index q: row / column shift and the element offset inside the source relative to the pixel's own
  // element 0 -- the same for every pixel, so it is tabulated once per block instead of divided out per element
  extern __shared__ int sh_tab[];                // [CP] offsets, [CP] packed (dh + 64) | (dw + 64) << 8 | valid << 16
  const int pad = (k - 1) / 2, KK = k * k * C, CPV = CP / 8;
  const int Hs = ups ? (H >> 1) : H, Ws = ups ? (W >> 1) : W;
  for (int q = threadIdx.x; q < CP; q += blockDim.x) {
    int tap = q / C, c = q - tap * C;
    int kh = tap / k, kw = tap - kh * k;
    int dh = sign * (kh - pad), dw = sign * (kw - pad);
    sh_tab[q] = (dh * Ws + dw) * C + c;          // used on the interior fast path of non-upsampled sources
    sh_tab[CP + q] = (dh + 64) | ((dw + 64) << 8) | ((q < KK ? 1 : 0) << 16) | (c << 20);
  }
  __syncthreads();
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < nchunks; i += (long long)gridDim.x * blockDim.x) {
    const int ch = (int)(i % CPV);
    long long p = i / CPV;
    const int w = (int)(p % W);
    p /= W;
    const int h = (int)(p % H);
    const long long n = p / H;
    const int q0 = ch * 8;
    uint4 o;
    if (!ups && C == 8 && q0 < KK) {
      // one chunk = the 8 channels of one tap: a single 16-byte load (the source rows are 16-byte aligned for C = 8)
      const int t = sh_tab[CP + q0];
      const int ih = h + (t & 0xFF) - 64, iw = w + ((t >> 8) & 0xFF) - 64;
      o = make_uint4(0, 0, 0, 0);
      if ((unsigned)ih < (unsigned)H && (unsigned)iw < (unsigned)W) {
        if (sizeof(T) == 2) {
          o = __ldg(reinterpret_cast<const uint4*>(x + ((n * H + ih) * W + iw) * 8));
        } else {
          const float* f = reinterpret_cast<const float*>(x) + ((n * H + ih) * W + iw) * 8;
          o = make_uint4(bf16x2_bits(f[0], f[1]), bf16x2_bits(f[2], f[3]), bf16x2_bits(f[4], f[5]), bf16x2_bits(f[6], f[7]));
        }
      }
    } else {
      float v[8];
      const bool interior = !ups && h >= pad && h < H - pad && w >= pad && w < W - pad;
      const T* px = x + ((n * Hs + (ups ? 0 : h)) * Ws + (ups ? 0 : w)) * C;
#pragma unroll
      for (int e = 0; e < 8; e++) {
        const int t = sh_tab[CP + q0 + e];
        float val = 0.f;
        if ((t >> 16) & 1) {
          if (interior) {
            val = ld1<T>(px + sh_tab[q0 + e]);
          } else {
            int ih = h + (t & 0xFF) - 64, iw = w + ((t >> 8) & 0xFF) - 64;
            if ((unsigned)ih < (unsigned)H && (unsigned)iw < (unsigned)W) {
              if (ups) { ih >>= 1; iw >>= 1; }
              val = ld1<T>(x + ((n * Hs + ih) * Ws + iw) * C + (t >> 20));
            }
          }
        }
        v[e] = val;
      }
      o = make_uint4(bf16x2_bits(v[0], v[1]), bf16x2_bits(v[2], v[3]), bf16x2_bits(v[4], v[5]), bf16x2_bits(v[6], v[7]));
    }
    *reinterpret_cast<uint4*>(out + i * 8) = o;
  }
}

// Row-tiled form of the patch tensor for sources that are not upsampled (every patch the training step asks for).  The
// per-element form above spends ~20 instructions per 2-byte element (table look-ups, bounds checks, 64-bit index splits) and
// runs at 1.1 TB/s on the 7x7 x 3-channel patches of the generator head's backward pass (717 MB written in 0.63 ms).  Here a
// block owns TH output rows of one image: the TH + k - 1 source rows it needs are staged once in shared memory as bf16 with
// the SAME padding materialised as zeros (left / right margins of >= pad*C elements, rows beyond the image zero), so a patch
// element is one unconditional 2-byte shared-memory read at  off(q) + hl*RS + w*C  -- off(q) depends on the thread's fixed
// chunk only and lives in registers.  Thread t owns chunk t % CPV of the pixels t / CPV, t / CPV + lanes, ...: consecutive
// threads write consecutive 16-byte chunks.  C = 8 (the stem features): a chunk is the 8 channels of one tap, one 16-byte read.
template <typename T>
__device__ __forceinline__ unsigned short bf16_bits_of(const T* p);
template <>
__device__ __forceinline__ unsigned short bf16_bits_of<__nv_bfloat16>(const __nv_bfloat16* p) {
  return __ldg(reinterpret_cast<const unsigned short*>(p));
}
template <>
__device__ __forceinline__ unsigned short bf16_bits_of<float>(const float* p) {
  return __bfloat16_as_ushort(__float2bfloat16_rn(__ldg(p)));
}

template <typename T, bool VEC8>
__global__ void __launch_bounds__(256) im2col_rows_kernel(const T* __restrict__ x, int H, int W, int C, int k, int CP, int sign,
                                                          int TH, int tiles_h, int LP, int RS, __nv_bfloat16* __restrict__ out) {
  extern __shared__ __align__(16) unsigned short sh_rows[];      // [TH + k - 1][RS]
  const int pad = (k - 1) / 2, KK = k * k * C, CPV = CP / 8, WC = W * C;
  const int n = blockIdx.x / tiles_h, h0 = (blockIdx.x - n * tiles_h) * TH;
  const int th = min(TH, H - h0), R = th + k - 1;
  if ((WC & 7) == 0 && (reinterpret_cast<uintptr_t>(x) & 15) == 0) {
    // rows of the source are 16-byte aligned and a multiple of 8 elements, like the margins: stage 8 elements per access
    const int RV = RS >> 3, LPV = LP >> 3, WCV = WC >> 3;
    for (int i = threadIdx.x; i < R * RV; i += blockDim.x) {
      const int r = i / RV, ev = i - r * RV;
      const int ih = h0 - pad + r, jv = ev - LPV;
      uint4 o = make_uint4(0, 0, 0, 0);
      if ((unsigned)ih < (unsigned)H && (unsigned)jv < (unsigned)WCV) {
        const T* src = x + ((long long)n * H + ih) * WC + jv * 8;
        if (sizeof(T) == 2) {
          o = __ldg(reinterpret_cast<const uint4*>(src));
        } else {
          const float4 f0 = __ldg(reinterpret_cast<const float4*>(src)), f1 = __ldg(reinterpret_cast<const float4*>(src) + 1);
          o = make_uint4(bf16x2_bits(f0.x, f0.y), bf16x2_bits(f0.z, f0.w), bf16x2_bits(f1.x, f1.y), bf16x2_bits(f1.z, f1.w));
        }
      }
      *reinterpret_cast<uint4*>(sh_rows + r * RS + ev * 8) = o;
    }
  } else {
    for (int r = 0; r < R; r++) {
      const int ih = h0 - pad + r;
      const bool row_in = (unsigned)ih < (unsigned)H;
      const T* src = x + ((long long)n * H + (row_in ? ih : 0)) * WC;
      unsigned short* dst = sh_rows + r * RS;
      for (int e = threadIdx.x; e < RS; e += blockDim.x) {
        const int j = e - LP;
        dst[e] = (row_in && (unsigned)j < (unsigned)WC) ? bf16_bits_of<T>(src + j) : (unsigned short)0;
      }
    }
  }
  __syncthreads();
  const int lanes = blockDim.x / CPV;
  const int ch = threadIdx.x % CPV, lane = threadIdx.x / CPV;
  if (lane >= lanes) return;
  int off[VEC8 ? 1 : 8];
  unsigned valid = 0;                              // bit e: flattened index ch*8 + e is a real (tap, channel), not the zero tail
#pragma unroll
  for (int e = 0; e < (VEC8 ? 1 : 8); e++) {
    const int q = ch * 8 + e;
    off[e] = 0;
    if (q < KK) {
      const int tap = q / C, c = q - tap * C;
      const int kh = tap / k, kw = tap - kh * k;
      off[e] = (pad + sign * (kh - pad)) * RS + LP + sign * (kw - pad) * C + c;
      valid |= 1u << e;
    }
  }
  int hl = lane / W, w = lane - hl * W;
  const int dh = lanes / W, dw = lanes - dh * W;
  __nv_bfloat16* obase = out + (((long long)n * H + h0) * W) * CP + ch * 8;
  while (hl < th) {
    const int pix = hl * RS + w * C;
    uint4 o;
    if (VEC8) {
      o = *reinterpret_cast<const uint4*>(sh_rows + pix + off[0]);
    } else {
      unsigned short v[8];
#pragma unroll
      for (int e = 0; e < 8; e++) v[e] = ((valid >> e) & 1u) ? sh_rows[pix + off[e]] : (unsigned short)0;
      o = make_uint4((uint32_t)v[0] | ((uint32_t)v[1] << 16), (uint32_t)v[2] | ((uint32_t)v[3] << 16),
                     (uint32_t)v[4] | ((uint32_t)v[5] << 16), (uint32_t)v[6] | ((uint32_t)v[7] << 16));
    }
    *reinterpret_cast<uint4*>(obase + ((long long)hl * W + w) * CP) = o;
    hl += dh;
    w += dw;
    if (w >= W) { w -= W; hl++; }
  }
}

// launches the row-tiled kernel when the layer fits it; false = not taken (the caller uses the per-element kernel)
template <typename T>
static bool im2col_rows_try(const T* x, int N, int H, int W, int C, int k, int CP, int sign, __nv_bfloat16* out, cudaStream_t s) {
  static int on = -1;
  if (on < 0) { const char* e = getenv("FGC_IM2COL_ROWS"); on = e ? atoi(e) : 1; }
  const int CPV = CP / 8, pad = (k - 1) / 2;
  if (!on || k < 3 || CPV > 256 || (long long)N * H > 0x7fffffffLL / 16) return false;
  const int LP = 8 * ((pad * C + 7) / 8);                       // left margin: >= pad*C zeros, interior 16-byte aligned
  const int RS = 8 * ((LP + W * C + pad * C + 7) / 8);
  int TH = 8;
  while (TH > 1 && (size_t)(TH + k - 1) * RS * 2 > 40 * 1024) TH >>= 1;
  if ((size_t)(TH + k - 1) * RS * 2 > 40 * 1024) return false;
  if (TH > H) TH = H;
  const int tiles_h = (H + TH - 1) / TH;
  const size_t smem = (size_t)(TH + k - 1) * RS * 2;
  const bool vec8 = C == 8 && (W * C) % 8 == 0;
  if (vec8)
    im2col_rows_kernel<T, true><<<N * tiles_h, 256, smem, s>>>(x, H, W, C, k, CP, sign, TH, tiles_h, LP, RS, out);
  else
    im2col_rows_kernel<T, false><<<N * tiles_h, 256, smem, s>>>(x, H, W, C, k, CP, sign, TH, tiles_h, LP, RS, out);
  return true;
}

int ew_grid(long long work, int threads);

// ------------------------------------------------------------------------------------------------------
// Skinny products: y[M <= 64 rows, Nout] = x[M, K] * W (+ bias, act) with fp32 accumulation on the CUDA cores.
// These are the per-time-step products of the caption encoder (word LSTM gates [N,2D]x[2D,4D], the row term of the
// multimodal LSTM, and their input gradients; models_collection.py:184-187,212-226): one 128-row tensor-core tile
// would be half empty and the 16 CTAs it yields spend their time on pipeline latency (75 us per call, 60+60 calls
// per iteration).  Here a CTA owns 32 output columns; its 8 warps split every 128-wide K chunk 8 ways, each thread
// keeps all 64 rows of one column in registers, x is staged in shared memory and read as warp-wide broadcasts.
// Weight addressing is the packer's: value(k, n) = w[k*k_stride + n*n_stride + base], so the forward product
// (n contiguous) and the input-gradient product (k contiguous, float4 loads) share the kernel.
// ------------------------------------------------------------------------------------------------------
struct RowsArgs {
  const void* src[kMaxSrc];
  int C[kMaxSrc], cbase[kMaxSrc];
  int nsrc;
  int M, K;
  const float* w;
  long long k_stride, n_stride, base;
  int nout;
  const float* bias;
  int act, accumulate;
  void* y;
  int y_dtype;
  int wvec;                  // k_stride == 1 and 16-byte aligned rows: float4 weight loads
};

template <typename T>
__global__ void __launch_bounds__(256) conv_rows_kernel(const __grid_constant__ RowsArgs a) {
  constexpr int KC = 128, NB = 32, NS = 8, MR = 64;
  __shared__ __align__(16) float xs[MR * KC];        // 32 KB: x chunk [64][128]; reused as the slice-reduction buffer
  const int tid = threadIdx.x, lane = tid & 31, slice = tid >> 5;
  const int n = blockIdx.x * NB + lane;
  const bool nvalid = n < a.nout;
  float acc[MR];
#pragma unroll
  for (int m = 0; m < MR; m++) acc[m] = 0.f;
  const float* wn = a.w + (long long)(nvalid ? n : 0) * a.n_stride + a.base;
  for (int k0 = 0; k0 < a.K; k0 += KC) {
    __syncthreads();
    for (int i = tid; i < MR * KC; i += 256) {
      const int m = i / KC, kk = i - m * KC;
      const int kg = k0 + kk;
      float v = 0.f;
      if (m < a.M && kg < a.K) {
        int s = 0;
        while (s + 1 < a.nsrc && kg >= a.cbase[s + 1]) s++;
        v = ld1<T>(reinterpret_cast<const T*>(a.src[s]) + (long long)m * a.C[s] + (kg - a.cbase[s]));
      }
      xs[i] = v;
    }
    __syncthreads();
#pragma unroll 1
    for (int kk = slice * (KC / NS); kk < (slice + 1) * (KC / NS); kk += 4) {
      const int kg = k0 + kk;
      float w0 = 0.f, w1 = 0.f, w2 = 0.f, w3 = 0.f;
      if (nvalid) {
        if (a.wvec && kg + 3 < a.K) {
          const float4 t = __ldg(reinterpret_cast<const float4*>(wn + kg));
          w0 = t.x; w1 = t.y; w2 = t.z; w3 = t.w;
        } else {
          if (kg < a.K) w0 = __ldg(wn + (long long)kg * a.k_stride);
          if (kg + 1 < a.K) w1 = __ldg(wn + (long long)(kg + 1) * a.k_stride);
          if (kg + 2 < a.K) w2 = __ldg(wn + (long long)(kg + 2) * a.k_stride);
          if (kg + 3 < a.K) w3 = __ldg(wn + (long long)(kg + 3) * a.k_stride);
        }
      }
#pragma unroll
      for (int m = 0; m < MR; m++) {
        const float4 x4 = *reinterpret_cast<const float4*>(xs + m * KC + kk);
        acc[m] += x4.x * w0 + x4.y * w1 + x4.z * w2 + x4.w * w3;
      }
    }
  }
  // combine the 8 K slices: two halves of 32 rows through shared memory [slice][row][column]
#pragma unroll
  for (int half = 0; half < 2; half++) {
    __syncthreads();
#pragma unroll
    for (int mm = 0; mm < 32; mm++) xs[(slice * 32 + mm) * NB + lane] = acc[half * 32 + mm];
    __syncthreads();
#pragma unroll
    for (int r = 0; r < 4; r++) {
      const int mm = slice * 4 + r, m = half * 32 + mm;
      float v = 0.f;
#pragma unroll
      for (int sl = 0; sl < NS; sl++) v += xs[(sl * 32 + mm) * NB + lane];
      if (m < a.M && nvalid) {
        if (a.bias) v += __ldg(a.bias + n);
        v = small_act(v, a.act);
        const long long o = (long long)m * a.nout + n;
        if (a.y_dtype == FGC_F32) {
          float* yp = reinterpret_cast<float*>(a.y);
          yp[o] = a.accumulate ? yp[o] + v : v;
        } else {
          __nv_bfloat16* yp = reinterpret_cast<__nv_bfloat16*>(a.y);
          yp[o] = __float2bfloat16_rn(a.accumulate ? __bfloat162float(yp[o]) + v : v);
        }
      }
    }
  }
}

static int rows_mode() {
  // Measured on B200 (profiles/r1d): 166 us per [64,1024]x[1024,2048] call against 60 us on the tensor-core path -- one
  // column per thread leaves the kernel bound by shared-memory return bandwidth (one LDS.128 per 4 FMAs).  Kept behind
  // FGC_ROWS=1 as the starting point for a register-tiled version; the default routing does not use it.
  static int mode = -1;
  if (mode < 0) { const char* e = getenv("FGC_ROWS"); mode = e ? atoi(e) : 0; }
  return mode;
}

// returns -1 when the product is not eligible, else the launch status
int conv_rows_try(const ConvGeom& g, int src_dtype, const float* w, long long tap_stride, long long k_stride, long long n_stride,
                  long long base, int nout, const float* bias, int act, int accumulate, void* y, int y_dtype, cudaStream_t s) {
  (void)tap_stride;
  if (!rows_mode() || g.k != 1 || g.stride != 1 || g.M > 64 || g.M < 1) return -1;
  RowsArgs a;
  a.nsrc = g.nsrc;
  for (int i = 0; i < g.nsrc; i++) {
    if (g.ups[i]) return -1;
    a.src[i] = g.src[i]; a.C[i] = g.C[i]; a.cbase[i] = g.cbase[i];
  }
  a.M = (int)g.M;
  a.K = g.cbase[g.nsrc - 1] + g.C[g.nsrc - 1];
  if (a.K < 64) return -1;
  a.w = w; a.k_stride = k_stride; a.n_stride = n_stride; a.base = base;
  a.nout = nout; a.bias = bias; a.act = act; a.accumulate = accumulate; a.y = y; a.y_dtype = y_dtype;
  a.wvec = (k_stride == 1 && (n_stride & 3) == 0 && (base & 3) == 0 && (reinterpret_cast<uintptr_t>(w) & 15) == 0) ? 1 : 0;
  const int grid = (nout + 31) / 32;
  if (src_dtype == FGC_F32) conv_rows_kernel<float><<<grid, 256, 0, s>>>(a);
  else conv_rows_kernel<__nv_bfloat16><<<grid, 256, 0, s>>>(a);
  g_conv_counts[2]++;
  count_launch();
  return check_launch("conv_rows");
}

static bool all_small(const ConvGeom& g) {
  for (int i = 0; i < g.nsrc; i++)
    if (g.big[i]) return false;
  return true;
}
int g_small_mode = -1;         // fgc_set_conv_flags / env FGC_SMALL; 0 sends the narrow layers through the tensor-core path
static int small_mode() {
  if (g_small_mode < 0) { const char* e = getenv("FGC_SMALL"); g_small_mode = e ? atoi(e) : 1; }
  return g_small_mode;
}

// returns -1 when the layer is not eligible, else the launch status
int conv_small_fwd_try(const ConvGeom& g, int src_dtype, const float* w, long long tap_stride, long long k_stride,
                       long long n_stride, long long base, int nout, const float* bias, int act, int accumulate, void* y,
                       int y_dtype, cudaStream_t s) {
  if (!small_mode() || !all_small(g) || nout > 8) return -1;
  if (g.sign < 0 && g.stride != 1) return -1;
  SmallArgs a;
  a.g = g;
  a.w = w; a.tap_stride = tap_stride; a.k_stride = k_stride; a.n_stride = n_stride; a.base = base;
  a.nout = nout; a.bias = bias; a.act = act; a.accumulate = accumulate; a.y = y; a.y_dtype = y_dtype;
  { static int dbg = -1; if (dbg < 0) { const char* e = getenv("FGC_SF_DBG"); dbg = e ? atoi(e) : 0; } a.dbg = dbg; }
  a.ctot = g.cbase[g.nsrc - 1] + g.C[g.nsrc - 1];
  a.tile_h = (kTS - 1) * g.stride + g.k;
  a.tile_w = (kTW - 1) * g.stride + g.k;
  a.tiles_x = (g.OW + kTW - 1) / kTW;
  a.tiles_y = (g.OH + kTS - 1) / kTS;
  size_t smem = sizeof(float) * ((size_t)g.k * g.k * a.ctot * 8 + (size_t)a.ctot * a.tile_h * a.tile_w);
  if (smem > 96 * 1024) return -1;
  long long blocks = (long long)g.N * a.tiles_x * a.tiles_y;
  if (blocks > 0x7FFFFFFFLL) return -1;
  if (src_dtype == FGC_F32) {
    static bool set = false;
    if (!set) { cudaFuncSetAttribute(conv_small_fwd_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024); set = true; }
    conv_small_fwd_kernel<float><<<(int)blocks, 256, smem, s>>>(a);
  } else {
    static bool set = false;
    if (!set) { cudaFuncSetAttribute(conv_small_fwd_kernel<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024); set = true; }
    conv_small_fwd_kernel<__nv_bfloat16><<<(int)blocks, 256, smem, s>>>(a);
  }
  g_conv_counts[2]++;
  count_launch();
  return check_launch("conv_small_fwd");
}

int conv_small_wgrad_try(const ConvGeom& g, int src_dtype, const void* gy, int Cin_total, int Cout, float* dw, cudaStream_t s) {
  if (!small_mode() || !all_small(g) || Cout > 8 || g.sign < 0) return -1;
  const int KK = g.k * g.k * Cin_total;
  if (KK > 512) return -1;
  SmallWgradArgs a;
  a.g = g;
  a.gy = gy; a.Cout = Cout; a.ctot = Cin_total; a.dw = dw;
  { static int dbg = -1; if (dbg < 0) { const char* e = getenv("FGC_SWG_DBG"); dbg = e ? atoi(e) : 0; } a.dbg = dbg; }
  a.tile_h = a.tile_w = (kTS - 1) * g.stride + g.k;
  a.tiles_x = (g.OW + kTS - 1) / kTS;
  a.tiles_y = (g.OH + kTS - 1) / kTS;
  long long ntiles = (long long)g.N * a.tiles_x * a.tiles_y;
  if (ntiles > 0x7FFFFFFFLL) return -1;
  a.ntiles = (int)ntiles;
  // rows of dW per thread.  Measured (profiles/r1i,r1j): register blocking over several rows (one gy read feeding up to 5
  // rows) is SLOWER here than one row per thread -- 7x7 stem 0.51 -> 0.81 ms -- because the block then holds 16 pixel
  // slices of 32 threads and every slice walks the same 256-pixel tile; one row per thread and 3 slices it stays.
  // With the half-row unrolled products (round 2) two rows per thread win on the filters of ~100 rows and more (7x7 stem 685 ->
  // 603 us, [8,3] -> 8 gate 728 -> 652 us; three rows lose again: profiles/r2at_small_conv_experiments.log); env FGC_SWG_NKK overrides.
  int nkk = KK >= 96 ? 2 : 1;
  { static int e_nkk = -1; if (e_nkk < 0) { const char* e = getenv("FGC_SWG_NKK"); e_nkk = e ? atoi(e) : 0; }
    if (e_nkk >= 1 && e_nkk <= 5 && KK >= 96) nkk = e_nkk; }
  a.KKP = ((((KK + nkk - 1) / nkk) + 31) / 32) * 32;   // threads per pixel slice
  a.S = 512 / a.KKP;
  if (a.S < 1) a.S = 1;
  if (a.S > 16) a.S = 16;
  if (a.KKP * nkk < KK || a.KKP > 512) return -1;
  const int threads = a.KKP * a.S;
  size_t smem = sizeof(float) * ((((size_t)Cin_total * a.tile_h * a.tile_w + 3) & ~(size_t)3) + 256 * 8);
  if (smem < sizeof(float) * 8 * (size_t)threads * nkk) smem = sizeof(float) * 8 * (size_t)threads * nkk;     // the final cross-slice sums
  if (smem > 96 * 1024) return -1;
  // persistent CTAs: exactly as many as are resident at once (a second wave would double the run time)
#define FGC_SW(T_, N_)                                                                                              \
  do {                                                                                                              \
    static bool set = false;                                                                                        \
    if (!set) { cudaFuncSetAttribute(conv_small_wgrad_kernel<T_, N_>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024); set = true; } \
    int per_sm = 1;                                                                                                 \
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, conv_small_wgrad_kernel<T_, N_>, threads, smem) != cudaSuccess || per_sm < 1) \
      per_sm = 1;                                                                                                   \
    int grid = num_sms() * per_sm;                                                                                  \
    if (grid > a.ntiles) grid = a.ntiles;                                                                           \
    conv_small_wgrad_kernel<T_, N_><<<grid, threads, smem, s>>>(a);                                                 \
  } while (0)
#define FGC_SWN(T_)                                    \
  switch (nkk) {                                       \
    case 1: FGC_SW(T_, 1); break;                      \
    case 2: FGC_SW(T_, 2); break;                      \
    case 3: FGC_SW(T_, 3); break;                      \
    case 4: FGC_SW(T_, 4); break;                      \
    default: FGC_SW(T_, 5); break;                     \
  }
  if (src_dtype == FGC_F32) { FGC_SWN(float) } else { FGC_SWN(__nv_bfloat16) }
#undef FGC_SWN
#undef FGC_SW
  g_conv_counts[3]++;
  count_launch();
  return check_launch("conv_small_wgrad");
}

}  // namespace fgc

extern "C" int fgc_im2col_small(const void* x, int dtype, int N, int H, int W, int C, int ups, int k, int mirror, void* out,
                                fgc_stream stream) {
  const int sign = mirror ? -1 : 1;
  using namespace fgc;
  FGC_REQUIRE(k % 2 == 1 && C > 0 && N > 0, "im2col_small: bad arguments");
  FGC_REQUIRE((reinterpret_cast<uintptr_t>(out) & 15) == 0, "im2col_small: out must be 16-byte aligned");
  if (ups) FGC_REQUIRE(H % 2 == 0 && W % 2 == 0, "im2col_small: upsampled source needs even H, W");
  const int CP = 8 * ((k * k * C + 7) / 8);
  const long long nchunks = (long long)N * H * W * (CP / 8);
  cudaStream_t s = as_stream(stream);
  if (dtype != FGC_F32 && dtype != FGC_BF16) { set_error("im2col_small: bad dtype %d", dtype); return FGC_EINVAL; }
  if (!ups && (dtype == FGC_F32 ? im2col_rows_try<float>((const float*)x, N, H, W, C, k, CP, sign, (__nv_bfloat16*)out, s)
                                : im2col_rows_try<__nv_bfloat16>((const __nv_bfloat16*)x, N, H, W, C, k, CP, sign,
                                                                 (__nv_bfloat16*)out, s))) {
    count_launch();
    return check_launch("im2col_small (rows)");
  }
  if (dtype == FGC_F32)
    im2col_small_kernel<float><<<ew_grid(nchunks, 256), 256, 2 * CP * sizeof(int), s>>>((const float*)x, nchunks, H, W, C, ups, k, CP, sign, (__nv_bfloat16*)out);
  else if (dtype == FGC_BF16)
    im2col_small_kernel<__nv_bfloat16><<<ew_grid(nchunks, 256), 256, 2 * CP * sizeof(int), s>>>((const __nv_bfloat16*)x, nchunks, H, W, C, ups, k, CP,
                                                                            sign, (__nv_bfloat16*)out);
  else { set_error("im2col_small: bad dtype %d", dtype); return FGC_EINVAL; }
  count_launch();
  return check_launch("im2col_small");
}
