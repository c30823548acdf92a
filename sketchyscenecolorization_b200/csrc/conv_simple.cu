// Plain CUDA-core (fp32 FMA) direct convolution: forward, input gradient, weight gradient.
//
// This is the library's own checker for the tcgen05 kernels in conv_tc.cu (selected with
// FGC_CONV_IMPL=simple): same C-ABI, same geometry decode, no tensor cores, no staging -- one thread per
// output element.  It is only meant for small shapes; the tcgen05 path is the product path.
#include "conv_geom.cuh"

namespace fgc {
int ew_grid(long long work, int threads);
int num_sms();

__device__ __forceinline__ float apply_act(float v, int act) {
  switch (act) {
    case FGC_ACT_LRELU: return v > 0.f ? v : 0.2f * v;
    case FGC_ACT_TANH: return tanhf(v);
    case FGC_ACT_MIU: return miu_relu(v);
    case FGC_ACT_RELU: return fmaxf(v, 0.f);
    default: return v;
  }
}

template <typename TS, typename TO>
__global__ void conv_fwd_simple_kernel(ConvGeom g, const float* __restrict__ w, int Cin_total, int Cout,
                                       const float* __restrict__ bias, int act, int accumulate, TO* __restrict__ y) {
  long long total = g.M * Cout;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    int co = (int)(idx % Cout);
    long long m = idx / Cout;
    int ow = (int)(m % g.OW);
    int oh = (int)((m / g.OW) % g.OH);
    int n = (int)(m / ((long long)g.OW * g.OH));
    float acc = bias ? bias[co] : 0.f;
    for (int kh = 0; kh < g.k; kh++) {
      int ih = oh * g.stride + g.sign * (kh - g.pad_t);
      if (ih < 0 || ih >= g.H) continue;
      for (int kw = 0; kw < g.k; kw++) {
        int iw = ow * g.stride + g.sign * (kw - g.pad_l);
        if (iw < 0 || iw >= g.W) continue;
        const float* wt = w + (long long)(kh * g.k + kw) * Cin_total * Cout;
        for (int s = 0; s < g.nsrc; s++) {
          int Hs = g.ups[s] ? g.H / 2 : g.H, Ws = g.ups[s] ? g.W / 2 : g.W;
          int hh = g.ups[s] ? ih >> 1 : ih, ww = g.ups[s] ? iw >> 1 : iw;
          const TS* p = (const TS*)g.src[s] + (((long long)n * Hs + hh) * Ws + ww) * g.C[s];
          const float* wc = wt + (long long)g.cbase[s] * Cout + co;
          for (int c = 0; c < g.C[s]; c++) acc += ld1<TS>(p + c) * wc[(long long)c * Cout];
        }
      }
    }
    float v = apply_act(acc, act);
    if (accumulate) v += (float)y[idx];        // plain load: y is written by this kernel
    st1<TO>(y + idx, v);
  }
}

// gx[n,h,w,ci] (=|+=) sum_{kh,kw,co} gy[n, h+pad_t-kh, w+pad_l-kw, co] * w[kh,kw,c_off+ci,co]; ups: 2x2-summed
template <typename TG, typename TO>
__global__ void conv_dgrad_simple_kernel(const TG* __restrict__ gy, int N, int H, int W, const float* __restrict__ w, int k,
                                         int Cin_total, int Cout, int c_off, int c_len, int ups, int accumulate,
                                         TO* __restrict__ gx) {
  int oh_ = ups ? H / 2 : H, ow_ = ups ? W / 2 : W;
  long long total = (long long)N * oh_ * ow_ * c_len;
  int pad = (k - 1) / 2;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    int ci = (int)(idx % c_len);
    long long p = idx / c_len;
    int x = (int)(p % ow_);
    int y = (int)((p / ow_) % oh_);
    int n = (int)(p / ((long long)ow_ * oh_));
    float acc = 0.f;
    int reps = ups ? 2 : 1;
    for (int dy = 0; dy < reps; dy++)
      for (int dx = 0; dx < reps; dx++) {
        int h = ups ? 2 * y + dy : y, ww = ups ? 2 * x + dx : x;
        for (int kh = 0; kh < k; kh++) {
          int gh = h + pad - kh;
          if (gh < 0 || gh >= H) continue;
          for (int kw = 0; kw < k; kw++) {
            int gw = ww + pad - kw;
            if (gw < 0 || gw >= W) continue;
            const TG* gp = gy + (((long long)n * H + gh) * W + gw) * Cout;
            const float* wp = w + ((long long)(kh * k + kw) * Cin_total + c_off + ci) * Cout;
            for (int co = 0; co < Cout; co++) acc += ld1<TG>(gp + co) * wp[co];
          }
        }
      }
    if (accumulate) acc += ld1<TO>(gx + idx);
    st1<TO>(gx + idx, acc);
  }
}

// one block per (tap, concat channel); threads over co; serial over all output pixels
template <typename TS>
__global__ void conv_wgrad_simple_kernel(ConvGeom g, const TS* __restrict__ gy, int Cin_total, int Cout, float* dw) {
  int tap = blockIdx.x / Cin_total, cg = blockIdx.x % Cin_total;
  int kh = tap / g.k, kw = tap % g.k;
  int s = 0;
  while (s + 1 < g.nsrc && cg >= g.cbase[s + 1]) s++;
  int c = cg - g.cbase[s];
  int Hs = g.ups[s] ? g.H / 2 : g.H, Ws = g.ups[s] ? g.W / 2 : g.W;
  const TS* src = (const TS*)g.src[s];
  for (int co = threadIdx.x; co < Cout; co += blockDim.x) {
    float acc = 0.f;
    for (long long m = 0; m < g.M; m++) {
      int ow = (int)(m % g.OW);
      int oh = (int)((m / g.OW) % g.OH);
      int n = (int)(m / ((long long)g.OW * g.OH));
      int ih = oh * g.stride + kh - g.pad_t, iw = ow * g.stride + kw - g.pad_l;
      if (ih < 0 || ih >= g.H || iw < 0 || iw >= g.W) continue;
      int hh = g.ups[s] ? ih >> 1 : ih, ww = g.ups[s] ? iw >> 1 : iw;
      acc += ld1<TS>(src + (((long long)n * Hs + hh) * Ws + ww) * g.C[s] + c) * ld1<TS>(gy + m * Cout + co);
    }
    dw[((long long)tap * Cin_total + cg) * Cout + co] += acc;
  }
}

// db[c] += sum_m gy[m,c]
template <typename T>
__global__ void colsum_kernel(const T* __restrict__ x, long long M, int C, int rows_per_block, float* out) {
  long long r0 = (long long)blockIdx.x * rows_per_block, r1 = r0 + rows_per_block;
  if (r1 > M) r1 = M;
  if (C <= (int)blockDim.x) {
    // row lanes of a block are combined in shared memory first: one atomic per channel and block (a 3-channel tensor used to
    // send 255 atomics per block to 3 addresses -- 228 us for the bias gradient of a 3-channel layer at 192 x 192 x 128)
    __shared__ float sh_part[256];
    int lanes = blockDim.x / C, c = threadIdx.x % C, rl = threadIdx.x / C;
    float s = 0.f;
    if (rl < lanes)
      for (long long r = r0 + rl; r < r1; r += lanes) s += ld1<T>(x + r * C + c);
    sh_part[threadIdx.x] = s;
    __syncthreads();
    if ((int)threadIdx.x < C) {
      float t = 0.f;
      for (int l = 0; l < lanes; l++) t += sh_part[l * C + threadIdx.x];
      atomicAdd(&out[threadIdx.x], t);
    }
  } else {
    for (int c = blockIdx.y * blockDim.x + threadIdx.x; c < C; c += gridDim.y * blockDim.x) {     // grid.y: column chunks
      float s = 0.f;
      for (long long r = r0; r < r1; r++) s += ld1<T>(x + r * C + c);
      atomicAdd(&out[c], s);
    }
  }
}

// vectorised variant: a thread owns V consecutive channels and walks rows, four independent loads in flight; the block
// combines its row lanes in shared memory and issues one atomic per channel
template <typename T, int V>
__global__ void colsum_vec_kernel(const T* __restrict__ x, long long M, int C, int rows_per_block, float* out) {
  extern __shared__ float sh_col[];              // [C]
  const int CV = C / V;
  const int lanes = blockDim.x / CV;
  const int v = threadIdx.x % CV, rl = threadIdx.x / CV;
  for (int i = threadIdx.x; i < C; i += blockDim.x) sh_col[i] = 0.f;
  __syncthreads();
  long long r0 = (long long)blockIdx.x * rows_per_block, r1 = r0 + rows_per_block;
  if (r1 > M) r1 = M;
  if (rl < lanes) {
    float s[V];
#pragma unroll
    for (int k = 0; k < V; k++) s[k] = 0.f;
    walk_rows<T, V, 4>(x + r0 * C + v * V, C, rl, (int)(r1 - r0), lanes, [&](int, const float (&a)[kMaxV]) {
#pragma unroll
      for (int k = 0; k < V; k++) s[k] += a[k];
    });
    // few channel vectors per row (CV a power of two below 32): lanes CV apart hold the same channels -- combine them with
    // shuffles first.  C = 8 is ONE vector: all 256 threads of a block used to queue on the same 8 shared-memory words
    // (174 us for a 75 MB tensor, 0.4 TB/s)
    int CVp = CV;
    if (CV < 32 && (CV & (CV - 1)) == 0 && lanes * CV == (int)blockDim.x) {
#pragma unroll
      for (int k = 0; k < V; k++)
        for (int o = 16; o >= CV; o >>= 1) s[k] += __shfl_xor_sync(0xffffffffu, s[k], o);
      CVp = 32;                                  // one depositor per channel vector and warp
    }
    if (CVp == CV || (threadIdx.x & 31) < CV) {
#pragma unroll
      for (int k = 0; k < V; k++) atomicAdd(&sh_col[v * V + k], s[k]);
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < C; i += blockDim.x) atomicAdd(&out[i], sh_col[i]);
}

int colsum_launch(const void* x, int dtype, long long M, int C, float* out, cudaStream_t s) {
  int vec = vec_width(x, C, dtype);
  if (vec >= 4 && C / vec <= 256 && M >= 4096) {
    const int V = vec >= 8 ? 8 : 4;
    const int CV = C / V;
    const int lanes = 256 / CV;
    long long want = (long long)num_sms() * 16;
    long long rpb = (M + want - 1) / want;
    if (rpb < lanes * 8LL) rpb = lanes * 8LL;
    int nblk = (int)((M + rpb - 1) / rpb);
    size_t sm = sizeof(float) * C;
    if (dtype == FGC_F32) {
      if (V == 8) colsum_vec_kernel<float, 8><<<nblk, 256, sm, s>>>((const float*)x, M, C, (int)rpb, out);
      else colsum_vec_kernel<float, 4><<<nblk, 256, sm, s>>>((const float*)x, M, C, (int)rpb, out);
    } else {
      if (V == 8) colsum_vec_kernel<__nv_bfloat16, 8><<<nblk, 256, sm, s>>>((const __nv_bfloat16*)x, M, C, (int)rpb, out);
      else colsum_vec_kernel<__nv_bfloat16, 4><<<nblk, 256, sm, s>>>((const __nv_bfloat16*)x, M, C, (int)rpb, out);
    }
    count_launch();
    return FGC_OK;
  }
  long long want = (long long)num_sms() * 4;
  long long rpb = (M + want - 1) / want;
  if (rpb < 32) rpb = 32;
  int nblk = (int)((M + rpb - 1) / rpb);
  // wide rows (the [N, 4D] gate gradients of the caption encoder: 64 rows): the columns spread over grid.y, 2 blocks became 64
  int ny = C > 256 ? (C + 255) / 256 : 1;
  if (ny > 64) ny = 64;
  FGC_DISPATCH_DTYPE(dtype, T, (colsum_kernel<T><<<dim3(nblk, ny), 256, 0, s>>>((const T*)x, M, C, (int)rpb, out)));
  count_launch();
  return FGC_OK;
}

int conv_fwd_simple(const ConvGeom& g, int src_dtype, const float* w, int Cin_total, int Cout, const float* bias, int act,
                    int accumulate, void* y, int y_dtype, cudaStream_t s) {
  long long total = g.M * Cout;
  int grid = ew_grid(total, 128);
#define FGC_L(TS, TO) conv_fwd_simple_kernel<TS, TO><<<grid, 128, 0, s>>>(g, w, Cin_total, Cout, bias, act, accumulate, (TO*)y)
  if (src_dtype == FGC_F32 && y_dtype == FGC_F32) FGC_L(float, float);
  else if (src_dtype == FGC_F32 && y_dtype == FGC_BF16) FGC_L(float, __nv_bfloat16);
  else if (src_dtype == FGC_BF16 && y_dtype == FGC_F32) FGC_L(__nv_bfloat16, float);
  else if (src_dtype == FGC_BF16 && y_dtype == FGC_BF16) FGC_L(__nv_bfloat16, __nv_bfloat16);
  else { set_error("conv_fwd: bad dtypes"); return FGC_EINVAL; }
#undef FGC_L
  count_launch();
  return FGC_OK;
}

int conv_dgrad_simple(const void* gy, int gy_dtype, int N, int H, int W, const float* w, int k, int Cin_total, int Cout,
                      int c_off, int c_len, int ups, int accumulate, void* gx, int gx_dtype, cudaStream_t s) {
  long long total = (long long)N * (ups ? H / 2 : H) * (ups ? W / 2 : W) * c_len;
  int grid = ew_grid(total, 128);
#define FGC_L(TG, TO) \
  conv_dgrad_simple_kernel<TG, TO><<<grid, 128, 0, s>>>((const TG*)gy, N, H, W, w, k, Cin_total, Cout, c_off, c_len, ups, accumulate, (TO*)gx)
  if (gy_dtype == FGC_F32 && gx_dtype == FGC_F32) FGC_L(float, float);
  else if (gy_dtype == FGC_F32 && gx_dtype == FGC_BF16) FGC_L(float, __nv_bfloat16);
  else if (gy_dtype == FGC_BF16 && gx_dtype == FGC_F32) FGC_L(__nv_bfloat16, float);
  else if (gy_dtype == FGC_BF16 && gx_dtype == FGC_BF16) FGC_L(__nv_bfloat16, __nv_bfloat16);
  else { set_error("conv_dgrad: bad dtypes"); return FGC_EINVAL; }
#undef FGC_L
  count_launch();
  return FGC_OK;
}

int conv_wgrad_simple(const ConvGeom& g, int src_dtype, const void* gy, int Cin_total, int Cout, float* dw, cudaStream_t s) {
  int blocks = g.k * g.k * Cin_total;
  int threads = Cout < 32 ? 32 : (Cout > 256 ? 256 : ((Cout + 31) / 32) * 32);
  if (src_dtype == FGC_F32) conv_wgrad_simple_kernel<float><<<blocks, threads, 0, s>>>(g, (const float*)gy, Cin_total, Cout, dw);
  else conv_wgrad_simple_kernel<__nv_bfloat16><<<blocks, threads, 0, s>>>(g, (const __nv_bfloat16*)gy, Cin_total, Cout, dw);
  count_launch();
  return FGC_OK;
}

}  // namespace fgc
