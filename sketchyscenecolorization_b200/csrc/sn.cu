// Spectral normalisation (sn.spectral_normed_weight, sn.py:12-52): one power iteration
//   a = u W^T, v = a/(|a|+eps), b = v W, u' = b/(|b|+eps), sigma = b u'^T, W_bar = W/sigma
// and its backward, differentiating through sigma AND through the power iteration (the reference has no
// stop_gradient).  W is [K,C] row-major (HWIO reshaped), fp32.  work = a[K] | b[C] | u_new[C] | scal[8].
#include "common.cuh"

namespace fgc {
int ew_grid(long long work, int threads);

#define SN_EPS 1e-12f

__global__ void sn_a_kernel(const float* __restrict__ w, const float* __restrict__ u, int K, int C, float* __restrict__ a,
                            float* scal) {
  int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  int lane = threadIdx.x & 31;
  if (row >= K) return;
  const float* wr = w + (long long)row * C;
  float s = 0.f;
  for (int c = lane; c < C; c += 32) s += wr[c] * u[c];
  s = warp_sum(s);
  if (lane == 0) { a[row] = s; atomicAdd(&scal[0], s * s); }
}
// b[c] += sum_{k in chunk} v[k] W[k,c]
__global__ void sn_b_kernel(const float* __restrict__ w, const float* __restrict__ a, int K, int C, int rows_per_block,
                            const float* __restrict__ scal, float* b) {
  float inv = 1.f / (sqrtf(scal[0]) + SN_EPS);
  int k0 = blockIdx.x * rows_per_block, k1 = min(k0 + rows_per_block, K);
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float s = 0.f;
    for (int k = k0; k < k1; k++) s += a[k] * inv * w[(long long)k * C + c];
    atomicAdd(&b[c], s);
  }
}
__global__ void sn_fin_kernel(const float* __restrict__ b, int C, float* __restrict__ u_new, float* scal) {
  __shared__ float red[32];
  __shared__ float nb_s;
  float s = 0.f;
  for (int c = threadIdx.x; c < C; c += blockDim.x) s += b[c] * b[c];
  s = block_sum(s, red);
  if (threadIdx.x == 0) {
    float nb = sqrtf(s);
    nb_s = nb;
    scal[1] = sqrtf(scal[0]);
    scal[2] = nb;
    scal[3] = s / (nb + SN_EPS);     // sigma = b . u_new
  }
  __syncthreads();
  float inv = 1.f / (nb_s + SN_EPS);
  for (int c = threadIdx.x; c < C; c += blockDim.x) u_new[c] = b[c] * inv;
}
__global__ void sn_scale_kernel(const float* __restrict__ w, long long n, const float* __restrict__ scal, float* __restrict__ wbar) {
  float sigma = scal[3];
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    wbar[i] = w[i] / sigma;
}

__global__ void sn_dot_kernel(const float* __restrict__ x, const float* __restrict__ y, long long n, float* out) {
  __shared__ float red[32];
  float s = 0.f;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) s += x[i] * y[i];
  s = block_sum(s, red);
  if (threadIdx.x == 0) atomicAdd(out, s);
}
__device__ __forceinline__ float sn_gb_coef(const float* scal) {
  float nb = scal[2], sigma = scal[3];
  float gsig = -scal[4] / (sigma * sigma);
  return gsig * (nb + 2.f * SN_EPS) / ((nb + SN_EPS) * (nb + SN_EPS));
}
// gv[k] = sum_c gb[c] W[k,c];  scal[6] += gv[k]*a[k]
__global__ void sn_gv_kernel(const float* __restrict__ w, const float* __restrict__ b, const float* __restrict__ a, int K, int C,
                             float* scal, float* __restrict__ gv) {
  int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  int lane = threadIdx.x & 31;
  if (row >= K) return;
  float coef = sn_gb_coef(scal);
  const float* wr = w + (long long)row * C;
  float s = 0.f;
  for (int c = lane; c < C; c += 32) s += coef * b[c] * wr[c];
  s = warp_sum(s);
  if (lane == 0) { gv[row] = s; atomicAdd(&scal[6], s * a[row]); }
}
__global__ void sn_dw_kernel(const float* __restrict__ gwbar, const float* __restrict__ a, const float* __restrict__ b,
                             const float* __restrict__ u, const float* __restrict__ gv, int K, int C,
                             const float* __restrict__ scal, float* __restrict__ dw) {
  long long n = (long long)K * C;
  float na = scal[1], sigma = scal[3];
  float coef = sn_gb_coef(scal);
  float inv_na = 1.f / (na + SN_EPS);
  float t = scal[6] / ((na + SN_EPS) * (na + SN_EPS) * na);
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    int k = (int)(i / C), c = (int)(i % C);
    float v = a[k] * inv_na;
    float ga = gv[k] * inv_na - a[k] * t;
    dw[i] += gwbar[i] / sigma + v * (coef * b[c]) + ga * u[c];
  }
}
}  // namespace fgc

using namespace fgc;

extern "C" {

int fgc_sn_fwd(const float* w, const float* u, int K, int C, float* wbar, float* work, fgc_stream stream) {
  cudaStream_t s = as_stream(stream);
  float* a = work;
  float* b = work + K;
  float* u_new = b + C;
  float* scal = u_new + C;
  cudaMemsetAsync(b, 0, sizeof(float) * (2 * C + 8), s);
  sn_a_kernel<<<cdiv(K, 8), 256, 0, s>>>(w, u, K, C, a, scal);
  int rpb = 32;
  sn_b_kernel<<<cdiv(K, rpb), 256, 0, s>>>(w, a, K, C, rpb, scal, b);
  sn_fin_kernel<<<1, 256, 0, s>>>(b, C, u_new, scal);
  long long n = (long long)K * C;
  sn_scale_kernel<<<ew_grid(n, 256), 256, 0, s>>>(w, n, scal, wbar);
  count_launch(4);
  FGC_LAUNCH_CHECK("sn_fwd");
  return FGC_OK;
}

int fgc_sn_bwd(const float* gwbar, const float* w, const float* u, int K, int C, float* work, float* gv, float* dw,
               fgc_stream stream) {
  cudaStream_t s = as_stream(stream);
  float* a = work;
  float* b = work + K;
  float* scal = b + 2 * C;
  cudaMemsetAsync(scal + 4, 0, sizeof(float) * 4, s);
  long long n = (long long)K * C;
  sn_dot_kernel<<<ew_grid(n, 256), 256, 0, s>>>(gwbar, w, n, scal + 4);
  sn_gv_kernel<<<cdiv(K, 8), 256, 0, s>>>(w, b, a, K, C, scal, gv);
  sn_dw_kernel<<<ew_grid(n, 256), 256, 0, s>>>(gwbar, a, b, u, gv, K, C, scal, dw);
  count_launch(3);
  FGC_LAUNCH_CHECK("sn_bwd");
  return FGC_OK;
}

}  // extern "C"
