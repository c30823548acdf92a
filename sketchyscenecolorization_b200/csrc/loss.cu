// Loss + gradient kernels and the fused Adam step.
//
// Reference: graph_single.get_losses (:317-581): softplus GAN (:401-402), ACGAN cross-entropy with the focal
// weight on the real branch (:343-352), smooth-L1 x100 (:552-555,575), l2 weight decay (:570-576);
// tf.train.AdamOptimizer(beta1=0, beta2=0.9) (:588) with lr*decay (:139-142).
// Every loss kernel adds its (weighted) value to lossbuf[slot] and to lossbuf[0] (the total).
#include "common.cuh"

namespace fgc {
int ew_grid(long long work, int threads);

__device__ __forceinline__ float softplusf(float x) { return fmaxf(x, 0.f) + log1pf(expf(-fabsf(x))); }

template <typename T>
__global__ void softplus_mean_kernel(const T* __restrict__ d, long long n, float sign, float* lossbuf, int slot, T* __restrict__ gd) {
  __shared__ float red[32];
  float acc = 0.f;
  float invn = 1.f / (float)n;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    float x = sign * ld1<T>(d + i);
    acc += softplusf(x);
    float sg = 1.f / (1.f + expf(-x));
    st1<T>(gd + i, sign * sg * invn);
  }
  float t = block_sum(acc, red);
  if (threadIdx.x == 0) { atomicAdd(&lossbuf[slot], t * invn); atomicAdd(&lossbuf[0], t * invn); }
}

template <typename T>
__global__ void ce_loss_kernel(const T* __restrict__ logits, const int32_t* __restrict__ labels, int N, int C, int focal,
                               float weight, float* lossbuf, int slot, T* __restrict__ glogits) {
  int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  const T* lg = logits + (long long)n * C;
  int lab = labels[n];
  float m = -INFINITY;
  for (int c = 0; c < C; c++) m = fmaxf(m, ld1<T>(lg + c));
  float se = 0.f;
  for (int c = 0; c < C; c++) se += expf(ld1<T>(lg + c) - m);
  float lse = m + logf(se);
  float ce = lse - ld1<T>(lg + lab);
  float pt = expf(-ce);
  float loss = focal ? (1.f - pt) * (1.f - pt) * ce : ce;
  float wn = weight / (float)N;
  for (int c = 0; c < C; c++) {
    float p = expf(ld1<T>(lg + c) - lse);
    float oh = (c == lab) ? 1.f : 0.f;
    float g;
    if (focal) g = (-2.f * (1.f - pt) * ce) * (pt * (oh - p)) + (1.f - pt) * (1.f - pt) * (p - oh);
    else g = p - oh;
    st1<T>(glogits + (long long)n * C + c, g * wn);
  }
  atomicAdd(&lossbuf[slot], loss * wn);
  atomicAdd(&lossbuf[0], loss * wn);
}

template <typename T>
__global__ void smooth_l1_kernel(const T* __restrict__ target, const T* __restrict__ gen, long long n, float weight,
                                 float* lossbuf, int slot, T* __restrict__ ggen) {
  __shared__ float red[32];
  float acc = 0.f;
  float wn = weight / (float)n;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    float d = ld1<T>(target + i) - ld1<T>(gen + i);
    float ab = fabsf(d);
    acc += ab < 1.f ? 0.5f * ab * ab : ab - 0.5f;
    st1<T>(ggen + i, -fminf(fmaxf(d, -1.f), 1.f) * wn);
  }
  float t = block_sum(acc, red);
  if (threadIdx.x == 0) { atomicAdd(&lossbuf[slot], t * wn); atomicAdd(&lossbuf[0], t * wn); }
}

__global__ void reg_loss_kernel(const float* __restrict__ flat, const long long* __restrict__ start, const int32_t* __restrict__ len,
                                const float* __restrict__ reg, float* lossbuf, int slot) {
  __shared__ float red[32];
  int ch = blockIdx.x;
  float r = reg[ch];
  if (r == 0.f) return;
  const float* p = flat + start[ch];
  int L = len[ch];
  float acc = 0.f;
  for (int i = threadIdx.x; i < L; i += blockDim.x) acc += p[i] * p[i];
  float t = block_sum(acc, red);
  if (threadIdx.x == 0) { float v = 0.5f * r * t; atomicAdd(&lossbuf[slot], v); atomicAdd(&lossbuf[0], v); }
}

__global__ void adam_step_kernel(float* __restrict__ flat, float* __restrict__ grad, float* __restrict__ v,
                                 const long long* __restrict__ start, const int32_t* __restrict__ len,
                                 const float* __restrict__ reg, float lr_t, const float* __restrict__ lr_t_dev, float beta2,
                                 float eps, int add_reg) {
  if (lr_t_dev) lr_t = *lr_t_dev;       // step size kept in device memory (CUDA-graph replay: no baked-in scalars)
  int ch = blockIdx.x;
  long long s0 = start[ch];
  int L = len[ch];
  float r = add_reg ? reg[ch] : 0.f;
  for (int i = threadIdx.x; i < L; i += blockDim.x) {
    float w = flat[s0 + i];
    float g = grad[s0 + i] + r * w;
    float vv = beta2 * v[s0 + i] + (1.f - beta2) * g * g;
    grad[s0 + i] = g;
    v[s0 + i] = vv;
    flat[s0 + i] = w - lr_t * g / (sqrtf(vv) + eps);
  }
}
// the reference's other optimisers (graph_single.get_optimizer, :584-593), same chunk table as Adam.
// kind 1 RMSProp (decay 0.9, momentum 0, eps 1e-10), 2 Adadelta (rho 0.95, eps 1e-8), 3 Adagrad
__global__ void opt_step_kernel(float* __restrict__ flat, float* __restrict__ grad, float* __restrict__ s1, float* __restrict__ s2,
                                const long long* __restrict__ start, const int32_t* __restrict__ len,
                                const float* __restrict__ reg, int kind, float lr, const float* __restrict__ lr_dev, int add_reg) {
  if (lr_dev) lr = *lr_dev;
  int ch = blockIdx.x;
  long long s0 = start[ch];
  int L = len[ch];
  float r = add_reg ? reg[ch] : 0.f;
  for (int i = threadIdx.x; i < L; i += blockDim.x) {
    float w = flat[s0 + i];
    float g = grad[s0 + i] + r * w;
    grad[s0 + i] = g;
    if (kind == 1) {
      float ms = 0.9f * s1[s0 + i] + 0.1f * g * g;
      s1[s0 + i] = ms;
      flat[s0 + i] = w - lr * g / sqrtf(ms + 1e-10f);
    } else if (kind == 2) {
      float acc = 0.95f * s1[s0 + i] + 0.05f * g * g;
      float au = s2[s0 + i];
      float upd = sqrtf(au + 1e-8f) * rsqrtf(acc + 1e-8f) * g;
      s1[s0 + i] = acc;
      s2[s0 + i] = 0.95f * au + 0.05f * upd * upd;
      flat[s0 + i] = w - lr * upd;
    } else {
      float acc = s1[s0 + i] + g * g;
      s1[s0 + i] = acc;
      flat[s0 + i] = w - lr * g / sqrtf(acc);
    }
  }
}
}  // namespace fgc

using namespace fgc;

extern "C" {

int fgc_softplus_mean(const void* d, int dtype, long long n, float sign, float* lossbuf, int slot, void* gd, fgc_stream stream) {
  cudaStream_t s = as_stream(stream);
  int grid = ew_grid(n, 256);
  if (grid > 64) grid = 64;
  FGC_DISPATCH_DTYPE(dtype, T, (softplus_mean_kernel<T><<<grid, 256, 0, s>>>((const T*)d, n, sign, lossbuf, slot, (T*)gd)));
  count_launch();
  FGC_LAUNCH_CHECK("softplus_mean");
  return FGC_OK;
}
int fgc_ce_loss(const void* logits, int dtype, const int32_t* labels, int N, int C, int focal, float weight,
                float* lossbuf, int slot, void* glogits, fgc_stream stream) {
  cudaStream_t s = as_stream(stream);
  FGC_DISPATCH_DTYPE(dtype, T,
                     (ce_loss_kernel<T><<<cdiv(N, 64), 64, 0, s>>>((const T*)logits, labels, N, C, focal, weight, lossbuf, slot, (T*)glogits)));
  count_launch();
  FGC_LAUNCH_CHECK("ce_loss");
  return FGC_OK;
}
int fgc_smooth_l1(const void* target, const void* gen, int dtype, long long n, float weight, float* lossbuf, int slot,
                  void* ggen, fgc_stream stream) {
  cudaStream_t s = as_stream(stream);
  int grid = ew_grid(n, 256);
  if (grid > 592) grid = 592;
  FGC_DISPATCH_DTYPE(dtype, T,
                     (smooth_l1_kernel<T><<<grid, 256, 0, s>>>((const T*)target, (const T*)gen, n, weight, lossbuf, slot, (T*)ggen)));
  count_launch();
  FGC_LAUNCH_CHECK("smooth_l1");
  return FGC_OK;
}
int fgc_reg_loss(const float* flat, const long long* start, const int32_t* len, const float* reg, int nchunks,
                 float* lossbuf, int slot, fgc_stream stream) {
  reg_loss_kernel<<<nchunks, 256, 0, as_stream(stream)>>>(flat, start, len, reg, lossbuf, slot);
  count_launch();
  FGC_LAUNCH_CHECK("reg_loss");
  return FGC_OK;
}
int fgc_adam_step(float* flat, float* grad, float* v, const long long* start, const int32_t* len, const float* reg,
                  int nchunks, float lr_t, const float* lr_t_dev, float beta2, float eps, int add_reg, fgc_stream stream) {
  adam_step_kernel<<<nchunks, 256, 0, as_stream(stream)>>>(flat, grad, v, start, len, reg, lr_t, lr_t_dev, beta2, eps, add_reg);
  count_launch();
  FGC_LAUNCH_CHECK("adam_step");
  return FGC_OK;
}
int fgc_opt_step(float* flat, float* grad, float* s1, float* s2, const long long* start, const int32_t* len, const float* reg,
                 int nchunks, int kind, float lr, const float* lr_dev, int add_reg, fgc_stream stream) {
  FGC_REQUIRE(kind >= 1 && kind <= 3 && (kind != 2 || s2 != nullptr), "opt_step: kind %d (1 RMSProp, 2 Adadelta with two slots, 3 Adagrad)", kind);
  opt_step_kernel<<<nchunks, 256, 0, as_stream(stream)>>>(flat, grad, s1, s2, start, len, reg, kind, lr, lr_dev, add_reg);
  count_launch();
  FGC_LAUNCH_CHECK("opt_step");
  return FGC_OK;
}

}  // extern "C"
