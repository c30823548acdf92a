// Caption-encoder kernels: l2-normalise, embedding gather/scatter, the fused BasicLSTMCell gate stage
// (with the <pad> skip of tf.cond), group sums, the atanh-like output transform.  All fp32.
//
// Reference: models_collection.encode_feat_with_text (:150-248); BasicLSTMCell semantics (gate order
// i, j, f, o; forget_bias 1.0; state [c,h]) as used at :184-187,213,226.
#include "common.cuh"

namespace fgc {
int ew_grid(long long work, int threads);

// one warp per row
__global__ void l2norm_rows_fwd_kernel(const float* __restrict__ x, int R, int D, float* __restrict__ y, float* __restrict__ inv) {
  int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  int lane = threadIdx.x & 31;
  if (row >= R) return;
  const float* xr = x + (long long)row * D;
  float s = 0.f;
  for (int c = lane; c < D; c += 32) { float v = xr[c]; s += v * v; }
  s = warp_sum(s);
  float iv = rsqrtf(fmaxf(s, 1e-12f));
  // one Newton step: rsqrtf is approximate (2 ulp); the reference is an exact-ish tf.rsqrt
  iv = iv * (1.5f - 0.5f * fmaxf(s, 1e-12f) * iv * iv);
  for (int c = lane; c < D; c += 32) y[(long long)row * D + c] = xr[c] * iv;
  if (lane == 0) inv[row] = iv;
}
__global__ void l2norm_rows_bwd_kernel(const float* __restrict__ gy, const float* __restrict__ y, const float* __restrict__ inv,
                                       int R, int D, float* __restrict__ gx) {
  int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  int lane = threadIdx.x & 31;
  if (row >= R) return;
  const float* g = gy + (long long)row * D;
  const float* yr = y + (long long)row * D;
  float s = 0.f;
  for (int c = lane; c < D; c += 32) s += g[c] * yr[c];
  s = warp_sum(s);
  float iv = inv[row];
  for (int c = lane; c < D; c += 32) gx[(long long)row * D + c] = (g[c] - yr[c] * s) * iv;
}

__global__ void embedding_fwd_kernel(const float* __restrict__ table, const int32_t* __restrict__ ids, int N, int T, int t, int D,
                                     float* __restrict__ out) {
  long long total = (long long)N * D;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    int n = (int)(i / D), c = (int)(i % D);
    out[i] = table[(long long)ids[n * T + t] * D + c];
  }
}
__global__ void embedding_bwd_kernel(const float* __restrict__ g, const int32_t* __restrict__ ids, int N, int T, int t, int D,
                                     float* dtable) {
  long long total = (long long)N * D;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    int n = (int)(i / D), c = (int)(i % D);
    atomicAdd(&dtable[(long long)ids[n * T + t] * D + c], g[i]);
  }
}

__device__ __forceinline__ float sigmoid_acc(float x) { return 1.f / (1.f + expf(-x)); }

__global__ void lstm_cell_fwd_kernel(const float* __restrict__ gates, const float* __restrict__ gates2,
                                     const float* __restrict__ grow, const float* __restrict__ c_prev,
                                     const float* __restrict__ h_prev, const int32_t* __restrict__ ids, int T, int t, int N, int P,
                                     int D, float* __restrict__ c, float* __restrict__ h, float* __restrict__ pre) {
  long long total = (long long)N * P * D;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    long long r = idx / D;
    int d = (int)(idx % D);
    int n = (int)(r / P);
    float p[4];
#pragma unroll
    for (int q = 0; q < 4; q++) {
      long long o = r * 4 * D + (long long)q * D + d;
      float v = gates[o];
      if (gates2) v += gates2[o];
      if (grow) v += grow[(long long)n * 4 * D + (long long)q * D + d];
      p[q] = v;
      pre[o] = v;
    }
    float cp = c_prev[idx], hp = h_prev[idx];
    if (ids[n * T + t] != 0) {
      float cn = cp * sigmoid_acc(p[2] + 1.0f) + sigmoid_acc(p[0]) * tanhf(p[1]);
      c[idx] = cn;
      h[idx] = tanhf(cn) * sigmoid_acc(p[3]);
    } else {
      c[idx] = cp;
      h[idx] = hp;
    }
  }
}
__global__ void lstm_cell_bwd_kernel(const float* __restrict__ gc, const float* __restrict__ gh, const float* __restrict__ pre,
                                     const float* __restrict__ c_prev, const int32_t* __restrict__ ids, int T, int t, int N, int P,
                                     int D, float* __restrict__ g_pre, float* __restrict__ g_c_prev, float* __restrict__ g_h_pass) {
  long long total = (long long)N * P * D;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    long long r = idx / D;
    int d = (int)(idx % D);
    int n = (int)(r / P);
    long long o = r * 4 * D + d;
    float gcv = gc[idx], ghv = gh[idx];
    if (ids[n * T + t] != 0) {
      float si = sigmoid_acc(pre[o]), tj = tanhf(pre[o + D]), sf = sigmoid_acc(pre[o + 2LL * D] + 1.0f),
            so = sigmoid_acc(pre[o + 3LL * D]);
      float cp = c_prev[idx];
      float cn = cp * sf + si * tj;
      float tc = tanhf(cn);
      float g_o = ghv * tc * so * (1.f - so);
      float g_cn = gcv + ghv * so * (1.f - tc * tc);
      g_pre[o] = g_cn * tj * si * (1.f - si);
      g_pre[o + D] = g_cn * si * (1.f - tj * tj);
      g_pre[o + 2LL * D] = g_cn * cp * sf * (1.f - sf);
      g_pre[o + 3LL * D] = g_o;
      g_c_prev[idx] = g_cn * sf;
      g_h_pass[idx] = 0.f;
    } else {
      g_pre[o] = 0.f; g_pre[o + D] = 0.f; g_pre[o + 2LL * D] = 0.f; g_pre[o + 3LL * D] = 0.f;
      g_c_prev[idx] = gcv;
      g_h_pass[idx] = ghv;
    }
  }
}

__global__ void rows_group_sum_kernel(const float* __restrict__ x, int N, int P, int C, float* __restrict__ out) {
  long long total = (long long)N * C;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    int n = (int)(i / C), c = (int)(i % C);
    float s = 0.f;
    for (int p = 0; p < P; p++) s += x[((long long)n * P + p) * C + c];
    out[i] = s;
  }
}

__global__ void atanh_relu_fwd_kernel(const float* __restrict__ h, long long n, float* __restrict__ y) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    float v = h[i];
    float t = 0.5f * (logf(1.001f + v) - logf(1.001f - v));
    y[i] = t > 0.f ? t : 0.f;
  }
}
__global__ void atanh_relu_bwd_kernel(const float* __restrict__ gy, const float* __restrict__ h, long long n, float* __restrict__ gx) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    float v = h[i];
    float t = 0.5f * (logf(1.001f + v) - logf(1.001f - v));
    gx[i] = t > 0.f ? gy[i] * 0.5f * (1.f / (1.001f + v) + 1.f / (1.001f - v)) : 0.f;
  }
}
}  // namespace fgc

using namespace fgc;

extern "C" {

int fgc_l2norm_rows_fwd(const float* x, int R, int D, float* y, float* inv, fgc_stream stream) {
  l2norm_rows_fwd_kernel<<<cdiv(R, 8), 256, 0, as_stream(stream)>>>(x, R, D, y, inv);
  count_launch();
  FGC_LAUNCH_CHECK("l2norm_rows_fwd");
  return FGC_OK;
}
int fgc_l2norm_rows_bwd(const float* gy, const float* y, const float* inv, int R, int D, float* gx, fgc_stream stream) {
  l2norm_rows_bwd_kernel<<<cdiv(R, 8), 256, 0, as_stream(stream)>>>(gy, y, inv, R, D, gx);
  count_launch();
  FGC_LAUNCH_CHECK("l2norm_rows_bwd");
  return FGC_OK;
}
int fgc_embedding_fwd(const float* table, const int32_t* ids, int N, int T, int t, int D, float* out, fgc_stream stream) {
  embedding_fwd_kernel<<<ew_grid((long long)N * D, 256), 256, 0, as_stream(stream)>>>(table, ids, N, T, t, D, out);
  count_launch();
  FGC_LAUNCH_CHECK("embedding_fwd");
  return FGC_OK;
}
int fgc_embedding_bwd(const float* g, const int32_t* ids, int N, int T, int t, int D, float* dtable, fgc_stream stream) {
  embedding_bwd_kernel<<<ew_grid((long long)N * D, 256), 256, 0, as_stream(stream)>>>(g, ids, N, T, t, D, dtable);
  count_launch();
  FGC_LAUNCH_CHECK("embedding_bwd");
  return FGC_OK;
}
int fgc_lstm_cell_fwd(const float* gates, const float* gates2, const float* grow, const float* c_prev,
                      const float* h_prev, const int32_t* ids, int T, int t, int N, int P, int D,
                      float* c, float* h, float* pre, fgc_stream stream) {
  lstm_cell_fwd_kernel<<<ew_grid((long long)N * P * D, 256), 256, 0, as_stream(stream)>>>(gates, gates2, grow, c_prev, h_prev, ids,
                                                                                         T, t, N, P, D, c, h, pre);
  count_launch();
  FGC_LAUNCH_CHECK("lstm_cell_fwd");
  return FGC_OK;
}
int fgc_lstm_cell_bwd(const float* gc, const float* gh, const float* pre, const float* c_prev,
                      const int32_t* ids, int T, int t, int N, int P, int D,
                      float* g_pre, float* g_c_prev, float* g_h_pass, fgc_stream stream) {
  lstm_cell_bwd_kernel<<<ew_grid((long long)N * P * D, 256), 256, 0, as_stream(stream)>>>(gc, gh, pre, c_prev, ids, T, t, N, P, D,
                                                                                         g_pre, g_c_prev, g_h_pass);
  count_launch();
  FGC_LAUNCH_CHECK("lstm_cell_bwd");
  return FGC_OK;
}
int fgc_rows_group_sum(const float* x, int N, int P, int C, float* out, fgc_stream stream) {
  rows_group_sum_kernel<<<ew_grid((long long)N * C, 256), 256, 0, as_stream(stream)>>>(x, N, P, C, out);
  count_launch();
  FGC_LAUNCH_CHECK("rows_group_sum");
  return FGC_OK;
}
int fgc_atanh_relu_fwd(const float* h, long long n, float* y, fgc_stream stream) {
  atanh_relu_fwd_kernel<<<ew_grid(n, 256), 256, 0, as_stream(stream)>>>(h, n, y);
  count_launch();
  FGC_LAUNCH_CHECK("atanh_relu_fwd");
  return FGC_OK;
}
int fgc_atanh_relu_bwd(const float* gy, const float* h, long long n, float* gx, fgc_stream stream) {
  atanh_relu_bwd_kernel<<<ew_grid(n, 256), 256, 0, as_stream(stream)>>>(gy, h, n, gx);
  count_launch();
  FGC_LAUNCH_CHECK("atanh_relu_bwd");
  return FGC_OK;
}

}  // extern "C"
