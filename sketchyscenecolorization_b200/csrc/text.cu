// Caption-encoder kernels: l2-normalise, embedding gather/scatter, the fused BasicLSTMCell gate stage
// (with the <pad> skip of tf.cond), group sums, the atanh-like output transform.  All fp32.
//
// Reference: models_collection.encode_feat_with_text (:150-248); BasicLSTMCell semantics (gate order
// i, j, f, o; forget_bias 1.0; state [c,h]) as used at :184-187,213,226.
#include "common.cuh"

namespace fgc {
int ew_grid(long long work, int threads);
int num_sms();

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
      if (pre) pre[o] = v;
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
// the same, four hidden units per thread (D % 4 == 0, 16-byte aligned buffers): the multimodal LSTM of the instance-matching
// model walks N*96*96 rows x 2000 gate columns per step -- 128-bit accesses, one row / sample decode per four units
__global__ void lstm_cell_fwd_vec4_kernel(const float4* __restrict__ gates, const float4* __restrict__ gates2,
                                          const float4* __restrict__ grow, const float4* __restrict__ c_prev,
                                          const float4* __restrict__ h_prev, const int32_t* __restrict__ ids, int T, int t, int N,
                                          int P, int D4, float4* __restrict__ c, float4* __restrict__ h, float4* __restrict__ pre) {
  const long long total = (long long)N * P * D4;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const long long r = idx / D4;
    const int d = (int)(idx - r * D4);
    const int n = (int)(r / P);
    float p[4][4];
#pragma unroll
    for (int q = 0; q < 4; q++) {
      const long long o = r * 4 * D4 + (long long)q * D4 + d;
      float4 v = __ldcs(gates + o);
      if (gates2) { float4 u = __ldg(gates2 + o); v.x += u.x; v.y += u.y; v.z += u.z; v.w += u.w; }
      if (grow) { float4 u = __ldg(grow + (long long)n * 4 * D4 + (long long)q * D4 + d); v.x += u.x; v.y += u.y; v.z += u.z; v.w += u.w; }
      p[q][0] = v.x; p[q][1] = v.y; p[q][2] = v.z; p[q][3] = v.w;
      if (pre) pre[o] = v;
    }
    const float4 cp = c_prev[idx], hp = h_prev[idx];
    if (ids[n * T + t] != 0) {
      const float cpv[4] = {cp.x, cp.y, cp.z, cp.w};
      float cn[4], hn[4];
#pragma unroll
      for (int k = 0; k < 4; k++) {
        cn[k] = cpv[k] * sigmoid_acc(p[2][k] + 1.0f) + sigmoid_acc(p[0][k]) * tanhf(p[1][k]);
        hn[k] = tanhf(cn[k]) * sigmoid_acc(p[3][k]);
      }
      c[idx] = make_float4(cn[0], cn[1], cn[2], cn[3]);
      h[idx] = make_float4(hn[0], hn[1], hn[2], hn[3]);
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


// ------------------------------------------------------------------------------------------------------
// The word LSTM's whole recurrence in ONE persistent launch  (models_collection.py:173-213; north_star: "fused warp-level
// GEMV + gate kernel").  The input half of the gate pre-activations, [e_t] @ K[0:D] + b for all T steps, does not depend on
// the recurrence and arrives precomputed (gx, one tensor-core product); what is sequential is h(t-1) @ K[D:2D], a
// [N, D] x [D, 4D] product per step that is pure latency as a launch of its own (62 us each through the conv path).  Here
// CTA b owns LS_UNITS = 4 hidden units (the i, j, f, o columns of those units: 16 of the 4D gate columns) and keeps its
// 16 x D slice of the recurrent weights in shared memory for the whole sequence; thread (n, u) accumulates the four gates of
// (sample n, unit u) in fp32 registers from the staged h(t-1) rows and applies the cell; the new h goes to global memory
// and a grid-wide barrier (one atomic counter; grid <= SM count, one CTA per SM, so all CTAs are co-resident) separates the
// steps.  All fp32 on CUDA cores: 2*N*D*4D flops per step (134 MFLOP at N 64, D 512) is ~2.5 us of the machine.
// The backward kernel runs BPTT the same way: phase A = cell backward for the CTA's units (gate gradients to global
// memory), barrier, phase B = g_h(t-1)[n, k] = sum_col g_pre[n, col] * K[D + k, col] for the CTA's four k (its rows of the
// recurrent weights in shared memory; the thread that produces g_h(n, k) is the one that consumes it in the next phase A,
// so it stays in a register and one barrier per step suffices).
// ------------------------------------------------------------------------------------------------------
constexpr int LS_UNITS = 4, LS_THREADS = 256, LS_ROWS = 64, LS_MAXCH = 4;
constexpr int LS_GC = 512;              // backward: columns of the gate gradients staged per pass

__device__ __forceinline__ unsigned int ld_acquire_u32(const unsigned int* p) {
  unsigned int v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
// 16-byte global -> shared copies that bypass registers (LDGSTS, L2 only): a thread keeps all its copies of a stage in flight
__device__ __forceinline__ void ls_cp_async16(void* smem_dst, const void* gsrc) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void ls_cp_async_wait_all() { asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void grid_barrier(unsigned int* bar, unsigned int target) {
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    atomicAdd(bar, 1u);
    while (ld_acquire_u32(bar) < target) __nanosleep(32);
    __threadfence();
  }
  __syncthreads();
}

__global__ void __launch_bounds__(LS_THREADS, 1)
lstm_seq_fwd_kernel(const float* __restrict__ gx, const float* __restrict__ kh, const int32_t* __restrict__ ids, int T, int N, int D,
                    float* h_all, float* c_all, float* __restrict__ pre_all, unsigned int* bar) {
  extern __shared__ __align__(16) float ls_smem[];
  float* Ws = ls_smem;                 // [D][unit][gate]
  float* Hs = ls_smem + (size_t)D * 16;   // [LS_ROWS][D + 4]
  const int HS = D + 4;
  const int tid = threadIdx.x, u = tid & 3, nl = tid >> 2;
  const int d0 = blockIdx.x * LS_UNITS, d = d0 + u;
  for (int i = tid; i < D * 16; i += LS_THREADS) {
    const int k = i >> 4, uu = (i >> 2) & 3, g = i & 3;
    Ws[i] = __ldg(kh + (size_t)k * 4 * D + (size_t)g * D + d0 + uu);
  }
  const int D4 = D >> 2;
  for (int t = 0; t < T; t++) {
    const float* hprev = h_all + (size_t)t * N * D;
    for (int n0 = 0; n0 < N; n0 += LS_ROWS) {
      const int nn = min(LS_ROWS, N - n0);
      __syncthreads();
      for (int i = tid; i < nn * D4; i += LS_THREADS) {      // all of a thread's copies in flight at once (one L2 round trip)
        const int r = i / D4, c4 = i - r * D4;
        ls_cp_async16(Hs + (size_t)r * HS + c4 * 4, hprev + (size_t)(n0 + r) * D + c4 * 4);
      }
      ls_cp_async_wait_all();
      __syncthreads();
      if (nl < nn) {
        const int n = n0 + nl;
        const size_t o = ((size_t)t * N + n) * 4 * D + d;
        // issued before the loop, consumed after it: the L2 latency of these four loads hides behind the dot products
        const float x0 = __ldg(gx + o), x1 = __ldg(gx + o + D), x2 = __ldg(gx + o + 2 * (size_t)D), x3 = __ldg(gx + o + 3 * (size_t)D);
        const float cp = __ldcg(c_all + ((size_t)t * N + n) * D + d);
        const int32_t tok = __ldg(ids + (size_t)n * T + t);
        float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
        const float* hr = Hs + (size_t)nl * HS;
        const float* wr = Ws + u * 4;
#pragma unroll 4
        for (int k = 0; k < D; k += 4) {
          const float4 hv = *reinterpret_cast<const float4*>(hr + k);
          const float4 w0 = *reinterpret_cast<const float4*>(wr + (k + 0) * 16);
          const float4 w1 = *reinterpret_cast<const float4*>(wr + (k + 1) * 16);
          const float4 w2 = *reinterpret_cast<const float4*>(wr + (k + 2) * 16);
          const float4 w3 = *reinterpret_cast<const float4*>(wr + (k + 3) * 16);
          a0 = fmaf(hv.x, w0.x, a0); a1 = fmaf(hv.x, w0.y, a1); a2 = fmaf(hv.x, w0.z, a2); a3 = fmaf(hv.x, w0.w, a3);
          a0 = fmaf(hv.y, w1.x, a0); a1 = fmaf(hv.y, w1.y, a1); a2 = fmaf(hv.y, w1.z, a2); a3 = fmaf(hv.y, w1.w, a3);
          a0 = fmaf(hv.z, w2.x, a0); a1 = fmaf(hv.z, w2.y, a1); a2 = fmaf(hv.z, w2.z, a2); a3 = fmaf(hv.z, w2.w, a3);
          a0 = fmaf(hv.w, w3.x, a0); a1 = fmaf(hv.w, w3.y, a1); a2 = fmaf(hv.w, w3.z, a2); a3 = fmaf(hv.w, w3.w, a3);
        }
        a0 += x0; a1 += x1; a2 += x2; a3 += x3;
        pre_all[o] = a0; pre_all[o + D] = a1; pre_all[o + 2 * (size_t)D] = a2; pre_all[o + 3 * (size_t)D] = a3;
        const size_t so = ((size_t)(t + 1) * N + n) * D + d;
        const float hp = hr[d];
        float cn = cp, hn = hp;
        if (tok != 0) {
          cn = cp * sigmoid_acc(a2 + 1.0f) + sigmoid_acc(a0) * tanhf(a1);
          hn = tanhf(cn) * sigmoid_acc(a3);
        }
        c_all[so] = cn;
        h_all[so] = hn;
      }
    }
    if (t + 1 < T) grid_barrier(bar, (unsigned int)(t + 1) * gridDim.x);
  }
}

__global__ void __launch_bounds__(LS_THREADS, 1)
lstm_seq_bwd_kernel(const float* __restrict__ g_hext, const float* __restrict__ pre_all, const float* __restrict__ c_all,
                    const float* __restrict__ kh, const int32_t* __restrict__ ids, int T, int N, int D, float* g_pre_all,
                    unsigned int* bar) {
  extern __shared__ __align__(16) float ls_smem[];
  float* WT = ls_smem;                 // [unit][4D + 4]: rows d0..d0+3 of the recurrent weights (row pitch skewed by 4 banks)
  float* GS = ls_smem + (size_t)LS_UNITS * (4 * D + 4);      // [LS_ROWS][LS_GC + 4]: stage of the gate gradients
  const int tid = threadIdx.x, u = tid & 3, nl = tid >> 2;
  const int d0 = blockIdx.x * LS_UNITS, d = d0 + u;
  const int C4 = 4 * D, WTS = C4 + 4;
  for (int i = tid; i < LS_UNITS * C4; i += LS_THREADS) WT[(i / C4) * WTS + (i % C4)] = __ldg(kh + (size_t)d0 * C4 + i);
  float g_c[LS_MAXCH], g_hrec[LS_MAXCH], g_pass[LS_MAXCH];
#pragma unroll
  for (int ci = 0; ci < LS_MAXCH; ci++) { g_c[ci] = 0.f; g_hrec[ci] = 0.f; g_pass[ci] = 0.f; }
  const int nch = (N + LS_ROWS - 1) / LS_ROWS;
  unsigned int epoch = 0;
  for (int t = T - 1; t >= 0; t--) {
#pragma unroll
    for (int ci = 0; ci < LS_MAXCH; ci++) {
      const int n = ci * LS_ROWS + nl;
      if (ci < nch && n < N) {
        const size_t si = ((size_t)t * N + n) * D + d;
        const size_t o = ((size_t)t * N + n) * C4 + d;
        const float gh = __ldg(g_hext + si) + g_hrec[ci];
        if (__ldg(ids + (size_t)n * T + t) != 0) {
          const float si_ = sigmoid_acc(__ldg(pre_all + o)), tj = tanhf(__ldg(pre_all + o + D)),
                      sf = sigmoid_acc(__ldg(pre_all + o + 2 * (size_t)D) + 1.0f), so = sigmoid_acc(__ldg(pre_all + o + 3 * (size_t)D));
          const float cp = __ldg(c_all + si);
          const float cn = cp * sf + si_ * tj;
          const float tc = tanhf(cn);
          const float g_cn = g_c[ci] + gh * so * (1.f - tc * tc);
          g_pre_all[o] = g_cn * tj * si_ * (1.f - si_);
          g_pre_all[o + D] = g_cn * si_ * (1.f - tj * tj);
          g_pre_all[o + 2 * (size_t)D] = g_cn * cp * sf * (1.f - sf);
          g_pre_all[o + 3 * (size_t)D] = gh * tc * so * (1.f - so);
          g_c[ci] = g_cn * sf;
          g_pass[ci] = 0.f;
        } else {
          g_pre_all[o] = 0.f; g_pre_all[o + D] = 0.f; g_pre_all[o + 2 * (size_t)D] = 0.f; g_pre_all[o + 3 * (size_t)D] = 0.f;
          g_pass[ci] = gh;
        }
      }
    }
    if (t == 0) break;                  // the initial state is a constant: nothing consumes g_h(-1)
    epoch++;
    grid_barrier(bar, epoch * gridDim.x);
    const float4* wt = reinterpret_cast<const float4*>(WT + (size_t)u * WTS);
    // g_h(t-1)[n, k] = sum_col g_pre[n, col] * K[D + k, col]: the 4D gate gradients of the chunk's rows stream through a
    // shared-memory stage in column blocks of LS_GC (coalesced cp.async, all in flight), four k per row from the stage
#pragma unroll
    for (int ci = 0; ci < LS_MAXCH; ci++) {
      if (ci >= nch) continue;
      const int n0 = ci * LS_ROWS, nn = min(LS_ROWS, N - n0);
      float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
      for (int cb = 0; cb < C4; cb += LS_GC) {
        const int gc4 = min(LS_GC, C4 - cb) >> 2;           // float4 columns of this pass (4D is a multiple of 16)
        __syncthreads();
        const float* gsrc = g_pre_all + ((size_t)t * N + n0) * C4 + cb;
        for (int i = tid; i < nn * gc4; i += LS_THREADS) {
          const int r = i / gc4, c4 = i - r * gc4;
          ls_cp_async16(GS + (size_t)r * (LS_GC + 4) + c4 * 4, gsrc + (size_t)r * C4 + c4 * 4);
        }
        ls_cp_async_wait_all();
        __syncthreads();
        if (nl < nn) {
          const float4* gp = reinterpret_cast<const float4*>(GS + (size_t)nl * (LS_GC + 4));
          const float4* wv4 = wt + (cb >> 2);
#pragma unroll 8
          for (int c = 0; c < gc4; c++) {
            const float4 gv = gp[c];
            const float4 wv = wv4[c];
            s0 = fmaf(gv.x, wv.x, s0); s1 = fmaf(gv.y, wv.y, s1); s2 = fmaf(gv.z, wv.z, s2); s3 = fmaf(gv.w, wv.w, s3);
          }
        }
      }
      if (nl < nn) g_hrec[ci] = (s0 + s1) + (s2 + s3) + g_pass[ci];
    }
  }
}

__global__ void embedding_all_fwd_kernel(const float* __restrict__ table, const int32_t* __restrict__ ids, int N, int T, int D,
                                         float* __restrict__ out) {
  long long total = (long long)T * N * D;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    int c = (int)(i % D);
    long long r = i / D;
    int n = (int)(r % N), t = (int)(r / N);
    out[i] = table[(long long)ids[n * T + t] * D + c];
  }
}
__global__ void embedding_all_bwd_kernel(const float* __restrict__ g, const int32_t* __restrict__ ids, int N, int T, int D,
                                         float* dtable) {
  long long total = (long long)T * N * D;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    int c = (int)(i % D);
    long long r = i / D;
    int n = (int)(r % N), t = (int)(r / N);
    float v = g[i];
    if (v != 0.f) atomicAdd(&dtable[(long long)ids[n * T + t] * D + c], v);
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
int fgc_embedding_all_fwd(const float* table, const int32_t* ids, int N, int T, int D, float* out, fgc_stream stream) {
  embedding_all_fwd_kernel<<<ew_grid((long long)T * N * D, 256), 256, 0, as_stream(stream)>>>(table, ids, N, T, D, out);
  count_launch();
  FGC_LAUNCH_CHECK("embedding_all_fwd");
  return FGC_OK;
}
int fgc_embedding_all_bwd(const float* g, const int32_t* ids, int N, int T, int D, float* dtable, fgc_stream stream) {
  embedding_all_bwd_kernel<<<ew_grid((long long)T * N * D, 256), 256, 0, as_stream(stream)>>>(g, ids, N, T, D, dtable);
  count_launch();
  FGC_LAUNCH_CHECK("embedding_all_bwd");
  return FGC_OK;
}
static int lstm_seq_check(int T, int N, int D) {
  FGC_REQUIRE(T >= 1 && N >= 1, "lstm_seq: empty sequence or batch");
  FGC_REQUIRE(D % LS_UNITS == 0 && D >= 16 && D <= 512, "lstm_seq: hidden size %d (multiple of 4, 16..512)", D);
  FGC_REQUIRE(D / LS_UNITS <= num_sms(), "lstm_seq: %d CTAs must be co-resident on %d SMs", D / LS_UNITS, num_sms());
  FGC_REQUIRE(N <= LS_ROWS * LS_MAXCH, "lstm_seq: at most %d samples", LS_ROWS * LS_MAXCH);
  return FGC_OK;
}
int fgc_lstm_seq_fwd(const float* gx, const float* kh, const int32_t* ids, int T, int N, int D, float* h_all, float* c_all,
                     float* pre_all, unsigned int* barrier, fgc_stream stream) {
  int e = lstm_seq_check(T, N, D);
  if (e) return e;
  const size_t smem = ((size_t)D * 16 + (size_t)LS_ROWS * (D + 4)) * sizeof(float);
  static bool attr = false;
  if (!attr) { cudaFuncSetAttribute(lstm_seq_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024); attr = true; }
  cudaMemsetAsync(barrier, 0, sizeof(unsigned int), as_stream(stream));
  cudaMemsetAsync(h_all, 0, (size_t)N * D * sizeof(float), as_stream(stream));      // slot 0: the zero initial state (:190,:196)
  cudaMemsetAsync(c_all, 0, (size_t)N * D * sizeof(float), as_stream(stream));
  lstm_seq_fwd_kernel<<<D / LS_UNITS, LS_THREADS, smem, as_stream(stream)>>>(gx, kh, ids, T, N, D, h_all, c_all, pre_all, barrier);
  count_launch();
  FGC_LAUNCH_CHECK("lstm_seq_fwd");
  return FGC_OK;
}
int fgc_lstm_seq_bwd(const float* g_hext, const float* pre_all, const float* c_all, const float* kh, const int32_t* ids, int T, int N,
                     int D, float* g_pre_all, unsigned int* barrier, fgc_stream stream) {
  int e = lstm_seq_check(T, N, D);
  if (e) return e;
  const size_t smem = ((size_t)LS_UNITS * (4 * D + 4) + (size_t)LS_ROWS * (LS_GC + 4)) * sizeof(float);
  static bool attr_b = false;
  if (!attr_b) { cudaFuncSetAttribute(lstm_seq_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024); attr_b = true; }
  cudaMemsetAsync(barrier, 0, sizeof(unsigned int), as_stream(stream));
  lstm_seq_bwd_kernel<<<D / LS_UNITS, LS_THREADS, smem, as_stream(stream)>>>(g_hext, pre_all, c_all, kh, ids, T, N, D, g_pre_all, barrier);
  count_launch();
  FGC_LAUNCH_CHECK("lstm_seq_bwd");
  return FGC_OK;
}
int fgc_lstm_cell_fwd(const float* gates, const float* gates2, const float* grow, const float* c_prev,
                      const float* h_prev, const int32_t* ids, int T, int t, int N, int P, int D,
                      float* c, float* h, float* pre, fgc_stream stream) {
  const uintptr_t al = (uintptr_t)gates | (uintptr_t)gates2 | (uintptr_t)grow | (uintptr_t)c_prev | (uintptr_t)h_prev | (uintptr_t)c |
                       (uintptr_t)h | (uintptr_t)pre;
  if (D % 4 == 0 && (al & 15) == 0)
    lstm_cell_fwd_vec4_kernel<<<ew_grid((long long)N * P * (D / 4), 256), 256, 0, as_stream(stream)>>>(
        (const float4*)gates, (const float4*)gates2, (const float4*)grow, (const float4*)c_prev, (const float4*)h_prev, ids, T, t, N, P,
        D / 4, (float4*)c, (float4*)h, (float4*)pre);
  else
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
