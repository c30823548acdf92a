// Read-only streaming sweep on a B200: what does a pure reduction pass over a 604 MB bf16 tensor need (block size, blocks per
// SM, independent 16-byte loads in flight per thread, cache hint) to reach the HBM read rate?  Build: nvcc -O3 -gencode
// arch=compute_100a,code=sm_100a -o readbw readbw.cu ; the result decides the shape of the statistics / min-max / gradient
// reduction kernels in csrc/elementwise.cu.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

template <int MODE>
__device__ __forceinline__ uint4 ld16(const uint4* p) {
  uint4 v;
  if (MODE == 0) {
    v = __ldg(p);
  } else if (MODE == 1) {
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
  } else {
    asm volatile("ld.global.cs.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
  }
  return v;
}

// each thread: U independent 16-byte loads per iteration, consecutive threads contiguous, blocks own contiguous chunks
template <int U, int MODE>
__global__ void read_kernel(const uint4* __restrict__ x, long long nvec, long long chunk, unsigned* out) {
  long long b0 = (long long)blockIdx.x * chunk, b1 = b0 + chunk;
  if (b1 > nvec) b1 = nvec;
  unsigned acc = 0;
  for (long long i = b0 + threadIdx.x; i < b1; i += (long long)blockDim.x * U) {
    uint4 v[U];
#pragma unroll
    for (int u = 0; u < U; u++) {
      long long j = i + (long long)u * blockDim.x;
      v[u] = j < b1 ? ld16<MODE>(x + j) : make_uint4(0, 0, 0, 0);
    }
#pragma unroll
    for (int u = 0; u < U; u++) acc += v[u].x ^ v[u].y ^ v[u].z ^ v[u].w;
  }
  if (acc == 0x12345678u) out[0] = acc;
}

template <int U, int MODE>
float run(const uint4* x, long long nvec, int threads, int blocks, unsigned* out) {
  long long chunk = (nvec + blocks - 1) / blocks;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  read_kernel<U, MODE><<<blocks, threads>>>(x, nvec, chunk, out);
  cudaEventRecord(e0);
  for (int r = 0; r < 5; r++) read_kernel<U, MODE><<<blocks, threads>>>(x, nvec, chunk, out);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  return ms / 5;
}

int main() {
  const long long bytes = 64LL * 192 * 192 * 128 * 2;
  const long long nvec = bytes / 16;
  uint4* x; unsigned* out;
  cudaMalloc(&x, bytes); cudaMalloc(&out, 4);
  cudaMemset(x, 1, bytes);
  int sms = 0; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  printf("SMs %d, %.0f MB\n", sms, bytes / 1e6);
  const int thr[] = {256, 512, 1024};
  const int bps[] = {1, 2, 4, 8, 16, 24};
  for (int ti = 0; ti < 3; ti++)
    for (int bi = 0; bi < 6; bi++) {
      int threads = thr[ti], blocks = sms * bps[bi];
      if (threads * bps[bi] < 512) continue;
      float m[12];
      m[0] = run<1, 0>(x, nvec, threads, blocks, out); m[1] = run<2, 0>(x, nvec, threads, blocks, out);
      m[2] = run<4, 0>(x, nvec, threads, blocks, out); m[3] = run<8, 0>(x, nvec, threads, blocks, out);
      m[4] = run<1, 1>(x, nvec, threads, blocks, out); m[5] = run<2, 1>(x, nvec, threads, blocks, out);
      m[6] = run<4, 1>(x, nvec, threads, blocks, out); m[7] = run<8, 1>(x, nvec, threads, blocks, out);
      m[8] = run<4, 2>(x, nvec, threads, blocks, out); m[9] = run<8, 2>(x, nvec, threads, blocks, out);
      printf("threads %4d blocks/SM %2d | ldg U1 %.2f U2 %.2f U4 %.2f U8 %.2f | nc.noalloc U1 %.2f U2 %.2f U4 %.2f U8 %.2f | cs U4 %.2f U8 %.2f TB/s\n",
             threads, bps[bi], bytes / m[0] / 1e9, bytes / m[1] / 1e9, bytes / m[2] / 1e9, bytes / m[3] / 1e9, bytes / m[4] / 1e9,
             bytes / m[5] / 1e9, bytes / m[6] / 1e9, bytes / m[7] / 1e9, bytes / m[8] / 1e9, bytes / m[9] / 1e9);
    }
  cudaError_t e = cudaDeviceSynchronize();
  printf("status %s\n", cudaGetErrorString(e));
  return 0;
}
