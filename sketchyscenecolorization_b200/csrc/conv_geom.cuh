// Geometry shared by the convolution kernels: how the reduction (K) axis of the implicit GEMM is cut into
// 64-wide "slabs" over the concatenated sources and the filter taps.
//
//   source with C >= 64 and C % 8 == 0 ("big"):  one run of slabs per tap, ceil(C/64) slabs each
//                                                (a partial last slab is zero filled);
//   any other source ("small", e.g. the 3-channel sketch pyramid, 8-channel stem features):
//                                                taps and channels are flattened, q = tap*C + c, and cut
//                                                into ceil(k*k*C/64) slabs.
// Slabs are ordered source by source (tf.concat order, mru.py:403,552,572).  The same decode is used by the
// weight packer, the forward/dgrad producers and the wgrad epilogue, so they cannot disagree.
#pragma once
#include "common.cuh"

namespace fgc {

constexpr int kMaxSrc = 4;

struct ConvGeom {
  const void* src[kMaxSrc];
  const void* patch[kMaxSrc];   // narrow sources: optional pre-flattened (tap, channel) bf16 tensor, 64 channels per slab
  int C[kMaxSrc];
  int ups[kMaxSrc];
  int cbase[kMaxSrc];        // first channel of the source inside the concatenated input
  int big[kMaxSrc];
  int slab_begin[kMaxSrc + 1];
  int nsrc;
  int N, H, W;               // logical (full-resolution) input size
  int OH, OW;
  int k, stride, pad_t, pad_l, sign;   // input row = oh*stride + sign*(kh - pad_t)
  int nslabs;
  int ow_bits, oh_bits;      // pixel packing: ow | oh << ow_bits | n << (ow_bits + oh_bits)
  long long M;               // N*OH*OW
};

struct SlabInfo {
  int s;        // source index
  int big;
  int tap;      // big: filter tap
  int c0;       // big: first channel inside the source
  int q0;       // small: first flattened (tap*C + c) index
};

__host__ __device__ inline SlabInfo decode_slab(const ConvGeom& g, int slab) {
  SlabInfo r;
  int s = 0;
  while (s + 1 < g.nsrc && slab >= g.slab_begin[s + 1]) s++;
  int local = slab - g.slab_begin[s];
  r.s = s;
  r.big = g.big[s];
  if (r.big) {
    int ncb = (g.C[s] + 63) / 64;
    r.tap = local / ncb;
    r.c0 = (local % ncb) * 64;
    r.q0 = 0;
  } else {
    r.tap = 0;
    r.c0 = 0;
    r.q0 = local * 64;
  }
  return r;
}

// (tap, channel-in-concat) of element kk of a slab, or false if it is padding
__host__ __device__ inline bool slab_elem(const ConvGeom& g, const SlabInfo& si, int kk, int* tap, int* cglob) {
  if (si.big) {
    int c = si.c0 + kk;
    if (c >= g.C[si.s]) return false;
    *tap = si.tap;
    *cglob = g.cbase[si.s] + c;
    return true;
  }
  int q = si.q0 + kk;
  if (q >= g.k * g.k * g.C[si.s]) return false;
  *tap = q / g.C[si.s];
  *cglob = g.cbase[si.s] + q % g.C[si.s];
  return true;
}

// does (n, oh, ow) fit the 32-bit pixel packing (0xFFFFFFFF is the invalid-row sentinel)?
inline bool geom_fits(const ConvGeom& g) {
  int nb = 32 - g.ow_bits - g.oh_bits;
  return nb >= 1 && (nb >= 32 ? (long long)g.N < 0xFFFFFFFFLL : (long long)g.N < (1LL << nb) - 1);
}

inline int finish_geom(ConvGeom& g) {
  int sb = 0, cb = 0;
  for (int s = 0; s < g.nsrc; s++) {
    g.cbase[s] = cb;
    g.big[s] = (g.C[s] >= 64 && g.C[s] % 8 == 0) ? 1 : 0;
    g.slab_begin[s] = sb;
    sb += g.big[s] ? g.k * g.k * ((g.C[s] + 63) / 64) : (g.k * g.k * g.C[s] + 63) / 64;
    cb += g.C[s];
  }
  g.slab_begin[g.nsrc] = sb;
  g.nslabs = sb;
  g.M = (long long)g.N * g.OH * g.OW;
  g.ow_bits = 0;
  while ((1 << g.ow_bits) < g.OW) g.ow_bits++;
  g.oh_bits = 0;
  while ((1 << g.oh_bits) < g.OH) g.oh_bits++;
  return cb;
}

}  // namespace fgc
