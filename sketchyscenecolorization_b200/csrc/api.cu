// Error reporting + launch accounting for libfgcolor.so
#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <mutex>

#include "common.cuh"
#include <stdlib.h>

namespace fgc {
static thread_local char g_err[512] = "";
static std::atomic<long long> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }
int check_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("%s: %s", what, cudaGetErrorString(e));
    return FGC_ECUDA;
  }
  return FGC_OK;
}
}  // namespace fgc

// CRC-32C (Castagnoli), slicing-by-8, host only: the checksum of TensorFlow's tensor-bundle snapshots (tf_bundle.py)
static uint32_t g_crc_tab[8][256];
static std::once_flag g_crc_once;     // the table walk is live on non-x86 hosts, called from several reader threads
static void crc_init() {
  for (uint32_t i = 0; i < 256; i++) {
    uint32_t c = i;
    for (int k = 0; k < 8; k++) c = (c & 1) ? (c >> 1) ^ 0x82F63B78u : c >> 1;
    g_crc_tab[0][i] = c;
  }
  for (uint32_t i = 0; i < 256; i++)
    for (int t = 1; t < 8; t++) g_crc_tab[t][i] = (g_crc_tab[t - 1][i] >> 8) ^ g_crc_tab[0][g_crc_tab[t - 1][i] & 0xFF];
}

#if defined(__x86_64__)
// SSE4.2 crc32 instructions (Castagnoli polynomial in hardware): ~8 GB/s on one dependency chain against ~1 GB/s for the
// table walk -- the TFRecord reader checks 0.9 MB of payload per sample (tfrecord_input.py)
static inline uint64_t crc32c_hw_u64(uint64_t c, uint64_t v) { __asm__("crc32q %1, %0" : "+r"(c) : "rm"(v)); return c; }
static inline uint32_t crc32c_hw_u8(uint32_t c, uint8_t v) { __asm__("crc32b %1, %0" : "+r"(c) : "rm"(v)); return c; }
static bool cpu_has_sse42() {
  unsigned a = 1, b = 0, c = 0, d = 0;
  __asm__ volatile("cpuid" : "+a"(a), "=b"(b), "+c"(c), "=d"(d));
  return (c >> 20) & 1u;
}
static uint32_t crc32c_hw(const uint8_t* p, size_t n, uint32_t c) {
  while (n && (reinterpret_cast<uintptr_t>(p) & 7)) { c = crc32c_hw_u8(c, *p++); n--; }
  uint64_t c64 = c;
  while (n >= 32) {
    c64 = crc32c_hw_u64(c64, *reinterpret_cast<const uint64_t*>(p));
    c64 = crc32c_hw_u64(c64, *reinterpret_cast<const uint64_t*>(p + 8));
    c64 = crc32c_hw_u64(c64, *reinterpret_cast<const uint64_t*>(p + 16));
    c64 = crc32c_hw_u64(c64, *reinterpret_cast<const uint64_t*>(p + 24));
    p += 32; n -= 32;
  }
  while (n >= 8) { c64 = crc32c_hw_u64(c64, *reinterpret_cast<const uint64_t*>(p)); p += 8; n -= 8; }
  c = (uint32_t)c64;
  while (n--) c = crc32c_hw_u8(c, *p++);
  return c;
}
#endif

extern "C" {
/* crc32c of data[0:n) continuing from `crc` (0 for a fresh checksum); unmasked */
unsigned int fgc_crc32c(const void* data, size_t n, unsigned int crc) {
  const uint8_t* p = static_cast<const uint8_t*>(data);
  uint32_t c = ~crc;
#if defined(__x86_64__)
  static const bool hw = cpu_has_sse42() && !getenv("FGC_CRC_TABLE");
  if (hw) return ~crc32c_hw(p, n, c);
#endif
  std::call_once(g_crc_once, crc_init);
  while (n && (reinterpret_cast<uintptr_t>(p) & 7)) { c = g_crc_tab[0][(c ^ *p++) & 0xFF] ^ (c >> 8); n--; }
  while (n >= 8) {
    uint64_t v = *reinterpret_cast<const uint64_t*>(p) ^ c;
    c = g_crc_tab[7][v & 0xFF] ^ g_crc_tab[6][(v >> 8) & 0xFF] ^ g_crc_tab[5][(v >> 16) & 0xFF] ^ g_crc_tab[4][(v >> 24) & 0xFF] ^
        g_crc_tab[3][(v >> 32) & 0xFF] ^ g_crc_tab[2][(v >> 40) & 0xFF] ^ g_crc_tab[1][(v >> 48) & 0xFF] ^ g_crc_tab[0][v >> 56];
    p += 8; n -= 8;
  }
  while (n--) c = g_crc_tab[0][(c ^ *p++) & 0xFF] ^ (c >> 8);
  return ~c;
}
const char* fgc_last_error(void) { return fgc::g_err; }
int fgc_version(void) { return 100; }
long long fgc_launch_count(void) { return fgc::g_launches.load(); }
}
