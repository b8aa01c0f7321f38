// common.cuh — shared device/host helpers for the txl_b200 C-ABI library (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <math.h>
#include "../../include/txl_b200.h"

typedef __nv_bfloat16 bf16;

// ------------------------------------------------------------------ error plumbing
void txl_set_error(const char* fmt, ...);
#define TXL_CHECK_ARG(cond, ...)                      \
  do {                                                \
    if (!(cond)) {                                    \
      txl_set_error(__VA_ARGS__);                     \
      return TXL_EINVAL;                              \
    }                                                 \
  } while (0)
#define TXL_CUDA(expr)                                                                  \
  do {                                                                                  \
    cudaError_t _e = (expr);                                                            \
    if (_e != cudaSuccess) {                                                            \
      txl_set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
      return TXL_ECUDA;                                                                 \
    }                                                                                   \
  } while (0)
extern unsigned long long g_txl_launches;
#define TXL_LAUNCH_CHECK()          \
  do {                              \
    ++g_txl_launches;               \
    TXL_CUDA(cudaGetLastError());   \
  } while (0)

// ------------------------------------------------------------------ programmatic dependent launch (decode step)
// With txl_set_pdl(1) the decode-step kernels are launched with cudaLaunchAttributeProgrammaticStreamSerialization: the next kernel's CTAs
// may start (barrier init, weight / ring prefetch) while the previous kernel drains; every such kernel executes griddepcontrol.wait before
// it touches anything a predecessor wrote or may still read.  Without the attribute both instructions are no-ops.
extern int g_txl_pdl;
#ifdef __CUDACC__
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
template <typename... KArgs, typename... Args>
static inline cudaError_t txl_launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute attr[1];
  if (g_txl_pdl) {
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
  }
  return cudaLaunchKernelEx(&cfg, kern, KArgs(args)...);
}
#endif

static inline int64_t cdiv64(int64_t a, int64_t b) { return (a + b - 1) / b; }
static inline int64_t imin64(int64_t a, int64_t b) { return a < b ? a : b; }
int txl_num_sms();

// ------------------------------------------------------------------ dtype helpers
template <typename T> __device__ __forceinline__ float to_f32(T v);
template <> __device__ __forceinline__ float to_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ float to_f32<bf16>(bf16 v) { return __bfloat162float(v); }
template <typename T> __device__ __forceinline__ T from_f32(float v);
template <> __device__ __forceinline__ float from_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ bf16 from_f32<bf16>(float v) { return __float2bfloat16_rn(v); }

// ------------------------------------------------------------------ integer geometry of the band  [A.2-4, A.4, A.5]
// These four functions ARE the closed forms of HF's uint8 triu+tril mask and pad/reshape _rel_shift;
// tests/test_index_maps.py checks them bit-exactly against the literal construction.
struct BandGeom {
  int T, mlen, klen, msl, clamp;  // msl = mask_shift_len (same_length) or a huge value (plain causal)
};
__host__ __device__ __forceinline__ BandGeom make_band(const TxlBand& b) {
  BandGeom g;
  g.T = b.T; g.mlen = b.mlen; g.klen = b.mlen + b.T; g.clamp = b.clamp_len;
  if (b.same_length) {
    int mask_len = g.klen - b.mem_len;
    g.msl = mask_len > 0 ? b.T - mask_len : b.T;
  } else {
    g.msl = 1 << 29;
  }
  return g;
}
// first / last live key (inclusive) of query row i
__host__ __device__ __forceinline__ int band_lo(const BandGeom& g, int i) {
  long lo = (long)i - g.msl + 1;
  return lo < 0 ? 0 : (int)lo;
}
__host__ __device__ __forceinline__ int band_hi(const BandGeom& g, int i) { return i + g.mlen; }
// row of the r-table used at (i, j): relative distance, clamped
__host__ __device__ __forceinline__ int band_ridx(const BandGeom& g, int i, int j) {
  int p = g.mlen + i - j;
  return (g.clamp > 0 && p > g.clamp) ? g.clamp : p;
}
// rows of the r-table a call needs: distances 0..P-1
__host__ __device__ __forceinline__ int band_num_r(const BandGeom& g) {
  int pmax = g.klen - 1;
  if (g.clamp > 0 && pmax > g.clamp) pmax = g.clamp;
  return pmax + 1;
}

// ------------------------------------------------------------------ counter-based dropout
// keep-mask for logical element `idx` of dropout site `site`; identical in forward and backward and in every kernel.
// One 32-bit avalanche hash (lowbias32) serves two neighbouring elements (16 random bits each).
__device__ __forceinline__ uint32_t dropout_key(uint64_t seed, uint32_t site) {
  return (uint32_t)seed * 0x9E3779B1u + (uint32_t)(seed >> 32) * 0x85EBCA77u + (site + 1u) * 0xC2B2AE3Du;
}
__device__ __forceinline__ uint32_t dropout_hash(uint32_t key, uint64_t pair_idx) {
  uint32_t x = (uint32_t)pair_idx + key + (uint32_t)(pair_idx >> 32) * 0x27D4EB2Fu;
  x ^= x >> 16; x *= 0x7FEB352Du;
  x ^= x >> 15; x *= 0x846CA68Bu;
  x ^= x >> 16;
  return x;
}
__device__ __forceinline__ uint32_t dropout_threshold(float p) { return (uint32_t)(p * 65536.0f); }
__device__ __forceinline__ float dropout_scale(uint64_t seed, uint32_t site, uint64_t idx, float p, float inv_keep) {
  const uint32_t h = dropout_hash(dropout_key(seed, site), idx >> 1);
  const uint32_t u = (idx & 1) ? (h >> 16) : (h & 0xFFFFu);
  return u >= dropout_threshold(p) ? inv_keep : 0.0f;
}

// keep-scales of 8 consecutive elements starting at the EVEN index idx0 (4 hashes instead of 8, key and threshold formed once)
__device__ __forceinline__ void dropout_scale8(uint32_t key, uint32_t thr, uint64_t idx0, float inv_keep, float* sc) {
  const uint64_t pair0 = idx0 >> 1;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const uint32_t h = dropout_hash(key, pair0 + j);
    sc[2 * j] = (h & 0xFFFFu) >= thr ? inv_keep : 0.0f;
    sc[2 * j + 1] = (h >> 16) >= thr ? inv_keep : 0.0f;
  }
}

// ------------------------------------------------------------------ warp reductions
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
