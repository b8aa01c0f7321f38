// simt_ops.cu — bandwidth-bound kernels of the Transformer-XL path: embedding, sinusoid table, residual+LayerNorm,
// dropout, log-softmax/NLL, parameter casts/transposes, AdamW, index maps.  All fp32 math, fp32 or bf16 storage.
#include "common.cuh"
#include <stdarg.h>
#include <string.h>

// ------------------------------------------------------------------ error plumbing
static thread_local char g_err[512] = "";
void txl_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
extern "C" const char* txl_last_error(void) { return g_err; }
extern "C" int txl_version(void) { return 100; }
unsigned long long g_txl_launches = 0;
int g_txl_pdl = 0;
extern "C" int txl_set_pdl(int on) { const int old = g_txl_pdl; g_txl_pdl = on ? 1 : 0; return old; }
extern "C" unsigned long long txl_launch_count(void) { return g_txl_launches; }

int txl_num_sms() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}
extern "C" int txl_device_ok(void) {
  int dev = 0, major = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) { txl_set_error("no CUDA device"); return TXL_ENODEV; }
  cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
  if (major != 10) { txl_set_error("device compute capability %d.x is not sm_100", major); return TXL_ENODEV; }
  return TXL_OK;
}

#define DISPATCH_DTYPE(dtype, ...)                                    \
  if ((dtype) == TXL_F32) { typedef float T; __VA_ARGS__; }           \
  else if ((dtype) == TXL_BF16) { typedef bf16 T; __VA_ARGS__; }      \
  else { txl_set_error("bad dtype %d", (int)(dtype)); return TXL_EINVAL; }

// ------------------------------------------------------------------ index maps
__global__ void index_map_kernel(TxlBand band, uint8_t* masked, int32_t* ridx, int32_t* lo, int32_t* hi) {
  BandGeom g = make_band(band);
  int64_t n = (int64_t)g.T * g.klen;
  for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < n; idx += (int64_t)gridDim.x * blockDim.x) {
    int i = (int)(idx / g.klen), j = (int)(idx % g.klen);
    bool live = j >= band_lo(g, i) && j <= band_hi(g, i);
    masked[idx] = live ? 0 : 1;
    ridx[idx] = live ? band_ridx(g, i, j) : -1;
    if (j == 0) { lo[i] = band_lo(g, i); hi[i] = min(band_hi(g, i), g.klen - 1); }
  }
}
extern "C" int txl_relattn_index_map(const TxlBand* band, uint8_t* masked, int32_t* ridx, int32_t* lo, int32_t* hi, void* stream) {
  TXL_CHECK_ARG(band && band->T > 0 && band->mlen >= 0, "index_map: bad band");
  index_map_kernel<<<64, 256, 0, (cudaStream_t)stream>>>(*band, masked, ridx, lo, hi);
  TXL_LAUNCH_CHECK();
  return TXL_OK;
}

// ------------------------------------------------------------------ embedding
template <typename T>
__global__ void embed_fwd_kernel(const int64_t* __restrict__ ids, const T* __restrict__ E, T* __restrict__ out, int64_t n_tok,
                                 int d, int V, float scale, float p, float inv_keep, uint64_t seed, uint32_t site) {
  int64_t total = n_tok * d;
  for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    int64_t n = idx / d; int c = (int)(idx % d);
    int64_t id = ids[n];
    float v = (id >= 0 && id < V) ? to_f32(E[id * d + c]) * scale : 0.f;
    if (p > 0.f) v *= dropout_scale(seed, site, (uint64_t)idx, p, inv_keep);
    out[idx] = from_f32<T>(v);
  }
}
template <typename T>
__global__ void embed_bwd_kernel(const int64_t* __restrict__ ids, const T* __restrict__ dOut, const T* __restrict__ dOut2, float* __restrict__ dE, int64_t n_tok,
                                 int d, int V, float scale, float p, float inv_keep, uint64_t seed, uint32_t site) {
  int64_t total = n_tok * d;
  for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    int64_t n = idx / d; int c = (int)(idx % d);
    int64_t id = ids[n];
    if (id < 0 || id >= V) continue;
    float v = to_f32(dOut[idx]);
    if (dOut2) v += to_f32(dOut2[idx]);
    v *= scale;
    if (p > 0.f) v *= dropout_scale(seed, site, (uint64_t)idx, p, inv_keep);
    atomicAdd(&dE[id * d + c], v);
  }
}
extern "C" int txl_embed_fwd(const int64_t* ids, const void* E, void* out, int64_t n_tok, int d, int V, float scale,
                             int dtype, float drop_p, uint64_t seed, uint32_t site, void* stream) {
  TXL_CHECK_ARG(n_tok > 0 && d > 0 && V > 0, "embed_fwd: bad sizes");
  int grid = (int)imin64(cdiv64(n_tok * d, 256), (int64_t)txl_num_sms() * 16);
  float ik = drop_p > 0.f ? 1.f / (1.f - drop_p) : 1.f;
  DISPATCH_DTYPE(dtype, (embed_fwd_kernel<T><<<grid, 256, 0, (cudaStream_t)stream>>>(ids, (const T*)E, (T*)out, n_tok, d, V, scale, drop_p, ik, seed, site)));
  TXL_LAUNCH_CHECK();
  return TXL_OK;
}
extern "C" int txl_embed_bwd(const int64_t* ids, const void* dOut, const void* dOut2, float* dE, int64_t n_tok, int d, int V, float scale,
                             int dtype, float drop_p, uint64_t seed, uint32_t site, void* stream) {
  TXL_CHECK_ARG(n_tok > 0 && d > 0 && V > 0, "embed_bwd: bad sizes");
  int grid = (int)imin64(cdiv64(n_tok * d, 256), (int64_t)txl_num_sms() * 16);
  float ik = drop_p > 0.f ? 1.f / (1.f - drop_p) : 1.f;
  DISPATCH_DTYPE(dtype, (embed_bwd_kernel<T><<<grid, 256, 0, (cudaStream_t)stream>>>(ids, (const T*)dOut, (const T*)dOut2, dE, n_tok, d, V, scale, drop_p, ik, seed, site)));
  TXL_LAUNCH_CHECK();
  return TXL_OK;
}

// ------------------------------------------------------------------ sinusoid table
template <typename T>
__global__ void posemb_kernel(T* out, int klen, int clamp, int d, float p, float inv_keep, uint64_t seed, uint32_t site) {
  int half = d / 2;
  int64_t total = (int64_t)klen * d;
  for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    int x = (int)(idx / d), c = (int)(idx % d);
    int pos = klen - 1 - x;                       // pos_seq = arange(klen-1, -1, -1)
    if (clamp > 0 && pos > clamp) pos = clamp;    // .clamp_(max=clamp_len)
    int k = c < half ? c : c - half;
    // inv_freq = 1 / 10000^(2k/d) evaluated like torch: fp32 pow then reciprocal; product in fp32
    float inv_freq = 1.0f / powf(10000.0f, (float)(2 * k) / (float)d);
    float a = (float)pos * inv_freq;
    float v = c < half ? sinf(a) : cosf(a);
    if (p > 0.f) v *= dropout_scale(seed, site, (uint64_t)idx, p, inv_keep);
    out[idx] = from_f32<T>(v);
  }
}
extern "C" int txl_posemb_table(void* out, int klen, int clamp_len, int d, int dtype, float drop_p, uint64_t seed, uint32_t site, void* stream) {
  TXL_CHECK_ARG(klen > 0 && d > 0 && d % 2 == 0, "posemb: bad sizes");
  int grid = (int)imin64(cdiv64((int64_t)klen * d, 256), 4096);
  float ik = drop_p > 0.f ? 1.f / (1.f - drop_p) : 1.f;
  DISPATCH_DTYPE(dtype, (posemb_kernel<T><<<grid, 256, 0, (cudaStream_t)stream>>>((T*)out, klen, clamp_len, d, drop_p, ik, seed, site)));
  TXL_LAUNCH_CHECK();
  return TXL_OK;
}

// ------------------------------------------------------------------ residual + LayerNorm (one warp per row, 16-byte vectors)
constexpr int LN_VEC = 8;           // elements per lane per vector step
constexpr int LN_MAX_STEPS = 4;     // d <= 1024
template <typename T> struct Vec8 { T v[LN_VEC]; };
template <typename T> __device__ __forceinline__ void ld8(const T* p, float* f) {
  if constexpr (sizeof(T) == 2) {
    uint4 u = *reinterpret_cast<const uint4*>(p);
    const bf16* e = reinterpret_cast<const bf16*>(&u);
#pragma unroll
    for (int k = 0; k < 8; ++k) f[k] = __bfloat162float(e[k]);
  } else {
    float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
    f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w; f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
  }
}
template <typename T> __device__ __forceinline__ void st8(T* p, const float* f) {
  if constexpr (sizeof(T) == 2) {
    uint4 u;
    __nv_bfloat162 a = __floats2bfloat162_rn(f[0], f[1]), b = __floats2bfloat162_rn(f[2], f[3]), c = __floats2bfloat162_rn(f[4], f[5]), d = __floats2bfloat162_rn(f[6], f[7]);
    u.x = *reinterpret_cast<uint32_t*>(&a); u.y = *reinterpret_cast<uint32_t*>(&b); u.z = *reinterpret_cast<uint32_t*>(&c); u.w = *reinterpret_cast<uint32_t*>(&d);
    *reinterpret_cast<uint4*>(p) = u;
  } else {
    *reinterpret_cast<float4*>(p) = make_float4(f[0], f[1], f[2], f[3]);
    *reinterpret_cast<float4*>(p + 4) = make_float4(f[4], f[5], f[6], f[7]);
  }
}
// rounds through the storage type so that forward statistics and backward see the same z
template <typename T> __device__ __forceinline__ float round_to(float v) { return to_f32(from_f32<T>(v)); }

template <typename T, int STEPS>
__global__ void __launch_bounds__(256, (STEPS <= 2 ? 4 : 2)) add_ln_fwd_kernel(const T* __restrict__ x, const T* __restrict__ r, const float* __restrict__ gamma,
                                  const float* __restrict__ beta, T* __restrict__ y, T* __restrict__ z, float* __restrict__ mean,
                                  float* __restrict__ rstd, int64_t rows, int d, float eps, float p, float inv_keep, uint64_t seed,
                                  uint32_t site) {
  const int lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
  const int64_t warp = blockIdx.x * (int64_t)wpb + (threadIdx.x >> 5);
  const uint32_t dkey = dropout_key(seed, site), dthr = dropout_threshold(p);
  for (int64_t row = warp; row < rows; row += (int64_t)gridDim.x * wpb) {
    float v[STEPS][LN_VEC];
    float s = 0.f;
#pragma unroll
    for (int e = 0; e < STEPS; ++e) {
      const int c = (e * 32 + lane) * LN_VEC;
      if (c < d) {
        float xv[LN_VEC], rv[LN_VEC], ks[LN_VEC];
        ld8(x + row * d + c, xv);
        if (r) ld8(r + row * d + c, rv);
        if (p > 0.f) dropout_scale8(dkey, dthr, (uint64_t)(row * d + c), inv_keep, ks);
#pragma unroll
        for (int k = 0; k < LN_VEC; ++k) {
          float rr = r ? rv[k] : 0.f;
          if (p > 0.f) rr *= ks[k];
          v[e][k] = round_to<T>(xv[k] + rr);
          s += v[e][k];
        }
        if (z) st8(z + row * d + c, v[e]);
      }
    }
    const float mu = warp_sum(s) / d;
    float q = 0.f;
#pragma unroll
    for (int e = 0; e < STEPS; ++e) {
      const int c = (e * 32 + lane) * LN_VEC;
      if (c < d) {
#pragma unroll
        for (int k = 0; k < LN_VEC; ++k) { float t = v[e][k] - mu; q += t * t; }
      }
    }
    const float rs = rsqrtf(warp_sum(q) / d + eps);
#pragma unroll
    for (int e = 0; e < STEPS; ++e) {
      const int c = (e * 32 + lane) * LN_VEC;
      if (c < d) {
        float gm[LN_VEC], bt[LN_VEC], o[LN_VEC];
        ld8(gamma + c, gm); ld8(beta + c, bt);
#pragma unroll
        for (int k = 0; k < LN_VEC; ++k) o[k] = (v[e][k] - mu) * rs * gm[k] + bt[k];
        st8(y + row * d + c, o);
      }
    }
    if (lane == 0) { if (mean) mean[row] = mu; if (rstd) rstd[row] = rs; }
  }
}
extern "C" int txl_add_ln_fwd(const void* x, const void* r, const float* gamma, const float* beta, void* y, void* z,
                              float* mean, float* rstd, int64_t rows, int d, float eps, int dtype, float drop_p, uint64_t seed,
                              uint32_t site, void* stream) {
  TXL_CHECK_ARG(rows > 0 && d % LN_VEC == 0 && d <= 32 * LN_VEC * LN_MAX_STEPS, "add_ln_fwd: d=%d must be a multiple of 8, <= 1024", d);
  int grid = (int)imin64(cdiv64(rows, 8), (int64_t)txl_num_sms() * 8);
  float ik = drop_p > 0.f ? 1.f / (1.f - drop_p) : 1.f;
#define LN_FWD_LAUNCH(S) add_ln_fwd_kernel<T, S><<<grid, 256, 0, (cudaStream_t)stream>>>((const T*)x, (const T*)r, gamma, beta, (T*)y, (T*)z, mean, rstd, rows, d, eps, drop_p, ik, seed, site)
  DISPATCH_DTYPE(dtype, { if (d <= 256) LN_FWD_LAUNCH(1); else if (d <= 512) LN_FWD_LAUNCH(2); else LN_FWD_LAUNCH(4); });
#undef LN_FWD_LAUNCH
  TXL_LAUNCH_CHECK();
  return TXL_OK;
}

// dy_total = dy (+ dy2);  dz = LN'(dy_total)
template <typename T, int STEPS>
__global__ void __launch_bounds__(256, (STEPS <= 2 ? 2 : 1)) add_ln_bwd_kernel(const T* __restrict__ dy, const T* __restrict__ dy2, const T* __restrict__ z, const float* __restrict__ gamma,
                                  const float* __restrict__ mean, const float* __restrict__ rstd, T* dx_out, int accumulate_dx,
                                  T* dr_out, float* __restrict__ dgamma, float* __restrict__ dbeta, int64_t rows, int d, float p,
                                  float inv_keep, uint64_t seed, uint32_t site) {
  extern __shared__ float sm[];  // [2][d] block partials
  const uint32_t dkey = dropout_key(seed, site), dthr = dropout_threshold(p);
  const int lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
  const int64_t warp = blockIdx.x * (int64_t)wpb + (threadIdx.x >> 5);
  for (int c = threadIdx.x; c < 2 * d; c += blockDim.x) sm[c] = 0.f;
  __syncthreads();
  float ag[STEPS][LN_VEC], ab[STEPS][LN_VEC];
#pragma unroll
  for (int e = 0; e < STEPS; ++e)
#pragma unroll
    for (int k = 0; k < LN_VEC; ++k) { ag[e][k] = 0.f; ab[e][k] = 0.f; }
  const int64_t row_step = (int64_t)gridDim.x * wpb;
  for (int64_t row = warp; row < rows; row += row_step) {
    // pull the next row of this warp towards L2 while this one is reduced and stored (pure HBM streaming kernel; one prefetch per 128-byte line)
    if (row + row_step < rows && (lane & 7) == 0) {
#pragma unroll
      for (int e = 0; e < STEPS; ++e) {
        const int c = (e * 32 + lane) * LN_VEC;
        if (c < d) {
          const int64_t o = (row + row_step) * d + c;
          asm volatile("prefetch.global.L2 [%0];" ::"l"(dy + o));
          if (dy2) asm volatile("prefetch.global.L2 [%0];" ::"l"(dy2 + o));
          asm volatile("prefetch.global.L2 [%0];" ::"l"(z + o));
        }
      }
    }
    const float mu = mean[row], rs = rstd[row];
    float g[STEPS][LN_VEC], xh[STEPS][LN_VEC];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int e = 0; e < STEPS; ++e) {
      const int c = (e * 32 + lane) * LN_VEC;
      if (c < d) {
        float dyv[LN_VEC], zv[LN_VEC], gm[LN_VEC];
        ld8(dy + row * d + c, dyv);
        if (dy2) { float t2[LN_VEC]; ld8(dy2 + row * d + c, t2);
#pragma unroll
          for (int k = 0; k < LN_VEC; ++k) dyv[k] += t2[k]; }
        ld8(z + row * d + c, zv); ld8(gamma + c, gm);
#pragma unroll
        for (int k = 0; k < LN_VEC; ++k) {
          xh[e][k] = (zv[k] - mu) * rs;
          g[e][k] = dyv[k] * gm[k];
          s1 += g[e][k]; s2 += g[e][k] * xh[e][k];
          ag[e][k] += dyv[k] * xh[e][k]; ab[e][k] += dyv[k];
        }
      }
    }
    s1 = warp_sum(s1) / d; s2 = warp_sum(s2) / d;
#pragma unroll
    for (int e = 0; e < STEPS; ++e) {
      const int c = (e * 32 + lane) * LN_VEC;
      if (c < d) {
        float dz[LN_VEC];
#pragma unroll
        for (int k = 0; k < LN_VEC; ++k) dz[k] = rs * (g[e][k] - s1 - xh[e][k] * s2);
        if (dr_out) {
          float o[LN_VEC], ks[LN_VEC];
          if (p > 0.f) dropout_scale8(dkey, dthr, (uint64_t)(row * d + c), inv_keep, ks);
#pragma unroll
          for (int k = 0; k < LN_VEC; ++k) o[k] = dz[k] * (p > 0.f ? ks[k] : 1.f);
          st8(dr_out + row * d + c, o);
        }
        if (dx_out) {
          if (accumulate_dx) { float old[LN_VEC]; ld8(dx_out + row * d + c, old);
#pragma unroll
            for (int k = 0; k < LN_VEC; ++k) dz[k] += old[k]; }
          st8(dx_out + row * d + c, dz);
        }
      }
    }
  }
#pragma unroll
  for (int e = 0; e < STEPS; ++e) {
    const int c = (e * 32 + lane) * LN_VEC;
    if (c < d) {
#pragma unroll
      for (int k = 0; k < LN_VEC; ++k) { atomicAdd(&sm[c + k], ag[e][k]); atomicAdd(&sm[d + c + k], ab[e][k]); }
    }
  }
  __syncthreads();
  for (int c = threadIdx.x; c < d; c += blockDim.x) {
    if (dgamma) atomicAdd(&dgamma[c], sm[c]);
    if (dbeta) atomicAdd(&dbeta[c], sm[d + c]);
  }
}
extern "C" int txl_add_ln_bwd(const void* dy, const void* dy2, const void* z, const float* gamma, const float* mean, const float* rstd,
                              void* dx_out, int accumulate_dx, void* dr_out, float* dgamma, float* dbeta, int64_t rows, int d,
                              int dtype, float drop_p, uint64_t seed, uint32_t site, void* stream) {
  TXL_CHECK_ARG(rows > 0 && d % LN_VEC == 0 && d <= 32 * LN_VEC * LN_MAX_STEPS, "add_ln_bwd: d=%d must be a multiple of 8, <= 1024", d);
  // dx_out may alias dy: each lane reads its elements of a row before any lane of the warp writes them
  int grid = (int)imin64(cdiv64(rows, 8), (int64_t)txl_num_sms() * 2);
  float ik = drop_p > 0.f ? 1.f / (1.f - drop_p) : 1.f;
  size_t smem = 2 * (size_t)d * sizeof(float);
#define LN_BWD_LAUNCH(S) add_ln_bwd_kernel<T, S><<<grid, 256, smem, (cudaStream_t)stream>>>((const T*)dy, (const T*)dy2, (const T*)z, gamma, mean, rstd, (T*)dx_out, accumulate_dx, (T*)dr_out, dgamma, dbeta, rows, d, drop_p, ik, seed, site)
  DISPATCH_DTYPE(dtype, { if (d <= 256) LN_BWD_LAUNCH(1); else if (d <= 512) LN_BWD_LAUNCH(2); else LN_BWD_LAUNCH(4); });
#undef LN_BWD_LAUNCH
  TXL_LAUNCH_CHECK();
  return TXL_OK;
}

// ------------------------------------------------------------------ column sums (bias gradients)
template <typename T>
__global__ void colsum_kernel(const T* __restrict__ X, int64_t M, int64_t N, int64_t ldx, float* __restrict__ out, int rows_per_block) {
  int64_t n = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (n >= N) return;
  int64_t m0 = (int64_t)blockIdx.y * rows_per_block, m1 = m0 + rows_per_block < M ? m0 + rows_per_block : M;
  float s = 0.f;
  for (int64_t m = m0; m < m1; ++m) s += to_f32(X[m * ldx + n]);
  atomicAdd(&out[n], s);
}
// bf16, N and ldx multiples of 8, 16-byte aligned rows: every thread owns 8 adjacent columns (one 16-byte load per row), a block walks
// its slice of rows with all lanes reading contiguous bytes, partial sums meet in shared memory and leave as one atomic per column per block
__global__ void __launch_bounds__(256) colsum_bf16_vec_kernel(const bf16* __restrict__ X, int64_t M, int64_t N, int64_t ldx, float* __restrict__ out) {
  __shared__ float red[256 * 8];
  const int tpr = (int)(N / 8);                         // threads per row (host guarantees tpr <= 256 and 256 % tpr == 0)
  const int rpi = 256 / tpr;                            // rows per block iteration
  const int col8 = threadIdx.x % tpr, rsub = threadIdx.x / tpr;
  float acc[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) acc[k] = 0.f;
  for (int64_t m = (int64_t)blockIdx.x * rpi + rsub; m < M; m += (int64_t)gridDim.x * rpi) {
    const uint4 u = *reinterpret_cast<const uint4*>(X + m * ldx + col8 * 8);
    const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
    for (int k = 0; k < 4; ++k) { acc[2 * k] += __uint_as_float(w[k] << 16); acc[2 * k + 1] += __uint_as_float(w[k] & 0xFFFF0000u); }
  }
#pragma unroll
  for (int k = 0; k < 8; ++k) red[threadIdx.x * 8 + k] = acc[k];
  __syncthreads();
  for (int c = threadIdx.x; c < (int)N; c += blockDim.x) {
    float s = 0.f;
    for (int r = 0; r < rpi; ++r) s += red[(r * tpr + c / 8) * 8 + (c & 7)];
    atomicAdd(&out[c], s);
  }
}
extern "C" int txl_colsum(const void* X, int64_t M, int64_t N, int64_t ldx, int dtype, float* out, void* stream) {
  TXL_CHECK_ARG(X && out && M > 0 && N > 0 && ldx >= N, "colsum: bad args");
  if (dtype == TXL_BF16 && N % 8 == 0 && ldx % 8 == 0 && N / 8 <= 256 && 256 % (N / 8) == 0 && (((uintptr_t)X) & 15) == 0 && M >= 1024) {
    const int rpi = 256 / (int)(N / 8);
    const int grid = (int)imin64(cdiv64(M, (int64_t)rpi * 4), (int64_t)txl_num_sms() * 4);
    colsum_bf16_vec_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>((const bf16*)X, M, N, ldx, out);
    TXL_LAUNCH_CHECK();
    return TXL_OK;
  }
  int rpb = 256;
  dim3 grid((unsigned)cdiv64(N, 128), (unsigned)cdiv64(M, rpb));
  DISPATCH_DTYPE(dtype, (colsum_kernel<T><<<grid, 128, 0, (cudaStream_t)stream>>>((const T*)X, M, N, ldx, out, rpb)));
  TXL_LAUNCH_CHECK();
  return TXL_OK;
}

// ------------------------------------------------------------------ plain dropout
template <typename T>
__global__ void dropout_kernel(const T* __restrict__ x, T* __restrict__ y, int64_t n, float p, float inv_keep, uint64_t seed, uint32_t site) {
  for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < n; idx += (int64_t)gridDim.x * blockDim.x)
    y[idx] = from_f32<T>(to_f32(x[idx]) * dropout_scale(seed, site, (uint64_t)idx, p, inv_keep));
}
extern "C" int txl_dropout(const void* x, void* y, int64_t n, int dtype, float drop_p, uint64_t seed, uint32_t site, void* stream) {
  TXL_CHECK_ARG(n > 0 && drop_p >= 0.f && drop_p < 1.f, "dropout: bad args");
  if (drop_p == 0.f) {
    if (x != y) TXL_CUDA(cudaMemcpyAsync(y, x, n * (dtype == TXL_F32 ? 4 : 2), cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
    return TXL_OK;
  }
  int grid = (int)imin64(cdiv64(n, 256), (int64_t)txl_num_sms() * 16);
  DISPATCH_DTYPE(dtype, (dropout_kernel<T><<<grid, 256, 0, (cudaStream_t)stream>>>((const T*)x, (T*)y, n, drop_p, 1.f / (1.f - drop_p), seed, site)));
  TXL_LAUNCH_CHECK();
  return TXL_OK;
}

// ------------------------------------------------------------------ log-softmax + NLL (one warp per row; V is the small music vocab)
template <typename T>
__global__ void lsm_nll_fwd_kernel(const T* __restrict__ logits, int64_t ldl, const int64_t* __restrict__ labels, float* __restrict__ losses,
                                   float* __restrict__ lse_out, float* __restrict__ logprobs, int64_t* __restrict__ argmax, int64_t N, int V) {
  int lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
  int64_t warp = blockIdx.x * (int64_t)wpb + (threadIdx.x >> 5);
  for (int64_t row = warp; row < N; row += (int64_t)gridDim.x * wpb) {
    const T* l = logits + row * ldl;
    float m = -INFINITY; int am = 0x7fffffff;
    for (int v = lane; v < V; v += 32) {
      float x = to_f32(l[v]);
      if (x > m) { m = x; am = v; }   // first max per lane (v ascending)
    }
    float mm = warp_max(m);
    // lowest index among lanes holding the max (torch.argmax returns the first maximal index)
    int cand = (m == mm) ? am : 0x7fffffff;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) cand = min(cand, __shfl_xor_sync(0xffffffffu, cand, o));
    float s = 0.f;
    for (int v = lane; v < V; v += 32) s += expf(to_f32(l[v]) - mm);
    float lse = mm + logf(warp_sum(s));
    if (logprobs)
      for (int v = lane; v < V; v += 32) logprobs[row * V + v] = to_f32(l[v]) - lse;
    if (lane == 0) {
      if (lse_out) lse_out[row] = lse;
      if (argmax) argmax[row] = cand;
      if (losses) {
        int64_t lab = labels ? labels[row] : -100;
        losses[row] = (lab >= 0 && lab < V) ? lse - to_f32(l[lab]) : 0.f;
      }
    }
  }
}
extern "C" int txl_logsoftmax_nll_fwd(const void* logits, int64_t ldl, const int64_t* labels, float* losses, float* lse,
                                      float* logprobs, int64_t* argmax, int64_t N, int V, int dtype, void* stream) {
  TXL_CHECK_ARG(N > 0 && V > 0 && ldl >= V, "lsm_nll_fwd: bad sizes");
  int grid = (int)imin64(cdiv64(N, 8), (int64_t)txl_num_sms() * 8);
  DISPATCH_DTYPE(dtype, (lsm_nll_fwd_kernel<T><<<grid, 256, 0, (cudaStream_t)stream>>>((const T*)logits, ldl, labels, losses, lse, logprobs, argmax, N, V)));
  TXL_LAUNCH_CHECK();
  return TXL_OK;
}
template <typename T, typename TO>
__global__ void lsm_nll_bwd_kernel(const T* __restrict__ logits, int64_t ldl, TO* __restrict__ dlogits, int64_t ldd, const int64_t* __restrict__ labels,
                                   const float* __restrict__ lse, const float* __restrict__ grow, int64_t N, int V) {
  int lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
  int64_t warp = blockIdx.x * (int64_t)wpb + (threadIdx.x >> 5);
  for (int64_t row = warp; row < N; row += (int64_t)gridDim.x * wpb) {
    const T* l = logits + row * ldl;
    TO* o = dlogits + row * ldd;
    int64_t lab = labels[row];
    bool valid = lab >= 0 && lab < V;
    float g = valid ? grow[row] : 0.f, ls = lse[row];
    for (int v = lane; v < ldd; v += 32) {
      float out = 0.f;
      if (v < V && g != 0.f) out = (expf(to_f32(l[v]) - ls) - (v == lab ? 1.f : 0.f)) * g;
      o[v] = from_f32<TO>(out);
    }
  }
}
extern "C" int txl_logsoftmax_nll_bwd(const void* logits, int64_t ldl, int dtype, void* dlogits, int64_t ldd, int dtype_out, const int64_t* labels,
                                      const float* lse, const float* grow, int64_t N, int V, void* stream) {
  TXL_CHECK_ARG(N > 0 && V > 0 && ldl >= V && ldd >= V && logits && dlogits, "lsm_nll_bwd: bad sizes");
  TXL_CHECK_ARG(logits != dlogits || (dtype == dtype_out && ldl == ldd), "lsm_nll_bwd: in-place needs identical dtype and pitch");
  int grid = (int)imin64(cdiv64(N, 8), (int64_t)txl_num_sms() * 8);
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == TXL_F32 && dtype_out == TXL_F32) lsm_nll_bwd_kernel<float, float><<<grid, 256, 0, st>>>((const float*)logits, ldl, (float*)dlogits, ldd, labels, lse, grow, N, V);
  else if (dtype == TXL_F32 && dtype_out == TXL_BF16) lsm_nll_bwd_kernel<float, bf16><<<grid, 256, 0, st>>>((const float*)logits, ldl, (bf16*)dlogits, ldd, labels, lse, grow, N, V);
  else if (dtype == TXL_BF16 && dtype_out == TXL_BF16) lsm_nll_bwd_kernel<bf16, bf16><<<grid, 256, 0, st>>>((const bf16*)logits, ldl, (bf16*)dlogits, ldd, labels, lse, grow, N, V);
  else { txl_set_error("lsm_nll_bwd: unsupported dtype pair"); return TXL_EINVAL; }
  TXL_LAUNCH_CHECK();
  return TXL_OK;
}

// ------------------------------------------------------------------ adaptive softmax, cluster path (HF ProjectedAdaptiveLogSoftmax, n_clusters > 0)
// A logits row holds the V token logits (columns 0..V-1, over the tied embedding) followed by the n_clusters cluster logits (columns V..V+nc-1,
// over crit.cluster_weight): ONE GEMM over the [V+nc, d] extended matrix produces it.  Head = tokens [0, c0) + the cluster columns (HF's head
// index of tail i is c0 + i - 1 = column V + i - 1 here); tail i = tokens [c_{i-1}, c_i).  log p(v) = head log-softmax for v < c0, else
// head log-prob of the tail's cluster column + tail log-softmax.  One warp per row; lse[row][0] = head, lse[row][i] = tail i.
struct Cuts { int nc; int c[5]; };   // c[0] = shortlist size, ..., c[nc] = V

__device__ __forceinline__ int cluster_of(const Cuts& q, int v) {
  int k = 0;
#pragma unroll
  for (int i = 0; i < 4; ++i) k += (i < q.nc && v >= q.c[i]) ? 1 : 0;
  return k;
}

template <typename T>
__global__ void adaptive_lsm_fwd_kernel(const T* __restrict__ logits, int64_t ldl, const int64_t* __restrict__ labels, float* __restrict__ losses,
                                        float* __restrict__ lse_out, float* __restrict__ logprobs, int64_t* __restrict__ argmax, int64_t N, int V,
                                        const Cuts q) {
  const int lane = threadIdx.x & 31, wpb = blockDim.x >> 5, nc = q.nc;
  const int64_t warp = blockIdx.x * (int64_t)wpb + (threadIdx.x >> 5);
  for (int64_t row = warp; row < N; row += (int64_t)gridDim.x * wpb) {
    const T* l = logits + row * ldl;
    float lse[5], clp[5];     // clp[i] = head log-prob of tail i's cluster column (clp[0] = 0)
    {
      float m = -INFINITY;
      for (int v = lane; v < q.c[0]; v += 32) m = fmaxf(m, to_f32(l[v]));
      if (lane < nc) m = fmaxf(m, to_f32(l[V + lane]));
      m = warp_max(m);
      float s = 0.f;
      for (int v = lane; v < q.c[0]; v += 32) s += expf(to_f32(l[v]) - m);
      if (lane < nc) s += expf(to_f32(l[V + lane]) - m);
      lse[0] = m + logf(warp_sum(s));
      clp[0] = 0.f;
    }
    for (int i = 1; i <= nc; ++i) {
      float m = -INFINITY;
      for (int v = q.c[i - 1] + lane; v < q.c[i]; v += 32) m = fmaxf(m, to_f32(l[v]));
      m = warp_max(m);
      float s = 0.f;
      for (int v = q.c[i - 1] + lane; v < q.c[i]; v += 32) s += expf(to_f32(l[v]) - m);
      lse[i] = m + logf(warp_sum(s));
      clp[i] = to_f32(l[V + i - 1]) - lse[0];
    }
    if (logprobs || argmax) {
      float best = -INFINITY; int am = 0x7fffffff;
      for (int v = lane; v < V; v += 32) {
        const int k = cluster_of(q, v);
        const float lp = clp[k] + (to_f32(l[v]) - lse[k]);
        if (logprobs) logprobs[row * V + v] = lp;
        if (lp > best) { best = lp; am = v; }
      }
      if (argmax) {
        const float bb = warp_max(best);
        int cand = (best == bb) ? am : 0x7fffffff;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) cand = min(cand, __shfl_xor_sync(0xffffffffu, cand, o));
        if (lane == 0) argmax[row] = cand;
      }
    }
    if (lane == 0) {
      if (lse_out)
        for (int i = 0; i <= nc; ++i) lse_out[row * (nc + 1) + i] = lse[i];
      if (losses) {
        const int64_t lab = labels ? labels[row] : -100;
        float out = 0.f;
        if (lab >= 0 && lab < V) {
          const int k = cluster_of(q, (int)lab);
          out = -(clp[k] + (to_f32(l[lab]) - lse[k]));
        }
        losses[row] = out;
      }
    }
  }
}

// d(-log p(label)) / d logits, times grow[row]:  head columns softmax_head - [v == label or v == label's cluster column];
// the label's tail: softmax_tail - [v == label]; other tails 0.
template <typename T, typename TO>
__global__ void adaptive_lsm_bwd_kernel(const T* __restrict__ logits, int64_t ldl, TO* __restrict__ dlogits, int64_t ldd, const int64_t* __restrict__ labels,
                                        const float* __restrict__ lse, const float* __restrict__ grow, int64_t N, int V, const Cuts q) {
  const int lane = threadIdx.x & 31, wpb = blockDim.x >> 5, nc = q.nc;
  const int64_t warp = blockIdx.x * (int64_t)wpb + (threadIdx.x >> 5);
  for (int64_t row = warp; row < N; row += (int64_t)gridDim.x * wpb) {
    const T* l = logits + row * ldl;
    TO* o = dlogits + row * ldd;
    const int64_t lab = labels[row];
    const bool valid = lab >= 0 && lab < V;
    const float g = valid ? grow[row] : 0.f;
    const int kl = valid ? cluster_of(q, (int)lab) : 0;
    const float lse_h = lse[row * (nc + 1)], lse_t = lse[row * (nc + 1) + kl];
    const int hot_head = kl == 0 ? (int)lab : V + kl - 1;
    for (int v = lane; v < ldd; v += 32) {
      float out = 0.f;
      if (g != 0.f && v < V + nc) {
        if (v < q.c[0] || v >= V) out = (expf(to_f32(l[v]) - lse_h) - (v == hot_head ? 1.f : 0.f)) * g;
        else if (kl > 0 && v >= q.c[kl - 1] && v < q.c[kl]) out = (expf(to_f32(l[v]) - lse_t) - (v == lab ? 1.f : 0.f)) * g;
      }
      o[v] = from_f32<TO>(out);
    }
  }
}

// HF builds the criterion with keep_order=False: the loss vector it returns is PACKED - the (b, t < T-1) positions whose label falls in
// cluster 0 first (in position order), then cluster 1, ..., ignored labels as trailing zeros (`out[offset : offset + n_i]`).  Stable
// partition by one CTA: per cluster, chunks of blockDim positions, ballot + prefix counts.  perm[k] = b*T + t of packed entry k (-1: none).
__global__ void pack_losses_kernel(const float* __restrict__ pos_losses, const int64_t* __restrict__ labels_shift, int B, int T, int V, const Cuts q,
                                   float* __restrict__ packed, int64_t* __restrict__ perm) {
  __shared__ int wcnt[32];
  __shared__ int base;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nw = blockDim.x >> 5;
  const int64_t n = (int64_t)B * (T - 1);
  if (tid == 0) base = 0;
  __syncthreads();
  for (int k = 0; k <= q.nc; ++k) {
    for (int64_t i0 = 0; i0 < n; i0 += blockDim.x) {
      const int64_t i = i0 + tid;
      bool hit = false; int64_t src = 0;
      if (i < n) {
        src = (i / (T - 1)) * T + (i % (T - 1));
        const int64_t lab = labels_shift[src];
        hit = lab >= 0 && lab < V && cluster_of(q, (int)lab) == k;
      }
      const unsigned m = __ballot_sync(0xffffffffu, hit);
      if (lane == 0) wcnt[warp] = __popc(m);
      __syncthreads();
      int before = 0, total = 0;
      for (int w = 0; w < nw; ++w) { if (w < warp) before += wcnt[w]; total += wcnt[w]; }
      if (hit) {
        const int dst = base + before + __popc(m & ((1u << lane) - 1u));
        packed[dst] = pos_losses[src];
        if (perm) perm[dst] = src;
      }
      __syncthreads();
      if (tid == 0) base += total;
      __syncthreads();
    }
  }
  for (int64_t i = base + tid; i < n; i += blockDim.x) { packed[i] = 0.f; if (perm) perm[i] = -1; }
}

static int make_cuts(Cuts& q, int V, int nc, const int* cutoffs, const char* who) {
  if (!(nc >= 1 && nc <= 4 && cutoffs)) { txl_set_error("%s: 1 <= n_clusters <= 4 and cutoffs required", who); return TXL_EINVAL; }
  q.nc = nc;
  int prev = 0;
  for (int i = 0; i < nc; ++i) {
    if (!(cutoffs[i] > prev && cutoffs[i] < V)) { txl_set_error("%s: cutoffs must be increasing inside (0, V)", who); return TXL_EINVAL; }
    q.c[i] = prev = cutoffs[i];
  }
  for (int i = nc; i < 5; ++i) q.c[i] = V;
  return TXL_OK;
}

extern "C" int txl_adaptive_lsm_nll_fwd(const void* logits, int64_t ldl, const int64_t* labels, float* losses, float* lse, float* logprobs,
                                        int64_t* argmax, int64_t N, int V, int n_clusters, const int* cutoffs, int dtype, void* stream) {
  TXL_CHECK_ARG(logits && N > 0 && V > 0 && ldl >= V + n_clusters, "adaptive_lsm_nll_fwd: bad sizes");
  Cuts q; int rc = make_cuts(q, V, n_clusters, cutoffs, "adaptive_lsm_nll_fwd"); if (rc) return rc;
  const int grid = (int)imin64(cdiv64(N, 8), (int64_t)txl_num_sms() * 8);
  DISPATCH_DTYPE(dtype, (adaptive_lsm_fwd_kernel<T><<<grid, 256, 0, (cudaStream_t)stream>>>((const T*)logits, ldl, labels, losses, lse, logprobs, argmax, N, V, q)));
  TXL_LAUNCH_CHECK();
  return TXL_OK;
}
extern "C" int txl_adaptive_lsm_nll_bwd(const void* logits, int64_t ldl, int dtype, void* dlogits, int64_t ldd, int dtype_out, const int64_t* labels,
                                        const float* lse, const float* grow, int64_t N, int V, int n_clusters, const int* cutoffs, void* stream) {
  TXL_CHECK_ARG(logits && dlogits && labels && lse && grow && N > 0 && V > 0 && ldl >= V + n_clusters && ldd >= V + n_clusters, "adaptive_lsm_nll_bwd: bad sizes");
  TXL_CHECK_ARG(logits != dlogits || (dtype == dtype_out && ldl == ldd), "adaptive_lsm_nll_bwd: in-place needs identical dtype and pitch");
  Cuts q; int rc = make_cuts(q, V, n_clusters, cutoffs, "adaptive_lsm_nll_bwd"); if (rc) return rc;
  const int grid = (int)imin64(cdiv64(N, 8), (int64_t)txl_num_sms() * 8);
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == TXL_F32 && dtype_out == TXL_F32) adaptive_lsm_bwd_kernel<float, float><<<grid, 256, 0, st>>>((const float*)logits, ldl, (float*)dlogits, ldd, labels, lse, grow, N, V, q);
  else if (dtype == TXL_F32 && dtype_out == TXL_BF16) adaptive_lsm_bwd_kernel<float, bf16><<<grid, 256, 0, st>>>((const float*)logits, ldl, (bf16*)dlogits, ldd, labels, lse, grow, N, V, q);
  else { txl_set_error("adaptive_lsm_nll_bwd: unsupported dtype pair"); return TXL_EINVAL; }
  TXL_LAUNCH_CHECK();
  return TXL_OK;
}
extern "C" int txl_pack_losses(const float* pos_losses, const int64_t* labels_shift, int B, int T, int V, int n_clusters, const int* cutoffs,
                               float* packed, int64_t* perm, void* stream) {
  TXL_CHECK_ARG(pos_losses && labels_shift && packed && B > 0 && T > 1 && V > 0, "pack_losses: bad args");
  Cuts q; int rc = make_cuts(q, V, n_clusters, cutoffs, "pack_losses"); if (rc) return rc;
  pack_losses_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(pos_losses, labels_shift, B, T, V, q, packed, perm);
  TXL_LAUNCH_CHECK();
  return TXL_OK;
}

// loss = mean(losses[losses != 0])  — single block, deterministic tree order
__global__ void masked_mean_kernel(const float* __restrict__ losses, int64_t N, float* loss_out, float* count_out) {
  __shared__ double ssum[32];
  __shared__ double scnt[32];
  double s = 0.0, c = 0.0;
  for (int64_t i = threadIdx.x; i < N; i += blockDim.x) {
    float v = losses[i];
    if (v != 0.f) { s += (double)v; c += 1.0; }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) { s += __shfl_xor_sync(0xffffffffu, s, o); c += __shfl_xor_sync(0xffffffffu, c, o); }
  if ((threadIdx.x & 31) == 0) { ssum[threadIdx.x >> 5] = s; scnt[threadIdx.x >> 5] = c; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double S = 0.0, C = 0.0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) { S += ssum[w]; C += scnt[w]; }
    *loss_out = (float)(S / C);   // C == 0 -> NaN, as torch's empty mean
    if (count_out) *count_out = (float)C;
  }
}
extern "C" int txl_masked_mean(const float* losses, int64_t N, float* loss_out, float* count_out, void* stream) {
  TXL_CHECK_ARG(N > 0, "masked_mean: N");
  masked_mean_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(losses, N, loss_out, count_out);
  TXL_LAUNCH_CHECK();
  return TXL_OK;
}

// next-token-prediction accuracy counts: position t predicts token t+1; pads (-100) are not counted
__global__ void ntp_acc_kernel(const int64_t* __restrict__ preds, int64_t ldp, const int64_t* __restrict__ labels, int64_t ldl, int B, int T,
                               unsigned long long* __restrict__ out) {
  unsigned long long hit = 0, cnt = 0;
  const int64_t total = (int64_t)B * (T - 1);
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t b = i / (T - 1), t = i % (T - 1);
    const int64_t lab = labels[b * ldl + t + 1];
    if (lab != -100) { ++cnt; hit += (preds[b * ldp + t] == lab) ? 1ull : 0ull; }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) { hit += __shfl_xor_sync(0xffffffffu, hit, o); cnt += __shfl_xor_sync(0xffffffffu, cnt, o); }
  if ((threadIdx.x & 31) == 0 && cnt) { atomicAdd(&out[0], hit); atomicAdd(&out[1], cnt); }
}
extern "C" int txl_ntp_acc(const int64_t* preds, int64_t ld_preds, const int64_t* labels, int64_t ld_labels, int B, int T, int64_t* out, void* stream) {
  TXL_CHECK_ARG(preds && labels && out && B > 0 && T > 0 && ld_preds >= T && ld_labels >= T, "ntp_acc: bad args");
  if (T < 2) return TXL_OK;
  const int grid = (int)imin64(cdiv64((int64_t)B * (T - 1), 256), (int64_t)txl_num_sms() * 4);
  ntp_acc_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(preds, ld_preds, labels, ld_labels, B, T, reinterpret_cast<unsigned long long*>(out));
  TXL_LAUNCH_CHECK();
  return TXL_OK;
}

// ------------------------------------------------------------------ formats either side of the path
// DataCollatorForLanguageModeling(mlm=False): labels = input_ids with every pad id replaced by -100
__global__ void clm_labels_kernel(const int64_t* __restrict__ ids, int64_t* __restrict__ labels, int64_t n, int64_t pad) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t t = ids[i];
    labels[i] = t == pad ? -100 : t;
  }
}
extern "C" int txl_clm_labels(const int64_t* ids, int64_t* labels, int64_t n, int64_t pad_id, void* stream) {
  TXL_CHECK_ARG(ids && labels && n > 0, "clm_labels: bad args");
  const int grid = (int)imin64(cdiv64(n, 256), (int64_t)txl_num_sms() * 8);
  clm_labels_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(ids, labels, n, pad_id);
  TXL_LAUNCH_CHECK();
  return TXL_OK;
}
// labels [B, T] (row stride ld) -> shift[b*T + t] = labels[b, t+1] (t < T-1), -100 in the last column: what HF's crit sees as `labels[..., 1:]`.
// Block 0 first applies the reference's fix-up (musicnlp/models/transformer_xl.py:176-182), with its own arithmetic: if
// sum(labels[0, 1:]) == (T-1) * -100 then labels[0, 1] = eos, IN PLACE (the caller's tensor is mutated as in the reference), no host sync.
// Also counts labels outside [0, V) that are not -100 into bad[0] (optional range check, HF raises on those).
__global__ void shift_labels_kernel(int64_t* __restrict__ labels, int64_t ld, int B, int T, int64_t eos, int V, int64_t* __restrict__ shift,
                                    int* __restrict__ bad) {
  __shared__ long long s_sum[32];
  __shared__ int s_fix;
  const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  int64_t* row = labels + (int64_t)b * ld;
  if (b == 0) {
    long long acc = 0;
    for (int t = 1 + tid; t < T; t += blockDim.x) acc += row[t];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) s_sum[warp] = acc;
    __syncthreads();
    if (tid == 0) {
      long long tot = 0;
      for (int w = 0; w < (int)(blockDim.x >> 5); ++w) tot += s_sum[w];
      s_fix = (T > 1 && tot == (long long)(T - 1) * -100) ? 1 : 0;
      if (s_fix) row[1] = eos;
    }
    __syncthreads();
  }
  int nbad = 0;
  for (int t = tid; t < T; t += blockDim.x) {
    int64_t v = -100;
    if (t + 1 < T) {
      v = (b == 0 && t == 0 && s_fix) ? eos : row[t + 1];
      if (v != -100 && (v < 0 || v >= V)) ++nbad;
    }
    shift[(int64_t)b * T + t] = v;
  }
  if (bad && nbad) atomicAdd(bad, nbad);
}
extern "C" int txl_shift_labels(int64_t* labels, int64_t ld, int B, int T, int64_t eos, int V, int64_t* shift, int* bad, void* stream) {
  TXL_CHECK_ARG(labels && shift && B > 0 && T > 0 && ld >= T && V > 0, "shift_labels: bad args");
  shift_labels_kernel<<<B, 256, 0, (cudaStream_t)stream>>>(labels, ld, B, T, eos, V, shift, bad);
  TXL_LAUNCH_CHECK();
  return TXL_OK;
}
// last position of `token` in every row (-1 if absent): one warp per row, scanning from the end
__global__ void last_index_kernel(const int64_t* __restrict__ ids, int64_t ld, int B, int T, int64_t token, int64_t* __restrict__ out) {
  const int lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
  for (int b = blockIdx.x * wpb + (threadIdx.x >> 5); b < B; b += gridDim.x * wpb) {
    int64_t found = -1;
    for (int base = ((T - 1) / 32) * 32; base >= 0 && found < 0; base -= 32) {
      const int t = base + lane;
      const bool hit = t < T && ids[(int64_t)b * ld + t] == token;
      const unsigned m = __ballot_sync(0xffffffffu, hit);
      if (m) found = base + (31 - __clz((int)m));
    }
    if (lane == 0) out[b] = found;
  }
}
extern "C" int txl_last_index_of(const int64_t* ids, int64_t ld, int B, int T, int64_t token, int64_t* out, void* stream) {
  TXL_CHECK_ARG(ids && out && B > 0 && T > 0 && ld >= T, "last_index_of: bad args");
  const int grid = (int)imin64(cdiv64(B, 8), (int64_t)txl_num_sms() * 4);
  last_index_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(ids, ld, B, T, token, out);
  TXL_LAUNCH_CHECK();
  return TXL_OK;
}

// ------------------------------------------------------------------ casts / transposes / layout
__global__ void cast_f2b_kernel(const float* __restrict__ s, bf16* __restrict__ d, int64_t n) {
  int64_t i = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) * 4;
  int64_t stride = (int64_t)gridDim.x * blockDim.x * 4;
  for (; i + 3 < n; i += stride) {
    float4 v = *reinterpret_cast<const float4*>(s + i);
    __nv_bfloat162 a = __floats2bfloat162_rn(v.x, v.y), b = __floats2bfloat162_rn(v.z, v.w);
    uint2 o; o.x = *reinterpret_cast<uint32_t*>(&a); o.y = *reinterpret_cast<uint32_t*>(&b);
    *reinterpret_cast<uint2*>(d + i) = o;
  }
  if (blockIdx.x == 0 && threadIdx.x == 0)
    for (int64_t t = n & ~(int64_t)3; t < n; ++t) d[t] = __float2bfloat16_rn(s[t]);
}
__global__ void cast_b2f_kernel(const bf16* __restrict__ s, float* __restrict__ d, int64_t n) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) d[i] = __bfloat162float(s[i]);
}
extern "C" int txl_cast_f32_to_bf16(const float* src, void* dst, int64_t n, void* stream) {
  TXL_CHECK_ARG(n > 0 && ((uintptr_t)src % 16 == 0) && ((uintptr_t)dst % 8 == 0), "cast: alignment");
  int grid = (int)imin64(cdiv64(n, 1024), (int64_t)txl_num_sms() * 16);
  cast_f2b_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(src, (bf16*)dst, n);
  TXL_LAUNCH_CHECK();
  return TXL_OK;
}
extern "C" int txl_cast_bf16_to_f32(const void* src, float* dst, int64_t n, void* stream) {
  TXL_CHECK_ARG(n > 0, "cast: n");
  int grid = (int)imin64(cdiv64(n, 256), (int64_t)txl_num_sms() * 16);
  cast_b2f_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>((const bf16*)src, dst, n);
  TXL_LAUNCH_CHECK();
  return TXL_OK;
}
template <typename T>
__global__ void transpose_kernel(const T* __restrict__ A, T* __restrict__ B, int64_t rows, int64_t cols) {
  __shared__ T tile[32][33];
  int64_t c0 = blockIdx.x * 32ll, r0 = blockIdx.y * 32ll;
  for (int dy = threadIdx.y; dy < 32; dy += blockDim.y) {
    int64_t r = r0 + dy, c = c0 + threadIdx.x;
    if (r < rows && c < cols) tile[dy][threadIdx.x] = A[r * cols + c];
  }
  __syncthreads();
  for (int dy = threadIdx.y; dy < 32; dy += blockDim.y) {
    int64_t c = c0 + dy, r = r0 + threadIdx.x;
    if (r < rows && c < cols) B[c * rows + r] = tile[threadIdx.x][dy];
  }
}
extern "C" int txl_transpose(const void* A, void* B, int64_t rows, int64_t cols, int dtype, void* stream) {
  TXL_CHECK_ARG(rows > 0 && cols > 0, "transpose: sizes");
  dim3 grid((unsigned)cdiv64(cols, 32), (unsigned)cdiv64(rows, 32)), block(32, 8);
  DISPATCH_DTYPE(dtype, (transpose_kernel<T><<<grid, block, 0, (cudaStream_t)stream>>>((const T*)A, (T*)B, rows, cols)));
  TXL_LAUNCH_CHECK();
  return TXL_OK;
}
// (rows,B,d) time-major <-> (B,rows,d) batch-major, with dtype conversion
template <typename TS, typename TD>
__global__ void tm_bm_kernel(const TS* __restrict__ src, TD* __restrict__ dst, int rows, int B, int d, int to_bm) {
  int64_t total = (int64_t)rows * B * d;
  for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    int c = (int)(idx % d); int64_t rb = idx / d;
    if (to_bm) {  // dst index is (b, r, c)
      int r = (int)(rb % rows), b = (int)(rb / rows);
      dst[idx] = from_f32<TD>(to_f32(src[((int64_t)r * B + b) * d + c]));
    } else {      // dst index is (r, b, c)
      int b = (int)(rb % B), r = (int)(rb / B);
      dst[idx] = from_f32<TD>(to_f32(src[((int64_t)b * rows + r) * d + c]));
    }
  }
}
static int tm_bm(const void* src, void* dst, int rows, int B, int d, int ds, int dd, int to_bm, void* stream) {
  TXL_CHECK_ARG(rows > 0 && B > 0 && d > 0, "tm_bm: sizes");
  int grid = (int)imin64(cdiv64((int64_t)rows * B * d, 256), (int64_t)txl_num_sms() * 16);
  cudaStream_t st = (cudaStream_t)stream;
  if (ds == TXL_F32 && dd == TXL_F32) tm_bm_kernel<float, float><<<grid, 256, 0, st>>>((const float*)src, (float*)dst, rows, B, d, to_bm);
  else if (ds == TXL_F32 && dd == TXL_BF16) tm_bm_kernel<float, bf16><<<grid, 256, 0, st>>>((const float*)src, (bf16*)dst, rows, B, d, to_bm);
  else if (ds == TXL_BF16 && dd == TXL_F32) tm_bm_kernel<bf16, float><<<grid, 256, 0, st>>>((const bf16*)src, (float*)dst, rows, B, d, to_bm);
  else if (ds == TXL_BF16 && dd == TXL_BF16) tm_bm_kernel<bf16, bf16><<<grid, 256, 0, st>>>((const bf16*)src, (bf16*)dst, rows, B, d, to_bm);
  else { txl_set_error("tm_bm: bad dtype"); return TXL_EINVAL; }
  TXL_LAUNCH_CHECK();
  return TXL_OK;
}
extern "C" int txl_tm_to_bm(const void* src, void* dst, int rows, int B, int d, int ds, int dd, void* stream) { return tm_bm(src, dst, rows, B, d, ds, dd, 1, stream); }
extern "C" int txl_bm_to_tm(const void* src, void* dst, int rows, int B, int d, int ds, int dd, void* stream) { return tm_bm(src, dst, rows, B, d, ds, dd, 0, stream); }

// ------------------------------------------------------------------ optimiser
__global__ void sumsq_kernel(const float* __restrict__ g, int64_t n, float* out) {
  float s = 0.f;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) s += g[i] * g[i];
  s = warp_sum(s);
  __shared__ float sm[32];
  if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x < 32) {
    float v = threadIdx.x < (blockDim.x >> 5) ? sm[threadIdx.x] : 0.f;
    v = warp_sum(v);
    if (threadIdx.x == 0) atomicAdd(out, v);
  }
}
extern "C" int txl_sumsq(const float* g, int64_t n, float* out, void* stream) {
  TXL_CHECK_ARG(n > 0, "sumsq: n");
  int grid = (int)imin64(cdiv64(n, 1024), (int64_t)txl_num_sms() * 4);
  sumsq_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(g, n, out);
  TXL_LAUNCH_CHECK();
  return TXL_OK;
}
__global__ void adamw_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
                             const uint8_t* __restrict__ decay_mask, int64_t n, float lr, float b1, float b2, float eps, float wd,
                             float bc1, float bc2, const float* __restrict__ grad_scale_dev, bf16* __restrict__ shadow) {
  float gs = grad_scale_dev ? *grad_scale_dev : 1.f;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    float gi = g[i] * gs, pi = p[i];
    if (!decay_mask || decay_mask[i]) pi *= (1.f - lr * wd);          // torch.optim.AdamW: decoupled decay first
    float mi = b1 * m[i] + (1.f - b1) * gi;
    float vi = b2 * v[i] + (1.f - b2) * gi * gi;
    m[i] = mi; v[i] = vi;
    float denom = sqrtf(vi) / sqrtf(bc2) + eps;
    pi -= (lr / bc1) * mi / denom;
    p[i] = pi;
    if (shadow) shadow[i] = __float2bfloat16_rn(pi);
  }
}
extern "C" int txl_adamw_step(float* p, const float* g, float* m, float* v, const uint8_t* decay_mask, int64_t n, float lr,
                              float beta1, float beta2, float eps, float weight_decay, int step, const float* grad_scale_dev,
                              void* bf16_shadow, void* stream) {
  TXL_CHECK_ARG(n > 0 && step >= 1, "adamw: bad args");
  float bc1 = 1.f - powf(beta1, (float)step), bc2 = 1.f - powf(beta2, (float)step);
  int grid = (int)imin64(cdiv64(n, 256), (int64_t)txl_num_sms() * 16);
  adamw_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(p, g, m, v, decay_mask, n, lr, beta1, beta2, eps, weight_decay, bc1, bc2, grad_scale_dev, (bf16*)bf16_shadow);
  TXL_LAUNCH_CHECK();
  return TXL_OK;
}
