// simt_relattn.cu — relative-position band attention, fp32-FMA kernels (fp32 parity mode, odd head sizes, T=1 decode
// fallback).  Fuses HF's AC + BD einsums, `_rel_shift`, the same_length band mask, softmax and P.V; nothing of size
// T x klen is ever materialised.  [A.3 steps 2-8, A.4, A.5]
#include "common.cuh"

int txl_relattn_fwd_tc(const void* q, const void* k_mem, const void* v_mem, const void* k_cur, const void* v_cur, const void* r,
                       const float* rwb, const float* rrb, void* out, float* lse, void* saved, const TxlAttnDims* dims, void* stream, int* handled);
int64_t txl_relattn_saved_bytes_tc(const TxlAttnDims* D);

int64_t txl_relattn_bwd_tc_workspace(const TxlAttnDims* D);
int txl_relattn_bwd_tc(const void* q, const void* k_mem, const void* v_mem, const void* k_cur, const void* v_cur, const void* r, const float* rwb,
                       const float* rrb, const void* out, const float* lse, const void* dout, void* dq, void* dk_mem, void* dv_mem, void* dk_cur,
                       void* dv_cur, float* dr, float* drwb, float* drrb, void* ws, const void* saved, const TxlAttnDims* D, void* stream, int* handled);

namespace {
constexpr int BQ = 32;   // query rows per CTA
constexpr int BKC = 32;  // keys per chunk (= one lane per key)
constexpr int RPW = 8;   // rows per warp (4 warps)

struct AttnPtrs {
  const void *q, *k_mem, *v_mem, *k_cur, *v_cur, *r;
  const float *rwb, *rrb;
};

template <typename T, int DH>
struct Smem {
  float Qw[BQ][DH];
  float Qr[BQ][DH];
  float Ks[BKC][DH + 1];
  float Vs[BKC][DH + 1];
  float Rs[BQ + BKC - 1][DH + 1];
};

// stage q (+biases), and per chunk K/V/R tiles
template <typename T, int DH>
__device__ __forceinline__ void load_q(Smem<T, DH>& s, const AttnPtrs& P, const TxlAttnDims& D, const BandGeom& g, int b, int h, int i0) {
  for (int e = threadIdx.x; e < BQ * DH; e += blockDim.x) {
    int rr = e / DH, c = e % DH, i = i0 + rr;
    float qv = 0.f;
    if (i < g.T) qv = to_f32(((const T*)P.q)[((int64_t)b * g.T + i) * D.ldq + h * DH + c]);
    s.Qw[rr][c] = qv + P.rwb[h * DH + c];
    s.Qr[rr][c] = qv + P.rrb[h * DH + c];
  }
}
template <typename T, int DH>
__device__ __forceinline__ void load_chunk(Smem<T, DH>& s, const AttnPtrs& P, const TxlAttnDims& D, const BandGeom& g, int b, int h,
                                           int i0, int jc) {
  for (int e = threadIdx.x; e < BKC * DH; e += blockDim.x) {
    int jj = e / DH, c = e % DH, j = jc + jj;
    float kv = 0.f, vv = 0.f;
    if (j < g.klen) {
      if (j < g.mlen) {
        int64_t off = ((int64_t)b * g.mlen + j) * D.ldkv_mem + h * DH + c;
        kv = to_f32(((const T*)P.k_mem)[off]); vv = to_f32(((const T*)P.v_mem)[off]);
      } else {
        int64_t off = ((int64_t)b * g.T + (j - g.mlen)) * D.ldkv_cur + h * DH + c;
        kv = to_f32(((const T*)P.k_cur)[off]); vv = to_f32(((const T*)P.v_cur)[off]);
      }
    }
    s.Ks[jj][c] = kv; s.Vs[jj][c] = vv;
  }
  // window w <-> distance p = mlen + i0 - jc - (BKC-1) + w   (w = rr - jj + BKC-1)
  // r row of distance p is x = klen-1-p (HF r_head_k order; the clamp is baked into the table)
  int pbase = g.mlen + i0 - jc - (BKC - 1);
  for (int e = threadIdx.x; e < (BQ + BKC - 1) * DH; e += blockDim.x) {
    int w = e / DH, c = e % DH, x = g.klen - 1 - (pbase + w);
    float rv = 0.f;
    if (x >= 0 && x < g.klen) rv = to_f32(((const T*)P.r)[(int64_t)x * D.H * DH + h * DH + c]);
    s.Rs[w][c] = rv;
  }
}

template <typename T, int DH>
__global__ void __launch_bounds__(128) relattn_fwd_kernel(AttnPtrs P, T* __restrict__ out, float* __restrict__ lse, TxlAttnDims D) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  Smem<T, DH>& s = *reinterpret_cast<Smem<T, DH>*>(smem_raw);
  const BandGeom g = make_band(D.band);
  const int i0 = blockIdx.x * BQ, h = blockIdx.y, b = blockIdx.z;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  constexpr int DPL = (DH + 31) / 32;
  const float scale = rsqrtf((float)DH);

  load_q<T, DH>(s, P, D, g, b, h, i0);
  float m[RPW], l[RPW], acc[RPW][DPL];
#pragma unroll
  for (int t = 0; t < RPW; ++t) { m[t] = -INFINITY; l[t] = 0.f;
#pragma unroll
    for (int dd = 0; dd < DPL; ++dd) acc[t][dd] = 0.f; }

  const int ilast = min(i0 + BQ - 1, g.T - 1);
  const int jlo = band_lo(g, i0), jhi = min(band_hi(g, ilast), g.klen - 1);
  for (int jc = (jlo / BKC) * BKC; jc <= jhi; jc += BKC) {
    __syncthreads();
    load_chunk<T, DH>(s, P, D, g, b, h, i0, jc);
    __syncthreads();
    float ac[RPW], bd[RPW];
#pragma unroll
    for (int t = 0; t < RPW; ++t) { ac[t] = 0.f; bd[t] = 0.f; }
#pragma unroll 4
    for (int c = 0; c < DH; ++c) {
      float kd = s.Ks[lane][c];
#pragma unroll
      for (int t = 0; t < RPW; ++t) {
        int rr = warp * RPW + t;
        ac[t] = fmaf(s.Qw[rr][c], kd, ac[t]);
        bd[t] = fmaf(s.Qr[rr][c], s.Rs[rr - lane + BKC - 1][c], bd[t]);
      }
    }
    const int j = jc + lane;
#pragma unroll
    for (int t = 0; t < RPW; ++t) {
      int rr = warp * RPW + t, i = i0 + rr;
      bool valid = i < g.T && j < g.klen && j >= band_lo(g, i) && j <= band_hi(g, i);
      float sc = valid ? (ac[t] + bd[t]) * scale : -INFINITY;
      float mn = fmaxf(m[t], warp_max(sc));
      if (mn == -INFINITY) continue;  // warp-uniform
      float p = valid ? __expf(sc - mn) : 0.f;
      float corr = __expf(m[t] - mn);  // m=-inf -> 0
      l[t] = l[t] * corr + warp_sum(p);
      m[t] = mn;
#pragma unroll
      for (int dd = 0; dd < DPL; ++dd) acc[t][dd] *= corr;
      for (int jj = 0; jj < BKC; ++jj) {
        float pj = __shfl_sync(0xffffffffu, p, jj);
#pragma unroll
        for (int dd = 0; dd < DPL; ++dd) {
          int c = lane + 32 * dd;
          if (c < DH) acc[t][dd] = fmaf(pj, s.Vs[jj][c], acc[t][dd]);
        }
      }
    }
  }
#pragma unroll
  for (int t = 0; t < RPW; ++t) {
    int i = i0 + warp * RPW + t;
    if (i >= g.T) continue;
    float inv = 1.f / l[t];
#pragma unroll
    for (int dd = 0; dd < DPL; ++dd) {
      int c = lane + 32 * dd;
      if (c < DH) out[((int64_t)b * g.T + i) * (D.H * DH) + h * DH + c] = from_f32<T>(acc[t][dd] * inv);
    }
    if (lane == 0) lse[((int64_t)b * D.H + h) * g.T + i] = m[t] + __logf(l[t]);
  }
}

// ------------------------------------------------------------------ backward
template <typename T, int DH>
struct SmemBwd {
  Smem<T, DH> f;
  float dOs[BQ][DH];
  float dKs[BKC][DH + 1];
  float dVs[BKC][DH + 1];
  float dRs[BQ + BKC - 1][DH + 1];
  float delta[BQ];
  float lse[BQ];
};

template <typename T, int DH>
__global__ void __launch_bounds__(128) relattn_bwd_kernel(AttnPtrs P, const T* __restrict__ out, const float* __restrict__ lse,
                                                          const T* __restrict__ dout, T* __restrict__ dq, float* __restrict__ dk_ws,
                                                          float* __restrict__ dv_ws, float* __restrict__ dr, float* __restrict__ drwb,
                                                          float* __restrict__ drrb, TxlAttnDims D) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  SmemBwd<T, DH>& s = *reinterpret_cast<SmemBwd<T, DH>*>(smem_raw);
  const BandGeom g = make_band(D.band);
  const int i0 = blockIdx.x * BQ, h = blockIdx.y, b = blockIdx.z;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  constexpr int DPL = (DH + 31) / 32;
  const float scale = rsqrtf((float)DH);
  const int HD = D.H * DH;

  load_q<T, DH>(s.f, P, D, g, b, h, i0);
  for (int e = threadIdx.x; e < BQ * DH; e += blockDim.x) {
    int rr = e / DH, c = e % DH, i = i0 + rr;
    s.dOs[rr][c] = i < g.T ? to_f32(dout[((int64_t)b * g.T + i) * HD + h * DH + c]) : 0.f;
  }
  __syncthreads();
  // delta_i = sum_c dO[i,c] * O[i,c]
  for (int rr = warp; rr < BQ; rr += 4) {
    int i = i0 + rr;
    float sd = 0.f;
    if (i < g.T)
      for (int c = lane; c < DH; c += 32) sd += s.dOs[rr][c] * to_f32(out[((int64_t)b * g.T + i) * HD + h * DH + c]);
    sd = warp_sum(sd);
    if (lane == 0) { s.delta[rr] = sd; s.lse[rr] = i < g.T ? lse[((int64_t)b * D.H + h) * g.T + i] : 0.f; }
  }
  float dqw[RPW][DPL], dqr[RPW][DPL];
#pragma unroll
  for (int t = 0; t < RPW; ++t)
#pragma unroll
    for (int dd = 0; dd < DPL; ++dd) { dqw[t][dd] = 0.f; dqr[t][dd] = 0.f; }

  const int ilast = min(i0 + BQ - 1, g.T - 1);
  const int jlo = band_lo(g, i0), jhi = min(band_hi(g, ilast), g.klen - 1);
  for (int jc = (jlo / BKC) * BKC; jc <= jhi; jc += BKC) {
    __syncthreads();
    load_chunk<T, DH>(s.f, P, D, g, b, h, i0, jc);
    for (int e = threadIdx.x; e < BKC * (DH + 1); e += blockDim.x) { (&s.dKs[0][0])[e] = 0.f; (&s.dVs[0][0])[e] = 0.f; }
    for (int e = threadIdx.x; e < (BQ + BKC - 1) * (DH + 1); e += blockDim.x) (&s.dRs[0][0])[e] = 0.f;
    __syncthreads();
    float ac[RPW], bd[RPW], dp[RPW];
#pragma unroll
    for (int t = 0; t < RPW; ++t) { ac[t] = 0.f; bd[t] = 0.f; dp[t] = 0.f; }
#pragma unroll 4
    for (int c = 0; c < DH; ++c) {
      float kd = s.f.Ks[lane][c], vd = s.f.Vs[lane][c];
#pragma unroll
      for (int t = 0; t < RPW; ++t) {
        int rr = warp * RPW + t;
        ac[t] = fmaf(s.f.Qw[rr][c], kd, ac[t]);
        bd[t] = fmaf(s.f.Qr[rr][c], s.f.Rs[rr - lane + BKC - 1][c], bd[t]);
        dp[t] = fmaf(s.dOs[rr][c], vd, dp[t]);
      }
    }
    const int j = jc + lane;
    float pr[RPW], ds[RPW];
#pragma unroll
    for (int t = 0; t < RPW; ++t) {
      int rr = warp * RPW + t, i = i0 + rr;
      bool valid = i < g.T && j < g.klen && j >= band_lo(g, i) && j <= band_hi(g, i);
      pr[t] = valid ? __expf((ac[t] + bd[t]) * scale - s.lse[rr]) : 0.f;
      ds[t] = pr[t] * (dp[t] - s.delta[rr]) * scale;
    }
    // dV, dK (per key = per lane), dR (per window row)
    for (int c = 0; c < DH; ++c) {
      float av = 0.f, ak = 0.f;
#pragma unroll
      for (int t = 0; t < RPW; ++t) {
        int rr = warp * RPW + t;
        av = fmaf(pr[t], s.dOs[rr][c], av);
        ak = fmaf(ds[t], s.f.Qw[rr][c], ak);
        atomicAdd(&s.dRs[rr - lane + BKC - 1][c], ds[t] * s.f.Qr[rr][c]);
      }
      atomicAdd(&s.dVs[lane][c], av);
      atomicAdd(&s.dKs[lane][c], ak);
    }
    // dQw, dQr (lane-parallel over the head dim)
#pragma unroll
    for (int t = 0; t < RPW; ++t) {
      int rr = warp * RPW + t;
      for (int jj = 0; jj < BKC; ++jj) {
        float dsj = __shfl_sync(0xffffffffu, ds[t], jj);
#pragma unroll
        for (int dd = 0; dd < DPL; ++dd) {
          int c = lane + 32 * dd;
          if (c < DH) {
            dqw[t][dd] = fmaf(dsj, s.f.Ks[jj][c], dqw[t][dd]);
            dqr[t][dd] = fmaf(dsj, s.f.Rs[rr - jj + BKC - 1][c], dqr[t][dd]);
          }
        }
      }
    }
    __syncthreads();
    // flush chunk accumulators
    for (int e = threadIdx.x; e < BKC * DH; e += blockDim.x) {
      int jj = e / DH, c = e % DH, jg = jc + jj;
      if (jg < g.klen) {
        int64_t off = ((int64_t)b * g.klen + jg) * HD + h * DH + c;
        float kv = s.dKs[jj][c], vv = s.dVs[jj][c];
        if (kv != 0.f) atomicAdd(&dk_ws[off], kv);
        if (vv != 0.f) atomicAdd(&dv_ws[off], vv);
      }
    }
    int pbase = g.mlen + i0 - jc - (BKC - 1);
    for (int e = threadIdx.x; e < (BQ + BKC - 1) * DH; e += blockDim.x) {
      int w = e / DH, c = e % DH, x = g.klen - 1 - (pbase + w);
      float v = s.dRs[w][c];
      if (x >= 0 && x < g.klen && v != 0.f) atomicAdd(&dr[(int64_t)x * HD + h * DH + c], v);
    }
  }
  __syncthreads();
  // dq = dQw + dQr; bias grads: reduce the CTA's rows through smem (reuse dKs/dVs rows 0: [DH])
  for (int e = threadIdx.x; e < DH; e += blockDim.x) { s.dKs[0][e] = 0.f; s.dVs[0][e] = 0.f; }
  __syncthreads();
#pragma unroll
  for (int dd = 0; dd < DPL; ++dd) {
    int c = lane + 32 * dd;
    if (c < DH) {
      float sw = 0.f, sr = 0.f;
#pragma unroll
      for (int t = 0; t < RPW; ++t) {
        int i = i0 + warp * RPW + t;
        if (i < g.T) {
          dq[((int64_t)b * g.T + i) * D.ldq + h * DH + c] = from_f32<T>(dqw[t][dd] + dqr[t][dd]);
          sw += dqw[t][dd]; sr += dqr[t][dd];
        }
      }
      atomicAdd(&s.dKs[0][c], sw);
      atomicAdd(&s.dVs[0][c], sr);
    }
  }
  __syncthreads();
  for (int c = threadIdx.x; c < DH; c += blockDim.x) {
    atomicAdd(&drwb[h * DH + c], s.dKs[0][c]);
    atomicAdd(&drrb[h * DH + c], s.dVs[0][c]);
  }
}

// fp32 [B, klen, HD] accumulators -> strided outputs
template <typename T>
__global__ void scatter_dkv_kernel(const float* __restrict__ dk_ws, const float* __restrict__ dv_ws, T* dk_mem, T* dv_mem, T* dk_cur,
                                   T* dv_cur, int B, int T_, int mlen, int HD, int64_t ld_mem, int64_t ld_cur) {
  int klen = mlen + T_;
  int64_t total = (int64_t)B * klen * HD;
  for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    int c = (int)(idx % HD); int64_t bj = idx / HD; int j = (int)(bj % klen), b = (int)(bj / klen);
    if (j < mlen) {
      if (dk_mem) {
        int64_t off = ((int64_t)b * mlen + j) * ld_mem + c;
        dk_mem[off] = from_f32<T>(dk_ws[idx]); dv_mem[off] = from_f32<T>(dv_ws[idx]);
      }
    } else {
      int64_t off = ((int64_t)b * T_ + (j - mlen)) * ld_cur + c;
      dk_cur[off] = from_f32<T>(dk_ws[idx]); dv_cur[off] = from_f32<T>(dv_ws[idx]);
    }
  }
}

int check_dims(const TxlAttnDims* D) {
  TXL_CHECK_ARG(D && D->B > 0 && D->H > 0 && D->band.T > 0 && D->band.mlen >= 0, "relattn: bad dims");
  TXL_CHECK_ARG(D->dh == 16 || D->dh == 32 || D->dh == 64 || D->dh == 128, "relattn: d_head %d not in {16,32,64,128}", D->dh);
  TXL_CHECK_ARG(!(D->band.same_length && D->band.mem_len <= 0), "relattn: same_length with mem_len<=0 masks every key");
  TXL_CHECK_ARG(D->ldq >= D->H * D->dh && D->ldkv_cur >= D->H * D->dh && (D->band.mlen == 0 || D->ldkv_mem >= D->H * D->dh), "relattn: bad strides");
  return TXL_OK;
}
}  // namespace

#define DISPATCH_T_DH(dtype, dh, ...)                                                            \
  if ((dtype) == TXL_F32) { typedef float T; DISPATCH_DH(dh, __VA_ARGS__) }                      \
  else if ((dtype) == TXL_BF16) { typedef bf16 T; DISPATCH_DH(dh, __VA_ARGS__) }                 \
  else { txl_set_error("relattn: bad dtype"); return TXL_EINVAL; }
#define DISPATCH_DH(dh, ...)                                  \
  switch (dh) {                                               \
    case 16: { constexpr int DH = 16; __VA_ARGS__; } break;   \
    case 32: { constexpr int DH = 32; __VA_ARGS__; } break;   \
    case 64: { constexpr int DH = 64; __VA_ARGS__; } break;   \
    default: { constexpr int DH = 128; __VA_ARGS__; } break;  \
  }

extern "C" int txl_relattn_fwd(const void* q, const void* k_mem, const void* v_mem, const void* k_cur, const void* v_cur,
                               const void* r, const float* rwb, const float* rrb, void* out, float* lse, void* saved,
                               const TxlAttnDims* D, void* stream) {
  int rc = check_dims(D);
  if (rc) return rc;
  TXL_CHECK_ARG(q && k_cur && v_cur && r && rwb && rrb && out && lse && (D->band.mlen == 0 || (k_mem && v_mem)), "relattn_fwd: null pointer");
  if (D->dtype == TXL_BF16) {
    int handled = 0;
    rc = txl_relattn_fwd_tc(q, k_mem, v_mem, k_cur, v_cur, r, rwb, rrb, out, lse, saved, D, stream, &handled);
    if (rc) return rc;
    if (handled) return TXL_OK;
  }
  if (saved) { txl_set_error("relattn_fwd: `saved` given but the tensor-core path cannot run this call (see txl_relattn_saved_bytes)"); return TXL_EINVAL; }
  AttnPtrs P{q, k_mem, v_mem, k_cur, v_cur, r, rwb, rrb};
  dim3 grid((D->band.T + BQ - 1) / BQ, D->H, D->B);
  cudaStream_t st = (cudaStream_t)stream;
  DISPATCH_T_DH(D->dtype, D->dh, {
    size_t smem = sizeof(Smem<T, DH>);
    TXL_CUDA(cudaFuncSetAttribute(relattn_fwd_kernel<T, DH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    relattn_fwd_kernel<T, DH><<<grid, 128, smem, st>>>(P, (T*)out, lse, *D);
  });
  TXL_LAUNCH_CHECK();
  return TXL_OK;
}

extern "C" int64_t txl_relattn_saved_bytes(const TxlAttnDims* D) {
  if (!D || check_dims(D)) return 0;
  return txl_relattn_saved_bytes_tc(D);
}

extern "C" int64_t txl_relattn_bwd_workspace(const TxlAttnDims* D) {
  if (!D) return 0;
  int64_t simt = 2ll * D->B * (D->band.mlen + D->band.T) * D->H * D->dh * (int64_t)sizeof(float);
  int64_t tc = txl_relattn_bwd_tc_workspace(D);
  return simt > tc ? simt : tc;
}

extern "C" int txl_relattn_bwd(const void* q, const void* k_mem, const void* v_mem, const void* k_cur, const void* v_cur,
                               const void* r, const float* rwb, const float* rrb, const void* out, const float* lse,
                               const void* dout, void* dq, void* dk_mem, void* dv_mem, void* dk_cur, void* dv_cur, float* dr,
                               float* drwb, float* drrb, void* ws, const void* saved, const TxlAttnDims* D, void* stream) {
  int rc = check_dims(D);
  if (rc) return rc;
  TXL_CHECK_ARG(q && k_cur && v_cur && r && rwb && rrb && out && lse && dout && dq && dk_cur && dv_cur && dr && drwb && drrb && ws,
                "relattn_bwd: null pointer");
  TXL_CHECK_ARG((dk_mem == nullptr) == (dv_mem == nullptr), "relattn_bwd: dk_mem/dv_mem must both be given or both NULL");
  if (D->dtype == TXL_BF16) {
    int handled = 0;
    rc = txl_relattn_bwd_tc(q, k_mem, v_mem, k_cur, v_cur, r, rwb, rrb, out, lse, dout, dq, dk_mem, dv_mem, dk_cur, dv_cur, dr, drwb, drrb, ws, saved, D, stream, &handled);
    if (rc) return rc;
    if (handled) return TXL_OK;
  }
  if (saved) { txl_set_error("relattn_bwd: `saved` given but the tensor-core path cannot run this call"); return TXL_EINVAL; }
  AttnPtrs P{q, k_mem, v_mem, k_cur, v_cur, r, rwb, rrb};
  cudaStream_t st = (cudaStream_t)stream;
  const int klen = D->band.mlen + D->band.T, HD = D->H * D->dh;
  int64_t n = (int64_t)D->B * klen * HD;
  float* dk_ws = (float*)ws; float* dv_ws = dk_ws + n;
  TXL_CUDA(cudaMemsetAsync(ws, 0, 2 * n * sizeof(float), st));
  dim3 grid((D->band.T + BQ - 1) / BQ, D->H, D->B);
  DISPATCH_T_DH(D->dtype, D->dh, {
    size_t smem = sizeof(SmemBwd<T, DH>);
    TXL_CUDA(cudaFuncSetAttribute(relattn_bwd_kernel<T, DH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    relattn_bwd_kernel<T, DH><<<grid, 128, smem, st>>>(P, (const T*)out, lse, (const T*)dout, (T*)dq, dk_ws, dv_ws, dr, drwb, drrb, *D);
    int g2 = (int)imin64(cdiv64(n, 256), (int64_t)txl_num_sms() * 16);
    ++g_txl_launches;
    scatter_dkv_kernel<T><<<g2, 256, 0, st>>>(dk_ws, dv_ws, (T*)dk_mem, (T*)dv_mem, (T*)dk_cur, (T*)dv_cur, D->B, D->band.T, D->band.mlen, HD, D->ldkv_mem, D->ldkv_cur);
  });
  TXL_LAUNCH_CHECK();
  return TXL_OK;
}
