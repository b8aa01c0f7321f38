// tc_relattn.cu — relative-position band attention FORWARD on tcgen05 tensor cores (bf16, d_head 64).
//
// One CTA = one (batch, head, 128-query tile); it walks the 64-key tiles of that tile's live band.  Per key tile:
//   S   [128 x  64] = Qw . K^T          (Qw = q + r_w_bias)                     tcgen05.mma, fp32 in TMEM cols   0.. 63
//   BD0 [128 x 192] = Qr . Rwin^T       (Qr = q + r_r_bias, Rwin = 192 rows of   tcgen05.mma, fp32 in TMEM cols  64..255
//                                        HF's r_head_k starting at x0 = T-128-i0+j0)
//   score[i, j] = S[i, j] + BD0[i, 127 - i + j]   <- HF `_rel_shift` as a per-row skew of the window (Appendix A.4):
//        the warp-uniform part of the skew (32*(3-warp)) is folded into the TMEM column address of tcgen05.ld,
//        the per-lane part (31-lane) is a 5-stage register barrel shifter (select ops, static register indices).
//   band mask from the integer closed forms (common.cuh), online softmax in the exp2 domain, P (bf16) -> swizzled smem,
//   O  += P . V                          tcgen05.mma into TMEM cols 0..63 (S region reused), accumulated in registers.
// K, V and the R window arrive by TMA (SWIZZLE_128B); Qw/Qr/P are written by the softmax threads in the same swizzle.
// 160 threads: warps 0-3 = one query row per thread (TMEM lane = row), warp 4 lane 0 = TMA producer + MMA issuer.
// TMEM 256 columns and ~91 KB of shared memory per CTA => two CTAs per SM overlap each other's MMA and softmax phases.
// Nothing of size T x klen is ever materialised.   [A.3 steps 2-8, A.4, A.5]
#include "tc_common.cuh"
#include <stdlib.h>

namespace {
constexpr int BQ = 128, BKV = 64, DH = 64, WIN = 192;
constexpr int NTHREADS = 160;
constexpr int OFF_QW = 0, OFF_QR = 16384, OFF_P = 32768, OFF_K = 49152, OFF_V = 57344, OFF_R = 65536, OFF_BAR = 90112;
constexpr int SMEM_BYTES = OFF_BAR + 128 + 1024;
constexpr int TMEM_COLS = 256;

struct FwdArgs {
  const bf16* q;
  const float *rwb, *rrb;
  bf16* out;
  float* lse;
  int B, H;
  TxlBand band;
  int64_t ldq;
  float scale_log2;
  float* m_tiles;      // saved-for-backward (optional): the reference m used by every (row, key tile), [B, H, nI, nt_max, 128] fp32
  int nt_max;          //   ... the bf16 P~ = exp2(score - m) tiles themselves go out through tmP, [B, H, nI, nt_max][128 x 64]
  int frozen_ref;      // saving mode: m is fixed by the row's first key tile that holds a live key, so P = P~ * exp2(m - lse) with ONE factor per row — the backward
                       //   folds it into dO and reads the P~ tiles directly instead of writing and re-reading a normalised copy (0.6 GB per layer).
                       //   Later scores may exceed m: P~ > 1 is fine in bf16 / fp32 (same exponent range); a row whose scores spread by more than
                       //   ~88 nats around its first tile's maximum overflows to inf / NaN loudly (the exact SIMT kernels have no such limit).
};

// w[j] <- w[j + sh] for j < OUT, 0 <= sh < 32; W = OUT + 31 valid inputs.  Select ops only, static register indices.
template <int OUT>
__device__ __forceinline__ void barrel_shift(float* w, int sh) {
  const bool b16 = sh & 16, b8 = sh & 8, b4 = sh & 4, b2 = sh & 2, b1 = sh & 1;
#pragma unroll
  for (int c = 0; c < OUT + 15; ++c) w[c] = b16 ? w[c + 16] : w[c];
#pragma unroll
  for (int c = 0; c < OUT + 7; ++c) w[c] = b8 ? w[c + 8] : w[c];
#pragma unroll
  for (int c = 0; c < OUT + 3; ++c) w[c] = b4 ? w[c + 4] : w[c];
#pragma unroll
  for (int c = 0; c < OUT + 1; ++c) w[c] = b2 ? w[c + 2] : w[c];
#pragma unroll
  for (int c = 0; c < OUT; ++c) w[c] = b1 ? w[c + 1] : w[c];
}

__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
  __nv_bfloat162 p = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&p);
}

__global__ void __launch_bounds__(NTHREADS, 2)
relattn_fwd_tc_kernel(const __grid_constant__ CUtensorMap tmKm, const __grid_constant__ CUtensorMap tmVm,
                      const __grid_constant__ CUtensorMap tmKc, const __grid_constant__ CUtensorMap tmVc,
                      const __grid_constant__ CUtensorMap tmR, const __grid_constant__ CUtensorMap tmP, const FwdArgs a) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  uint8_t* sm = smem_raw + ((1024u - (raw & 1023u)) & 1023u);
  uint64_t* bars = reinterpret_cast<uint64_t*>(sm + OFF_BAR);
  uint64_t *kr_full = bars + 0, *v_full = bars + 1, *s_full = bars + 2, *p_full = bars + 3, *o_full = bars + 4, *o_read = bars + 5;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 6);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const BandGeom g = make_band(a.band);
  const int i0 = blockIdx.x * BQ, h = blockIdx.y, b = blockIdx.z;
  const int HD = a.H * DH;
  const int ilast = min(i0 + BQ - 1, g.T - 1);
  const int jt0 = band_lo(g, i0) / BKV;
  const int jt1 = min(band_hi(g, ilast), g.klen - 1) / BKV;
  const int ntiles = jt1 - jt0 + 1;
  const int tile0 = ((b * a.H + h) * (int)gridDim.x + (int)blockIdx.x) * a.nt_max;   // first saved tile of this CTA

  if (tid == 0) {
    mbar_init(kr_full, 1); mbar_init(v_full, 1); mbar_init(s_full, 1);
    mbar_init(p_full, 128); mbar_init(o_full, 1); mbar_init(o_read, 128);
    fence_barrier_init();
  }
  if (warp == 4) tmem_alloc<TMEM_COLS>(tmem_slot);
  if (warp < 4) {
    // ---- stage Qw = q + r_w_bias and Qr = q + r_r_bias as swizzled K-major A operands (row = query, 128 B per row)
    const int r = tid, i = i0 + r;
    const uint4* src = reinterpret_cast<const uint4*>(a.q + ((int64_t)b * g.T + i) * a.ldq + h * DH);
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      uint4 raw4 = make_uint4(0, 0, 0, 0);
      if (i < g.T) raw4 = src[c];
      const bf16* e = reinterpret_cast<const bf16*>(&raw4);
      float f[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) f[k] = __bfloat162float(e[k]);
      const float4 w0 = *reinterpret_cast<const float4*>(a.rwb + h * DH + c * 8), w1 = *reinterpret_cast<const float4*>(a.rwb + h * DH + c * 8 + 4);
      const float4 r0 = *reinterpret_cast<const float4*>(a.rrb + h * DH + c * 8), r1 = *reinterpret_cast<const float4*>(a.rrb + h * DH + c * 8 + 4);
      uint4 ow, orr;
      ow.x = pack_bf16(f[0] + w0.x, f[1] + w0.y); ow.y = pack_bf16(f[2] + w0.z, f[3] + w0.w);
      ow.z = pack_bf16(f[4] + w1.x, f[5] + w1.y); ow.w = pack_bf16(f[6] + w1.z, f[7] + w1.w);
      orr.x = pack_bf16(f[0] + r0.x, f[1] + r0.y); orr.y = pack_bf16(f[2] + r0.z, f[3] + r0.w);
      orr.z = pack_bf16(f[4] + r1.x, f[5] + r1.y); orr.w = pack_bf16(f[6] + r1.z, f[7] + r1.w);
      const uint32_t off = r * 128 + ((c ^ (r & 7)) << 4);
      *reinterpret_cast<uint4*>(sm + OFF_QW + off) = ow;
      *reinterpret_cast<uint4*>(sm + OFF_QR + off) = orr;
    }
    fence_proxy_async_smem();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 4) {
    // ======================= TMA producer + MMA issuer: the whole warp runs the loop converged, one elected lane issues
    {
      const uint32_t idesc_s = umma_idesc_bf16(BQ, BKV, 0, 0);
      const uint32_t idesc_bd = umma_idesc_bf16(BQ, WIN, 0, 0);
      const uint32_t idesc_pv = umma_idesc_bf16(BQ, DH, 0, 1);
      const uint32_t qw_addr = smem_u32(sm + OFF_QW), qr_addr = smem_u32(sm + OFF_QR), p_addr = smem_u32(sm + OFF_P);
      const uint32_t k_addr = smem_u32(sm + OFF_K), v_addr = smem_u32(sm + OFF_V), r_addr = smem_u32(sm + OFF_R);
      auto load_kr = [&](int n) {
        if (lane == 0) {
          const int j0 = (jt0 + n) * BKV;
          mbar_expect_tx(kr_full, BKV * DH * 2 + WIN * DH * 2);
          if (j0 < g.mlen) tma_load_2d(sm + OFF_K, &tmKm, kr_full, h * DH, b * g.mlen + j0);
          else tma_load_2d(sm + OFF_K, &tmKc, kr_full, h * DH, b * g.T + (j0 - g.mlen));
          tma_load_2d(sm + OFF_R, &tmR, kr_full, h * DH, g.T - BQ - i0 + j0);     // x0 = T-128-i0+j0 (rows past klen: zero fill)
        }
        __syncwarp();
      };
      auto load_v = [&](int n) {
        if (lane == 0) {
          const int j0 = (jt0 + n) * BKV;
          mbar_expect_tx(v_full, BKV * DH * 2);
          if (j0 < g.mlen) tma_load_2d(sm + OFF_V, &tmVm, v_full, h * DH, b * g.mlen + j0);
          else tma_load_2d(sm + OFF_V, &tmVc, v_full, h * DH, b * g.T + (j0 - g.mlen));
        }
        __syncwarp();
      };
      load_kr(0);
      load_v(0);
      for (int n = 0; n < ntiles; ++n) {
        const uint32_t ph = n & 1;
        if (n > 0) mbar_wait(o_read, (n - 1) & 1);   // softmax threads have drained the previous P.V out of TMEM cols 0..63
        mbar_wait(kr_full, ph);
        tc_fence_after();
#pragma unroll
        for (int kk = 0; kk < DH / 16; ++kk)
          umma_bf16_warp(tmem_base, umma_smem_desc(qw_addr + kk * 32, 16, 1024), umma_smem_desc(k_addr + kk * 32, 16, 1024), idesc_s, kk > 0);
#pragma unroll
        for (int kk = 0; kk < DH / 16; ++kk)
          umma_bf16_warp(tmem_base + 64, umma_smem_desc(qr_addr + kk * 32, 16, 1024), umma_smem_desc(r_addr + kk * 32, 16, 1024), idesc_bd, kk > 0);
        umma_commit_warp(s_full);
        mbar_wait(s_full, ph);                        // K and R smem are free again: prefetch the next tile behind the softmax
        if (n + 1 < ntiles) load_kr(n + 1);
        mbar_wait(p_full, ph);
        if (a.m_tiles) {      // save the bf16 P~ tile for the backward pass, straight from the swizzled MMA operand
          if (lane == 0) {
            asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(reinterpret_cast<uint64_t>(&tmP)), "r"(p_addr), "r"(0), "r"((tile0 + n) * BQ) : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
          }
          __syncwarp();
        }
        mbar_wait(v_full, ph);
        tc_fence_after();
#pragma unroll
        for (int kk = 0; kk < BKV / 16; ++kk)
          umma_bf16_warp(tmem_base, umma_smem_desc(p_addr + kk * 32, 16, 1024), umma_smem_desc(v_addr + kk * 2048, 8192, 1024), idesc_pv, kk > 0);
        umma_commit_warp(o_full);
        mbar_wait(o_full, ph);
        if (n + 1 < ntiles) load_v(n + 1);
        if (a.m_tiles) {                              // P smem is rewritten only after the next s_full
          if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
          __syncwarp();
        }
      }
      if (a.m_tiles && lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
      __syncwarp();
    }
  } else {
    // ======================= softmax threads: thread r owns query row i0 + r (= TMEM lane r)
    const int r = tid, i = i0 + r;
    const uint32_t lane_base = (uint32_t)(warp * 32) << 16;
    const int sh = 31 - lane;
    int lo_i = 1, hi_i = 0;    // rows past T: everything masked
    if (i < g.T) { lo_i = band_lo(g, i); hi_i = min(band_hi(g, i), g.klen - 1); }
    const int lo_all = band_lo(g, ilast), hi_all = band_hi(g, i0);   // keys live for EVERY row of the tile (if tile rows all < T)
    float m_run = -INFINITY, l_run = 0.f, m_ref = 0.f;
    float O[DH];
#pragma unroll
    for (int c = 0; c < DH; ++c) O[c] = 0.f;

    for (int n = 0; n < ntiles; ++n) {
      const uint32_t ph = n & 1;
      const int j0 = (jt0 + n) * BKV;
      mbar_wait(s_full, ph);
      tc_fence_after();
      float t[BKV];
      tmem_ld_32x32(tmem_base + lane_base, t);
      tmem_ld_32x32(tmem_base + lane_base + 32, t + 32);
      tmem_ld_wait();
#pragma unroll
      for (int qd = 0; qd < 4; ++qd) {
        float w[48];
        const uint32_t cb = 64 + 32 * (3 - warp) + 16 * qd;
        tmem_ld_32x16(tmem_base + lane_base + cb, w);
        tmem_ld_32x16(tmem_base + lane_base + cb + 16, w + 16);
        tmem_ld_32x16(tmem_base + lane_base + cb + 32, w + 32);
        tmem_ld_wait();
        barrel_shift<16>(w, sh);
#pragma unroll
        for (int jj = 0; jj < 16; ++jj) t[16 * qd + jj] += w[jj];
      }
      const bool tile_full = (i0 + BQ <= g.T) && j0 >= lo_all && j0 + BKV - 1 <= hi_all;
      if (!tile_full) {
#pragma unroll
        for (int jj = 0; jj < BKV; ++jj) {
          const int j = j0 + jj;
          if (j < lo_i || j > hi_i) t[jj] = -INFINITY;
        }
      }
      float m_new, m_safe, corr;
      if (a.frozen_ref && m_run != -INFINITY) { m_new = m_run; m_safe = m_ref; corr = 1.f; }   // reference fixed by the row's first key tile with a live key: no maximum needed
      else {
        float mx = t[0];
#pragma unroll
        for (int jj = 1; jj < BKV; ++jj) mx = fmaxf(mx, t[jj]);
        m_new = fmaxf(m_run, mx * a.scale_log2);
        m_safe = (m_new == -INFINITY) ? 0.f : m_new;
        corr = exp2f(m_run - m_safe);
        m_ref = m_safe;
      }
      float sum = 0.f;
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        float p[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) { p[k] = exp2f(fmaf(t[c * 8 + k], a.scale_log2, -m_safe)); sum += p[k]; }
        uint4 o;
        o.x = pack_bf16(p[0], p[1]); o.y = pack_bf16(p[2], p[3]); o.z = pack_bf16(p[4], p[5]); o.w = pack_bf16(p[6], p[7]);
        *reinterpret_cast<uint4*>(sm + OFF_P + r * 128 + ((c ^ (r & 7)) << 4)) = o;
      }
      l_run = l_run * corr + sum;
      m_run = m_new;
      if (a.m_tiles) a.m_tiles[(int64_t)(tile0 + n) * BQ + r] = m_safe;
      fence_proxy_async_smem();
      tc_fence_before();
      mbar_arrive(p_full);
      mbar_wait(o_full, ph);
      tc_fence_after();
      float pv[DH];
      tmem_ld_32x32(tmem_base + lane_base, pv);
      tmem_ld_32x32(tmem_base + lane_base + 32, pv + 32);
      tmem_ld_wait();
      tc_fence_before();
      mbar_arrive(o_read);
#pragma unroll
      for (int c = 0; c < DH; ++c) O[c] = fmaf(O[c], corr, pv[c]);
    }
    if (i < g.T) {
      const float inv = 1.f / l_run;
      uint4* dst = reinterpret_cast<uint4*>(a.out + ((int64_t)b * g.T + i) * HD + h * DH);
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        uint4 o;
        o.x = pack_bf16(O[c * 8 + 0] * inv, O[c * 8 + 1] * inv); o.y = pack_bf16(O[c * 8 + 2] * inv, O[c * 8 + 3] * inv);
        o.z = pack_bf16(O[c * 8 + 4] * inv, O[c * 8 + 5] * inv); o.w = pack_bf16(O[c * 8 + 6] * inv, O[c * 8 + 7] * inv);
        dst[c] = o;
      }
      a.lse[((int64_t)b * a.H + h) * g.T + i] = (a.frozen_ref ? m_ref : m_run) * 0.6931471805599453f + logf(l_run);
      // the row's reference where the backward's prep kernel looks for it: the first tile's slot (a row whose first tile was fully masked has
      // P~ = 0 there, so any m serves that tile)
      if (a.m_tiles && a.frozen_ref) a.m_tiles[(int64_t)tile0 * BQ + r] = m_ref;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 4) tmem_dealloc<TMEM_COLS>(tmem_base);
}
}  // namespace

int txl_relattn_fwd_tc(const void* q, const void* k_mem, const void* v_mem, const void* k_cur, const void* v_cur, const void* r,
                       const float* rwb, const float* rrb, void* out, float* lse, void* saved, const TxlAttnDims* D, void* stream, int* handled) {
  *handled = 0;
  static int disabled = -1;
  if (disabled < 0) { const char* e = getenv("TXL_DISABLE_TC_ATTN"); const char* e2 = getenv("TXL_DISABLE_TC"); disabled = ((e && e[0] == '1') || (e2 && e2[0] == '1')) ? 1 : 0; }
  if (disabled) return TXL_OK;
  const int T = D->band.T, mlen = D->band.mlen, klen = T + mlen, HD = D->H * D->dh;
  auto al16 = [](const void* p) { return ((uintptr_t)p & 15) == 0; };
  if (D->dh != DH || (T % BKV) || (mlen % BKV) || klen < WIN) return TXL_OK;
  if ((D->ldq % 8) || (D->ldkv_cur % 8) || (mlen > 0 && (D->ldkv_mem % 8))) return TXL_OK;
  if (!al16(q) || !al16(k_cur) || !al16(v_cur) || !al16(r) || !al16(out) || !al16(rwb) || !al16(rrb) || (mlen > 0 && (!al16(k_mem) || !al16(v_mem)))) return TXL_OK;

  CUtensorMap tmKm, tmVm, tmKc, tmVc, tmR;
  int rc;
  if ((rc = txl_make_tmap_2d(&tmKc, k_cur, (uint64_t)D->B * T, (uint64_t)HD, (uint64_t)D->ldkv_cur, BKV, DH))) return rc;
  if ((rc = txl_make_tmap_2d(&tmVc, v_cur, (uint64_t)D->B * T, (uint64_t)HD, (uint64_t)D->ldkv_cur, BKV, DH))) return rc;
  if (mlen > 0) {
    if ((rc = txl_make_tmap_2d(&tmKm, k_mem, (uint64_t)D->B * mlen, (uint64_t)HD, (uint64_t)D->ldkv_mem, BKV, DH))) return rc;
    if ((rc = txl_make_tmap_2d(&tmVm, v_mem, (uint64_t)D->B * mlen, (uint64_t)HD, (uint64_t)D->ldkv_mem, BKV, DH))) return rc;
  } else { tmKm = tmKc; tmVm = tmVc; }
  if ((rc = txl_make_tmap_2d(&tmR, r, (uint64_t)klen, (uint64_t)HD, (uint64_t)HD, WIN, DH))) return rc;
  CUtensorMap tmP = tmR;
  FwdArgs a;
  a.m_tiles = nullptr; a.nt_max = 0; a.frozen_ref = 0;
  if (saved) {
    if (!al16(saved) || txl_relattn_saved_bytes_tc(D) == 0) return TXL_OK;
    const int64_t trows = txl_relattn_tile_rows(D);
    a.nt_max = txl_relattn_nt_max(&D->band);
    a.m_tiles = reinterpret_cast<float*>(reinterpret_cast<bf16*>(saved) + trows * BKV);
    if ((rc = txl_make_tmap_2d(&tmP, saved, (uint64_t)trows, (uint64_t)BKV, (uint64_t)BKV, BQ, BKV))) return rc;
    a.frozen_ref = txl_relattn_frozen_ref();
  }

  a.q = (const bf16*)q; a.rwb = rwb; a.rrb = rrb; a.out = (bf16*)out; a.lse = lse; a.B = D->B; a.H = D->H; a.band = D->band; a.ldq = D->ldq;
  a.scale_log2 = 1.4426950408889634f / sqrtf((float)DH);
  static bool attr_set = false;
  if (!attr_set) {
    TXL_CUDA(cudaFuncSetAttribute(relattn_fwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    attr_set = true;
  }
  dim3 grid((T + BQ - 1) / BQ, D->H, D->B);
  relattn_fwd_tc_kernel<<<grid, NTHREADS, SMEM_BYTES, (cudaStream_t)stream>>>(tmKm, tmVm, tmKc, tmVc, tmR, tmP, a);
  TXL_LAUNCH_CHECK();
  *handled = 1;
  return TXL_OK;
}
