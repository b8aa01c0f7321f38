// tc_relattn_bwd.cu — relative-position band attention BACKWARD on tcgen05 tensor cores (bf16, d_head 64).
//
// One kernel template, three accumulation orders over the same (batch b, 128-query tile I, 64-key tile J) band tiles:
//   MODE_DQ : CTA = (b, h, I), loops J.   acc0 = dQw = sum_J dS.K        acc1 = dQr = sum_J dBD0.Rwin   -> dq, d r_w_bias, d r_r_bias
//   MODE_DKV: CTA = (b, h, J), loops I.   acc0 = dK  = sum_I dS^T.Qw     acc1 = dV  = sum_I P^T.dO      -> dk, dv
//   MODE_DR : CTA = (h, diagonal J-2I), loops (I, b) — the R window is the same for all of them.
//                                          acc0/acc1 = dRwin = sum dBD0^T.Qr                             -> dr (fp32 atomics, ~2 M per layer)
// Every tile recomputes, on the tensor cores, S = Qw.K^T, BD0 = Qr.Rwin^T and dP = dO.V^T into TMEM; 512 "softmax" threads
// (4 per query row, 16 keys each) apply HF's `_rel_shift` exactly as the forward kernel does (TMEM column offset + register
// barrel shifter), rebuild P = exp2(score - lse) from the saved log-sum-exp, form dS = P (dP - delta) / sqrt(dh), and hand the
// tensor cores bf16 tiles in shared memory: P and dS K-major (also read transposed, MN-major, for dV/dK), and dBD0 — dS
// un-shifted back into window space by a per-row element offset on the shared-memory store (the inverse skew costs no ALU).
// No atomics are used for dq/dk/dv; nothing of size T x klen touches HBM.
// Operands arrive by TMA through a 2-stage ring (warp 8), one thread of warp 9 issues all tcgen05.mma.
// Supported: mlen == mem_len with same_length (every query sees exactly mem_len keys — the training configuration);
// anything else takes the SIMT kernel.   [A.3, A.4, A.5 backward]
#include "tc_common.cuh"
#include <stdlib.h>

namespace {
constexpr int BQ = 128, BKV = 64, DH = 64, WIN = 192;
constexpr int MODE_DQ = 0, MODE_DKV = 1, MODE_DR = 2;
constexpr int KPT = 16;                       // keys per softmax thread: 4 threads share a query row
constexpr int N_SOFTMAX = 128 * (BKV / KPT), NTHREADS = N_SOFTMAX + 96;
constexpr int W_PROD = N_SOFTMAX / 32, W_MMA = W_PROD + 1, W_ST = W_PROD + 2;   // TMA-load, MMA-issue and tile-store warps
constexpr int TM_S = 0, TM_DP = 64, TM_BD = 128, TM_ACC0 = 320, TM_ACC1 = 384, TMEM_COLS = 512;
constexpr int SZ_Q = 16384, SZ_KV = 8192, SZ_R = 24576, SZ_DBD = 49152;

// shared-memory plan per mode: resident operands, a 2-stage ring of streamed operands, work tiles written by the softmax threads
template <int MODE> struct Plan;
template <> struct Plan<MODE_DQ> {   // resident Qw Qr dO | ring {K V R} | dS dBD
  static constexpr int QW = 0, QR = 16384, DO = 32768, RING = 49152, STAGE = SZ_KV * 2 + SZ_R;
  static constexpr int K = 0, V = SZ_KV, R = 2 * SZ_KV;                     // offsets inside a stage
  static constexpr int DS = RING + 2 * STAGE, DBD = DS + SZ_Q, P = DBD + SZ_DBD, BAR = P + SZ_Q;   // P tile only for the tile store
  static constexpr int RES_BYTES = 3 * SZ_Q, STAGE_TX = STAGE;
};
template <> struct Plan<MODE_DKV> {  // resident K V | ring {Qw Qr dO R} | P dS pad
  static constexpr int K = 0, V = 8192, RING = 16384, STAGE = 3 * SZ_Q + SZ_R;
  static constexpr int QW = 0, QR = SZ_Q, DO = 2 * SZ_Q, R = 3 * SZ_Q;
  static constexpr int P = RING + 2 * STAGE, DS = P + SZ_Q, DBD = -1, BAR = DS + 2 * SZ_Q;   // 16 KB pad after dS: 2nd MN atom of dS^T
  static constexpr int RES_BYTES = 2 * SZ_KV, STAGE_TX = STAGE;
};
template <> struct Plan<MODE_DR> {   // resident R | ring {Qw Qr dO K V} | dBD
  static constexpr int R = 0, RING = 24576, STAGE = 3 * SZ_Q + 2 * SZ_KV;
  static constexpr int QW = 0, QR = SZ_Q, DO = 2 * SZ_Q, K = 3 * SZ_Q, V = 3 * SZ_Q + SZ_KV;
  static constexpr int DBD = RING + 2 * STAGE, DS = -1, P = -1, BAR = DBD + SZ_DBD;
  static constexpr int RES_BYTES = SZ_R, STAGE_TX = STAGE;
};
template <int MODE> constexpr int smem_bytes() { return Plan<MODE>::BAR + 128 + 1024; }

struct BwdArgs {
  const float *lse, *delta;      // [B, H, T]
  bf16 *dq, *dk_mem, *dv_mem, *dk_cur, *dv_cur;
  float *dr, *drwb, *drrb;
  int B, H;
  TxlBand band;
  int64_t ldq, ldkv_mem, ldkv_cur;
  float scale, scale_log2;
  int delta_min, n_delta;        // MODE_DR: diagonals J - 2I
  int store_tiles, nt_max;       // MODE_DQ: also write the bf16 P and dS tiles of every band tile (consumed by the lite dK/dV and dR kernels)
  int store_p;                   // dQ pass over saved tiles: 0 = the dK/dV pass reads the forward's P~ tiles (per-row reference), only dS is written
  const float* m_tiles;          // saved by the forward kernel: the running max behind every (row, key tile) of its P~ tiles
  int abl;                       // TXL_ABL timing ablations (results invalid when non-zero)
};

struct Tile { int b, I, J; };

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
__device__ __forceinline__ uint32_t pack2(float a, float b) {
  __nv_bfloat162 p = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&p);
}
__device__ __forceinline__ void tma_store_tile(const CUtensorMap* m, const void* src, int row) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(src)), "r"(0), "r"(row) : "memory");
}
__device__ __forceinline__ int q_tile_first_kt(const BandGeom& g, int I) { return band_lo(g, I * BQ) / BKV; }
__device__ __forceinline__ int q_tile_last_kt(const BandGeom& g, int I) {
  return min(band_hi(g, min(I * BQ + BQ - 1, g.T - 1)), g.klen - 1) / BKV;
}

// tile enumeration of one CTA
template <int MODE>
struct TileIter {
  int count, b, h, I, J, Ilo, delta, nI;
  __device__ void init(const BwdArgs& a, const BandGeom& g) {
    nI = (g.T + BQ - 1) / BQ;
    if (MODE == MODE_DQ) {
      I = blockIdx.x; h = blockIdx.y; b = blockIdx.z;
      J = q_tile_first_kt(g, I);
      count = q_tile_last_kt(g, I) - J + 1;
    } else if (MODE == MODE_DKV) {
      J = blockIdx.x; h = blockIdx.y; b = blockIdx.z;
      Ilo = -1; count = 0;
      for (int i = 0; i < nI; ++i)
        if (q_tile_first_kt(g, i) <= J && J <= q_tile_last_kt(g, i)) { if (Ilo < 0) Ilo = i; ++count; }
    } else {
      delta = a.delta_min + (int)blockIdx.x; h = blockIdx.y; b = 0;
      Ilo = -1; count = 0;
      for (int i = 0; i < nI; ++i) {
        int j = 2 * i + delta;
        if (q_tile_first_kt(g, i) <= j && j <= q_tile_last_kt(g, i)) { if (Ilo < 0) Ilo = i; ++count; }
      }
      count *= a.B;
    }
  }
  __device__ Tile get(int n, const BwdArgs& a) const {
    Tile t;
    if (MODE == MODE_DQ) { t.b = b; t.I = I; t.J = J + n; }
    else if (MODE == MODE_DKV) { t.b = b; t.I = Ilo + n; t.J = J; }
    else { t.I = Ilo + n / a.B; t.b = n % a.B; t.J = 2 * t.I + delta; }
    return t;
  }
};

struct Maps { CUtensorMap qw, qr, dO, km, vm, kc, vc, r, pst, dst, psv, r64, pst2, dst2, pkv, pkv2, dOkv; };   // pst/dst: [tiles*128, 64] bf16 tile stores (P, dS); psv: P~ saved by the forward; r64: R in 64-row boxes; pst2/dst2: two tiles per box;
// pkv / pkv2 / dOkv: what the lite dK/dV kernels multiply for dV — (P, dO), or with the forward's per-row reference (P~, dO scaled by exp2(m - lse))

template <int MODE>
__global__ void __launch_bounds__(NTHREADS, 1) relattn_bwd_tc_kernel(const __grid_constant__ Maps M, const BwdArgs a) {
  using PL = Plan<MODE>;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  uint8_t* sm = smem_raw + ((1024u - (raw & 1023u)) & 1023u);
  uint64_t* bars = reinterpret_cast<uint64_t*>(sm + PL::BAR);
  uint64_t *full = bars, *empty = bars + 2, *res_full = bars + 4, *f_full = bars + 5, *b_ready = bars + 6, *b_done = bars + 7, *acc_full = bars + 8, *t_free = bars + 9;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 10);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const BandGeom g = make_band(a.band);
  const int HD = a.H * DH;
  TileIter<MODE> it;
  it.init(a, g);
  const int h = it.h;
  if (MODE == MODE_DKV && it.J * BKV < g.mlen && a.dk_mem == nullptr) return;   // gradients of detached all-zero mems are not needed
  if (it.count == 0) {
    if (MODE == MODE_DKV && tid < BKV) {   // key tile outside every band: zero gradient
      const int j = it.J * BKV + tid;
      bf16* dk = j < g.mlen ? a.dk_mem + ((int64_t)it.b * g.mlen + j) * a.ldkv_mem : a.dk_cur + ((int64_t)it.b * g.T + (j - g.mlen)) * a.ldkv_cur;
      bf16* dv = j < g.mlen ? a.dv_mem + ((int64_t)it.b * g.mlen + j) * a.ldkv_mem : a.dv_cur + ((int64_t)it.b * g.T + (j - g.mlen)) * a.ldkv_cur;
      for (int c = 0; c < DH; ++c) { dk[h * DH + c] = __float2bfloat16(0.f); dv[h * DH + c] = __float2bfloat16(0.f); }
    }
    return;
  }

  if (tid == 0) {
    mbar_init(&full[0], 1); mbar_init(&full[1], 1); mbar_init(&empty[0], 1); mbar_init(&empty[1], 1);
    mbar_init(res_full, 1); mbar_init(f_full, 1); mbar_init(b_ready, N_SOFTMAX); mbar_init(b_done, (MODE == MODE_DQ && a.store_tiles) ? 2 : 1); mbar_init(acc_full, 1); mbar_init(t_free, N_SOFTMAX);
    fence_barrier_init();
  }
  if (warp == W_MMA) tmem_alloc<TMEM_COLS>(tmem_slot);
  if (PL::DBD >= 0) {   // window-space tile: only positions [127-r+32hh, +32) of row r are ever written, the rest must read as 0
    for (int e = tid; e < SZ_DBD / 16; e += NTHREADS) reinterpret_cast<uint4*>(sm + PL::DBD)[e] = make_uint4(0, 0, 0, 0);
  }
  if (MODE == MODE_DKV) {   // pad behind dS (second MN atom of dS^T): finite values only
    for (int e = tid; e < SZ_Q / 16; e += NTHREADS) reinterpret_cast<uint4*>(sm + PL::DS + SZ_Q)[e] = make_uint4(0, 0, 0, 0);
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  auto k_coords = [&](const Tile& t, const CUtensorMap*& mk, const CUtensorMap*& mv, int& row) {
    const int j0 = t.J * BKV;
    if (j0 < g.mlen) { mk = &M.km; mv = &M.vm; row = t.b * g.mlen + j0; }
    else { mk = &M.kc; mv = &M.vc; row = t.b * g.T + (j0 - g.mlen); }
  };

  if (warp == W_PROD) {
    if (lane == 0) {
      // ======================= TMA producer
      {
        const Tile t0 = it.get(0, a);
        mbar_expect_tx(res_full, PL::RES_BYTES);
        if (MODE == MODE_DQ) {
          const int row = t0.b * g.T + t0.I * BQ;
          tma_load_2d(sm + PL::QW, &M.qw, res_full, h * DH, row);
          tma_load_2d(sm + PL::QR, &M.qr, res_full, h * DH, row);
          tma_load_2d(sm + PL::DO, &M.dO, res_full, h * DH, row);
        } else if (MODE == MODE_DKV) {
          const CUtensorMap *mk, *mv; int row;
          k_coords(t0, mk, mv, row);
          tma_load_2d(sm + PL::K, mk, res_full, h * DH, row);
          tma_load_2d(sm + PL::V, mv, res_full, h * DH, row);
        } else {
          tma_load_2d(sm + PL::R, &M.r, res_full, h * DH, g.T - BQ + BKV * it.delta);   // x0 = T-128-128 I+64 J = T-128+64 delta
        }
      }
      for (int n = 0; n < it.count; ++n) {
        const int s = n & 1; const uint32_t rph = (n >> 1) & 1;
        const Tile t = it.get(n, a);
        mbar_wait(&empty[s], rph ^ 1);
        uint8_t* st = sm + PL::RING + s * PL::STAGE;
        mbar_expect_tx(&full[s], PL::STAGE_TX);
        if (MODE != MODE_DQ) {
          const int row = t.b * g.T + t.I * BQ;
          tma_load_2d(st + PL::QW, &M.qw, &full[s], h * DH, row);
          tma_load_2d(st + PL::QR, &M.qr, &full[s], h * DH, row);
          tma_load_2d(st + PL::DO, &M.dO, &full[s], h * DH, row);
        }
        if (MODE != MODE_DKV) {
          const CUtensorMap *mk, *mv; int row;
          k_coords(t, mk, mv, row);
          tma_load_2d(st + PL::K, mk, &full[s], h * DH, row);
          tma_load_2d(st + PL::V, mv, &full[s], h * DH, row);
        }
        if (MODE != MODE_DR) tma_load_2d(st + PL::R, &M.r, &full[s], h * DH, g.T - BQ - t.I * BQ + t.J * BKV);
      }
    }
  } else if (warp == W_MMA) {
    {   // whole warp, converged: one elected lane issues (umma_bf16_warp)
      // ======================= MMA issuer
      const uint32_t id_s = umma_idesc_bf16(BQ, BKV, 0, 0), id_bd = umma_idesc_bf16(BQ, WIN, 0, 0);
      const uint32_t id_kn = umma_idesc_bf16(BQ, DH, 0, 1);   // A K-major, B MN-major
      const uint32_t id_nn = umma_idesc_bf16(BQ, DH, 1, 1);   // A MN-major (transposed tile), B MN-major
      const uint32_t base = smem_u32(sm);
      mbar_wait(res_full, 0);
      auto stage_base = [&](int n) -> uint32_t { return base + PL::RING + (n & 1) * PL::STAGE; };
      auto front = [&](int n) {     // S, dP, BD0 of band tile n into TMEM
        const uint32_t st = stage_base(n);
        const uint32_t qw = (MODE == MODE_DQ ? base : st) + PL::QW, qr = (MODE == MODE_DQ ? base : st) + PL::QR;
        const uint32_t dO = (MODE == MODE_DQ ? base : st) + PL::DO;
        const uint32_t kk_ = (MODE == MODE_DKV ? base : st) + PL::K, vv = (MODE == MODE_DKV ? base : st) + PL::V;
        const uint32_t rr = (MODE == MODE_DR ? base : st) + PL::R;
        mbar_wait(&full[n & 1], (n >> 1) & 1);
        tc_fence_after();
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_bf16_warp(tmem_base + TM_S, umma_smem_desc(qw + k * 32, 16, 1024), umma_smem_desc(kk_ + k * 32, 16, 1024), id_s, k > 0);
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_bf16_warp(tmem_base + TM_DP, umma_smem_desc(dO + k * 32, 16, 1024), umma_smem_desc(vv + k * 32, 16, 1024), id_s, k > 0);
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_bf16_warp(tmem_base + TM_BD, umma_smem_desc(qr + k * 32, 16, 1024), umma_smem_desc(rr + k * 32, 16, 1024), id_bd, k > 0);
        umma_commit_warp(f_full);
      };
      front(0);
      for (int n = 0; n < it.count; ++n) {
        const int s = n & 1; const uint32_t ph = n & 1;
        const uint32_t st = stage_base(n);
        const uint32_t qw = (MODE == MODE_DQ ? base : st) + PL::QW, qr = (MODE == MODE_DQ ? base : st) + PL::QR;
        const uint32_t dO = (MODE == MODE_DQ ? base : st) + PL::DO;
        const uint32_t kk_ = (MODE == MODE_DKV ? base : st) + PL::K;
        const uint32_t rr = (MODE == MODE_DR ? base : st) + PL::R;
        // as soon as the softmax threads have pulled S/dP/BD0 of tile n into registers (long before they finish the arithmetic),
        // the tensor core starts the front end of tile n+1: it then overlaps tile n's softmax phase
        if (n + 1 < it.count) {
          mbar_wait(t_free, ph);
          tc_fence_after();
          front(n + 1);
        }
        // softmax threads have published the bf16 tiles of tile n
        mbar_wait(b_ready, ph);
        tc_fence_after();
        const uint32_t accum0 = n > 0;
        if (MODE == MODE_DQ) {
          const uint32_t ds = base + PL::DS, dbd = base + PL::DBD;
#pragma unroll
          for (int k = 0; k < 4; ++k)      // dQw += dS . K
            umma_bf16_warp(tmem_base + TM_ACC0, umma_smem_desc(ds + k * 32, 16, 1024), umma_smem_desc(kk_ + k * 2048, 8192, 1024), id_kn, accum0 | (k > 0));
#pragma unroll
          for (int k = 0; k < 12; ++k)     // dQr += dBD0 . Rwin
            umma_bf16_warp(tmem_base + TM_ACC1, umma_smem_desc(dbd + (k >> 2) * 16384 + (k & 3) * 32, 16, 1024), umma_smem_desc(rr + k * 2048, 8192, 1024), id_kn, accum0 | (k > 0));
        } else if (MODE == MODE_DKV) {
          const uint32_t ds = base + PL::DS, pp = base + PL::P;
#pragma unroll
          for (int k = 0; k < 8; ++k)      // dK += dS^T . Qw   (rows 64..127 of the accumulator are a second, ignored MN atom)
            umma_bf16_warp(tmem_base + TM_ACC0, umma_smem_desc(ds + k * 2048, 16384, 1024), umma_smem_desc(qw + k * 2048, 8192, 1024), id_nn, accum0 | (k > 0));
#pragma unroll
          for (int k = 0; k < 8; ++k)      // dV += P^T . dO
            umma_bf16_warp(tmem_base + TM_ACC1, umma_smem_desc(pp + k * 2048, 16384, 1024), umma_smem_desc(dO + k * 2048, 8192, 1024), id_nn, accum0 | (k > 0));
        } else {
          const uint32_t dbd = base + PL::DBD;
#pragma unroll
          for (int k = 0; k < 8; ++k)      // dRwin[0..127]   += dBD0[:, 0..127]^T . Qr
            umma_bf16_warp(tmem_base + TM_ACC0, umma_smem_desc(dbd + k * 2048, 16384, 1024), umma_smem_desc(qr + k * 2048, 8192, 1024), id_nn, accum0 | (k > 0));
#pragma unroll
          for (int k = 0; k < 8; ++k)      // dRwin[64..191]  += dBD0[:, 64..191]^T . Qr   (lanes 64..127 hold window rows 128..191)
            umma_bf16_warp(tmem_base + TM_ACC1, umma_smem_desc(dbd + 16384 + k * 2048, 16384, 1024), umma_smem_desc(qr + k * 2048, 8192, 1024), id_nn, accum0 | (k > 0));
        }
        umma_commit_warp(b_done);
        umma_commit_warp(&empty[s]);
      }
      umma_commit_warp(acc_full);
    }
  } else if (warp == W_ST) {
    // ======================= tile-store thread (MODE_DQ): bf16 P and dS of every band tile -> global, straight from the swizzled smem tiles
    if (lane == 0 && MODE == MODE_DQ && a.store_tiles) {
      for (int n = 0; n < it.count; ++n) {
        mbar_wait(b_ready, n & 1);
        const int row = ((((it.b * a.H + h) * it.nI + it.I) * a.nt_max) + n) * BQ;
        tma_store_tile(&M.pst, sm + PL::P, row);
        tma_store_tile(&M.dst, sm + PL::DS, row);
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        mbar_arrive(b_done);          // second arrival: the work tiles may be overwritten only after the stores have read them
      }
      asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    }
  } else {
    // ======================= softmax threads: row r = 32*(warp%4)+lane, keys [KPT*qd, KPT*qd+KPT) of the tile, qd = warp/4
    const int r = 32 * (warp & 3) + lane, qd = warp >> 2;
    const uint32_t lane_base = (uint32_t)(32 * (warp & 3)) << 16;
    const int sh = 31 - lane;
    const int c0 = 127 - r + KPT * qd;     // first window column this thread writes in the dBD0 tile
    // tile-invariant shared-memory offsets of this thread: its two 16-byte chunks of the K-major tiles, and its bf16 pairs in the dBD0 tile
    uint32_t tile_off[KPT / 8], pair_off[KPT / 2], single_off[2];
    {
      auto dbd_off = [&](int c) -> uint32_t { return (uint32_t)((c >> 6) * 16384 + r * 128 + ((((c & 63) >> 3) ^ (r & 7)) << 4) + (c & 7) * 2); };
#pragma unroll
      for (int c = 0; c < KPT / 8; ++c) tile_off[c] = (uint32_t)(r * 128 + ((((KPT / 8) * qd + c) ^ (r & 7)) << 4));
      const int ce = c0 + (c0 & 1);
#pragma unroll
      for (int k = 0; k < KPT / 2; ++k) pair_off[k] = dbd_off(ce + 2 * k);      // for odd c0 the last pair is unused
      single_off[0] = dbd_off(c0); single_off[1] = dbd_off(c0 + KPT - 1);
    }
    Tile t = it.get(0, a);
    float lse2 = 0.f, dlt = 0.f;
    auto load_row = [&](const Tile& tt, float& l2, float& dl) {
      const int i = tt.I * BQ + r;
      l2 = 0.f; dl = 0.f;
      if (i < g.T) {
        const int64_t o = ((int64_t)tt.b * a.H + h) * g.T + i;
        l2 = a.lse[o] * 1.4426950408889634f; dl = a.delta[o];
      }
    };
    load_row(t, lse2, dlt);
    for (int n = 0; n < it.count; ++n) {
      const uint32_t ph = n & 1;
      Tile tn = t; float lse2n = lse2, dltn = dlt;
      if (n + 1 < it.count) { tn = it.get(n + 1, a); if (MODE != MODE_DQ) load_row(tn, lse2n, dltn); }
      const int i0 = t.I * BQ, j0 = t.J * BKV, i = i0 + r;
      int lo_i = 1, hi_i = 0;
      if (i < g.T) { lo_i = band_lo(g, i); hi_i = min(band_hi(g, i), g.klen - 1); }
      const int ilast = min(i0 + BQ - 1, g.T - 1);
      const bool tile_full = (i0 + BQ <= g.T) && j0 >= band_lo(g, ilast) && j0 + BKV - 1 <= band_hi(g, i0);

      mbar_wait(f_full, ph);
      tc_fence_after();
      float p[KPT], ds[KPT], w[48];
      const uint32_t cb = TM_BD + 32 * (3 - (warp & 3)) + KPT * qd;
      tmem_ld_32x16(tmem_base + lane_base + TM_S + KPT * qd, p);
      tmem_ld_32x16(tmem_base + lane_base + TM_DP + KPT * qd, ds);
      tmem_ld_32x16(tmem_base + lane_base + cb, w);
      tmem_ld_32x16(tmem_base + lane_base + cb + 16, w + 16);
      tmem_ld_32x16(tmem_base + lane_base + cb + 32, w + 32);
      tmem_ld_wait();
      tc_fence_before();
      mbar_arrive(t_free);                   // S / dP / BD0 of this tile now live in registers: TMEM may be overwritten
      barrel_shift<16>(w, sh);
      if (tile_full) {      // interior band tile (the common case): no mask arithmetic at all
#pragma unroll
        for (int jj = 0; jj < KPT; ++jj) {
          const float pj = exp2f(fmaf(p[jj] + w[jj], a.scale_log2, -lse2));
          p[jj] = pj;
          ds[jj] = pj * (ds[jj] - dlt) * a.scale;
        }
      } else {
#pragma unroll
        for (int jj = 0; jj < KPT; ++jj) {
          const int j = j0 + KPT * qd + jj;
          const bool valid = j >= lo_i && j <= hi_i;
          const float pj = valid ? exp2f(fmaf(p[jj] + w[jj], a.scale_log2, -lse2)) : 0.f;
          p[jj] = pj;
          ds[jj] = pj * (ds[jj] - dlt) * a.scale;
        }
      }
      // previous tile's back-end MMAs must have finished reading the work tiles before they are overwritten
      if (n > 0) mbar_wait(b_done, (n - 1) & 1);
      if (PL::P >= 0 && (MODE != MODE_DQ || a.store_tiles)) {
#pragma unroll
        for (int c = 0; c < KPT / 8; ++c) {
          uint4 o; o.x = pack2(p[c * 8], p[c * 8 + 1]); o.y = pack2(p[c * 8 + 2], p[c * 8 + 3]); o.z = pack2(p[c * 8 + 4], p[c * 8 + 5]); o.w = pack2(p[c * 8 + 6], p[c * 8 + 7]);
          *reinterpret_cast<uint4*>(sm + PL::P + tile_off[c]) = o;
        }
      }
      if (PL::DS >= 0) {
#pragma unroll
        for (int c = 0; c < KPT / 8; ++c) {
          uint4 o; o.x = pack2(ds[c * 8], ds[c * 8 + 1]); o.y = pack2(ds[c * 8 + 2], ds[c * 8 + 3]); o.z = pack2(ds[c * 8 + 4], ds[c * 8 + 5]); o.w = pack2(ds[c * 8 + 6], ds[c * 8 + 7]);
          *reinterpret_cast<uint4*>(sm + PL::DS + tile_off[c]) = o;
        }
      }
      if (PL::DBD >= 0) {
        // inverse _rel_shift: dBD0[r, c0 + jj] = dS[r, KPT qd + jj]; the (tile-invariant) shared-memory addresses were computed once
        uint8_t* dbd = sm + PL::DBD;
        if ((c0 & 1) == 0) {
#pragma unroll
          for (int k = 0; k < KPT / 2; ++k) *reinterpret_cast<uint32_t*>(dbd + pair_off[k]) = pack2(ds[2 * k], ds[2 * k + 1]);
        } else {
          *reinterpret_cast<bf16*>(dbd + single_off[0]) = __float2bfloat16_rn(ds[0]);
#pragma unroll
          for (int k = 0; k < KPT / 2 - 1; ++k) *reinterpret_cast<uint32_t*>(dbd + pair_off[k]) = pack2(ds[2 * k + 1], ds[2 * k + 2]);
          *reinterpret_cast<bf16*>(dbd + single_off[1]) = __float2bfloat16_rn(ds[KPT - 1]);
        }
      }
      fence_proxy_async_smem();
      tc_fence_before();
      mbar_arrive(b_ready);
      t = tn; lse2 = lse2n; dlt = dltn;
    }
    // ======================= drain the accumulators (each thread: its row, KPT of the 64 columns)
    mbar_wait(acc_full, 0);
    tc_fence_after();
    float v0[KPT], v1[KPT];
    tmem_ld_32x16(tmem_base + lane_base + TM_ACC0 + KPT * qd, v0);
    tmem_ld_32x16(tmem_base + lane_base + TM_ACC1 + KPT * qd, v1);
    tmem_ld_wait();
    const Tile tl = it.get(0, a);
    if (MODE == MODE_DQ) {
      const int i = tl.I * BQ + r;
      if (i < g.T) {
        uint4* dst = reinterpret_cast<uint4*>(a.dq + ((int64_t)tl.b * g.T + i) * a.ldq + h * DH + KPT * qd);
#pragma unroll
        for (int c = 0; c < KPT / 8; ++c) {
          uint4 o;
          o.x = pack2(v0[c * 8] + v1[c * 8], v0[c * 8 + 1] + v1[c * 8 + 1]); o.y = pack2(v0[c * 8 + 2] + v1[c * 8 + 2], v0[c * 8 + 3] + v1[c * 8 + 3]);
          o.z = pack2(v0[c * 8 + 4] + v1[c * 8 + 4], v0[c * 8 + 5] + v1[c * 8 + 5]); o.w = pack2(v0[c * 8 + 6] + v1[c * 8 + 6], v0[c * 8 + 7] + v1[c * 8 + 7]);
          dst[c] = o;
        }
      } else {
#pragma unroll
        for (int c = 0; c < KPT; ++c) { v0[c] = 0.f; v1[c] = 0.f; }
      }
      // bias gradients: sums over this warp's 32 rows, one atomic per column per warp
#pragma unroll
      for (int c = 0; c < KPT; ++c) {
        const float s0 = warp_sum(v0[c]), s1 = warp_sum(v1[c]);
        if (lane == c) { atomicAdd(&a.drwb[h * DH + KPT * qd + c], s0); atomicAdd(&a.drrb[h * DH + KPT * qd + c], s1); }
      }
    } else if (MODE == MODE_DKV) {
      if (r < BKV) {
        const int j = tl.J * BKV + r;
        bf16 *dk, *dv;
        if (j < g.mlen) { dk = a.dk_mem + ((int64_t)tl.b * g.mlen + j) * a.ldkv_mem; dv = a.dv_mem + ((int64_t)tl.b * g.mlen + j) * a.ldkv_mem; }
        else { dk = a.dk_cur + ((int64_t)tl.b * g.T + (j - g.mlen)) * a.ldkv_cur; dv = a.dv_cur + ((int64_t)tl.b * g.T + (j - g.mlen)) * a.ldkv_cur; }
        uint4* pk = reinterpret_cast<uint4*>(dk + h * DH + KPT * qd);
        uint4* pv = reinterpret_cast<uint4*>(dv + h * DH + KPT * qd);
#pragma unroll
        for (int c = 0; c < KPT / 8; ++c) {
          uint4 o; o.x = pack2(v0[c * 8], v0[c * 8 + 1]); o.y = pack2(v0[c * 8 + 2], v0[c * 8 + 3]); o.z = pack2(v0[c * 8 + 4], v0[c * 8 + 5]); o.w = pack2(v0[c * 8 + 6], v0[c * 8 + 7]);
          pk[c] = o;
          uint4 u; u.x = pack2(v1[c * 8], v1[c * 8 + 1]); u.y = pack2(v1[c * 8 + 2], v1[c * 8 + 3]); u.z = pack2(v1[c * 8 + 4], v1[c * 8 + 5]); u.w = pack2(v1[c * 8 + 6], v1[c * 8 + 7]);
          pv[c] = u;
        }
      }
    } else {
      const int x0 = g.T - BQ + BKV * it.delta;
      const int xa = x0 + r;                 // acc0: window row r
      if (xa >= 0 && xa < g.klen) {
        float* d = a.dr + (int64_t)xa * HD + h * DH + KPT * qd;
#pragma unroll
        for (int c = 0; c < KPT; ++c) atomicAdd(d + c, v0[c]);
      }
      const int xb = x0 + 64 + r;            // acc1: lanes 64..127 hold window rows 128..191
      if (r >= 64 && xb >= 0 && xb < g.klen) {
        float* d = a.dr + (int64_t)xb * HD + h * DH + KPT * qd;
#pragma unroll
        for (int c = 0; c < KPT; ++c) atomicAdd(d + c, v1[c]);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == W_MMA) tmem_dealloc<TMEM_COLS>(tmem_base);
}


// ------------------------------------------------------------------ lite passes over the stored P / dS tiles
// dK[J] = sum_I dS(I,J)^T . Qw(I),  dV[J] = sum_I P(I,J)^T . dO(I): pure TMA -> tcgen05 pipeline, no recomputation, no thread math.
constexpr int LITE_STAGES = 3, LITE_STAGE = 4 * SZ_Q, LITE_BAR = LITE_STAGES * LITE_STAGE, LITE_SMEM = LITE_BAR + 256 + 1024;
__global__ void __launch_bounds__(192, 1) relattn_bwd_dkv_lite_kernel(const __grid_constant__ Maps M, const BwdArgs a) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  uint8_t* sm = smem_raw + ((1024u - (raw & 1023u)) & 1023u);
  uint64_t* bars = reinterpret_cast<uint64_t*>(sm + LITE_BAR);
  uint64_t *full = bars, *empty = bars + LITE_STAGES, *acc_full = bars + 2 * LITE_STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * LITE_STAGES + 1);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const BandGeom g = make_band(a.band);
  const int J = blockIdx.x, h = blockIdx.y, b = blockIdx.z, nI = (g.T + BQ - 1) / BQ;
  if (J * BKV < g.mlen && a.dk_mem == nullptr) return;
  int Ilo = -1, count = 0;
  for (int i = 0; i < nI; ++i)
    if (q_tile_first_kt(g, i) <= J && J <= q_tile_last_kt(g, i)) { if (Ilo < 0) Ilo = i; ++count; }
  auto out_ptrs = [&](int j, bf16*& dk, bf16*& dv) {
    if (j < g.mlen) { dk = a.dk_mem + ((int64_t)b * g.mlen + j) * a.ldkv_mem; dv = a.dv_mem + ((int64_t)b * g.mlen + j) * a.ldkv_mem; }
    else { dk = a.dk_cur + ((int64_t)b * g.T + (j - g.mlen)) * a.ldkv_cur; dv = a.dv_cur + ((int64_t)b * g.T + (j - g.mlen)) * a.ldkv_cur; }
  };
  if (count == 0) {
    if (tid < BKV) { bf16 *dk, *dv; out_ptrs(J * BKV + tid, dk, dv); for (int c = 0; c < DH; ++c) { dk[h * DH + c] = __float2bfloat16(0.f); dv[h * DH + c] = __float2bfloat16(0.f); } }
    return;
  }
  if (tid == 0) {
    for (int s = 0; s < LITE_STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    mbar_init(acc_full, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<128>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (warp == 0) {
    if (lane == 0) {
      for (int n = 0; n < count; ++n) {
        const int s = n % LITE_STAGES; const uint32_t rph = (n / LITE_STAGES) & 1;
        const int I = Ilo + n;
        mbar_wait(&empty[s], rph ^ 1);
        uint8_t* st = sm + s * LITE_STAGE;
        mbar_expect_tx(&full[s], LITE_STAGE);
        const int trow = ((((b * a.H + h) * nI + I) * a.nt_max) + (J - q_tile_first_kt(g, I))) * BQ;
        const int qrow = b * g.T + I * BQ;
        tma_load_2d(st, &M.pkv, &full[s], 0, trow);                    // P  (I,J)   [128 q x 64 keys]  (or P~, with dO pre-scaled per row)
        tma_load_2d(st + SZ_Q, &M.dst, &full[s], 0, trow);             // dS (I,J)
        tma_load_2d(st + 2 * SZ_Q, &M.qw, &full[s], h * DH, qrow);     // Qw (I)     [128 q x 64]
        tma_load_2d(st + 3 * SZ_Q, &M.dOkv, &full[s], h * DH, qrow);   // dO (I)
      }
    }
  } else if (warp == 1) {
    {   // whole warp, converged: one elected lane issues (umma_bf16_warp)
      const uint32_t id_nn = umma_idesc_bf16(BQ, DH, 1, 1);
      for (int n = 0; n < count; ++n) {
        const int s = n % LITE_STAGES; const uint32_t rph = (n / LITE_STAGES) & 1;
        mbar_wait(&full[s], rph);
        tc_fence_after();
        const uint32_t st = smem_u32(sm + s * LITE_STAGE);
#pragma unroll
        for (int k = 0; k < 8; ++k)      // dK += dS^T . Qw   (second MN atom of the A operand = the Qw tile behind dS: rows 64..127 ignored)
          umma_bf16_warp(tmem_base, umma_smem_desc(st + SZ_Q + k * 2048, 16384, 1024), umma_smem_desc(st + 2 * SZ_Q + k * 2048, 8192, 1024), id_nn, (n > 0) | (k > 0));
#pragma unroll
        for (int k = 0; k < 8; ++k)      // dV += P^T . dO
          umma_bf16_warp(tmem_base + 64, umma_smem_desc(st + k * 2048, 16384, 1024), umma_smem_desc(st + 3 * SZ_Q + k * 2048, 8192, 1024), id_nn, (n > 0) | (k > 0));
        umma_commit_warp(&empty[s]);
      }
      umma_commit_warp(acc_full);
    }
  } else {
    const int q4 = warp & 3, r = 32 * q4 + lane;     // warps 2..5 -> lane quadrants 2,3,0,1
    mbar_wait(acc_full, 0);
    tc_fence_after();
    if (r < BKV) {
      bf16 *dk, *dv;
      out_ptrs(J * BKV + r, dk, dv);
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        float v0[32], v1[32];
        tmem_ld_32x32(tmem_base + ((uint32_t)(32 * q4) << 16) + 32 * half, v0);
        tmem_ld_32x32(tmem_base + ((uint32_t)(32 * q4) << 16) + 64 + 32 * half, v1);
        tmem_ld_wait();
        uint4* pk = reinterpret_cast<uint4*>(dk + h * DH + 32 * half);
        uint4* pv = reinterpret_cast<uint4*>(dv + h * DH + 32 * half);
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          uint4 o; o.x = pack2(v0[c * 8], v0[c * 8 + 1]); o.y = pack2(v0[c * 8 + 2], v0[c * 8 + 3]); o.z = pack2(v0[c * 8 + 4], v0[c * 8 + 5]); o.w = pack2(v0[c * 8 + 6], v0[c * 8 + 7]);
          pk[c] = o;
          uint4 u; u.x = pack2(v1[c * 8], v1[c * 8 + 1]); u.y = pack2(v1[c * 8 + 2], v1[c * 8 + 3]); u.z = pack2(v1[c * 8 + 4], v1[c * 8 + 5]); u.w = pack2(v1[c * 8 + 6], v1[c * 8 + 7]);
          pv[c] = u;
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<128>(tmem_base);
}

// The same for TWO adjacent key tiles (128 keys) per CTA: the P / dS tiles of (I, 2Jp) and (I, 2Jp+1) sit next to each other in the
// tile store, one 256-row TMA box brings both, and side by side they are exactly the two MN atoms of a full M = 128 A operand.
// The 64-key kernel above wastes half of every MMA (rows 64..127 of its accumulators are ignored) and re-reads Qw / dO per key tile;
// it is bound by shared-memory bandwidth (160 KB per key tile), this one moves 96 KB per key tile.  Needs first_kt(I) and the tile
// count of every query tile to be even (mlen a multiple of 128).
constexpr int PAIR_STAGES = 2, PAIR_STAGE = 6 * SZ_Q, PAIR_BAR = PAIR_STAGES * PAIR_STAGE, PAIR_SMEM = PAIR_BAR + 128 + 1024;
__global__ void __launch_bounds__(192, 1) relattn_bwd_dkv_pair_kernel(const __grid_constant__ Maps M, const BwdArgs a) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  uint8_t* sm = smem_raw + ((1024u - (raw & 1023u)) & 1023u);
  uint64_t* bars = reinterpret_cast<uint64_t*>(sm + PAIR_BAR);
  uint64_t *full = bars, *empty = bars + PAIR_STAGES, *acc_full = bars + 2 * PAIR_STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * PAIR_STAGES + 1);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const BandGeom g = make_band(a.band);
  const int Jp = blockIdx.x, J = 2 * Jp, h = blockIdx.y, b = blockIdx.z, nI = (g.T + BQ - 1) / BQ;
  if (J * BKV + 2 * BKV <= g.mlen && a.dk_mem == nullptr) return;      // both key tiles lie in the detached mems
  int Ilo = -1, count = 0;
  for (int i = 0; i < nI; ++i)
    if (q_tile_first_kt(g, i) <= J && J <= q_tile_last_kt(g, i)) { if (Ilo < 0) Ilo = i; ++count; }
  auto out_ptrs = [&](int j, bf16*& dk, bf16*& dv) -> bool {
    if (j < g.mlen) {
      if (a.dk_mem == nullptr) return false;
      dk = a.dk_mem + ((int64_t)b * g.mlen + j) * a.ldkv_mem; dv = a.dv_mem + ((int64_t)b * g.mlen + j) * a.ldkv_mem;
    } else { dk = a.dk_cur + ((int64_t)b * g.T + (j - g.mlen)) * a.ldkv_cur; dv = a.dv_cur + ((int64_t)b * g.T + (j - g.mlen)) * a.ldkv_cur; }
    return true;
  };
  if (count == 0) {
    if (tid < 2 * BKV) { bf16 *dk, *dv; if (out_ptrs(J * BKV + tid, dk, dv)) for (int c = 0; c < DH; ++c) { dk[h * DH + c] = __float2bfloat16(0.f); dv[h * DH + c] = __float2bfloat16(0.f); } }
    return;
  }
  if (tid == 0) {
    for (int s = 0; s < PAIR_STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    mbar_init(acc_full, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<128>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (warp == 0) {
    if (lane == 0) {
      for (int n = 0; n < count; ++n) {
        const int s = n % PAIR_STAGES; const uint32_t rph = (n / PAIR_STAGES) & 1;
        const int I = Ilo + n;
        mbar_wait(&empty[s], rph ^ 1);
        uint8_t* st = sm + s * PAIR_STAGE;
        mbar_expect_tx(&full[s], PAIR_STAGE);
        const int trow = ((((b * a.H + h) * nI + I) * a.nt_max) + (J - q_tile_first_kt(g, I))) * BQ;
        const int qrow = b * g.T + I * BQ;
        tma_load_2d(st, &M.pkv2, &full[s], 0, trow);                    // P  (I,J), (I,J+1)   2 x [128 q x 64 keys]  (or P~)
        tma_load_2d(st + 2 * SZ_Q, &M.dst2, &full[s], 0, trow);         // dS (I,J), (I,J+1)
        tma_load_2d(st + 4 * SZ_Q, &M.qw, &full[s], h * DH, qrow);      // Qw (I)     [128 q x 64]
        tma_load_2d(st + 5 * SZ_Q, &M.dOkv, &full[s], h * DH, qrow);    // dO (I)
      }
    }
  } else if (warp == 1) {
    {   // whole warp, converged: one elected lane issues (umma_bf16_warp)
      const uint32_t id_nn = umma_idesc_bf16(BQ, DH, 1, 1);
      for (int n = 0; n < count; ++n) {
        const int s = n % PAIR_STAGES; const uint32_t rph = (n / PAIR_STAGES) & 1;
        mbar_wait(&full[s], rph);
        tc_fence_after();
        const uint32_t st = smem_u32(sm + s * PAIR_STAGE);
#pragma unroll
        for (int k = 0; k < 8; ++k)      // dK[128 keys] += dS^T . Qw   (the two key tiles are the two MN atoms of A, 16 KB apart)
          umma_bf16_warp(tmem_base, umma_smem_desc(st + 2 * SZ_Q + k * 2048, 16384, 1024), umma_smem_desc(st + 4 * SZ_Q + k * 2048, 8192, 1024), id_nn, (n > 0) | (k > 0));
#pragma unroll
        for (int k = 0; k < 8; ++k)      // dV[128 keys] += P^T . dO
          umma_bf16_warp(tmem_base + 64, umma_smem_desc(st + k * 2048, 16384, 1024), umma_smem_desc(st + 5 * SZ_Q + k * 2048, 8192, 1024), id_nn, (n > 0) | (k > 0));
        umma_commit_warp(&empty[s]);
      }
      umma_commit_warp(acc_full);
    }
  } else {
    const int q4 = warp & 3, r = 32 * q4 + lane;     // warps 2..5 -> lane quadrants 2,3,0,1; TMEM lane r = key J*64 + r
    mbar_wait(acc_full, 0);
    tc_fence_after();
    bf16 *dk, *dv;
    const bool wr = out_ptrs(J * BKV + r, dk, dv);
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      float v0[32], v1[32];
      tmem_ld_32x32(tmem_base + ((uint32_t)(32 * q4) << 16) + 32 * half, v0);
      tmem_ld_32x32(tmem_base + ((uint32_t)(32 * q4) << 16) + 64 + 32 * half, v1);
      tmem_ld_wait();
      if (wr) {
        uint4* pk = reinterpret_cast<uint4*>(dk + h * DH + 32 * half);
        uint4* pv = reinterpret_cast<uint4*>(dv + h * DH + 32 * half);
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          uint4 o; o.x = pack2(v0[c * 8], v0[c * 8 + 1]); o.y = pack2(v0[c * 8 + 2], v0[c * 8 + 3]); o.z = pack2(v0[c * 8 + 4], v0[c * 8 + 5]); o.w = pack2(v0[c * 8 + 6], v0[c * 8 + 7]);
          pk[c] = o;
          uint4 u; u.x = pack2(v1[c * 8], v1[c * 8 + 1]); u.y = pack2(v1[c * 8 + 2], v1[c * 8 + 3]); u.z = pack2(v1[c * 8 + 4], v1[c * 8 + 5]); u.w = pack2(v1[c * 8 + 6], v1[c * 8 + 7]);
          pv[c] = u;
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<128>(tmem_base);
}

// dRwin(diagonal) = sum_{I on the diagonal, b} dBD0^T . Qr, with dBD0 rebuilt from the stored dS tile by the inverse _rel_shift
// (a per-row element offset on a shared-memory to shared-memory copy).
constexpr int DRL_STAGES = 3, DRL_STAGE = 2 * SZ_Q, DRL_DBD = DRL_STAGES * DRL_STAGE, DRL_BAR = DRL_DBD + 2 * SZ_DBD, DRL_SMEM = DRL_BAR + 128 + 1024;
constexpr int DRL_THREADS = N_SOFTMAX + 64;   // no store warp here
__global__ void __launch_bounds__(DRL_THREADS, 1) relattn_bwd_dr_lite_kernel(const __grid_constant__ Maps M, const BwdArgs a) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  uint8_t* sm = smem_raw + ((1024u - (raw & 1023u)) & 1023u);
  uint64_t* bars = reinterpret_cast<uint64_t*>(sm + DRL_BAR);
  uint64_t *full = bars, *empty = bars + DRL_STAGES, *b_ready = bars + 2 * DRL_STAGES, *b_done = b_ready + 2, *acc_full = b_ready + 4;   // b_ready/b_done: one per dBD buffer
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(b_ready + 5);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const BandGeom g = make_band(a.band);
  const int HD = a.H * DH, nI = (g.T + BQ - 1) / BQ;
  const int delta = a.delta_min + (int)blockIdx.x, h = blockIdx.y;
  int Ilo = -1, cntI = 0;
  for (int i = 0; i < nI; ++i) {
    int j = 2 * i + delta;
    if (q_tile_first_kt(g, i) <= j && j <= q_tile_last_kt(g, i)) { if (Ilo < 0) Ilo = i; ++cntI; }
  }
  const int count = cntI * a.B;
  if (count == 0) return;
  if (tid == 0) {
    for (int s = 0; s < DRL_STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1 + N_SOFTMAX); }
    mbar_init(&b_ready[0], N_SOFTMAX); mbar_init(&b_ready[1], N_SOFTMAX); mbar_init(&b_done[0], 1); mbar_init(&b_done[1], 1); mbar_init(acc_full, 1);
    fence_barrier_init();
  }
  if (warp == W_MMA) tmem_alloc<128>(tmem_slot);
  for (int e = tid; e < 2 * SZ_DBD / 16; e += DRL_THREADS) reinterpret_cast<uint4*>(sm + DRL_DBD)[e] = make_uint4(0, 0, 0, 0);
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (warp == W_PROD) {
    if (lane == 0) {
      for (int n = 0; n < count; ++n) {
        const int s = n % DRL_STAGES; const uint32_t rph = (n / DRL_STAGES) & 1;
        const int I = Ilo + n / a.B, b = n % a.B, J = 2 * I + delta;
        mbar_wait(&empty[s], rph ^ 1);
        uint8_t* st = sm + s * DRL_STAGE;
        mbar_expect_tx(&full[s], DRL_STAGE);
        const int trow = ((((b * a.H + h) * nI + I) * a.nt_max) + (J - q_tile_first_kt(g, I))) * BQ;
        tma_load_2d(st, &M.dst, &full[s], 0, trow);                              // dS (I,J)
        tma_load_2d(st + SZ_Q, &M.qr, &full[s], h * DH, b * g.T + I * BQ);       // Qr (I)
      }
    }
  } else if (warp == W_MMA) {
    {   // whole warp, converged: one elected lane issues (umma_bf16_warp)
      const uint32_t id_nn = umma_idesc_bf16(BQ, DH, 1, 1);
      for (int n = 0; n < count; ++n) {
        const int s = n % DRL_STAGES, buf = n & 1;
        const uint32_t dbd = smem_u32(sm + DRL_DBD + buf * SZ_DBD);
        const uint32_t qr = smem_u32(sm + s * DRL_STAGE + SZ_Q);
        mbar_wait(&b_ready[buf], (n >> 1) & 1);
        tc_fence_after();
#pragma unroll
        for (int k = 0; k < 8; ++k)
          umma_bf16_warp(tmem_base, umma_smem_desc(dbd + k * 2048, 16384, 1024), umma_smem_desc(qr + k * 2048, 8192, 1024), id_nn, (n > 0) | (k > 0));
#pragma unroll
        for (int k = 0; k < 8; ++k)
          umma_bf16_warp(tmem_base + 64, umma_smem_desc(dbd + 16384 + k * 2048, 16384, 1024), umma_smem_desc(qr + k * 2048, 8192, 1024), id_nn, (n > 0) | (k > 0));
        umma_commit_warp(&b_done[buf]);
        umma_commit_warp(&empty[s]);
      }
      umma_commit_warp(acc_full);
    }
  } else {
    const int r = 32 * (warp & 3) + lane, qd = warp >> 2;
    const uint32_t lane_base = (uint32_t)(32 * (warp & 3)) << 16;
    const int c0 = 127 - r + KPT * qd;
    for (int n = 0; n < count; ++n) {
      const int s = n % DRL_STAGES, buf = n & 1; const uint32_t rph = (n / DRL_STAGES) & 1;
      uint8_t* dbd = sm + DRL_DBD + buf * SZ_DBD;
      auto addr = [&](int c) -> uint8_t* { return dbd + (c >> 6) * 16384 + r * 128 + ((((c & 63) >> 3) ^ (r & 7)) << 4) + (c & 7) * 2; };
      mbar_wait(&full[s], rph);
      // this thread's 16 dS values (bf16) out of the swizzled K-major tile
      const uint8_t* tile = sm + s * DRL_STAGE;
      uint4 u0 = *reinterpret_cast<const uint4*>(tile + r * 128 + (((2 * qd) ^ (r & 7)) << 4));
      uint4 u1 = *reinterpret_cast<const uint4*>(tile + r * 128 + (((2 * qd + 1) ^ (r & 7)) << 4));
      mbar_arrive(&empty[s]);                        // dS tile consumed by this thread (Qr is released by the MMA commit)
      uint32_t wv[8] = {u0.x, u0.y, u0.z, u0.w, u1.x, u1.y, u1.z, u1.w};
      if (n > 1) mbar_wait(&b_done[buf], ((n - 2) >> 1) & 1);      // the MMAs that read this dBD buffer two tiles ago are done
      if ((c0 & 1) == 0) {
#pragma unroll
        for (int k = 0; k < 8; ++k) *reinterpret_cast<uint32_t*>(addr(c0 + 2 * k)) = wv[k];
      } else {
        *reinterpret_cast<uint16_t*>(addr(c0)) = (uint16_t)(wv[0] & 0xFFFFu);
#pragma unroll
        for (int k = 0; k < 7; ++k) *reinterpret_cast<uint32_t*>(addr(c0 + 1 + 2 * k)) = (wv[k] >> 16) | (wv[k + 1] << 16);
        *reinterpret_cast<uint16_t*>(addr(c0 + 15)) = (uint16_t)(wv[7] >> 16);
      }
      fence_proxy_async_smem();
      mbar_arrive(&b_ready[buf]);
    }
    mbar_wait(acc_full, 0);
    tc_fence_after();
    float v0[KPT], v1[KPT];
    tmem_ld_32x16(tmem_base + lane_base + KPT * qd, v0);
    tmem_ld_32x16(tmem_base + lane_base + 64 + KPT * qd, v1);
    tmem_ld_wait();
    const int x0 = g.T - BQ + BKV * delta;
    const int xa = x0 + r;
    if (xa >= 0 && xa < g.klen) {
      float* d = a.dr + (int64_t)xa * HD + h * DH + KPT * qd;
#pragma unroll
      for (int c = 0; c < KPT; ++c) atomicAdd(d + c, v0[c]);
    }
    const int xb = x0 + 64 + r;
    if (r >= 64 && xb >= 0 && xb < g.klen) {
      float* d = a.dr + (int64_t)xb * HD + h * DH + KPT * qd;
#pragma unroll
      for (int c = 0; c < KPT; ++c) atomicAdd(d + c, v1[c]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == W_MMA) tmem_dealloc<128>(tmem_base);
}

// The same for TWO adjacent band diagonals per CTA (delta, delta+1 with delta even): their windows overlap in 128 of 192 rows, the union
// is 256 window rows = two fully used M = 128 accumulators (the single-diagonal kernel computes rows 64..127 twice), the dS tiles
// (I, J) and (I, J+1) arrive in one 256-row TMA box, and Qr(I) is staged once for both.  104 KB of shared-memory traffic per band tile
// instead of 160 KB.  The (I, b) loop is split over gridDim.z CTAs to keep every SM busy; partial dR sums meet in fp32 atomics.
constexpr int DRP_STAGES = 2, DRP_STAGE = 3 * SZ_Q, DRP_SZ_DBD = 4 * SZ_Q, DRP_DBD = DRP_STAGES * DRP_STAGE, DRP_BAR = DRP_DBD + 2 * DRP_SZ_DBD,
              DRP_SMEM = DRP_BAR + 128 + 1024;
static_assert(DRP_SMEM <= 232448, "dR pair kernel shared-memory plan exceeds 227 KB");
__global__ void __launch_bounds__(DRL_THREADS, 1) relattn_bwd_dr_pair_kernel(const __grid_constant__ Maps M, const BwdArgs a) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  uint8_t* sm = smem_raw + ((1024u - (raw & 1023u)) & 1023u);
  uint64_t* bars = reinterpret_cast<uint64_t*>(sm + DRP_BAR);
  uint64_t *full = bars, *empty = bars + DRP_STAGES, *b_ready = bars + 2 * DRP_STAGES, *b_done = b_ready + 2, *acc_full = b_ready + 4;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(b_ready + 5);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const BandGeom g = make_band(a.band);
  const int HD = a.H * DH, nI = (g.T + BQ - 1) / BQ;
  const int delta = a.delta_min + 2 * (int)blockIdx.x, h = blockIdx.y;
  const int nb = (a.B + (int)gridDim.z - 1) / (int)gridDim.z, b_lo = (int)blockIdx.z * nb, b_cnt = min(nb, a.B - b_lo);
  int Ilo = -1, cntI = 0;
  for (int i = 0; i < nI; ++i) {
    int j = 2 * i + delta;
    if (q_tile_first_kt(g, i) <= j && j <= q_tile_last_kt(g, i)) { if (Ilo < 0) Ilo = i; ++cntI; }
  }
  const int count = b_cnt > 0 ? cntI * b_cnt : 0;
  if (count == 0) return;
  if (tid == 0) {
    for (int s = 0; s < DRP_STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1 + N_SOFTMAX / 32); }
    mbar_init(&b_ready[0], N_SOFTMAX / 32); mbar_init(&b_ready[1], N_SOFTMAX / 32); mbar_init(&b_done[0], 1); mbar_init(&b_done[1], 1); mbar_init(acc_full, 1);
    fence_barrier_init();
  }
  if (warp == W_MMA) tmem_alloc<128>(tmem_slot);
  for (int e = tid; e < 2 * DRP_SZ_DBD / 16; e += DRL_THREADS) reinterpret_cast<uint4*>(sm + DRP_DBD)[e] = make_uint4(0, 0, 0, 0);
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (warp == W_PROD) {
    if (lane == 0) {
      for (int n = 0; n < count; ++n) {
        const int s = n % DRP_STAGES; const uint32_t rph = (n / DRP_STAGES) & 1;
        const int I = Ilo + n / b_cnt, b = b_lo + n % b_cnt, J = 2 * I + delta;
        mbar_wait(&empty[s], rph ^ 1);
        uint8_t* st = sm + s * DRP_STAGE;
        mbar_expect_tx(&full[s], DRP_STAGE);
        const int trow = ((((b * a.H + h) * nI + I) * a.nt_max) + (J - q_tile_first_kt(g, I))) * BQ;
        tma_load_2d(st, &M.dst2, &full[s], 0, trow);                                 // dS (I,J), (I,J+1)
        tma_load_2d(st + 2 * SZ_Q, &M.qr, &full[s], h * DH, b * g.T + I * BQ);       // Qr (I)
      }
    }
  } else if (warp == W_MMA) {
    {   // whole warp, converged: one elected lane issues (umma_bf16_warp)
      const uint32_t id_nn = umma_idesc_bf16(BQ, DH, 1, 1);
      for (int n = 0; n < count; ++n) {
        const int s = n % DRP_STAGES, buf = n & 1;
        const uint32_t dbd = smem_u32(sm + DRP_DBD + buf * DRP_SZ_DBD);
        const uint32_t qr = smem_u32(sm + s * DRP_STAGE + 2 * SZ_Q);
        mbar_wait(&b_ready[buf], (n >> 1) & 1);
        tc_fence_after();
#pragma unroll
        for (int k = 0; k < 8; ++k)      // dRwin[  0..127] += dBD0[:,   0..127]^T . Qr
          umma_bf16_warp(tmem_base, umma_smem_desc(dbd + k * 2048, 16384, 1024), umma_smem_desc(qr + k * 2048, 8192, 1024), id_nn, (n > 0) | (k > 0));
#pragma unroll
        for (int k = 0; k < 8; ++k)      // dRwin[128..255] += dBD0[:, 128..255]^T . Qr
          umma_bf16_warp(tmem_base + 64, umma_smem_desc(dbd + 2 * 16384 + k * 2048, 16384, 1024), umma_smem_desc(qr + k * 2048, 8192, 1024), id_nn, (n > 0) | (k > 0));
        umma_commit_warp(&b_done[buf]);
        umma_commit_warp(&empty[s]);
      }
      umma_commit_warp(acc_full);
    }
  } else {
    const int r = 32 * (warp & 3) + lane, qd = warp >> 2;
    const uint32_t lane_base = (uint32_t)(32 * (warp & 3)) << 16;
    const int c0 = 127 - r + KPT * qd;
    // tile-invariant shared-memory offsets: this thread's two 16-byte chunks of a K-major dS tile, and where its bf16 pairs land in the
    // window-space tile (tile t of the pair sits 64 window columns further right)
    const uint32_t ld_off0 = (uint32_t)(r * 128 + (((2 * qd) ^ (r & 7)) << 4)), ld_off1 = (uint32_t)(r * 128 + (((2 * qd + 1) ^ (r & 7)) << 4));
    auto dbd_off = [&](int c) -> uint32_t { return (uint32_t)((c >> 6) * 16384 + r * 128 + ((((c & 63) >> 3) ^ (r & 7)) << 4) + (c & 7) * 2); };
    const bool odd = c0 & 1;
    uint32_t w_off[9];       // word (or half-word at the two ends when c0 is odd) positions for tile 0; tile 1 = +16384 bytes (one 64-column atom)
#pragma unroll
    for (int k = 0; k < 9; ++k) w_off[k] = dbd_off(odd ? (k == 0 ? c0 : c0 + 2 * k - 1) : c0 + 2 * (k < 8 ? k : 7));
    for (int n = 0; n < count; ++n) {
      const int s = n % DRP_STAGES, buf = n & 1; const uint32_t rph = (n / DRP_STAGES) & 1;
      uint8_t* dbd = sm + DRP_DBD + buf * DRP_SZ_DBD;
      mbar_wait(&full[s], rph);
      const uint8_t* tile = sm + s * DRP_STAGE;
      uint32_t wv[2][8];
#pragma unroll
      for (int t = 0; t < 2; ++t) {
        const uint4 u0 = *reinterpret_cast<const uint4*>(tile + t * SZ_Q + ld_off0);
        const uint4 u1 = *reinterpret_cast<const uint4*>(tile + t * SZ_Q + ld_off1);
        wv[t][0] = u0.x; wv[t][1] = u0.y; wv[t][2] = u0.z; wv[t][3] = u0.w; wv[t][4] = u1.x; wv[t][5] = u1.y; wv[t][6] = u1.z; wv[t][7] = u1.w;
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty[s]);         // dS tiles consumed by this warp (Qr is released by the MMA commit)
      if (n > 1) mbar_wait(&b_done[buf], ((n - 2) >> 1) & 1);      // the MMAs that read this dBD buffer two steps ago are done
#pragma unroll
      for (int t = 0; t < 2; ++t) {
        uint8_t* base = dbd + t * 16384;
        if (!odd) {
#pragma unroll
          for (int k = 0; k < 8; ++k) *reinterpret_cast<uint32_t*>(base + w_off[k]) = wv[t][k];
        } else {
          *reinterpret_cast<uint16_t*>(base + w_off[0]) = (uint16_t)(wv[t][0] & 0xFFFFu);
#pragma unroll
          for (int k = 0; k < 7; ++k) *reinterpret_cast<uint32_t*>(base + w_off[k + 1]) = (wv[t][k] >> 16) | (wv[t][k + 1] << 16);
          *reinterpret_cast<uint16_t*>(base + w_off[8]) = (uint16_t)(wv[t][7] >> 16);
        }
      }
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(&b_ready[buf]);
    }
    mbar_wait(acc_full, 0);
    tc_fence_after();
    float v0[KPT], v1[KPT];
    tmem_ld_32x16(tmem_base + lane_base + KPT * qd, v0);
    tmem_ld_32x16(tmem_base + lane_base + 64 + KPT * qd, v1);
    tmem_ld_wait();
    const int x0 = g.T - BQ + BKV * delta;
    const int xa = x0 + r, xb = x0 + BQ + r;
    if (xa >= 0 && xa < g.klen) {
      float4* d = reinterpret_cast<float4*>(a.dr + (int64_t)xa * HD + h * DH + KPT * qd);
#pragma unroll
      for (int c = 0; c < KPT / 4; ++c) atomicAdd(d + c, make_float4(v0[4 * c], v0[4 * c + 1], v0[4 * c + 2], v0[4 * c + 3]));
    }
    if (xb >= 0 && xb < g.klen) {
      float4* d = reinterpret_cast<float4*>(a.dr + (int64_t)xb * HD + h * DH + KPT * qd);
#pragma unroll
      for (int c = 0; c < KPT / 4; ++c) atomicAdd(d + c, make_float4(v1[4 * c], v1[4 * c + 1], v1[4 * c + 2], v1[4 * c + 3]));
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == W_MMA) tmem_dealloc<128>(tmem_base);
}

// ------------------------------------------------------------------ dQ pass over the P~ tiles the forward kernel saved
// The forward pass leaves, per band tile, the bf16 numerators P~ = exp2(score - m) it fed to P.V and the running max m it used.
// With the final log-sum-exp, P = P~ * exp2(m - lse): no S / BD0 recomputation, no _rel_shift, no mask arithmetic, no exp per
// element.  Per tile only dP = dO.V^T is formed on the tensor cores (double-buffered in TMEM); the 512 softmax threads turn
// (P~, dP) into the bf16 P / dS / dBD0 work tiles, and dQw += dS.K, dQr += dBD0.Rwin accumulate in TMEM as before.
// (Tried and slower, see profiles/README.md: P~ through per-thread global loads with double-buffered work tiles; P / dS tile stores
// straight from registers; the lite dK/dV kernel rebuilding P.dO from P~ by scaling dO rows.)
// Operand traffic is what bounds this pass (a 2-stage ring put the ~2300-cycle TMA round trip on the critical path: measured
// period = TMA latency + back-end MMA issue), so: K/V/P~ stream through a 3-stage ring, and consecutive R windows — which
// overlap by 128 of their 192 rows — live in a 4-slot ring of 64-row chunks, one new 8 KB chunk per tile instead of 24 KB.
struct PlanS {   // resident dO | R ring 4 x 64 rows | ring {K V P~} x 3 | dS dBD P
  static constexpr int NST = 3;
  static constexpr int DO = 0, RRING = SZ_Q, RING = RRING + 4 * SZ_KV, STAGE = 2 * SZ_KV + SZ_Q;
  static constexpr int K = 0, V = SZ_KV, PT = 2 * SZ_KV;                                     // offsets inside a stage
  static constexpr int DS = RING + NST * STAGE, DBD = DS + SZ_Q, P = DBD + SZ_DBD, BAR = P + SZ_Q;
  static constexpr int SMEM = BAR + 256 + 1024;
};
static_assert(PlanS::SMEM <= 232448, "dQ pass shared-memory plan exceeds 227 KB");
constexpr int TS_DP = 0, TS_ACC0 = 128, TS_ACC1 = 192, TS_COLS = 256;
__device__ long long g_ts[3][32][8];     // TXL_ABL & 256: per-tile timestamps of one CTA — softmax thread 0, MMA warp, TMA producer
#define TS(role, n, k) do { if ((a.abl & 256) && blockIdx.x == 3 && blockIdx.y == 0 && blockIdx.z == 0 && (n) < 32) g_ts[role][n][k] = clock64(); } while (0)

__global__ void __launch_bounds__(NTHREADS, 1) relattn_bwd_dq_saved_kernel(const __grid_constant__ Maps M, const BwdArgs a) {
  using PL = PlanS;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  uint8_t* sm = smem_raw + ((1024u - (raw & 1023u)) & 1023u);
  uint64_t* bars = reinterpret_cast<uint64_t*>(sm + PL::BAR);
  uint64_t *full = bars, *empty = bars + 3, *rfull = bars + 6, *rfree = bars + 10, *f_full = bars + 14, *t_free = bars + 16;
  uint64_t *res_full = bars + 18, *acc_full = bars + 19, *b_ready = bars + 20, *b_done = bars + 21, *dbd_done = bars + 22;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 24);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const BandGeom g = make_band(a.band);
  const int I = blockIdx.x, h = blockIdx.y, b = blockIdx.z, nI = gridDim.x;
  const int J0 = q_tile_first_kt(g, I), count = q_tile_last_kt(g, I) - J0 + 1;
  const int tile0 = ((b * a.H + h) * nI + I) * a.nt_max;
  const int xbase = g.T - BQ - I * BQ + J0 * BKV;          // R row of window column 0 of band tile 0; chunk c = rows [xbase + 64 c, +64)

  if (tid == 0) {
    for (int s = 0; s < PL::NST; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    for (int s = 0; s < 4; ++s) { mbar_init(&rfull[s], 1); mbar_init(&rfree[s], 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(&f_full[s], 1); mbar_init(&t_free[s], N_SOFTMAX / 32); }
    mbar_init(res_full, 1); mbar_init(acc_full, 1);
    mbar_init(b_ready, N_SOFTMAX / 32); mbar_init(b_done, 2);    // b_done: back-end MMAs finished + the tile store has read the work tiles
    mbar_init(dbd_done, 1);                                      // the 12 dQr MMAs (issued first) are done with the window tile
    fence_barrier_init();
  }
  if (warp == W_MMA) tmem_alloc<TS_COLS>(tmem_slot);
  for (int e = tid; e < SZ_DBD / 16; e += NTHREADS) reinterpret_cast<uint4*>(sm + PL::DBD)[e] = make_uint4(0, 0, 0, 0);
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == W_PROD) {
    if (lane == 0) {
      auto load_chunk = [&](int c) {
        const int sl = c & 3;
        if (c >= 4) mbar_wait(&rfree[sl], ((c >> 2) - 1) & 1);      // the window that last used this slot (tile c-4) is done
        mbar_expect_tx(&rfull[sl], SZ_KV);
        tma_load_2d(sm + PL::RRING + sl * SZ_KV, &M.r64, &rfull[sl], h * DH, xbase + BKV * c);   // rows outside [0, klen): zero fill
      };
      auto load_stage = [&](int n) {
        const int s = n % PL::NST;
        const int j0 = (J0 + n) * BKV;
        if (n >= PL::NST) mbar_wait(&empty[s], ((n / PL::NST) - 1) & 1);   // back end of tile n-3 has read this stage
        TS(2, n, 0);
        uint8_t* st = sm + PL::RING + s * PL::STAGE;
        mbar_expect_tx(&full[s], PL::STAGE);
        if (j0 < g.mlen) { tma_load_2d(st + PL::K, &M.km, &full[s], h * DH, b * g.mlen + j0); tma_load_2d(st + PL::V, &M.vm, &full[s], h * DH, b * g.mlen + j0); }
        else { tma_load_2d(st + PL::K, &M.kc, &full[s], h * DH, b * g.T + (j0 - g.mlen)); tma_load_2d(st + PL::V, &M.vc, &full[s], h * DH, b * g.T + (j0 - g.mlen)); }
        tma_load_2d(st + PL::PT, &M.psv, &full[s], 0, (tile0 + n) * BQ);
      };
      mbar_expect_tx(res_full, SZ_Q);
      tma_load_2d(sm + PL::DO, &M.dO, res_full, h * DH, b * g.T + I * BQ);
      // band tile n reads window chunks n, n+1, n+2: chunks 0 .. count+1 in all.  Loads are issued in the order their slots free up:
      // the back end of tile m releases stage m mod 3 (-> tile m+3) and window chunk m (-> chunk m+4)
      const int nchunks = count + 2;
      for (int c = 0; c < 4 && c < nchunks; ++c) load_chunk(c);
      for (int n = 0; n < PL::NST && n < count; ++n) load_stage(n);
      for (int m = 0; m < count; ++m) {
        if (m + PL::NST < count) load_stage(m + PL::NST);
        if (m + 4 < nchunks) load_chunk(m + 4);
      }
    }
  } else if (warp == W_MMA) {
    {   // whole warp, converged: one elected lane issues (umma_bf16_warp)
      const uint32_t id_s = umma_idesc_bf16(BQ, BKV, 0, 0), id_kn = umma_idesc_bf16(BQ, DH, 0, 1);
      const uint32_t base = smem_u32(sm);
      mbar_wait(res_full, 0);
      auto front = [&](int n) {     // dP of band tile n into its TMEM buffer
        const int s = n % PL::NST, tb = n & 1;
        const uint32_t vv = base + PL::RING + s * PL::STAGE + PL::V, dO = base + PL::DO;
        mbar_wait(&full[s], (n / PL::NST) & 1);
        TS(1, n, 0);
        if (n >= 2) mbar_wait(&t_free[tb], ((n - 2) >> 1) & 1);
        TS(1, n, 1);
        tc_fence_after();
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_bf16_warp(tmem_base + TS_DP + 64 * tb, umma_smem_desc(dO + k * 32, 16, 1024), umma_smem_desc(vv + k * 32, 16, 1024), id_s, k > 0);
        umma_commit_warp(&f_full[tb]);
      };
      front(0);
      int fronts = 1;
      for (int n = 0; n < count; ++n) {
        const int s = n % PL::NST;
        const uint32_t st = base + PL::RING + s * PL::STAGE;
        // front end of tile n+1 ahead of tile n's back end only if its operands have already landed: never park behind a TMA
        // round trip while a finished set of work tiles waits for its MMAs
        if (fronts == n + 1 && n + 1 < count) {
          const int s1 = (n + 1) % PL::NST;
          if (__shfl_sync(0xffffffffu, mbar_test_wait(&full[s1], ((n + 1) / PL::NST) & 1), 0)) { front(n + 1); ++fronts; }
        }
        mbar_wait(b_ready, n & 1);
        if (n == 0) { mbar_wait(&rfull[0], 0); mbar_wait(&rfull[1], 0); mbar_wait(&rfull[2], 0); }
        else mbar_wait(&rfull[(n + 2) & 3], ((n + 2) >> 2) & 1);
        TS(1, n, 2);
        tc_fence_after();
        const uint32_t accum0 = n > 0;
        const uint32_t ds = base + PL::DS, dbd = base + PL::DBD, kk_ = st + PL::K, rr = base + PL::RRING;
#pragma unroll
        for (int k = 0; k < 12; ++k)     // dQr += dBD0 . Rwin, window chunk k/4 lives in ring slot (n + k/4) mod 4
          umma_bf16_warp(tmem_base + TS_ACC1, umma_smem_desc(dbd + (k >> 2) * 16384 + (k & 3) * 32, 16, 1024),
                         umma_smem_desc(rr + ((n + (k >> 2)) & 3) * SZ_KV + (k & 3) * 2048, 8192, 1024), id_kn, accum0 | (k > 0));
        umma_commit_warp(dbd_done);      // the window tile may be rewritten while the dQw MMAs and the tile store still run
#pragma unroll
        for (int k = 0; k < 4; ++k)      // dQw += dS . K
          umma_bf16_warp(tmem_base + TS_ACC0, umma_smem_desc(ds + k * 32, 16, 1024), umma_smem_desc(kk_ + k * 2048, 8192, 1024), id_kn, accum0 | (k > 0));
        umma_commit_warp(b_done);
        umma_commit_warp(&empty[s]);
        umma_commit_warp(&rfree[n & 3]);          // chunk n was the first chunk of this window: no later tile reads it
        TS(1, n, 3);
        if (fronts == n + 1 && n + 1 < count) { front(n + 1); ++fronts; }
      }
      umma_commit_warp(acc_full);
    }
  } else if (warp == W_ST) {
    if (lane == 0) {     // bf16 P and dS of every band tile -> global (consumed by the lite dK/dV and dR kernels)
      for (int n = 0; n < count; ++n) {
        mbar_wait(b_ready, n & 1);
        if (a.store_p) tma_store_tile(&M.pst, sm + PL::P, (tile0 + n) * BQ);
        tma_store_tile(&M.dst, sm + PL::DS, (tile0 + n) * BQ);
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        mbar_arrive(b_done);
      }
      asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    }
  } else {
    const int r = 32 * (warp & 3) + lane, qd = warp >> 2;
    const uint32_t lane_base = (uint32_t)(32 * (warp & 3)) << 16;
    const int c0 = 127 - r + KPT * qd;
    uint32_t tile_off[KPT / 8], pair_off[KPT / 2], single_off[2];
    {
      auto dbd_off = [&](int c) -> uint32_t { return (uint32_t)((c >> 6) * 16384 + r * 128 + ((((c & 63) >> 3) ^ (r & 7)) << 4) + (c & 7) * 2); };
#pragma unroll
      for (int c = 0; c < KPT / 8; ++c) tile_off[c] = (uint32_t)(r * 128 + ((((KPT / 8) * qd + c) ^ (r & 7)) << 4));
      const int ce = c0 + (c0 & 1);
#pragma unroll
      for (int k = 0; k < KPT / 2; ++k) pair_off[k] = dbd_off(ce + 2 * k);
      single_off[0] = dbd_off(c0); single_off[1] = dbd_off(c0 + KPT - 1);
    }
    const int i = I * BQ + r;
    float lse2 = 0.f, dlt = 0.f;
    if (i < g.T) {
      const int64_t o = ((int64_t)b * a.H + h) * g.T + i;
      lse2 = a.lse[o] * 1.4426950408889634f; dlt = a.delta[o];
    }
    const float* mt = a.m_tiles + (int64_t)tile0 * BQ + r;
    float m_next = mt[0];
    for (int n = 0; n < count; ++n) {
      const int s = n % PL::NST, tb = n & 1; const uint32_t sph = (n / PL::NST) & 1, tph = (n >> 1) & 1;
      const float f = exp2f(m_next - lse2);
      if (n + 1 < count) m_next = mt[(n + 1) * BQ];
      if (tid == 0) TS(0, n, 0);
      mbar_wait(&full[s], sph);
      if (tid == 0) TS(0, n, 1);
      const uint8_t* pt = sm + PL::RING + s * PL::STAGE + PL::PT;
      uint32_t pw[KPT / 2];
#pragma unroll
      for (int c = 0; c < KPT / 8; ++c) {
        const uint4 u = *reinterpret_cast<const uint4*>(pt + tile_off[c]);
        pw[4 * c] = u.x; pw[4 * c + 1] = u.y; pw[4 * c + 2] = u.z; pw[4 * c + 3] = u.w;
      }
      mbar_wait(&f_full[tb], tph);
      if (tid == 0) TS(0, n, 2);
      tc_fence_after();
      float p[KPT], ds[KPT];
      tmem_ld_32x16(tmem_base + lane_base + TS_DP + 64 * tb + KPT * qd, ds);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&t_free[tb]);
      if (tid == 0) TS(0, n, 3);
#pragma unroll
      for (int k = 0; k < KPT / 2; ++k) {
        p[2 * k] = __uint_as_float(pw[k] << 16) * f;
        p[2 * k + 1] = __uint_as_float(pw[k] & 0xFFFF0000u) * f;
      }
      const float fs = a.scale;
#pragma unroll
      for (int jj = 0; jj < KPT; ++jj) ds[jj] = p[jj] * (ds[jj] - dlt) * fs;
      if (tid == 0) TS(0, n, 4);
      if (n > 0) mbar_wait(dbd_done, (n - 1) & 1);    // the dQr MMAs of tile n-1 have read the window tile
      if (tid == 0) TS(0, n, 5);
      {
        uint8_t* dbd = sm + PL::DBD;
        if ((c0 & 1) == 0) {
#pragma unroll
          for (int k = 0; k < KPT / 2; ++k) *reinterpret_cast<uint32_t*>(dbd + pair_off[k]) = pack2(ds[2 * k], ds[2 * k + 1]);
        } else {
          *reinterpret_cast<bf16*>(dbd + single_off[0]) = __float2bfloat16_rn(ds[0]);
#pragma unroll
          for (int k = 0; k < KPT / 2 - 1; ++k) *reinterpret_cast<uint32_t*>(dbd + pair_off[k]) = pack2(ds[2 * k + 1], ds[2 * k + 2]);
          *reinterpret_cast<bf16*>(dbd + single_off[1]) = __float2bfloat16_rn(ds[KPT - 1]);
        }
      }
      if (n > 0) mbar_wait(b_done, (n - 1) & 1);      // ... and its dQw MMAs + tile store have read P / dS
#pragma unroll
      for (int c = 0; c < KPT / 8; ++c) {
        uint4 o; o.x = pack2(p[c * 8], p[c * 8 + 1]); o.y = pack2(p[c * 8 + 2], p[c * 8 + 3]); o.z = pack2(p[c * 8 + 4], p[c * 8 + 5]); o.w = pack2(p[c * 8 + 6], p[c * 8 + 7]);
        uint4 d; d.x = pack2(ds[c * 8], ds[c * 8 + 1]); d.y = pack2(ds[c * 8 + 2], ds[c * 8 + 3]); d.z = pack2(ds[c * 8 + 4], ds[c * 8 + 5]); d.w = pack2(ds[c * 8 + 6], ds[c * 8 + 7]);
        if (a.store_p) *reinterpret_cast<uint4*>(sm + PL::P + tile_off[c]) = o;
        *reinterpret_cast<uint4*>(sm + PL::DS + tile_off[c]) = d;
      }
      fence_proxy_async_smem();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(b_ready);
      if (tid == 0) TS(0, n, 6);
    }
    mbar_wait(acc_full, 0);
    tc_fence_after();
    float v0[KPT], v1[KPT];
    tmem_ld_32x16(tmem_base + lane_base + TS_ACC0 + KPT * qd, v0);
    tmem_ld_32x16(tmem_base + lane_base + TS_ACC1 + KPT * qd, v1);
    tmem_ld_wait();
    if (i < g.T) {
      uint4* dst = reinterpret_cast<uint4*>(a.dq + ((int64_t)b * g.T + i) * a.ldq + h * DH + KPT * qd);
#pragma unroll
      for (int c = 0; c < KPT / 8; ++c) {
        uint4 o;
        o.x = pack2(v0[c * 8] + v1[c * 8], v0[c * 8 + 1] + v1[c * 8 + 1]); o.y = pack2(v0[c * 8 + 2] + v1[c * 8 + 2], v0[c * 8 + 3] + v1[c * 8 + 3]);
        o.z = pack2(v0[c * 8 + 4] + v1[c * 8 + 4], v0[c * 8 + 5] + v1[c * 8 + 5]); o.w = pack2(v0[c * 8 + 6] + v1[c * 8 + 6], v0[c * 8 + 7] + v1[c * 8 + 7]);
        dst[c] = o;
      }
    } else {
#pragma unroll
      for (int c = 0; c < KPT; ++c) { v0[c] = 0.f; v1[c] = 0.f; }
    }
#pragma unroll
    for (int c = 0; c < KPT; ++c) {
      const float s0 = warp_sum(v0[c]), s1 = warp_sum(v1[c]);
      if (lane == c) { atomicAdd(&a.drwb[h * DH + KPT * qd + c], s0); atomicAdd(&a.drrb[h * DH + KPT * qd + c], s1); }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == W_MMA) tmem_dealloc<TS_COLS>(tmem_base);
}

// qw = q + r_w_bias, qr = q + r_r_bias (bf16, [B*T, HD]); delta[b,h,i] = sum_c dO.O;
// dos (optional) = dO * exp2(m - lse) per (row, head), m = the forward's per-row soft-max reference: P^T dO = P~^T dos, so the dK/dV pass
// multiplies the forward's P~ tiles as they are.
// One thread per 8 adjacent columns (16-byte loads / stores), the 8 threads of a (row, head) meet in three shuffles.
__global__ void __launch_bounds__(256) relattn_bwd_prep_kernel(const bf16* __restrict__ q, int64_t ldq, const float* __restrict__ rwb, const float* __restrict__ rrb,
                                        const bf16* __restrict__ out, const bf16* __restrict__ dout, bf16* __restrict__ qw, bf16* __restrict__ qr,
                                        float* __restrict__ delta, int B, int T, int H, bf16* __restrict__ dos, const float* __restrict__ lse,
                                        const float* __restrict__ m_tiles, int nt_max) {
  const int HD = H * DH, tpr = HD / 8;                         // threads per row
  const int64_t total = (int64_t)B * T * tpr;
  for (int64_t g = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; g < total; g += (int64_t)gridDim.x * blockDim.x) {   // total is a multiple of 8: groups stay whole
    const int64_t n = g / tpr; const int c = (int)(g % tpr) * 8;
    const uint4 qv = *reinterpret_cast<const uint4*>(q + n * ldq + c);
    const uint4 ov = *reinterpret_cast<const uint4*>(out + n * HD + c);
    const uint4 dv = *reinterpret_cast<const uint4*>(dout + n * HD + c);
    const float4 w0 = *reinterpret_cast<const float4*>(rwb + c), w1 = *reinterpret_cast<const float4*>(rwb + c + 4);
    const float4 r0 = *reinterpret_cast<const float4*>(rrb + c), r1 = *reinterpret_cast<const float4*>(rrb + c + 4);
    const float wb[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w}, rb[8] = {r0.x, r0.y, r0.z, r0.w, r1.x, r1.y, r1.z, r1.w};
    const uint32_t qs[4] = {qv.x, qv.y, qv.z, qv.w}, os[4] = {ov.x, ov.y, ov.z, ov.w}, ds[4] = {dv.x, dv.y, dv.z, dv.w};
    uint32_t ow[4], orr[4];
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float q0 = __uint_as_float(qs[k] << 16), q1 = __uint_as_float(qs[k] & 0xFFFF0000u);
      ow[k] = pack2(q0 + wb[2 * k], q1 + wb[2 * k + 1]);
      orr[k] = pack2(q0 + rb[2 * k], q1 + rb[2 * k + 1]);
      s += __uint_as_float(os[k] << 16) * __uint_as_float(ds[k] << 16) + __uint_as_float(os[k] & 0xFFFF0000u) * __uint_as_float(ds[k] & 0xFFFF0000u);
    }
    *reinterpret_cast<uint4*>(qw + n * HD + c) = make_uint4(ow[0], ow[1], ow[2], ow[3]);
    *reinterpret_cast<uint4*>(qr + n * HD + c) = make_uint4(orr[0], orr[1], orr[2], orr[3]);
    s += __shfl_xor_sync(0xffffffffu, s, 1); s += __shfl_xor_sync(0xffffffffu, s, 2); s += __shfl_xor_sync(0xffffffffu, s, 4);   // DH = 64 = 8 threads
    const int64_t b = n / T, i = n % T; const int hh = c / DH;
    if ((c & (DH - 1)) == 0) delta[(b * H + hh) * T + i] = s;
    if (dos) {
      const int nI = (T + BQ - 1) / BQ;
      const float m = m_tiles[(((b * H + hh) * nI + i / BQ) * (int64_t)nt_max) * BQ + (i % BQ)];
      const float f = exp2f(m - lse[(b * H + hh) * T + i] * 1.4426950408889634f);
      uint32_t od[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) od[k] = pack2(__uint_as_float(ds[k] << 16) * f, __uint_as_float(ds[k] & 0xFFFF0000u) * f);
      *reinterpret_cast<uint4*>(dos + n * HD + c) = make_uint4(od[0], od[1], od[2], od[3]);
    }
  }
}

template <int MODE>
int launch_mode(const Maps& M, const BwdArgs& a, dim3 grid, cudaStream_t st) {
  static bool attr_set = false;
  if (!attr_set) {
    TXL_CUDA(cudaFuncSetAttribute(relattn_bwd_tc_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes<MODE>()));
    attr_set = true;
  }
  relattn_bwd_tc_kernel<MODE><<<grid, NTHREADS, smem_bytes<MODE>(), st>>>(M, a);
  TXL_LAUNCH_CHECK();
  return TXL_OK;
}
}  // namespace

int bwd_nt_max_impl(const TxlBand& band);
static int bwd_nt_max(const TxlBand& band) { return bwd_nt_max_impl(band); }
int bwd_nt_max_impl(const TxlBand& band) {
  const BandGeom g = make_band(band);
  const int nI = (band.T + BQ - 1) / BQ;
  int mx = 0;
  for (int I = 0; I < nI; ++I) {
    int j0 = band_lo(g, I * BQ) / BKV;
    int ilast = I * BQ + BQ - 1 < band.T - 1 ? I * BQ + BQ - 1 : band.T - 1;
    int hi = band_hi(g, ilast) < g.klen - 1 ? band_hi(g, ilast) : g.klen - 1;
    int cnt = hi / BKV - j0 + 1;
    if (cnt > mx) mx = cnt;
  }
  return mx;
}
static int64_t bwd_tile_rows(const TxlAttnDims* D) {
  const int nI = (D->band.T + BQ - 1) / BQ;
  return (int64_t)D->B * D->H * nI * bwd_nt_max(D->band) * BQ;
}
int64_t txl_relattn_tile_rows(const TxlAttnDims* D) { return bwd_tile_rows(D); }
int txl_relattn_nt_max(const TxlBand* band) { return bwd_nt_max(*band); }
static bool tc_disabled(const char* name) {
  const char* e = getenv(name); const char* e2 = getenv("TXL_DISABLE_TC");
  return (e && e[0] == '1') || (e2 && e2[0] == '1');
}
// bytes of forward state (bf16 P~ tiles + fp32 row maxima) the tensor-core backward can reuse; 0 when that path does not apply
int64_t txl_relattn_saved_bytes_tc(const TxlAttnDims* D) {
  static int off = -1;
  if (off < 0) { const char* e = getenv("TXL_ATTN_BWD_RECOMPUTE"); off = (tc_disabled("TXL_DISABLE_TC_ATTN") || tc_disabled("TXL_DISABLE_TC_ATTN_BWD") || (e && e[0] == '1')) ? 1 : 0; }
  if (off || D->dtype != TXL_BF16) return 0;
  const int T = D->band.T, mlen = D->band.mlen;
  if (D->dh != DH || (T % BKV) || (mlen % BKV) || T + mlen < WIN || mlen <= 0) return 0;
  if (!(D->band.same_length && mlen == D->band.mem_len)) return 0;
  if ((D->ldq % 8) || (D->ldkv_cur % 8) || (D->ldkv_mem % 8)) return 0;
  const int64_t trows = bwd_tile_rows(D);
  if (trows >= (1ll << 31)) return 0;
  return trows * BKV * 2 + trows * 4;
}
int64_t txl_relattn_bwd_tc_workspace(const TxlAttnDims* D) {
  const int64_t n = (int64_t)D->B * D->band.T * D->H * D->dh;
  const int64_t base = 2 * n * 2 + (int64_t)D->B * D->H * D->band.T * 4 + 1024;
  return base + 2 * bwd_tile_rows(D) * BKV * 2 + 2048;     // + bf16 P and dS tile stores
}

int txl_relattn_frozen_ref() {
  static int on = -1;
  if (on < 0) { const char* e = getenv("TXL_ATTN_FROZEN_REF"); on = (e && e[0] == '0') ? 0 : 1; }
  return on;
}

// Probe switches of the backward (profiling only; never set on the product path): read from TXL_DBG / TXL_ABL ONCE at load time, changed
// afterwards only through txl_relattn_bwd_probe — no getenv on the per-call path.
static int env_int_once(const char* name) { const char* e = getenv(name); return e ? atoi(e) : 0; }
static int g_bwd_probe_dbg = env_int_once("TXL_DBG");
static int g_bwd_probe_abl = env_int_once("TXL_ABL");
extern "C" int txl_relattn_bwd_probe(int dbg, int abl) {
  const int old = g_bwd_probe_dbg;
  g_bwd_probe_dbg = dbg;
  g_bwd_probe_abl = abl;
  return old;
}

int txl_relattn_bwd_tc(const void* q, const void* k_mem, const void* v_mem, const void* k_cur, const void* v_cur, const void* r, const float* rwb,
                       const float* rrb, const void* out, const float* lse, const void* dout, void* dq, void* dk_mem, void* dv_mem, void* dk_cur,
                       void* dv_cur, float* dr, float* drwb, float* drrb, void* ws, const void* saved, const TxlAttnDims* D, void* stream, int* handled) {
  *handled = 0;
  static int disabled = -1;
  if (disabled < 0) { const char* e = getenv("TXL_DISABLE_TC_ATTN_BWD"); const char* e2 = getenv("TXL_DISABLE_TC"); disabled = ((e && e[0] == '1') || (e2 && e2[0] == '1')) ? 1 : 0; }
  if (disabled) return TXL_OK;
  const int T = D->band.T, mlen = D->band.mlen, klen = T + mlen, HD = D->H * D->dh;
  auto al16 = [](const void* p) { return ((uintptr_t)p & 15) == 0; };
  if (D->dh != DH || (T % BKV) || (mlen % BKV) || klen < WIN || mlen <= 0) return TXL_OK;
  if (!(D->band.same_length && mlen == D->band.mem_len)) return TXL_OK;       // dense distance-space band only
  if ((D->ldq % 8) || (D->ldkv_cur % 8) || (D->ldkv_mem % 8)) return TXL_OK;
  if (!al16(q) || !al16(k_cur) || !al16(v_cur) || !al16(k_mem) || !al16(v_mem) || !al16(r) || !al16(out) || !al16(dout) || !al16(dq) || !al16(dk_cur) ||
      !al16(dv_cur) || !al16(ws) || (dk_mem && (!al16(dk_mem) || !al16(dv_mem))))
    return TXL_OK;
  cudaStream_t st = (cudaStream_t)stream;
  const int dbg = g_bwd_probe_dbg;   // timing probes only (txl_relattn_bwd_probe): 16 = run the dQ pass alone, 32 = skip the prep kernel
  const int64_t n = (int64_t)D->B * T * HD;
  bf16* qw = (bf16*)ws; bf16* qr = qw + n; float* delta = (float*)(qr + n);
  // the forward's saved tiles carry ONE soft-max reference per row (frozen_ref): the dK/dV pass reads them directly against dO scaled per row,
  // and the dQ pass writes no normalised P tiles.  The scaled dO lives where those P tiles would have gone.
  const int64_t trows_ = bwd_tile_rows(D);
  const bool saved_ok = saved && txl_relattn_saved_bytes_tc(D) > 0 && al16(saved) && trows_ < (1ll << 31);
  const bool frozen = saved_ok && txl_relattn_frozen_ref();
  bf16* dos = nullptr;
  const float* m_saved = saved_ok ? reinterpret_cast<const float*>(reinterpret_cast<const bf16*>(saved) + trows_ * BKV) : nullptr;
  if (frozen) dos = (bf16*)((((uintptr_t)(delta + (int64_t)D->B * D->H * T)) + 1023) & ~(uintptr_t)1023);
  if (!(dbg & 32)) {
    const int64_t items = (int64_t)D->B * T * (HD / 8);
    relattn_bwd_prep_kernel<<<(unsigned)imin64(cdiv64(items, 256), (int64_t)txl_num_sms() * 16), 256, 0, st>>>((const bf16*)q, D->ldq, rwb, rrb, (const bf16*)out, (const bf16*)dout, qw, qr, delta, D->B, T, D->H,
                                                                                                                     dos, lse, m_saved, bwd_nt_max(D->band));
    TXL_LAUNCH_CHECK();
  }
  Maps M;
  int rc;
  if ((rc = txl_make_tmap_2d(&M.qw, qw, (uint64_t)D->B * T, (uint64_t)HD, (uint64_t)HD, BQ, DH))) return rc;
  if ((rc = txl_make_tmap_2d(&M.qr, qr, (uint64_t)D->B * T, (uint64_t)HD, (uint64_t)HD, BQ, DH))) return rc;
  if ((rc = txl_make_tmap_2d(&M.dO, dout, (uint64_t)D->B * T, (uint64_t)HD, (uint64_t)HD, BQ, DH))) return rc;
  if ((rc = txl_make_tmap_2d(&M.kc, k_cur, (uint64_t)D->B * T, (uint64_t)HD, (uint64_t)D->ldkv_cur, BKV, DH))) return rc;
  if ((rc = txl_make_tmap_2d(&M.vc, v_cur, (uint64_t)D->B * T, (uint64_t)HD, (uint64_t)D->ldkv_cur, BKV, DH))) return rc;
  if ((rc = txl_make_tmap_2d(&M.km, k_mem, (uint64_t)D->B * mlen, (uint64_t)HD, (uint64_t)D->ldkv_mem, BKV, DH))) return rc;
  if ((rc = txl_make_tmap_2d(&M.vm, v_mem, (uint64_t)D->B * mlen, (uint64_t)HD, (uint64_t)D->ldkv_mem, BKV, DH))) return rc;
  if ((rc = txl_make_tmap_2d(&M.r, r, (uint64_t)klen, (uint64_t)HD, (uint64_t)HD, WIN, DH))) return rc;

  BwdArgs a;
  a.lse = lse; a.delta = delta; a.dq = (bf16*)dq; a.dk_mem = (bf16*)dk_mem; a.dv_mem = (bf16*)dv_mem; a.dk_cur = (bf16*)dk_cur; a.dv_cur = (bf16*)dv_cur;
  a.dr = dr; a.drwb = drwb; a.drrb = drrb; a.B = D->B; a.H = D->H; a.band = D->band; a.ldq = D->ldq; a.ldkv_mem = D->ldkv_mem; a.ldkv_cur = D->ldkv_cur;
  a.scale = 1.f / sqrtf((float)DH); a.scale_log2 = a.scale * 1.4426950408889634f;
  // diagonals J - 2I touched by the band
  const BandGeom g = make_band(D->band);
  const int nI = (T + BQ - 1) / BQ;
  int dmin = 1 << 30, dmax = -(1 << 30);
  for (int I = 0; I < nI; ++I) {
    int j0 = band_lo(g, I * BQ) / BKV;
    int ilast = I * BQ + BQ - 1 < T - 1 ? I * BQ + BQ - 1 : T - 1;
    int hi = band_hi(g, ilast) < klen - 1 ? band_hi(g, ilast) : klen - 1;
    int j1 = hi / BKV;
    if (j0 - 2 * I < dmin) dmin = j0 - 2 * I;
    if (j1 - 2 * I > dmax) dmax = j1 - 2 * I;
  }
  a.delta_min = dmin; a.n_delta = dmax - dmin + 1;
  static int recompute = -1;     // TXL_ATTN_BWD_RECOMPUTE=1: three recompute passes instead of one recompute pass + two lite passes
  if (recompute < 0) { const char* e = getenv("TXL_ATTN_BWD_RECOMPUTE"); recompute = (e && e[0] == '1') ? 1 : 0; }
  const int64_t trows = bwd_tile_rows(D);
  a.nt_max = bwd_nt_max(D->band);
  a.store_tiles = (!recompute && trows * 1 < (1ll << 31)) ? 1 : 0;
  if (a.store_tiles) {
    uintptr_t pws = ((uintptr_t)(delta + (int64_t)D->B * D->H * T) + 1023) & ~(uintptr_t)1023;
    bf16* pstore = (bf16*)pws; bf16* dstore = pstore + trows * BKV;
    if ((rc = txl_make_tmap_2d(&M.pst, pstore, (uint64_t)trows, (uint64_t)BKV, (uint64_t)BKV, BQ, BKV))) return rc;
    if ((rc = txl_make_tmap_2d(&M.dst, dstore, (uint64_t)trows, (uint64_t)BKV, (uint64_t)BKV, BQ, BKV))) return rc;
  } else { M.pst = M.qw; M.dst = M.qw; }
  a.m_tiles = nullptr; M.psv = M.qw; M.r64 = M.r; M.pst2 = M.qw; M.dst2 = M.qw;
  M.pkv = M.pst; M.pkv2 = M.qw; M.dOkv = M.dO; a.store_p = 1;
  a.abl = g_bwd_probe_abl;
  if (saved && a.store_tiles && txl_relattn_saved_bytes_tc(D) > 0 && al16(saved)) {
    a.m_tiles = reinterpret_cast<const float*>(reinterpret_cast<const bf16*>(saved) + trows * BKV);
    if ((rc = txl_make_tmap_2d(&M.psv, saved, (uint64_t)trows, (uint64_t)BKV, (uint64_t)BKV, BQ, BKV))) return rc;
    if ((rc = txl_make_tmap_2d(&M.r64, r, (uint64_t)klen, (uint64_t)HD, (uint64_t)HD, BKV, DH))) return rc;
    if (frozen) {
      a.store_p = 0;
      M.pkv = M.psv;
      if ((rc = txl_make_tmap_2d(&M.dOkv, dos, (uint64_t)D->B * T, (uint64_t)HD, (uint64_t)HD, BQ, DH))) return rc;
    }
    static bool attr_s = false;
    if (!attr_s) { TXL_CUDA(cudaFuncSetAttribute(relattn_bwd_dq_saved_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, PlanS::SMEM)); attr_s = true; }
    relattn_bwd_dq_saved_kernel<<<dim3(nI, D->H, D->B), NTHREADS, PlanS::SMEM, st>>>(M, a);
    TXL_LAUNCH_CHECK();
    if (a.abl & 256) {
      static int printed = 0;
      if (!printed++) {
        cudaDeviceSynchronize();
        static long long ts[3][32][8];
        cudaMemcpyFromSymbol(ts, g_ts, sizeof(ts));
        const long long t0 = ts[2][0][0];
        for (int n = 0; n < 19; ++n) {
          fprintf(stderr, "tile %2d  prod: empty %6lld | mma: full %6lld tfree %6lld bready %6lld issued %6lld | smx: top %6lld full %6lld ffull %6lld ld %6lld math %6lld bdone %6lld arrive %6lld\n", n,
                  ts[2][n][0] - t0, ts[1][n][0] - t0, ts[1][n][1] - t0, ts[1][n][2] - t0, ts[1][n][3] - t0, ts[0][n][0] - t0, ts[0][n][1] - t0, ts[0][n][2] - t0,
                  ts[0][n][3] - t0, ts[0][n][4] - t0, ts[0][n][5] - t0, ts[0][n][6] - t0);
        }
      }
    }
  } else {
    if ((rc = launch_mode<MODE_DQ>(M, a, dim3(nI, D->H, D->B), st))) return rc;
  }
  if (dbg & 16) { *handled = 1; return TXL_OK; }
  if (a.store_tiles) {
    static bool attr_set = false;
    if (!attr_set) {
      TXL_CUDA(cudaFuncSetAttribute(relattn_bwd_dkv_lite_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, LITE_SMEM));
      TXL_CUDA(cudaFuncSetAttribute(relattn_bwd_dr_lite_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, DRL_SMEM));
      attr_set = true;
    }
    // two key tiles per CTA when every query tile's band starts on an even key tile and spans an even number of them
    bool pairs = (klen % (2 * BKV)) == 0;
    for (int I = 0; I < nI && pairs; ++I) {
      const int j0 = band_lo(g, I * BQ) / BKV;
      const int ilast = I * BQ + BQ - 1 < T - 1 ? I * BQ + BQ - 1 : T - 1;
      const int hi = band_hi(g, ilast) < klen - 1 ? band_hi(g, ilast) : klen - 1;
      if ((j0 & 1) || ((hi / BKV - j0 + 1) & 1)) pairs = false;
    }
    if (pairs) {
      uintptr_t pws2 = ((uintptr_t)(delta + (int64_t)D->B * D->H * T) + 1023) & ~(uintptr_t)1023;
      bf16* pstore2 = (bf16*)pws2; bf16* dstore2 = pstore2 + trows * BKV;
      if ((rc = txl_make_tmap_2d(&M.pst2, pstore2, (uint64_t)trows, (uint64_t)BKV, (uint64_t)BKV, 2 * BQ, BKV))) return rc;
      if ((rc = txl_make_tmap_2d(&M.dst2, dstore2, (uint64_t)trows, (uint64_t)BKV, (uint64_t)BKV, 2 * BQ, BKV))) return rc;
      M.pkv2 = M.pst2;
      if (!a.store_p) { if ((rc = txl_make_tmap_2d(&M.pkv2, saved, (uint64_t)trows, (uint64_t)BKV, (uint64_t)BKV, 2 * BQ, BKV))) return rc; }
      static bool attr_p = false;
      if (!attr_p) { TXL_CUDA(cudaFuncSetAttribute(relattn_bwd_dkv_pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, PAIR_SMEM)); attr_p = true; }
      relattn_bwd_dkv_pair_kernel<<<dim3(klen / (2 * BKV), D->H, D->B), 192, PAIR_SMEM, st>>>(M, a);
    } else {
      relattn_bwd_dkv_lite_kernel<<<dim3(klen / BKV, D->H, D->B), 192, LITE_SMEM, st>>>(M, a);
    }
    TXL_LAUNCH_CHECK();
    if (pairs && (a.n_delta % 2) == 0 && (((uintptr_t)dr) & 15) == 0) {     // diagonals pair up exactly like the key tiles (delta_min is even then)
      static bool attr_d = false;
      if (!attr_d) { TXL_CUDA(cudaFuncSetAttribute(relattn_bwd_dr_pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, DRP_SMEM)); attr_d = true; }
      const int npairs = a.n_delta / 2, heads = D->H;
      int zsplit = txl_num_sms() / (npairs * heads);           // split the (I, b) loop until every SM has a CTA
      if (zsplit < 1) zsplit = 1;
      if (zsplit > D->B) zsplit = D->B;
      relattn_bwd_dr_pair_kernel<<<dim3(npairs, heads, zsplit), DRL_THREADS, DRP_SMEM, st>>>(M, a);
    } else {
      relattn_bwd_dr_lite_kernel<<<dim3(a.n_delta, D->H, 1), DRL_THREADS, DRL_SMEM, st>>>(M, a);
    }
    TXL_LAUNCH_CHECK();
  } else {
    if ((rc = launch_mode<MODE_DKV>(M, a, dim3(klen / BKV, D->H, D->B), st))) return rc;
    if ((rc = launch_mode<MODE_DR>(M, a, dim3(a.n_delta, D->H, 1), st))) return rc;
  }
  *handled = 1;
  return TXL_OK;
}
