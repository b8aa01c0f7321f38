// decode_cluster.cu — the bf16 decode step, fourth generation: a thread-block CLUSTER of 8 CTAs (one per attention head) carries up to 8
// sequences through ALL layers and the LM head on its own.  No grid-wide synchronisation, no kernel boundary between stages: what crosses
// CTAs crosses distributed shared memory behind hardware cluster barriers (8 per layer).
//
// Why (profiles/r02_decode_ab.txt): a decode step is 12 x ~7 dependent all-to-all stages of ~1 us of work each.  As separate kernels a stage
// costs ~3.5 us of launch latency (second generation, 86 launches); as one persistent kernel with grid barriers ~4 us (barrier ~2 us + the L2
// round trips of handing activations from 148 CTAs to 148 CTAs; csrc/decode_persist.cu).  Inside a cluster the hand-over is an SM-to-SM read
// of a few KB and the barrier is a hardware one, and clusters never wait for each other.  The price: every cluster streams the whole weight
// set (from L2: 83 MB of weights + 12 MB of r tables stay resident there), which is what bounds small batches.
//
// Sequences are the N = 8 columns of mma.sync tiles whose M = 16 rows are weight rows (no padded rows); the cache is HF's own hidden-state
// `mems` used as a ring with the key / value projections absorbed into the query / output side (see decode_persist.cu: qt_h = W_k,h^T (q_h +
// r_w_bias_h), out_h = W_v,h sum_s p hid_s; BD from a [seq, head, distance] table).  Per layer, in CTA r (= head r):      [A.3-A.6 at T=1]
//   1  q_r = x W_q,r^T                        -> qa = q_r + r_w_bias_r, qb = q_r + r_r_bias_r                      (local)
//   2  qt_r = qa W_k,r  (rows = d columns)    -> local, pulled by the attention CTAs;  bd_r = qb r[:, r]^T -> global table;  barrier 1
//   3  attention item (sequence, key part): TMA-swizzled ring stages, transposed mma.sync tiles (as decode_persist.cu), partial context for
//      all heads -> local;  barrier 2
//   4  ctx_r = merge of the parts (pulled);  v_r = ctx_r W_v,r^T;  partial_r = v_r W_o[:, r]^T (all d columns) -> local;  barrier 3
//   5  CTA j sums the 8 partials of its 64 columns (pulled) + residual, row statistics -> local;  barrier 4;  LayerNorm of its columns with
//      the combined statistics -> its slice of y1;  barrier 5;  all-gather of y1 (pulled)
//   6  h1_r = relu(y1 W_1[r-slice]^T + b_1)   (d_inner / 8 features, local)
//   7  partial_r = h1_r W_2[:, r-slice]^T -> local;  barrier 6;  8 = 5 with y1 as residual: barriers 7, 8 -> x; ring slot of the next layer
// Weights stream through warp-private cp.async rings (16 rows x 128 columns per stage, 3 stages): no block-level synchronisation inside a GEMM.
#include "tc_common.cuh"
#include <cooperative_groups.h>
#include <string.h>
#include <stdlib.h>
namespace cg = cooperative_groups;

namespace {
constexpr int DC_PARTS = 16;                        // key parts per sequence in the partial-context scratch (>= the largest cluster)
constexpr int DC_THREADS = 256, DC_WARPS = 8;
constexpr int DC_D = 512, DC_DH = 64, DC_H = 8;
constexpr int DC_NS = 8;                            // sequences per cluster = N of the weight MMAs
constexpr int DC_KS = 64, DC_NST = 2;               // attention: keys per stage, stages
constexpr int DC_SPITCH = 72;
constexpr int DC_XP = DC_D * 2 + 64;                // pitch of a [sequence][d] activation tile (64 mod 128: conflict-free 16-byte fragment loads)
constexpr int DC_P64 = DC_DH * 2 + 64;
constexpr int DC_QP = DC_D * 2 + 16;                // pitch of the staged qt rows (attention B operand)
constexpr int DC_WSTG = 5120;                       // a warp's weight stage: 16 rows x (128 columns + pad)
constexpr int DC_STAGE = (DC_D / 64) * 8192;        // an attention stage: 64 keys x d

struct DcLayer {
  const bf16 *wq, *wkT, *wv, *wo, *w1, *w2, *r;
  const float *b1, *b2, *rwb, *rrb, *ln1w, *ln1b, *ln2w, *ln2b;
  bf16* ring;
};
struct DcArgs {
  const DcLayer* layers;
  const CUtensorMap* tmaps;
  const bf16* E;
  const float* out_bias;
  bf16* x;                    // [B, d]: embedding rows in, final hidden rows out
  float* bd;                  // [B, H, MLP]
  bf16* pctx;                 // [B, 8 key parts, H, d]  normalised partial contexts (L2: too many to keep in shared memory)
  float* pml;                 // [B, 8, H, 2]            (max, sum) of each part
  float* logits;
  const int32_t* pos;
  unsigned long long* tstamp;
  int64_t ldl;
  int B, ML, MLP, L, Vx, di, NS;
  int hints;                  // bit 0: weights evict_last, bit 1: ring rows evict_first (L2 cache-policy operands)
  float eps, scale_log2;
};

// shared-memory map (bytes from the 1024-aligned base)
struct DcSmem {
  unsigned char *xs, *ys, *ctxA, *qa, *qb, *vs, *h1s;      // [8 seqs][pitch] bf16 operand tiles
  unsigned char *qtl;                                       // local qt_r [8 seqs][d] bf16 (pitch d*2), pulled by the attention CTAs
  unsigned char *qsm;                                       // attention: staged qt of its sequence [8 heads][DC_QP]
  float *partl;                                             // [8 seqs][d] fp32 partial sums of this CTA's head / inner slice
  float *statl;                                             // [8 seqs][2] (mean, M2) of this CTA's 64 columns
  float *ssm;                                               // attention score tiles
  unsigned char *r1;                                        // attention stages | weight rings
};

__device__ __forceinline__ void mma16816(float* c, uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t wsel(const uint4& v, int i) { return i == 0 ? v.x : i == 1 ? v.y : i == 2 ? v.z : v.w; }
__device__ __forceinline__ void cpa16(void* dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst)), "l"(src) : "memory");
}
// L2 policy: the weights (83 MB + 12 MB of r tables) are re-read by every cluster each step and must stay in the 126 MB L2; the ring rows
// (0.8 GB per step at 64 sequences) are read once and must not push them out.
__device__ __forceinline__ uint64_t l2_policy_evict_last() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ uint64_t l2_policy_evict_first() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ void cpa16_hint(void* dst, const void* src, uint64_t pol) {
  asm volatile("cp.async.cg.shared.global.L2::cache_hint [%0], [%1], 16, %2;" ::"r"(smem_u32(dst)), "l"(src), "l"(pol) : "memory");
}
__device__ __forceinline__ void tma_load_2d_hint(void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, uint64_t pol) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%3, %4}], [%2], %5;" ::"r"(smem_u32(dst)),
               "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "l"(pol)
               : "memory");
}
__device__ __forceinline__ void cpa_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cpa_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void fence_proxy_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }
__device__ __forceinline__ uint32_t pack2(float a, float b) {
  __nv_bfloat162 p = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&p);
}

// One warp: C[n, s] = sum_k W[n, k] X[s, k] for its 16-row tiles tile0, tile0 + tstride, ... of W [n_rows, K] (row pitch ldw, global), X [8
// sequences][K] in shared memory (pitch pitchX).  The weight rows stream through the warp's private 3-stage cp.async ring in units of 16 rows
// x min(K, 128) columns; a 16-byte piece per row is two k16 steps' worth of fragment registers (the k index inside an MMA is only a label:
// the same relabelling on both operands).  epi(tile, c, b0, b1): c[0], c[1] = row 16 tile + g, sequences 2t, 2t+1; c[2], c[3] = row + 8;
// b0 / b1 = bias[row] / bias[row + 8] (0 without a bias), loaded when the tile starts so that the epilogue never waits for L2.
// The ring is private to the warp, so the FIRST TWO UNITS OF THE NEXT GEMM are requested (wg_prefetch) as soon as this one's last unit has
// been consumed - across __syncthreads and cluster barriers: a GEMM then starts with its pipeline full instead of two cold round trips.
struct WgSpec {
  const bf16* W;
  int64_t ldw;
  int K, n_rows, tile0, tstride;
};
// a unit = 16 rows x KC columns; lane -> (row r0 + i * rpp, 16-byte piece v) for i < npass: every per-unit quantity is a shift or an add
// (the copy loop used to spend ~400 instructions per 5 KB unit on divisions - it, not L2, bounded the stream)
struct WgPlan {
  int KC, lg_nkc, pitchS, nunits, rpp, npass, r0, v;
};
__device__ __forceinline__ WgPlan wg_plan(const WgSpec& w, int lane) {
  WgPlan p;
  p.KC = w.K < 128 ? w.K : 128;
  const int nkc = w.K / p.KC;                         // 1, 2 or 4
  p.lg_nkc = nkc == 4 ? 2 : nkc == 2 ? 1 : 0;
  p.pitchS = p.KC * 2 + 64;
  const int ntiles = (w.n_rows + 15) >> 4;
  const int mine = w.tile0 < ntiles ? (ntiles - w.tile0 + w.tstride - 1) / w.tstride : 0;
  p.nunits = mine << p.lg_nkc;
  const int lg_vpr = p.KC == 128 ? 4 : p.KC == 64 ? 3 : 2;      // 16-byte pieces per row: KC in {128, 64, 32}
  p.rpp = 32 >> lg_vpr;                               // rows per pass of the warp
  p.npass = 16 / p.rpp;
  p.r0 = lane >> lg_vpr;
  p.v = lane & ((1 << lg_vpr) - 1);
  return p;
}
__device__ __forceinline__ void wg_issue(const WgSpec& w, const WgPlan& p, int u, unsigned char* ring) {
  if (u < p.nunits) {
    const int tile = w.tile0 + (u >> p.lg_nkc) * w.tstride, kc = u & ((1 << p.lg_nkc) - 1);
    unsigned char* dst = ring + (u % 3) * DC_WSTG + p.r0 * p.pitchS + p.v * 16;
    const bf16* src = w.W + kc * p.KC + p.v * 8;
    int row = tile * 16 + p.r0;
#pragma unroll 4
    for (int i = 0; i < p.npass; ++i) {
      cpa16(dst, src + (int64_t)min(row, w.n_rows - 1) * w.ldw);
      dst += p.rpp * p.pitchS;
      row += p.rpp;
    }
  }
  cpa_commit();                                      // (possibly empty) one group per unit keeps wait_group counting uniform
}
__device__ __forceinline__ void wg_prefetch(const WgSpec& w, unsigned char* ring, int lane) {
  const WgPlan p = wg_plan(w, lane);
  wg_issue(w, p, 0, ring);
  wg_issue(w, p, 1, ring);
}
template <typename Epi>
__device__ __forceinline__ void warp_gemm(const unsigned char* X, int pitchX, const WgSpec& w, const float* bias, unsigned char* ring, int lane, bool prefetched,
                                          Epi&& epi) {
  const int g = lane >> 2, t = lane & 3;
  const WgPlan p = wg_plan(w, lane);
  const int KC = p.KC, nkc = 1 << p.lg_nkc, pitchS = p.pitchS, nunits = p.nunits;
  if (!prefetched) { wg_issue(w, p, 0, ring); wg_issue(w, p, 1, ring); }
  float acc[2][4];
  float bv0 = 0.f, bv1 = 0.f;
  for (int u = 0; u < nunits; ++u) {
    wg_issue(w, p, u + 2, ring);                     // into the stage unit u - 1 was read from (all lanes passed the __syncwarp below)
    const int kc = u & (nkc - 1), tile = w.tile0 + (u >> p.lg_nkc) * w.tstride;
    if (kc == 0) {
#pragma unroll
      for (int q = 0; q < 4; ++q) { acc[0][q] = 0.f; acc[1][q] = 0.f; }
      if (bias) { bv0 = bias[min(tile * 16 + g, w.n_rows - 1)]; bv1 = bias[min(tile * 16 + g + 8, w.n_rows - 1)]; }
    }
    cpa_wait<2>();
    __syncwarp();
    const unsigned char* st = ring + (u % 3) * DC_WSTG;
    const unsigned char* xr = X + g * pitchX + kc * KC * 2 + t * 16;
    for (int kb = 0; kb < (KC >> 5); ++kb) {
      const uint4 w0 = *reinterpret_cast<const uint4*>(st + g * pitchS + kb * 64 + t * 16);
      const uint4 w1 = *reinterpret_cast<const uint4*>(st + (g + 8) * pitchS + kb * 64 + t * 16);
      const uint4 xv = *reinterpret_cast<const uint4*>(xr + kb * 64);
#pragma unroll
      for (int s = 0; s < 2; ++s)
        mma16816(acc[s], wsel(w0, 2 * s), wsel(w1, 2 * s), wsel(w0, 2 * s + 1), wsel(w1, 2 * s + 1), wsel(xv, 2 * s), wsel(xv, 2 * s + 1));
    }
    if (kc == nkc - 1) {
      float c[4] = {acc[0][0] + acc[1][0], acc[0][1] + acc[1][1], acc[0][2] + acc[1][2], acc[0][3] + acc[1][3]};
      epi(tile, c, bv0, bv1);
    }
    __syncwarp();
  }
  cpa_wait<0>();
}

// ------------------------------------------------------------------------------------------------------------------------------ attention items
// The stage loop of decode_persist.cu's att_item (TMA boxes of 64 keys x 128 bytes with SWIZZLE_128B; phase A S^T[key, head] on mma.sync with
// keys as M, phase B ctx^T[col, head]; two stages), run over ALL items of this CTA as ONE pipeline: item k = keys [sb, se) of sequence
// seq_first + k; the ring stages of item k + 1 are requested while item k is still being consumed, and its qt rows (pulled from the eight
// head owners' shared memory) land in the other half of sm.qsm meanwhile.  Per item: normalised partial context of all heads ->
// pctx[seq, part] [8][d] bf16, (max, sum) -> pml[seq, part] [8][2] (global: read by the head owners after the cluster barrier).
template <int CS>
__device__ void dc_attention(cg::cluster_group& cluster, const DcArgs& a, const CUtensorMap* tm, const DcSmem& sm, uint64_t* full, uint32_t& uses,
                             int n_items, int seq_first, int part, int sb, int se, int cur) {
  constexpr int D = DC_D, NB = D / 64, NKS = D / 16, KH = NKS / 2, CW = D / DC_WARPS, NMT = CW / 16;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, g = lane >> 2, t = lane & 3;
  const int ML = a.ML, H = DC_H;
  const int nst = (se - sb + DC_KS - 1) / DC_KS, ntot = n_items * nst;
  constexpr int RPH = CS / 8, VPS = 64 / RPH;      // CTAs per head; 16-byte pieces of a qt row held by each of them
  const int seq_loc0 = seq_first - (blockIdx.x / CS) * a.NS;          // row of the first item's sequence in the owners' qtl
  unsigned char* stage0 = sm.r1;
  const uint64_t pol = l2_policy_evict_first();
  auto issue = [&](int j) {
    if (lane == 0 && j < ntot) {
      const uint32_t u = uses + j, st = u % DC_NST;
      const int row = (seq_first + j / nst) * ML + sb + (j % nst) * DC_KS;
      if (warp == 0) mbar_expect_tx(&full[st], DC_STAGE);
      for (int cb = warp; cb < NB; cb += DC_WARPS) tma_load_2d_hint(stage0 + (size_t)st * DC_STAGE + cb * 8192, tm, &full[st], cb * 64, row, pol);
    }
  };
  // qt[sequence of the item][head] from the eight head owners -> sm.qsm half (item & 1): two 16-byte pieces per thread, loaded into registers
  // when the previous item starts and stored just before the barrier of its last stage (a remote read takes ~1 us: never waited for)
  static_assert(8 * (D / 8) == 2 * DC_THREADS, "two qt pieces per thread");
  uint4 qpre[2];
  auto qt_load = [&](int item) {
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      const int e = tid + q * DC_THREADS, hh = e / (D / 8), v = e % (D / 8);
      const unsigned char* rp = cluster.map_shared_rank(sm.qtl, hh * RPH + v / VPS);
      qpre[q] = *reinterpret_cast<const uint4*>(rp + ((size_t)(seq_loc0 + item) * VPS + (v % VPS)) * 16);
    }
  };
  auto qt_store = [&](int item) {
    unsigned char* dst = sm.qsm + (size_t)(item & 1) * 8 * DC_QP;
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      const int e = tid + q * DC_THREADS, hh = e / (D / 8), v = e % (D / 8);
      *reinterpret_cast<uint4*>(dst + (size_t)hh * DC_QP + v * 16) = qpre[q];
    }
  };
#pragma unroll
  for (int j = 0; j < DC_NST; ++j) issue(j);
  qt_load(0);
  qt_store(0);
  __syncthreads();
  const int mt = warp & 3, kh = warp >> 2;
  float acc[NMT][4];
  float m_run = -INFINITY, l_run = 0.f;
  const unsigned char* qrow = nullptr;
  const float *bd0 = nullptr, *bd1 = nullptr;
  for (int j = 0; j < ntot; ++j) {
    const int item = j / nst, i = j - item * nst;
    if (i == 0) {
#pragma unroll
      for (int jj = 0; jj < NMT; ++jj)
#pragma unroll
        for (int q = 0; q < 4; ++q) acc[jj][q] = 0.f;
      m_run = -INFINITY; l_run = 0.f;
      qrow = sm.qsm + (size_t)(item & 1) * 8 * DC_QP + (size_t)g * DC_QP + t * 4;
      bd0 = a.bd + ((int64_t)(seq_first + item) * H + 2 * t) * a.MLP;
      bd1 = bd0 + a.MLP;
      if (item + 1 < n_items) qt_load(item + 1);
    }
    const uint32_t u = uses + j, st = u % DC_NST;
    const int s0 = sb + i * DC_KS;
    float* S = sm.ssm + (size_t)(j & 1) * 2 * 8 * DC_SPITCH;
    float bdv[4] = {0.f, 0.f, 0.f, 0.f};
    if (kh == 0) {
#pragma unroll
      for (int hf = 0; hf < 2; ++hf) {
        const int sk = s0 + mt * 16 + g + hf * 8;
        if (sk < se) {
          const int x = sk <= cur ? ML - cur + sk : sk - cur;
          bdv[hf * 2] = __ldcg(bd0 + x);
          bdv[hf * 2 + 1] = __ldcg(bd1 + x);
        }
      }
    }
    mbar_wait(&full[st], (u / DC_NST) & 1);
    const uint32_t sbase = smem_u32(stage0 + (size_t)st * DC_STAGE);
    {
      float c4[2][4];
#pragma unroll
      for (int q = 0; q < 2; ++q) { c4[q][0] = 0.f; c4[q][1] = 0.f; c4[q][2] = 0.f; c4[q][3] = 0.f; }
      const int row = mt * 16 + (lane & 7) + ((lane >> 3) & 1) * 8;
      const uint32_t rbase = sbase + row * 128, rsw = row & 7, chi = lane >> 4;
#pragma unroll
      for (int k2 = 0; k2 < KH; k2 += 2) {
#pragma unroll
        for (int q = 0; q < 2; ++q) {
          const int ks = kh * KH + k2 + q;
          uint32_t a0, a1, a2, a3;
          asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
                       : "=r"(a0), "=r"(a1), "=r"(a2), "=r"(a3)
                       : "r"(rbase + (ks >> 2) * 8192 + (((((ks & 3) << 1) + chi) ^ rsw) << 4)));
          const uint32_t b0 = *reinterpret_cast<const uint32_t*>(qrow + ks * 32), b1 = *reinterpret_cast<const uint32_t*>(qrow + ks * 32 + 16);
          mma16816(c4[q], a0, a1, a2, a3, b0, b1);
        }
      }
      float* Sh = S + (size_t)kh * 8 * DC_SPITCH + mt * 16 + g;
      Sh[(2 * t) * DC_SPITCH] = c4[0][0] + c4[1][0] + bdv[0];
      Sh[(2 * t + 1) * DC_SPITCH] = c4[0][1] + c4[1][1] + bdv[1];
      Sh[(2 * t) * DC_SPITCH + 8] = c4[0][2] + c4[1][2] + bdv[2];
      Sh[(2 * t + 1) * DC_SPITCH + 8] = c4[0][3] + c4[1][3] + bdv[3];
    }
    if (i == nst - 1 && item + 1 < n_items) qt_store(item + 1);      // (that half of qsm was last read two items ago)
    __syncthreads();
    if (j >= 1) issue(j - 1 + DC_NST);                // every warp is done with the previous stage's buffer
    {
      float sc[4][4];
      float mx = -INFINITY;
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) {
        const float* p0 = S + g * DC_SPITCH + ks * 16 + 2 * t;
        const float2 lo0 = *reinterpret_cast<const float2*>(p0), hi0 = *reinterpret_cast<const float2*>(p0 + 8);
        const float2 lo1 = *reinterpret_cast<const float2*>(p0 + 8 * DC_SPITCH), hi1 = *reinterpret_cast<const float2*>(p0 + 8 * DC_SPITCH + 8);
        const int kb = s0 + ks * 16 + 2 * t;
        sc[ks][0] = kb < se ? (lo0.x + lo1.x) * a.scale_log2 : -INFINITY;
        sc[ks][1] = kb + 1 < se ? (lo0.y + lo1.y) * a.scale_log2 : -INFINITY;
        sc[ks][2] = kb + 8 < se ? (hi0.x + hi1.x) * a.scale_log2 : -INFINITY;
        sc[ks][3] = kb + 9 < se ? (hi0.y + hi1.y) * a.scale_log2 : -INFINITY;
        mx = fmaxf(fmaxf(mx, fmaxf(sc[ks][0], sc[ks][1])), fmaxf(sc[ks][2], sc[ks][3]));
      }
      mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 1));
      mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 2));
      const float m_new = fmaxf(m_run, mx);
      const float m_safe = (m_new == -INFINITY) ? 0.f : m_new;
      const float corr = exp2f(m_run - m_safe);
      uint32_t pb0[4], pb1[4];
      float ls = 0.f;
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) {
        const float p0 = exp2f(sc[ks][0] - m_safe), p1 = exp2f(sc[ks][1] - m_safe), p2 = exp2f(sc[ks][2] - m_safe), p3 = exp2f(sc[ks][3] - m_safe);
        ls += (p0 + p1) + (p2 + p3);
        pb0[ks] = pack2(p0, p1); pb1[ks] = pack2(p2, p3);
      }
      ls += __shfl_xor_sync(0xffffffffu, ls, 1);
      ls += __shfl_xor_sync(0xffffffffu, ls, 2);
      l_run = l_run * corr + ls;
      m_run = m_new;
      const float ce = __shfl_sync(0xffffffffu, corr, 8 * t), co = __shfl_sync(0xffffffffu, corr, 8 * t + 4);
#pragma unroll
      for (int jj = 0; jj < NMT; ++jj) { acc[jj][0] *= ce; acc[jj][1] *= co; acc[jj][2] *= ce; acc[jj][3] *= co; }
      const int krow = (lane & 7) + (lane >> 4) * 8;
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) {
        const int row = ks * 16 + krow;
        const uint32_t roff = row * 128, rsw = row & 7;
#pragma unroll
        for (int jj = 0; jj < NMT; ++jj) {
          const int col0 = warp * CW + jj * 16;
          const uint32_t ch = ((col0 & 63) >> 3) + ((lane >> 3) & 1);
          uint32_t a0, a1, a2, a3;
          asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
                       : "=r"(a0), "=r"(a1), "=r"(a2), "=r"(a3)
                       : "r"(sbase + (col0 >> 6) * 8192 + roff + ((ch ^ rsw) << 4)));
          mma16816(acc[jj], a0, a1, a2, a3, pb0[ks], pb1[ks]);
        }
      }
    }
    if (i == nst - 1) {
      const float inv = l_run > 0.f ? 1.f / l_run : 0.f;
      const float ie = __shfl_sync(0xffffffffu, inv, 8 * t), io = __shfl_sync(0xffffffffu, inv, 8 * t + 4);
      bf16* de = a.pctx + (((int64_t)(seq_first + item) * DC_PARTS + part) * H + 2 * t) * D;
      bf16* dodd = de + D;
#pragma unroll
      for (int jj = 0; jj < NMT; ++jj) {
        const int c = warp * CW + jj * 16 + g;
        de[c] = __float2bfloat16_rn(acc[jj][0] * ie); de[c + 8] = __float2bfloat16_rn(acc[jj][2] * ie);
        dodd[c] = __float2bfloat16_rn(acc[jj][1] * io); dodd[c + 8] = __float2bfloat16_rn(acc[jj][3] * io);
      }
      if (warp == 0 && t == 0) {
        float* ml = a.pml + (((int64_t)(seq_first + item) * DC_PARTS + part) * H + g) * 2;
        ml[0] = m_run; ml[1] = l_run;
      }
    }
  }
  uses += ntot;
}

// ------------------------------------------------------------------------------------------------------------------------------ residual + LayerNorm
// CTA j owns COLS = d / CS columns.  z = residual + sum over the CS CTAs' partials (pulled) (+ bias); (mean, M2) of its columns per row ->
// statl; barrier; combined statistics (Chan) -> normalise -> its slice of `dst` (local); barrier; all-gather of the other slices (pulled).
template <int CS>
__device__ void dc_residual_ln(cg::cluster_group& cluster, const DcSmem& sm, unsigned char* res, unsigned char* dst, const float* bias, const float* gamma,
                               const float* beta, float eps, int rank, unsigned long long* ts = nullptr) {
  constexpr int COLS = DC_D / CS, CPL = COLS / 32;
  auto tick = [&](int k) { if (ts && blockIdx.x == 0 && threadIdx.x == 0) { unsigned long long tt; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(tt)); ts[k] = tt; } };
  tick(0);                 // columns per CTA, per lane (2 or 1)
  const int tid = threadIdx.x, lane = tid & 31, s = tid >> 5;      // one warp per sequence
  const int c = rank * COLS + lane * CPL;
  float gm[2] = {0.f, 0.f}, bt[2] = {0.f, 0.f}, bs[2] = {0.f, 0.f}, z[2] = {0.f, 0.f};
#pragma unroll
  for (int k = 0; k < CPL; ++k) { gm[k] = gamma[c + k]; bt[k] = beta[c + k]; if (bias) bs[k] = bias[c + k]; }
  float pv[CS][2];
#pragma unroll
  for (int r = 0; r < CS; ++r) {                                   // every remote read is issued before the first add
    const float* rp = cluster.map_shared_rank(sm.partl, r) + s * DC_D + c;
    if (CPL == 2) { const float2 v = *reinterpret_cast<const float2*>(rp); pv[r][0] = v.x; pv[r][1] = v.y; }
    else { pv[r][0] = rp[0]; pv[r][1] = 0.f; }
  }
#pragma unroll
  for (int k = 0; k < CPL; ++k) z[k] = __bfloat162float(reinterpret_cast<const bf16*>(res + (size_t)s * DC_XP)[c + k]) + bs[k];
#pragma unroll
  for (int r = 0; r < CS; ++r)
#pragma unroll
    for (int k = 0; k < CPL; ++k) z[k] += pv[r][k];
  const float mj = warp_sum(CPL == 2 ? z[0] + z[1] : z[0]) * (1.f / COLS);
  const float m2j = warp_sum(CPL == 2 ? (z[0] - mj) * (z[0] - mj) + (z[1] - mj) * (z[1] - mj) : (z[0] - mj) * (z[0] - mj));
  if (lane == 0) { sm.statl[s * 2] = mj; sm.statl[s * 2 + 1] = m2j; }
  tick(1);
  cluster.sync();
  tick(2);
  float mean = 0.f, m2 = 0.f;
  {
    float mr[CS], qr[CS];
#pragma unroll
    for (int r = 0; r < CS; ++r) {
      const float* rp = cluster.map_shared_rank(sm.statl, r);
      mr[r] = rp[s * 2]; qr[r] = rp[s * 2 + 1];
    }
#pragma unroll
    for (int r = 0; r < CS; ++r) mean += mr[r];
    mean *= (1.f / CS);
#pragma unroll
    for (int r = 0; r < CS; ++r) m2 += qr[r] + (float)COLS * (mr[r] - mean) * (mr[r] - mean);
  }
  const float rs = rsqrtf(m2 * (1.f / DC_D) + eps);
  bf16* drow = reinterpret_cast<bf16*>(dst + (size_t)s * DC_XP);
#pragma unroll
  for (int k = 0; k < CPL; ++k) drow[c + k] = __float2bfloat16_rn((z[k] - mean) * rs * gm[k] + bt[k]);
  tick(3);
  cluster.sync();
  tick(4);
  // all-gather: the other CTAs' column slices, 16 bytes per thread and step (8 sequences x COLS / 8 vectors per slice)
  constexpr int VPR = COLS / 8;
  for (int e = tid; e < (CS - 1) * 8 * VPR; e += DC_THREADS) {
    int r = e / (8 * VPR);
    const int ss = (e / VPR) & 7, v = e % VPR;
    r += (r >= rank);
    const unsigned char* rp = cluster.map_shared_rank(dst, r);
    const size_t off = (size_t)ss * DC_XP + r * COLS * 2 + v * 16;
    *reinterpret_cast<uint4*>(dst + off) = *reinterpret_cast<const uint4*>(rp + off);
  }
  __syncthreads();
  tick(5);
}

// ------------------------------------------------------------------------------------------------------------------------------ the step
// CS = CTAs per cluster: 8 (CTA = head) or 16 (two CTAs per head: both compute the head's q, each takes half of its qt columns / bd rows /
// v features, and a 1/16 slice of everything else - half the weight bytes per SM).  Launched with the cluster dimension as an attribute.
template <int CS>
__global__ void __launch_bounds__(DC_THREADS, 1) decode_cluster_kernel(const DcArgs a) {
  constexpr int RPH = CS / 8;                         // CTAs per head
  constexpr int DC_CS = CS;
  extern __shared__ __align__(128) unsigned char dc_smem[];
  __shared__ uint64_t full[DC_NST];
  cg::cluster_group cluster = cg::this_cluster();
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, g = lane >> 2, t = lane & 3;
  const int rank = (int)cluster.block_rank();
  const int head = rank / RPH, sub = rank % RPH;
  const int cl = blockIdx.x / DC_CS;
  const int seq0 = cl * a.NS, nsc = min(a.NS, a.B - seq0);         // this cluster's sequences
  const int ML = a.ML, DS = a.di / DC_CS;                           // DS = inner features of this CTA
  DcSmem sm;
  {
    unsigned char* p = dc_smem + ((1024u - (smem_u32(dc_smem) & 1023u)) & 1023u);
    sm.r1 = p; p += DC_NST * DC_STAGE;                              // (>= 8 warps x 3 x DC_WSTG)
    sm.xs = p; p += DC_NS * DC_XP;
    sm.ys = p; p += DC_NS * DC_XP;
    sm.ctxA = p; p += DC_NS * DC_XP;
    sm.qa = p; p += DC_NS * DC_P64;
    sm.qb = p; p += DC_NS * DC_P64;
    sm.vs = p; p += DC_NS * DC_P64;
    sm.h1s = p; p += DC_NS * (DS * 2 + 64);
    sm.qtl = p; p += DC_NS * DC_D * 2;
    sm.qsm = p; p += 2 * 8 * DC_QP;
    sm.partl = reinterpret_cast<float*>(p); p += DC_NS * DC_D * 4;
    sm.ssm = reinterpret_cast<float*>(p); p += 2 * 2 * 8 * DC_SPITCH * 4;
    sm.statl = reinterpret_cast<float*>(p); p += DC_NS * 2 * 4;
  }
  unsigned char* wring = sm.r1 + (size_t)warp * 3 * DC_WSTG;
  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < DC_NST; ++s) mbar_init(&full[s], 1);
    fence_barrier_init();
  }
  // operand tiles start as zeros (rows of absent sequences stay zero: they only meet discarded MMA columns)
  for (int e = tid; e < (int)((sm.qtl - sm.xs) / 16); e += DC_THREADS) reinterpret_cast<uint4*>(sm.xs)[e] = make_uint4(0, 0, 0, 0);
  __syncthreads();
  uint32_t uses = 0;
  const int p = *a.pos;
  const int cur = p % ML;
  // attention items of this cluster: (sequence, key part).  1, 2, 4 or 8 sequences: 8 / nsc parts per sequence, one item per CTA; otherwise
  // 8 parts per sequence and CTA r takes part r of every sequence - either way the eight CTAs stream the same number of keys
  const bool one_item = nsc > 0 && (DC_CS % nsc) == 0;
  const int KPC = one_item ? DC_CS / nsc : DC_CS;
  int per = (ML + KPC - 1) / KPC;
  per = (per + DC_KS - 1) / DC_KS * DC_KS;
  // x rows of the cluster's sequences (every CTA holds all of them)
  for (int e = tid; e < nsc * (DC_D / 8); e += DC_THREADS) {
    const int s = e / (DC_D / 8), v = e % (DC_D / 8);
    *reinterpret_cast<uint4*>(sm.xs + (size_t)s * DC_XP + v * 16) = __ldcg(reinterpret_cast<const uint4*>(a.x + (int64_t)(seq0 + s) * DC_D + v * 8));
  }
  __syncthreads();
  int nstamp = 0;
  auto stamp = [&]() {
    if (a.tstamp && blockIdx.x == 0 && tid == 0) { unsigned long long tt; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(tt)); a.tstamp[nstamp++] = tt; }
  };
  stamp();
  // the GEMMs of a layer (this CTA's slices) as streaming specs; each one's first two units are requested while the previous one drains
  const int XR = (((a.ML + 1 + RPH - 1) / RPH) + 15) / 16 * 16;         // bd rows (distances) per CTA of a head
  auto spec_q = [&](const DcLayer& ly) { return WgSpec{ly.wq + (int64_t)head * DC_DH * DC_D, DC_D, DC_D, DC_DH, warp, DC_WARPS}; };
  auto spec_qt = [&](const DcLayer& ly) { return WgSpec{ly.wkT + ((int64_t)head * DC_D + sub * (DC_D / RPH)) * DC_DH, DC_DH, DC_DH, DC_D / RPH, warp, DC_WARPS}; };
  auto spec_bd = [&](const DcLayer& ly) { return WgSpec{ly.r + (int64_t)sub * XR * DC_D + head * DC_DH, DC_D, DC_DH, min(XR, ML + 1 - sub * XR), warp, DC_WARPS}; };
  auto spec_v = [&](const DcLayer& ly) { return WgSpec{ly.wv + ((int64_t)head * DC_DH + sub * (DC_DH / RPH)) * DC_D, DC_D, DC_D, DC_DH / RPH, warp, DC_WARPS}; };
  auto spec_o = [&](const DcLayer& ly) { return WgSpec{ly.wo + head * DC_DH + sub * (DC_DH / RPH), DC_D, DC_DH / RPH, DC_D, warp, DC_WARPS}; };
  auto spec_f1 = [&](const DcLayer& ly) { return WgSpec{ly.w1 + (int64_t)rank * DS * DC_D, DC_D, DC_D, DS, warp, DC_WARPS}; };
  auto spec_f2 = [&](const DcLayer& ly) { return WgSpec{ly.w2 + rank * DS, a.di, DS, DC_D, warp, DC_WARPS}; };
  const WgSpec spec_lm = WgSpec{a.E, DC_D, DC_D, a.Vx, rank * DC_WARPS + warp, DC_CS * DC_WARPS};
  bool pre = false;                                   // the upcoming GEMM's first units are already in flight
  for (int l = 0; l < a.L; ++l) {
    const DcLayer ly = a.layers[l];
    // ring slot of this step: CTA r stores the row of sequence r (TMA reads it after barrier 1)
    if (rank < nsc && tid < DC_D / 8) {
      *reinterpret_cast<uint4*>(ly.ring + ((int64_t)(seq0 + rank) * ML + cur) * DC_D + tid * 8) = *reinterpret_cast<const uint4*>(sm.xs + (size_t)rank * DC_XP + tid * 16);
      fence_proxy_async_all();
    }
    // ---- 1: q_r -> qa, qb   (bias pairs of this lane's rows: r_w_bias through the hoisted slot, r_r_bias loaded next to it)
    {
      const float* rrb = ly.rrb + head * DC_DH;
      warp_gemm(sm.xs, DC_XP, spec_q(ly), ly.rwb + head * DC_DH, wring, lane, pre, [&](int tile, const float* c, float bw0, float bw1) {
        const int n = tile * 16 + g;
        const float br0 = rrb[n], br1 = rrb[n + 8];
        bf16* qa0 = reinterpret_cast<bf16*>(sm.qa + (size_t)(2 * t) * DC_P64); bf16* qa1 = reinterpret_cast<bf16*>(sm.qa + (size_t)(2 * t + 1) * DC_P64);
        bf16* qb0 = reinterpret_cast<bf16*>(sm.qb + (size_t)(2 * t) * DC_P64); bf16* qb1 = reinterpret_cast<bf16*>(sm.qb + (size_t)(2 * t + 1) * DC_P64);
        qa0[n] = __float2bfloat16_rn(c[0] + bw0); qa1[n] = __float2bfloat16_rn(c[1] + bw0); qa0[n + 8] = __float2bfloat16_rn(c[2] + bw1); qa1[n + 8] = __float2bfloat16_rn(c[3] + bw1);
        qb0[n] = __float2bfloat16_rn(c[0] + br0); qb1[n] = __float2bfloat16_rn(c[1] + br0); qb0[n + 8] = __float2bfloat16_rn(c[2] + br1); qb1[n + 8] = __float2bfloat16_rn(c[3] + br1);
      });
    }
    wg_prefetch(spec_qt(ly), wring, lane);
    __syncthreads();
    // ---- 2: qt_r (local) and bd_r (global table)
    warp_gemm(sm.qa, DC_P64, spec_qt(ly), nullptr, wring, lane, true, [&](int tile, const float* c, float, float) {
      const int n = tile * 16 + g;
      bf16* q0 = reinterpret_cast<bf16*>(sm.qtl) + (size_t)(2 * t) * (DC_D / RPH); bf16* q1 = q0 + DC_D / RPH;       // this CTA's columns of qt_head
      q0[n] = __float2bfloat16_rn(c[0]); q1[n] = __float2bfloat16_rn(c[1]); q0[n + 8] = __float2bfloat16_rn(c[2]); q1[n + 8] = __float2bfloat16_rn(c[3]);
    });
    wg_prefetch(spec_bd(ly), wring, lane);
    warp_gemm(sm.qb, DC_P64, spec_bd(ly), nullptr, wring, lane, true, [&](int tile, const float* c, float, float) {
#pragma unroll
      for (int hf = 0; hf < 2; ++hf) {
        const int x = sub * XR + tile * 16 + g + hf * 8;
        if (x <= ML && tile * 16 + g + hf * 8 < XR) {
          if (2 * t < nsc) a.bd[((int64_t)(seq0 + 2 * t) * DC_H + head) * a.MLP + x] = c[hf * 2];
          if (2 * t + 1 < nsc) a.bd[((int64_t)(seq0 + 2 * t + 1) * DC_H + head) * a.MLP + x] = c[hf * 2 + 1];
        }
      }
    });
    cluster.sync();                                                                                          // barrier 1
    stamp();
    // ---- 3: attention items of this CTA (one pipeline over all of them)
    {
      const int n_items = one_item ? (rank < nsc * KPC ? 1 : 0) : nsc;
      const int it_s0 = one_item ? rank / KPC : 0, part = one_item ? rank % KPC : rank;
      const int sb = part * per, se = min(sb + per, ML);
      if (n_items > 0) {
        if (sb < se) {
          fence_proxy_async_all();                    // the stage memory held generic-proxy data (weight rings): TMA writes follow
          __syncthreads();
          dc_attention<CS>(cluster, a, &a.tmaps[l], sm, full, uses, n_items, seq0 + it_s0, part, sb, se, cur);
        } else if (tid < 8) {                         // an empty key part (mem_len shorter than the parts): weight 0 (its rows are never read)
          for (int k = 0; k < n_items; ++k) {
            float* ml = a.pml + (((int64_t)(seq0 + it_s0 + k) * DC_PARTS + part) * DC_H + tid) * 2;
            ml[0] = -INFINITY; ml[1] = 0.f;
          }
        }
      }
      __syncthreads();                                // every warp is done with the stage memory: the weight rings take it back
    }
    wg_prefetch(spec_v(ly), wring, lane);
    cluster.sync();                                                                                          // barrier 2
    stamp();
    // ---- 4: merge the parts of this CTA's head for every sequence -> ctxA; v (its features of the head); partial over those features
    for (int e = tid; e < nsc * (DC_D / 8); e += DC_THREADS) {
      const int s = e / (DC_D / 8), v = e % (DC_D / 8);
      const float* mlp = a.pml + ((int64_t)(seq0 + s) * DC_PARTS * DC_H + head) * 2;
      const bf16* cxp = a.pctx + ((int64_t)(seq0 + s) * DC_PARTS * DC_H + head) * DC_D + v * 8;
      float lz[DC_PARTS];
      float M = -INFINITY, Ls = 0.f;
      {
        float2 ml[DC_PARTS];
#pragma unroll
        for (int kp = 0; kp < DC_PARTS; ++kp) ml[kp] = kp < KPC ? __ldcg(reinterpret_cast<const float2*>(mlp + (size_t)kp * DC_H * 2)) : make_float2(-INFINITY, 0.f);
#pragma unroll
        for (int kp = 0; kp < DC_PARTS; ++kp) M = fmaxf(M, ml[kp].x);
#pragma unroll
        for (int kp = 0; kp < DC_PARTS; ++kp) { lz[kp] = (ml[kp].x == -INFINITY) ? 0.f : exp2f(ml[kp].x - M) * ml[kp].y; Ls += lz[kp]; }
      }
      const float inv = Ls > 0.f ? 1.f / Ls : 0.f;
      float acc8[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int k0 = 0; k0 < DC_PARTS; k0 += 8) {        // eight loads in flight at a time
        if (k0 < KPC) {
          uint4 cu[8];
#pragma unroll
          for (int q = 0; q < 8; ++q)
            cu[q] = (k0 + q < KPC && lz[k0 + q] > 0.f) ? __ldcg(reinterpret_cast<const uint4*>(cxp + (size_t)(k0 + q) * DC_H * DC_D)) : make_uint4(0, 0, 0, 0);
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            const bf16* eb = reinterpret_cast<const bf16*>(&cu[q]);
            const float cf = lz[k0 + q] * inv;
#pragma unroll
            for (int k = 0; k < 8; ++k) acc8[k] = fmaf(cf, __bfloat162float(eb[k]), acc8[k]);
          }
        }
      }
      uint4 o;
      o.x = pack2(acc8[0], acc8[1]); o.y = pack2(acc8[2], acc8[3]); o.z = pack2(acc8[4], acc8[5]); o.w = pack2(acc8[6], acc8[7]);
      *reinterpret_cast<uint4*>(sm.ctxA + (size_t)s * DC_XP + v * 16) = o;
    }
    __syncthreads();
    warp_gemm(sm.ctxA, DC_XP, spec_v(ly), nullptr, wring, lane, true, [&](int tile, const float* c, float, float) {
      const int n = tile * 16 + g;
      bf16* v0 = reinterpret_cast<bf16*>(sm.vs + (size_t)(2 * t) * DC_P64); bf16* v1 = reinterpret_cast<bf16*>(sm.vs + (size_t)(2 * t + 1) * DC_P64);
      v0[n] = __float2bfloat16_rn(c[0]); v1[n] = __float2bfloat16_rn(c[1]); v0[n + 8] = __float2bfloat16_rn(c[2]); v1[n + 8] = __float2bfloat16_rn(c[3]);
    });
    wg_prefetch(spec_o(ly), wring, lane);
    __syncthreads();
    warp_gemm(sm.vs, DC_P64, spec_o(ly), nullptr, wring, lane, true, [&](int tile, const float* c, float, float) {
      const int n = tile * 16 + g;
      float* p0 = sm.partl + (size_t)(2 * t) * DC_D; float* p1 = p0 + DC_D;
      p0[n] = c[0]; p1[n] = c[1]; p0[n + 8] = c[2]; p1[n + 8] = c[3];
    });
    wg_prefetch(spec_f1(ly), wring, lane);
    cluster.sync();                                                                                          // barrier 3
    stamp();
    // ---- 5: y1 = LayerNorm(x + attention output)  (barriers 4, 5)
    dc_residual_ln<CS>(cluster, sm, sm.xs, sm.ys, nullptr, ly.ln1w, ly.ln1b, a.eps, rank, (a.tstamp && l == a.L - 1) ? a.tstamp + 130 : nullptr);
    stamp();
    // ---- 6: h1_r
    warp_gemm(sm.ys, DC_XP, spec_f1(ly), ly.b1 + rank * DS, wring, lane, true, [&](int tile, const float* c, float b0, float b1) {
      const int n = tile * 16 + g;
      bf16* h0 = reinterpret_cast<bf16*>(sm.h1s + (size_t)(2 * t) * (DS * 2 + 64)); bf16* h1 = reinterpret_cast<bf16*>(sm.h1s + (size_t)(2 * t + 1) * (DS * 2 + 64));
      if (n < DS) { h0[n] = __float2bfloat16_rn(fmaxf(c[0] + b0, 0.f)); h1[n] = __float2bfloat16_rn(fmaxf(c[1] + b0, 0.f)); }
      if (n + 8 < DS) { h0[n + 8] = __float2bfloat16_rn(fmaxf(c[2] + b1, 0.f)); h1[n + 8] = __float2bfloat16_rn(fmaxf(c[3] + b1, 0.f)); }
    });
    wg_prefetch(spec_f2(ly), wring, lane);
    __syncthreads();
    // ---- 7: partial_r of CoreNet.3
    warp_gemm(sm.h1s, DS * 2 + 64, spec_f2(ly), nullptr, wring, lane, true, [&](int tile, const float* c, float, float) {
      const int n = tile * 16 + g;
      float* p0 = sm.partl + (size_t)(2 * t) * DC_D; float* p1 = p0 + DC_D;
      p0[n] = c[0]; p1[n] = c[1]; p0[n + 8] = c[2]; p1[n + 8] = c[3];
    });
    if (l + 1 < a.L) wg_prefetch(spec_q(a.layers[l + 1]), wring, lane); else wg_prefetch(spec_lm, wring, lane);
    pre = true;
    cluster.sync();                                                                                          // barrier 6
    stamp();
    // ---- 8: x = LayerNorm(y1 + FF output + b2)  (barriers 7, 8)
    dc_residual_ln<CS>(cluster, sm, sm.ys, sm.xs, ly.b2, ly.ln2w, ly.ln2b, a.eps, rank);
    stamp();
  }
  // ---- LM head: the 16-row tiles of [E ; cluster_weight] round-robin over the 64 warps of the cluster; final hidden rows back to global
  warp_gemm(sm.xs, DC_XP, spec_lm, a.out_bias, wring, lane, true, [&](int tile, const float* c, float b0, float b1) {
#pragma unroll
    for (int hf = 0; hf < 2; ++hf) {
      const int n = tile * 16 + g + hf * 8;
      if (n < a.Vx) {
        const float b = hf ? b1 : b0;
        if (2 * t < nsc) a.logits[(int64_t)(seq0 + 2 * t) * a.ldl + n] = c[hf * 2] + b;
        if (2 * t + 1 < nsc) a.logits[(int64_t)(seq0 + 2 * t + 1) * a.ldl + n] = c[hf * 2 + 1] + b;
      }
    }
  });
  if (rank < nsc && tid < DC_D / 8)
    *reinterpret_cast<uint4*>(a.x + (int64_t)(seq0 + rank) * DC_D + tid * 8) = *reinterpret_cast<const uint4*>(sm.xs + (size_t)rank * DC_XP + tid * 16);
  stamp();
  cluster.sync();      // no CTA may exit while a peer can still read its shared memory
}

struct DcHostTable {
  DcLayer layers[64];
  CUtensorMap tm[64];
};

size_t dc_smem_bytes(int di, int CS) {
  const int DS = di / CS;
  size_t n = 1024 + (size_t)DC_NST * DC_STAGE + 3 * DC_NS * DC_XP + 3 * DC_NS * DC_P64 + (size_t)DC_NS * (DS * 2 + 64) + DC_NS * DC_D * 2 + 2 * 8 * DC_QP +
             DC_NS * DC_D * 4 + 2 * 2 * 8 * DC_SPITCH * 4 + DC_NS * 2 * 4;
  return n + 128;
}

int g_dc_max_clusters[2] = {0, 0};                  // co-resident clusters of 8 / 16 CTAs (cudaOccupancyMaxActiveClusters)

template <int CS>
void dc_query_clusters(int di) {
  int& slot = g_dc_max_clusters[CS == 16];
  if (slot != 0) return;
  slot = -1;
  const size_t smem = dc_smem_bytes(di, CS);
  if (cudaFuncSetAttribute(decode_cluster_kernel<CS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) { cudaGetLastError(); return; }
  if (CS > 8 && cudaFuncSetAttribute(decode_cluster_kernel<CS>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) != cudaSuccess) { cudaGetLastError(); return; }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(CS * (144 / CS)); cfg.blockDim = dim3(DC_THREADS); cfg.dynamicSmemBytes = smem;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = CS; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  int n = 0;
  if (cudaOccupancyMaxActiveClusters(&n, decode_cluster_kernel<CS>, &cfg) == cudaSuccess && n > 0) slot = n;
  else cudaGetLastError();
}

bool dc_shape_ok(int di, int CS) {
  if (di % (CS * 32) || di / CS > 512) return false;
  const int ds = di / CS;                            // K of the CoreNet.3 slices: 32, 64 or whole 128-column units
  if (!(ds == 32 || ds == 64 || ds % 128 == 0)) return false;
  return dc_smem_bytes(di, CS) <= 232448;
}

// clusters and sequences per cluster for B sequences with clusters of CS CTAs; 0 if that does not fit one wave of co-resident clusters
int dc_clusters(int B, int di, int CS, int* NS) {
  if (!dc_shape_ok(di, CS)) return 0;
  if (CS == 16) dc_query_clusters<16>(di); else dc_query_clusters<8>(di);
  const int ncl_max = g_dc_max_clusters[CS == 16];
  if (ncl_max < 1) return 0;
  int ncl = B < ncl_max ? B : ncl_max;
  const int ns = (B + ncl - 1) / ncl;
  if (ns > DC_NS) return 0;
  ncl = (B + ns - 1) / ns;                           // with ns sequences per cluster fewer clusters may do
  *NS = ns;
  return ncl;
}

// cluster size for B sequences: TXL_DC_CS forces 8 or 16; otherwise 16 CTAs per cluster (half the weight bytes per SM: the bound of small
// batches) whenever one wave of them holds the batch, else 8
int dc_pick(int B, int di, int* NS, int* ncl) {
  static const int forced = [] { const char* e = getenv("TXL_DC_CS"); return e ? atoi(e) : 0; }();
  for (int CS : {16, 8}) {
    if (forced && forced != CS) continue;
    const int n = dc_clusters(B, di, CS, NS);
    if (n > 0) { *ncl = n; return CS; }
  }
  return 0;
}
}  // namespace

static unsigned long long* g_dc_tstamp = nullptr;
extern "C" int txl_decode_cluster_set_timestamps(unsigned long long* dev_buf) { g_dc_tstamp = dev_buf; return TXL_OK; }
extern "C" int txl_decode_cluster_max_clusters(int di) {
  if (dc_shape_ok(di, 8)) dc_query_clusters<8>(di);
  if (dc_shape_ok(di, 16)) dc_query_clusters<16>(di);
  return (g_dc_max_clusters[0] > 0 ? g_dc_max_clusters[0] : 0) + 100 * (g_dc_max_clusters[1] > 0 ? g_dc_max_clusters[1] : 0);
}

extern "C" int txl_decode_cluster_supported(int B, int H, int dh, int d, int di, int ML, int L, int Vx) {
  if (!(B >= 1 && H == DC_H && dh == DC_DH && d == DC_D && ML >= 1 && L >= 1 && L <= 64 && Vx >= 1)) return 0;
  int ns, ncl;
  return dc_pick(B, di, &ns, &ncl) != 0 ? 1 : 0;
}

extern "C" int64_t txl_decode_cluster_ws_bytes(int B, int H, int dh, int d, int di, int ML, int L, int Vx) {
  if (!txl_decode_cluster_supported(B, H, dh, d, di, ML, L, Vx)) return 0;
  const int MLP = (ML + 1 + 3) / 4 * 4;
  return (int64_t)sizeof(DcHostTable) + 256 + (int64_t)B * H * MLP * 4 + 256 + (int64_t)B * DC_PARTS * H * d * 2 + 256 + (int64_t)B * DC_PARTS * H * 2 * 4 + 256;
}

// One decode step of all layers + the LM-head GEMM by clusters of 8 CTAs.  build_table = 1: upload the per-layer pointer table and the ring
// tensor maps into `ws` (synchronises the stream; no launch) - once per generation before the first step.
extern "C" int txl_decode_cluster_step(const void* const* wqkv, const void* const* wkT, const void* const* wo, const void* const* w1,
                                       const void* const* w2, const void* const* rtab, const float* const* b1, const float* const* b2,
                                       const float* const* rwb, const float* const* rrb, const float* const* ln1w, const float* const* ln1b,
                                       const float* const* ln2w, const float* const* ln2b, void* const* ring, const void* E, const float* out_bias,
                                       void* x, const int32_t* pos, float* logits, int64_t ldl, void* ws, int build_table, int B, int H, int dh, int d,
                                       int di, int ML, int L, int Vx, float eps, void* stream) {
  TXL_CHECK_ARG(txl_decode_cluster_supported(B, H, dh, d, di, ML, L, Vx),
                "decode_cluster: unsupported geometry (needs 8 heads of 64, d_model 512, d_inner a multiple of 256 up to 4096, B <= 8 x co-resident clusters)");
  TXL_CHECK_ARG(ws && ((uintptr_t)ws & 255) == 0, "decode_cluster: workspace must be 256-byte aligned");
  cudaStream_t st = (cudaStream_t)stream;
  const int MLP = (ML + 1 + 3) / 4 * 4;
  unsigned char* w = (unsigned char*)ws;
  DcHostTable* tab = (DcHostTable*)w;
  w += (sizeof(DcHostTable) + 255) / 256 * 256;
  float* bd = (float*)w;
  w += ((int64_t)B * H * MLP * 4 + 255) / 256 * 256;
  bf16* pctx = (bf16*)w;
  w += ((int64_t)B * DC_PARTS * H * d * 2 + 255) / 256 * 256;
  float* pml = (float*)w;
  if (build_table) {
    TXL_CHECK_ARG(wqkv && wkT && wo && w1 && w2 && rtab && b1 && b2 && rwb && rrb && ln1w && ln1b && ln2w && ln2b && ring, "decode_cluster: null table");
    static DcHostTable h;
    memset(&h, 0, sizeof(h));
    for (int l = 0; l < L; ++l) {
      DcLayer& y = h.layers[l];
      y.wq = (const bf16*)wqkv[l]; y.wv = (const bf16*)wqkv[l] + (int64_t)2 * d * d; y.wkT = (const bf16*)wkT[l]; y.wo = (const bf16*)wo[l];
      y.w1 = (const bf16*)w1[l]; y.w2 = (const bf16*)w2[l]; y.r = (const bf16*)rtab[l];
      y.b1 = b1[l]; y.b2 = b2[l]; y.rwb = rwb[l]; y.rrb = rrb[l]; y.ln1w = ln1w[l]; y.ln1b = ln1b[l]; y.ln2w = ln2w[l]; y.ln2b = ln2b[l];
      y.ring = (bf16*)ring[l];
      TXL_CHECK_ARG(y.wq && y.wkT && y.wo && y.w1 && y.w2 && y.r && y.ring && ((uintptr_t)y.ring & 15) == 0 && ((uintptr_t)y.wkT & 15) == 0 &&
                        ((uintptr_t)y.wq & 15) == 0 && ((uintptr_t)y.wo & 15) == 0 && ((uintptr_t)y.w1 & 15) == 0 && ((uintptr_t)y.w2 & 15) == 0 &&
                        ((uintptr_t)y.r & 15) == 0,
                    "decode_cluster: layer %d pointers (16-byte alignment)", l);
      int rc = txl_make_tmap_2d(&h.tm[l], y.ring, (uint64_t)B * ML, (uint64_t)d, (uint64_t)d, 64, 64);
      if (rc) return rc;
    }
    TXL_CUDA(cudaMemcpyAsync(tab, &h, sizeof(h), cudaMemcpyHostToDevice, st));
    TXL_CUDA(cudaStreamSynchronize(st));
    return TXL_OK;
  }
  TXL_CHECK_ARG(E && out_bias && x && pos && logits && ldl >= Vx && ((uintptr_t)x & 15) == 0 && ((uintptr_t)E & 15) == 0, "decode_cluster: bad step args");
  DcArgs a;
  a.layers = tab->layers; a.tmaps = tab->tm; a.E = (const bf16*)E; a.out_bias = out_bias; a.x = (bf16*)x; a.bd = bd; a.pctx = pctx; a.pml = pml; a.logits = logits; a.pos = pos;
  a.tstamp = g_dc_tstamp; a.ldl = ldl; a.B = B; a.ML = ML; a.MLP = MLP; a.L = L; a.Vx = Vx; a.di = di;
  a.eps = eps; a.scale_log2 = 1.4426950408889634f / sqrtf((float)dh);
  a.hints = 2;
  int ncl = 0;
  const int CS = dc_pick(B, di, &a.NS, &ncl);
  TXL_CHECK_ARG(CS != 0, "decode_cluster: no cluster configuration holds %d sequences in one wave", B);
  const size_t smem = dc_smem_bytes(di, CS);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)ncl * CS); cfg.blockDim = dim3(DC_THREADS); cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = CS; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  // the occupancy query set the function attributes for the d_inner it was first asked about: grow the shared-memory size when needed
  static size_t attr[2][64] = {{0}, {0}};
  int dev = 0;
  TXL_CUDA(cudaGetDevice(&dev));
  if (smem > attr[CS == 16][dev & 63]) {
    if (CS == 16) {
      TXL_CUDA(cudaFuncSetAttribute(decode_cluster_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      TXL_CUDA(cudaFuncSetAttribute(decode_cluster_kernel<16>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
    } else {
      TXL_CUDA(cudaFuncSetAttribute(decode_cluster_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    }
    attr[CS == 16][dev & 63] = smem;
  }
  if (CS == 16) { TXL_CUDA(cudaLaunchKernelEx(&cfg, decode_cluster_kernel<16>, a)); }
  else { TXL_CUDA(cudaLaunchKernelEx(&cfg, decode_cluster_kernel<8>, a)); }
  TXL_LAUNCH_CHECK();
  return TXL_OK;
}
