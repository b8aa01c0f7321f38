// decode_persist.cu — the bf16 decode step, third generation: ALL L layers of one T=1 step as ONE persistent cooperative kernel over a ring of
// cached HIDDEN STATES (HF's `mems` themselves), with the per-head key / value projections absorbed into the query / output side.
//
// Why (profiles/r01_decode_ab.txt): the second-generation step is 86 dependent launches (250 us of launch latency at 64 sequences,
// 420 us at 8 where the roofline budget is 30 us) over a ring of PROJECTED k|v rows that holds twice the bytes SURVEY 8d counts.  Here
//   * one CTA per SM stays resident for the whole step; the ~7 stages of a layer are separated by a grid barrier (one L2 atomic + an acquire
//     poll, ~0.5 us) instead of a kernel boundary (~3.5 us inside a CUDA graph);
//   * the cache is HF's own `mems[l]` [B, mem_len, d] used as a ring (slot pos % mem_len is overwritten with the layer input: for T=1 and
//     mlen == mem_len the same_length band IS the ring after that write), read ONCE per layer for all heads:
//        AC[h, s]  = (q_h + r_w_bias_h) . (W_k,h hid_s)  = qt_h . hid_s        qt_h = W_k,h^T (q_h + r_w_bias_h)   (d-vector per head)
//        out_h     = sum_s p[h, s] (W_v,h hid_s)         = W_v,h (sum_s p[h, s] hid_s)
//     so a key costs d*2 bytes of HBM instead of 2*d*2, and the 8 heads are the M rows of mma.sync tiles;
//   * BD[b, h, x] = (q_h + r_r_bias_h) . r[x, h] is one small GEMM per (head, 128 distances) for ALL sequences (the r table is shared), so
//     the attention stage reads 4 bytes per (head, key) instead of a d_head-wide r row per (sequence, key).
// Stages of layer l (each ends in a grid barrier):                                                              [A.3-A.6 at T=1, A.8']
//   QTBD  q_h = x W_q,h^T (recomputed per item: per-head locality instead of a barrier), then  qt[b,h,c-chunk] = (q_h + rwb_h) W_k,h[:, c]
//         or bd[b,h,x-chunk] = (q_h + rrb_h) r[x, h]^T
//   ATT   item (sequence, ring split): bulk-copy pipelined stream of ring rows (3 stages x 64 keys, issued by warp 0), scores on mma.sync (A = qt, rows =
//         heads), + bd, online softmax, P.hid on mma.sync (ldmatrix.trans); normalised partial context + (max, sum) per split
//   VAON  ctx_h = merge of the splits; v_h = ctx_h W_v,h^T; plane[h][b, n-chunk] = v_h W_o[n, h]^T        (again chained per head)
//   LN1   y1 = LayerNorm(x + sum_h plane[h])
//   FF1   h1 = relu(y1 W_1^T + b_1)                     (16 features per item, K split over the warps)
//   FF2   plane[ks] = h1[:, ks] W_2[:, ks]^T            (16 features x 512-wide K slice per item)
//   LN2   x = LayerNorm(y1 + sum_ks plane[ks] + b_2); also written to ring slot `cur` of layer l+1
// then the LM-head GEMM (token + adaptive-softmax cluster logits).  Log-softmax, sampling and the next embedding stay in txl_decode_tail.
// Everything that crosses CTAs is read with L1-bypassing loads (ld.global.cg / cp.async.cg / cp.async.bulk); reductions have a fixed order
// (no atomics on data), so a step is bit-reproducible.  Barrier waits are bounded (trap instead of hang).
#include "tc_common.cuh"
#include <string.h>

namespace {

constexpr int DP_CWARPS = 8;                        // MMA / consumer warps
constexpr int DP_THREADS = DP_CWARPS * 32;          // two warps per SM sub-partition: up to 255 registers per thread (a ninth warp caps them at 168)
constexpr int DP_DH = 64;                           // d_head
constexpr int DP_KS = 64;                           // keys per ring stage
constexpr int DP_NST = 3;                           // ring stages
constexpr int DP_MAXS = 16;                         // ring splits per sequence
constexpr int DP_SPITCH = DP_KS + 8;                // floats per head row of the stage score tile

struct DpLayer {
  const bf16 *wq, *wkT, *wv, *wo, *w1, *w2, *r;     // wq = qkv rows [0,d), wv = qkv rows [2d,3d); wkT [H][d][64]; r [ML+1, d]
  const float *b1, *b2, *rwb, *rrb, *ln1w, *ln1b, *ln2w, *ln2b;
  bf16* ring;                                       // [B, ML, d]
};

struct DpArgs {
  const DpLayer* layers;
  const CUtensorMap* tmaps;   // per layer: the ring as a [B*ML, d] bf16 tensor, box 64 rows x 64 columns, SWIZZLE_128B
  const bf16* E;              // [Vx, d]  (token rows, then adaptive-softmax cluster rows)
  const float* out_bias;      // [Vx]
  bf16* x;                    // [B, d]   layer input / output (in: embedding of the current token)
  bf16* qt;                   // [B, H, d]
  float* bd;                  // [B, H, MLP]
  bf16* pctx;                 // [B, S, H, d]   normalised partial contexts
  float* pml;                 // [B, S, H, 2]   (max in the exp2 domain, sum)
  float* planes;              // [max(H, KSL)][B, d]
  bf16* y1;                   // [B, d]
  bf16* h1;                   // [B, di]
  float* logits;              // [B, ldl]
  unsigned long long* bar;    // grid barrier counter (monotonic)
  const int32_t* pos;
  unsigned long long* tstamp;   // profiling only: %globaltimer of CTA 0 at the step's start and after every grid barrier (NULL: off)
  int64_t ldl;
  int B, H, d, di, ML, MLP, L, Vx, S, per, KW, KSL;   // per = keys per split; KW = K-slice width of FF2, KSL = number of slices
  float eps, scale_log2;
};

// ------------------------------------------------------------------------------------------------------------------------------ primitives
__device__ __forceinline__ void mma16816(float* c, uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t wsel(const uint4& v, int i) { return i == 0 ? v.x : i == 1 ? v.y : i == 2 ? v.z : v.w; }
__device__ __forceinline__ void cpa16(void* dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cpa_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cpa_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void ldsm_x2(uint32_t& r0, uint32_t& r1, uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x2.shared.b16 {%0,%1}, [%2];" : "=r"(r0), "=r"(r1) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3, uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void bulk_row(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)), "l"(src), "r"(bytes),
               "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void fence_proxy_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }
__device__ __forceinline__ uint32_t pack2(float a, float b) {
  __nv_bfloat162 p = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&p);
}
__device__ __forceinline__ void cbar() { __syncthreads(); }

// Grid barrier number `k` (0-based, counted over the whole generation): every CTA adds 1, all wait for (k+1)*G.  Writes made before it by
// any CTA are visible after it to L1-bypassing loads.  (Data handed to the async proxy - ring rows read by TMA - is fenced where it is
// written and where the copies are issued, not here: fence.proxy.async on every barrier is not free.)
__device__ __forceinline__ void grid_barrier(unsigned long long* bar, unsigned long long k) {
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    atomicAdd(bar, 1ull);
    const unsigned long long target = (k + 1ull) * gridDim.x;
    const long long t0 = clock64();
    for (;;) {
      unsigned long long v;
      asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(bar) : "memory");
      if (v >= target) break;
      if (clock64() - t0 > 6000000000ll) __trap();       // ~3 s: a protocol bug becomes a CUDA error, not a hung GPU
    }
  }
  __syncthreads();
}

// rows [0, nrows) x cols [0, ncols) of a bf16 matrix (row pitch ld elements) -> shared memory with row pitch `pitch` bytes, 16-byte cp.async
// pieces; source rows are clamped to [0, src_rows-1] (padding rows repeat the last one: their results are never stored)
__device__ __forceinline__ void load_tile(unsigned char* dst, int pitch, const bf16* src, int64_t ld, int nrows, int src_rows, int ncols, int tid0,
                                          int nthr) {
  const int vpr = ncols >> 3;
  for (int e = tid0; e < nrows * vpr; e += nthr) {
    const int r = e / vpr, v = e - r * vpr;
    cpa16(dst + (size_t)r * pitch + v * 16, src + (int64_t)min(r, src_rows - 1) * ld + v * 8);
  }
}

// acc[i][j][:] += A[16 i .., K] W[8 (nt0 + j) .., K]^T over the whole K for this warp's n-tiles (no cross-warp reduction).  Both operands in
// shared memory, row pitch = 64 mod 128 bytes: one 16-byte piece per row is two k16 steps' worth of fragment registers (the k index inside
// an MMA is only a label - the same relabelling on both operands, as in dec_linear_kernel).
template <int MT, int NTW>
__device__ __forceinline__ void gemm_nsplit(float (&acc)[MT][NTW][4], const unsigned char* A, int pitchA, const unsigned char* W, int pitchW, int K,
                                            int nt0, int lane) {
  const int g = lane >> 2, t = lane & 3;
  for (int kb = 0; kb < K / 32; ++kb) {
    uint4 av[MT][2], wv[NTW];
#pragma unroll
    for (int i = 0; i < MT; ++i) {
      av[i][0] = *reinterpret_cast<const uint4*>(A + (size_t)(i * 16 + g) * pitchA + kb * 64 + t * 16);
      av[i][1] = *reinterpret_cast<const uint4*>(A + (size_t)(i * 16 + g + 8) * pitchA + kb * 64 + t * 16);
    }
#pragma unroll
    for (int j = 0; j < NTW; ++j) wv[j] = *reinterpret_cast<const uint4*>(W + (size_t)((nt0 + j) * 8 + g) * pitchW + kb * 64 + t * 16);
#pragma unroll
    for (int s = 0; s < 2; ++s)
#pragma unroll
      for (int i = 0; i < MT; ++i)
#pragma unroll
        for (int j = 0; j < NTW; ++j)
          mma16816(acc[i][j], wsel(av[i][0], 2 * s), wsel(av[i][1], 2 * s), wsel(av[i][0], 2 * s + 1), wsel(av[i][1], 2 * s + 1), wsel(wv[j], 2 * s),
                   wsel(wv[j], 2 * s + 1));
  }
}

struct Smem {
  unsigned char* base;
  int pitchA;                 // d*2 + 64
  unsigned char *A, *W, *A2, *W2;
  float* red;
};

// ------------------------------------------------------------------------------------------------------------------------------ chained stage
// item = (head h, chunk c).  GEMM1: t_h [B, 64] = A1 [B, d] W1[h*64.., d]^T (+ bias1) -> bf16 (shared memory) ; GEMM2: out [B, N2] = t_h W2c^T.
// KIND 0 = QT (A1 = x, W1 = W_q, bias rwb, W2c = W_kT[h][c*64.., 64], out -> qt bf16), KIND 1 = BD (bias rrb, W2c = r[c*128.., h*64..], out -> bd),
// KIND 2 = VAON (A1 = merged context of head h, W1 = W_v, W2c = W_o[c*64.., h*64..], out -> planes[h]).
// the weight slices of a chained item (they do not depend on other CTAs: issued BEFORE the grid barrier that precedes the stage)
template <int KIND>
__device__ __forceinline__ void chain_weights(const DpArgs& a, const DpLayer& ly, const Smem& sm, int h, int c) {
  constexpr int P2 = DP_DH * 2 + 64;
  const int tid = threadIdx.x, d = a.d;
  const bf16* w1 = (KIND == 2 ? ly.wv : ly.wq) + (int64_t)h * DP_DH * d;
  load_tile(sm.W, sm.pitchA, w1, d, DP_DH, DP_DH, d, tid, DP_THREADS);
  if (KIND == 0) load_tile(sm.W2, P2, ly.wkT + ((int64_t)h * d + c * 64) * DP_DH, DP_DH, 64, 64, DP_DH, tid, DP_THREADS);
  if (KIND == 1) load_tile(sm.W2, P2, ly.r + (int64_t)c * 128 * d + h * DP_DH, d, 128, a.ML + 1 - c * 128, DP_DH, tid, DP_THREADS);
  if (KIND == 2) load_tile(sm.W2, P2, ly.wo + (int64_t)c * 64 * d + h * DP_DH, d, 64, 64, DP_DH, tid, DP_THREADS);
  cpa_commit();
}

template <int MT, int KIND>
__device__ void chain_item(const DpArgs& a, const DpLayer& ly, const Smem& sm, int h, int c, bool w_ready) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, g = lane >> 2, t = lane & 3;
  const int d = a.d, B = a.B, H = a.H;
  constexpr int P2 = DP_DH * 2 + 64;              // pitch of the K = 64 operands
  constexpr int N2 = KIND == 1 ? 128 : 64, NTW2 = N2 / 64;
  // every global load of the item is issued up front (a dependent L2 round trip costs ~0.8 us here)
  float bias0 = 0.f, bias1 = 0.f;
  if (KIND != 2) {
    const float* bias = KIND == 0 ? ly.rwb : ly.rrb;
    bias0 = bias[h * DP_DH + warp * 8 + 2 * t]; bias1 = bias[h * DP_DH + warp * 8 + 2 * t + 1];
  }
  if (tid < DP_CWARPS * 32) {
    if (!w_ready) chain_weights<KIND>(a, ly, sm, h, c);
    if (KIND == 2) {
      // A1 = merge of the ring splits of head h:  ctx = sum_z coef_z pctx[b, z, h, :],  coef_z = exp2(m_z - M) l_z / sum(...).
      // Dependent L2 round trips are what this costs, so: (1) one row per thread builds the coefficients with all its (m, l) loads in flight,
      // (2) the vectors are merged 4 at a time with the loads of up to 4 splits each in flight.
      float* coef = sm.red;                             // [MT*16][DP_MAXS]
      for (int b = tid; b < MT * 16; b += DP_THREADS) {
        const int bb = min(b, B - 1);
        float2 ml[DP_MAXS];
#pragma unroll
        for (int z = 0; z < DP_MAXS; ++z)
          ml[z] = z < a.S ? __ldcg(reinterpret_cast<const float2*>(a.pml + (((int64_t)bb * a.S + z) * H + h) * 2)) : make_float2(-INFINITY, 0.f);
        float M = -INFINITY;
#pragma unroll
        for (int z = 0; z < DP_MAXS; ++z) M = fmaxf(M, ml[z].x);
        float cz[DP_MAXS], Lsum = 0.f;
#pragma unroll
        for (int z = 0; z < DP_MAXS; ++z) { cz[z] = (ml[z].x == -INFINITY) ? 0.f : exp2f(ml[z].x - M) * ml[z].y; Lsum += cz[z]; }
        const float inv = Lsum > 0.f ? 1.f / Lsum : 0.f;
#pragma unroll
        for (int z = 0; z < DP_MAXS; ++z) coef[b * DP_MAXS + z] = cz[z] * inv;
      }
      __syncthreads();
      const int vpr = d / 8, nvec = MT * 16 * vpr;
      for (int v0 = tid; v0 < nvec; v0 += 4 * DP_THREADS) {
        float acc8[4][8];
#pragma unroll
        for (int q = 0; q < 4; ++q)
#pragma unroll
          for (int k = 0; k < 8; ++k) acc8[q][k] = 0.f;
        for (int z0 = 0; z0 < a.S; z0 += 4) {
          uint4 u[4][4];
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const int v = v0 + q * DP_THREADS, b = min(v / vpr, MT * 16 - 1), bb = min(b, B - 1), c0 = (v % vpr) * 8;
#pragma unroll
            for (int zz = 0; zz < 4; ++zz)
              u[q][zz] = (v < nvec && z0 + zz < a.S) ? __ldcg(reinterpret_cast<const uint4*>(a.pctx + (((int64_t)bb * a.S + z0 + zz) * H + h) * d + c0))
                                                    : make_uint4(0, 0, 0, 0);
          }
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const int v = v0 + q * DP_THREADS, b = min(v / vpr, MT * 16 - 1);
#pragma unroll
            for (int zz = 0; zz < 4; ++zz) {
              const float cf = z0 + zz < a.S ? coef[b * DP_MAXS + z0 + zz] : 0.f;
              const bf16* e = reinterpret_cast<const bf16*>(&u[q][zz]);
#pragma unroll
              for (int k = 0; k < 8; ++k) acc8[q][k] = fmaf(cf, __bfloat162float(e[k]), acc8[q][k]);
            }
          }
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int v = v0 + q * DP_THREADS;
          if (v < nvec) {
            uint4 o;
            o.x = pack2(acc8[q][0], acc8[q][1]); o.y = pack2(acc8[q][2], acc8[q][3]); o.z = pack2(acc8[q][4], acc8[q][5]); o.w = pack2(acc8[q][6], acc8[q][7]);
            *reinterpret_cast<uint4*>(sm.A + (size_t)(v / vpr) * sm.pitchA + (v % vpr) * 16) = o;
          }
        }
      }
    } else {
      load_tile(sm.A, sm.pitchA, a.x, d, MT * 16, B, d, tid, DP_CWARPS * 32);
    }
    cpa_commit();
    cpa_wait_all();
  }
  __syncthreads();
  if (tid < DP_CWARPS * 32) {
    // ---- GEMM1: warp w owns features 8w .. 8w+7 of the head
    float acc[MT][1][4];
#pragma unroll
    for (int i = 0; i < MT; ++i)
#pragma unroll
      for (int q = 0; q < 4; ++q) acc[i][0][q] = 0.f;
    gemm_nsplit<MT, 1>(acc, sm.A, sm.pitchA, sm.W, sm.pitchA, d, warp, lane);
    const int n = warp * 8 + 2 * t;
    const float b0 = bias0, b1 = bias1;
#pragma unroll
    for (int i = 0; i < MT; ++i) {
      *reinterpret_cast<uint32_t*>(sm.A2 + (size_t)(i * 16 + g) * P2 + n * 2) = pack2(acc[i][0][0] + b0, acc[i][0][1] + b1);
      *reinterpret_cast<uint32_t*>(sm.A2 + (size_t)(i * 16 + g + 8) * P2 + n * 2) = pack2(acc[i][0][2] + b0, acc[i][0][3] + b1);
    }
  }
  __syncthreads();
  if (tid < DP_CWARPS * 32) {
    // ---- GEMM2 over K = 64: warp w owns n-tiles NTW2 w ..
    float acc[MT][NTW2][4];
#pragma unroll
    for (int i = 0; i < MT; ++i)
#pragma unroll
      for (int j = 0; j < NTW2; ++j)
#pragma unroll
        for (int q = 0; q < 4; ++q) acc[i][j][q] = 0.f;
    gemm_nsplit<MT, NTW2>(acc, sm.A2, P2, sm.W2, P2, DP_DH, warp * NTW2, lane);
#pragma unroll
    for (int i = 0; i < MT; ++i)
#pragma unroll
      for (int j = 0; j < NTW2; ++j)
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          const int b = i * 16 + g + half * 8;
          const int n = (warp * NTW2 + j) * 8 + 2 * t;
          if (b >= B) continue;
          const float v0 = acc[i][j][half * 2], v1 = acc[i][j][half * 2 + 1];
          if (KIND == 0) {
            *reinterpret_cast<uint32_t*>(a.qt + ((int64_t)b * H + h) * d + c * 64 + n) = pack2(v0, v1);
          } else if (KIND == 1) {
            const int x = c * 128 + n;
            float* dst = a.bd + ((int64_t)b * H + h) * a.MLP + x;
            if (x <= a.ML) dst[0] = v0;
            if (x + 1 <= a.ML) dst[1] = v1;
          } else {
            *reinterpret_cast<float2*>(a.planes + ((int64_t)h * B + b) * d + c * 64 + n) = make_float2(v0, v1);
          }
        }
  }
  __syncthreads();
}

// ------------------------------------------------------------------------------------------------------------------------------ linear stage
// out[B, 16] = A[B, K] (row pitch lda, columns [k0, k0+K)) W[n0.., k0..]^T: the 8 warps split K in 32-wide blocks (warp w takes block w of every
// 256-column chunk) and meet in shared memory.  MODE 0: + bias, ReLU, bf16 -> h1 ; MODE 1: fp32 plane ; MODE 2: + bias, fp32 -> logits
__device__ __forceinline__ void lin_weights(const Smem& sm, const bf16* W, int64_t ldw, int N, int K, int n0, int k0) {
  load_tile(sm.W, K * 2 + 64, W + (int64_t)n0 * ldw + k0, ldw, 16, N - n0, K, threadIdx.x, DP_THREADS);
  cpa_commit();
}

template <int MT, int MODE>
__device__ void lin_item(const DpArgs& a, const Smem& sm, const bf16* A, int64_t lda, const bf16* W, int64_t ldw, const float* bias, int N, int K, int n0,
                         int k0, void* out, int64_t ldo, bool w_ready) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, g = lane >> 2, t = lane & 3;
  const int B = a.B;
  const int pitch = K * 2 + 64;
  // epilogue bias of this thread's output column (e = tid + 256 i: column tid & 15 for every i), loaded before anything else
  const float bias_e = (MODE != 1 && n0 + (tid & 15) < N) ? bias[n0 + (tid & 15)] : 0.f;
  if (!w_ready) lin_weights(sm, W, ldw, N, K, n0, k0);
  load_tile(sm.A, pitch, A + k0, lda, MT * 16, B, K, tid, DP_THREADS);
  cpa_commit();
  cpa_wait_all();
  __syncthreads();
  float (*red)[MT * 16][17] = reinterpret_cast<float (*)[MT * 16][17]>(sm.red);
  if (tid < DP_CWARPS * 32) {
    float acc[MT][2][4];
#pragma unroll
    for (int i = 0; i < MT; ++i)
#pragma unroll
      for (int j = 0; j < 2; ++j)
#pragma unroll
        for (int q = 0; q < 4; ++q) acc[i][j][q] = 0.f;
    for (int kb = warp; kb < K / 32; kb += DP_CWARPS) {
      uint4 av[MT][2], wv[2];
#pragma unroll
      for (int i = 0; i < MT; ++i) {
        av[i][0] = *reinterpret_cast<const uint4*>(sm.A + (size_t)(i * 16 + g) * pitch + kb * 64 + t * 16);
        av[i][1] = *reinterpret_cast<const uint4*>(sm.A + (size_t)(i * 16 + g + 8) * pitch + kb * 64 + t * 16);
      }
#pragma unroll
      for (int j = 0; j < 2; ++j) wv[j] = *reinterpret_cast<const uint4*>(sm.W + (size_t)(j * 8 + g) * pitch + kb * 64 + t * 16);
#pragma unroll
      for (int s = 0; s < 2; ++s)
#pragma unroll
        for (int i = 0; i < MT; ++i)
#pragma unroll
          for (int j = 0; j < 2; ++j)
            mma16816(acc[i][j], wsel(av[i][0], 2 * s), wsel(av[i][1], 2 * s), wsel(av[i][0], 2 * s + 1), wsel(av[i][1], 2 * s + 1), wsel(wv[j], 2 * s),
                     wsel(wv[j], 2 * s + 1));
    }
#pragma unroll
    for (int i = 0; i < MT; ++i)
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        red[warp][i * 16 + g][j * 8 + 2 * t] = acc[i][j][0];
        red[warp][i * 16 + g][j * 8 + 2 * t + 1] = acc[i][j][1];
        red[warp][i * 16 + g + 8][j * 8 + 2 * t] = acc[i][j][2];
        red[warp][i * 16 + g + 8][j * 8 + 2 * t + 1] = acc[i][j][3];
      }
  }
  __syncthreads();
  for (int e = tid; e < MT * 16 * 16; e += DP_THREADS) {
    const int m = e >> 4, nl = e & 15, n = n0 + nl;
    if (m < B && n < N) {
      float v = 0.f;
#pragma unroll
      for (int w = 0; w < DP_CWARPS; ++w) v += red[w][m][nl];
      if (MODE == 0) {
        v = fmaxf(v + bias_e, 0.f);
        reinterpret_cast<bf16*>(out)[(int64_t)m * ldo + n] = __float2bfloat16_rn(v);
      } else if (MODE == 1) {
        reinterpret_cast<float*>(out)[(int64_t)m * ldo + n] = v;
      } else {
        reinterpret_cast<float*>(out)[(int64_t)m * ldo + n] = v + bias_e;
      }
    }
  }
  __syncthreads();
}

// ------------------------------------------------------------------------------------------------------------------------------ LayerNorm rows
// y[b] = LayerNorm(res[b] + sum_p planes[p][b] (+ bias)) * gamma + beta, fp32 statistics (two-pass variance).  One 8-column vector per thread
// (D/8 threads per row, 2048/D rows per CTA) so that EVERY load of the stage - residual, up to 8 planes, bias, gamma, beta - is in flight
// at once: one L2 round trip instead of five.  Optionally also stored to the next layer's ring slot of this step (then fenced for the TMA).
template <int D>
__device__ void ln_rows(const DpArgs& a, const bf16* res, int nplanes, const float* bias, const float* gamma, const float* beta, bf16* y, bf16* ring,
                        int cur, float* lnred) {
  constexpr int LPR = D / 8;                        // threads per row
  constexpr int RPC = DP_THREADS / LPR;             // rows per CTA and pass
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int B = a.B;
  const int rl = tid / LPR, c = (tid % LPR) * 8;
  for (int b0 = blockIdx.x * RPC; b0 < B; b0 += gridDim.x * RPC) {          // (uniform per CTA: the barriers below are safe)
    const int b = b0 + rl;
    const bool on = b < B;
    const int bb = on ? b : B - 1;
    const uint4 u = __ldcg(reinterpret_cast<const uint4*>(res + (int64_t)bb * D + c));
    float4 pl[8][2];
#pragma unroll
    for (int p = 0; p < 8; ++p) {
      pl[p][0] = p < nplanes ? __ldcg(reinterpret_cast<const float4*>(a.planes + ((int64_t)p * B + bb) * D + c)) : make_float4(0.f, 0.f, 0.f, 0.f);
      pl[p][1] = p < nplanes ? __ldcg(reinterpret_cast<const float4*>(a.planes + ((int64_t)p * B + bb) * D + c + 4)) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    const float4 g0 = *reinterpret_cast<const float4*>(gamma + c), g1 = *reinterpret_cast<const float4*>(gamma + c + 4);
    const float4 e0 = *reinterpret_cast<const float4*>(beta + c), e1 = *reinterpret_cast<const float4*>(beta + c + 4);
    float4 s0 = make_float4(0.f, 0.f, 0.f, 0.f), s1 = s0;
    if (bias) { s0 = *reinterpret_cast<const float4*>(bias + c); s1 = *reinterpret_cast<const float4*>(bias + c + 4); }
    float v[8];
    {
      const bf16* eb = reinterpret_cast<const bf16*>(&u);
#pragma unroll
      for (int k = 0; k < 8; ++k) v[k] = __bfloat162float(eb[k]);
    }
#pragma unroll
    for (int p = 0; p < 8; ++p) {
      v[0] += pl[p][0].x; v[1] += pl[p][0].y; v[2] += pl[p][0].z; v[3] += pl[p][0].w;
      v[4] += pl[p][1].x; v[5] += pl[p][1].y; v[6] += pl[p][1].z; v[7] += pl[p][1].w;
    }
    v[0] += s0.x; v[1] += s0.y; v[2] += s0.z; v[3] += s0.w; v[4] += s1.x; v[5] += s1.y; v[6] += s1.z; v[7] += s1.w;
    // row sums over the LPR threads of the row: inside a warp by shuffles; rows wider than a warp meet in shared memory
    auto rowsum = [&](float x) {
      if (LPR >= 32) {
        x = warp_sum(x);
        if (LPR > 32) {
          constexpr int WPR = LPR / 32 > 0 ? LPR / 32 : 1;
          __syncthreads();
          if (lane == 0) lnred[warp] = x;
          __syncthreads();
          float tt = 0.f;
#pragma unroll
          for (int w = 0; w < WPR; ++w) tt += lnred[(warp / WPR) * WPR + w];
          x = tt;
        }
      } else {
#pragma unroll
        for (int o = LPR / 2; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
      }
      return x;
    };
    float sum = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) sum += v[k];
    const float mu = rowsum(sum) / D;
    float q = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) { const float tt = v[k] - mu; q += tt * tt; }
    const float rs = rsqrtf(rowsum(q) / D + a.eps);
    if (on) {
      uint4 o;
      o.x = pack2((v[0] - mu) * rs * g0.x + e0.x, (v[1] - mu) * rs * g0.y + e0.y);
      o.y = pack2((v[2] - mu) * rs * g0.z + e0.z, (v[3] - mu) * rs * g0.w + e0.w);
      o.z = pack2((v[4] - mu) * rs * g1.x + e1.x, (v[5] - mu) * rs * g1.y + e1.y);
      o.w = pack2((v[6] - mu) * rs * g1.z + e1.z, (v[7] - mu) * rs * g1.w + e1.w);
      *reinterpret_cast<uint4*>(y + (int64_t)b * D + c) = o;
      if (ring) *reinterpret_cast<uint4*>(ring + ((int64_t)b * a.ML + cur) * D + c) = o;
    }
  }
  if (ring) fence_proxy_async_all();                  // the ring rows are read by TMA (async proxy) in the next layer's attention stage
}

// ------------------------------------------------------------------------------------------------------------------------------ attention
// D = d_model.  item = (sequence b, split z): keys [z per, min((z+1) per, ML)) in stages of 64.  A stage = D/64 TMA boxes of 64 rows x 128
// bytes (SWIZZLE_128B: 16-byte chunk c of row r sits at chunk c ^ (r & 7)), 8 copies instead of one per row (per-row bulk copies stalled the
// issuing warp for ~2000 cycles per stage).  Rows past the split belong to later positions (finite, they meet p = 0) or are zero-filled.
// The MMAs run "transposed" so that no tile row is padding (mma.sync issues one HMMA per ~16 cycles and sub-partition: it is the bound):
//   phase A  S^T[key, head] = hid[key, :] . qt[head, :]      M = 16 keys (ldmatrix), N = 8 heads, warp = (key tile, half of d)
//   phase B  ctx^T[col, head] += hid[key, col] p[head, key]  M = 16 columns (ldmatrix.trans), N = 8 heads, K = keys, warp = D/8 columns
template <int D>
__device__ void att_item(const DpArgs& a, const CUtensorMap* tm, unsigned char* stage0, float* ssm, unsigned char* qsm, uint64_t* full, uint32_t& uses,
                         int b, int z, int cur) {
  constexpr int NB = D / 64;                        // 128-byte column blocks of a row
  constexpr int STAGE = NB * 8192;
  constexpr int NKS = D / 16;                       // k16 steps of a score row
  constexpr int KH = NKS / 2;                       //   ... per d-half
  constexpr int CW = D / DP_CWARPS;                 // context columns owned by a warp
  constexpr int NMT = CW / 16;                      // m-tiles of those
  constexpr int QPITCH = D * 2 + 16;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, g = lane >> 2, t = lane & 3;
  const int H = a.H, ML = a.ML;
  const int s_begin = z * a.per, s_end = min(s_begin + a.per, ML);
  const int nst = (s_end - s_begin + DP_KS - 1) / DP_KS;
  // one 8 KB box per warp (a thread that issued all eight spent ~800 cycles on it): the transaction count is armed by warp 0; a box that
  // completes before that only drives the count negative - the phase cannot flip before the arming arrival
  auto issue = [&](int i) {
    if (lane == 0 && i < nst) {
      const uint32_t u = uses + i, st = u % DP_NST;
      if (warp == 0) mbar_expect_tx(&full[st], STAGE);
      for (int cb = warp; cb < NB; cb += DP_CWARPS) tma_load_2d(stage0 + (size_t)st * STAGE + cb * 8192, tm, &full[st], cb * 64, b * ML + s_begin + i * DP_KS);
    }
  };
#pragma unroll
  for (int i = 0; i < DP_NST; ++i) issue(i);
  // B operand of the scores: qt[b] (rows = heads, zero rows past H) staged in shared memory
  for (int e = tid; e < 8 * (D / 8); e += DP_THREADS) {
    const int hh = e / (D / 8), v = e % (D / 8);
    uint4 u = make_uint4(0, 0, 0, 0);
    if (hh < H) u = __ldcg(reinterpret_cast<const uint4*>(a.qt + ((int64_t)b * H + hh) * D + v * 8));
    *reinterpret_cast<uint4*>(qsm + (size_t)hh * QPITCH + v * 16) = u;
  }
  __syncthreads();
  const int mt = warp & 3, kh = warp >> 2;          // phase A: keys 16 mt .., k16 steps [kh KH, (kh+1) KH)
  const unsigned char* qrow = qsm + (size_t)g * QPITCH + t * 4;
  float acc[NMT][4];
#pragma unroll
  for (int j = 0; j < NMT; ++j)
#pragma unroll
    for (int q = 0; q < 4; ++q) acc[j][q] = 0.f;
  float m_run = -INFINITY, l_run = 0.f;              // of head g (replicated over t and over the warps: same inputs, same operations)
  long long pc[6] = {0, 0, 0, 0, 0, 0};              // profiling (a.tstamp set): cycles before the wait / in it / phase A / barrier / phase B
  const bool prof = a.tstamp && blockIdx.x == 0 && tid == 0;
  long long tprev = prof ? clock64() : 0;
  auto lap = [&](int k) { if (prof) { const long long tn = clock64(); pc[k] += tn - tprev; tprev = tn; } };
  const float* bd0 = a.bd + ((int64_t)b * H + min(2 * t, H - 1)) * a.MLP;
  const float* bd1 = a.bd + ((int64_t)b * H + min(2 * t + 1, H - 1)) * a.MLP;
  for (int i = 0; i < nst; ++i) {
    const uint32_t u = uses + i, st = u % DP_NST;
    const int s0 = s_begin + i * DP_KS;
    float* S = ssm + (size_t)(i & 1) * 2 * 8 * DP_SPITCH;
    // position term of this thread's keys (16 mt + g, + 8) and heads (2t, 2t+1): r row x = ML - ((cur - s) mod ML); added by the kh = 0 warps
    float bdv[4] = {0.f, 0.f, 0.f, 0.f};
    if (kh == 0) {
#pragma unroll
      for (int hf = 0; hf < 2; ++hf) {
        const int sk = s0 + mt * 16 + g + hf * 8;
        if (sk < s_end) {
          const int x = sk <= cur ? ML - cur + sk : sk - cur;
          if (2 * t < H) bdv[hf * 2] = __ldcg(bd0 + x);
          if (2 * t + 1 < H) bdv[hf * 2 + 1] = __ldcg(bd1 + x);
        }
      }
    }
    lap(0);
    mbar_wait(&full[st], (u / DP_NST) & 1);
    lap(1);
    const uint32_t sbase = smem_u32(stage0 + (size_t)st * STAGE);
    {
      // phase A
      float c4[2][4];
#pragma unroll
      for (int q = 0; q < 2; ++q) { c4[q][0] = 0.f; c4[q][1] = 0.f; c4[q][2] = 0.f; c4[q][3] = 0.f; }
      const int row = mt * 16 + (lane & 7) + ((lane >> 3) & 1) * 8;
      const uint32_t rbase = sbase + row * 128, rsw = row & 7, chi = lane >> 4;
#pragma unroll
      for (int k2 = 0; k2 < KH; k2 += 2) {            // two independent accumulator chains
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
      // c: (key 16 mt + g, heads 2t, 2t+1), (key + 8, same heads)
      float* Sh = S + (size_t)kh * 8 * DP_SPITCH + mt * 16 + g;
      Sh[(2 * t) * DP_SPITCH] = c4[0][0] + c4[1][0] + bdv[0];
      Sh[(2 * t + 1) * DP_SPITCH] = c4[0][1] + c4[1][1] + bdv[1];
      Sh[(2 * t) * DP_SPITCH + 8] = c4[0][2] + c4[1][2] + bdv[2];
      Sh[(2 * t + 1) * DP_SPITCH + 8] = c4[0][3] + c4[1][3] + bdv[3];
    }
    lap(2);
    cbar();
    lap(3);
    // every warp is past phase A of this stage, hence done with the previous stage's buffer: refill it with stage i - 1 + NST
    if (i >= 1) issue(i - 1 + DP_NST);
    {
      // phase B: this thread's probabilities of head g (k16 block ks: keys 16 ks + 2t, +1, +8, +9) are the B fragments of every m-tile
      float sc[4][4];
      float mx = -INFINITY;
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) {
        const float* p0 = S + g * DP_SPITCH + ks * 16 + 2 * t;
        const float2 lo0 = *reinterpret_cast<const float2*>(p0), hi0 = *reinterpret_cast<const float2*>(p0 + 8);
        const float2 lo1 = *reinterpret_cast<const float2*>(p0 + 8 * DP_SPITCH), hi1 = *reinterpret_cast<const float2*>(p0 + 8 * DP_SPITCH + 8);
        const int kb = s0 + ks * 16 + 2 * t;
        sc[ks][0] = kb < s_end ? (lo0.x + lo1.x) * a.scale_log2 : -INFINITY;
        sc[ks][1] = kb + 1 < s_end ? (lo0.y + lo1.y) * a.scale_log2 : -INFINITY;
        sc[ks][2] = kb + 8 < s_end ? (hi0.x + hi1.x) * a.scale_log2 : -INFINITY;
        sc[ks][3] = kb + 9 < s_end ? (hi0.y + hi1.y) * a.scale_log2 : -INFINITY;
        mx = fmaxf(fmaxf(mx, fmaxf(sc[ks][0], sc[ks][1])), fmaxf(sc[ks][2], sc[ks][3]));
      }
      mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 1));
      mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 2));
      const float m_new = fmaxf(m_run, mx);
      const float m_safe = (m_new == -INFINITY) ? 0.f : m_new;
      const float corr = exp2f(m_run - m_safe);         // m_run = -inf -> 0
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
      // the accumulators hold heads 2t, 2t+1: their rescale factors live in lanes 8t and 8t + 4
      const float ce = __shfl_sync(0xffffffffu, corr, 8 * t), co = __shfl_sync(0xffffffffu, corr, 8 * t + 4);
#pragma unroll
      for (int j = 0; j < NMT; ++j) { acc[j][0] *= ce; acc[j][1] *= co; acc[j][2] *= ce; acc[j][3] *= co; }
      // A fragments = hid^T through ldmatrix.trans: matrices (keys 0-7 | cols 0-7), (keys 0-7 | cols 8-15), (keys 8-15 | cols 0-7), (keys 8-15 | 8-15)
      const int krow = (lane & 7) + (lane >> 4) * 8;
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) {                // k16 block outer, m-tiles inner: NMT independent accumulator chains
        const int row = ks * 16 + krow;
        const uint32_t roff = row * 128, rsw = row & 7;
#pragma unroll
        for (int j = 0; j < NMT; ++j) {
          const int col0 = warp * CW + j * 16;
          const uint32_t ch = ((col0 & 63) >> 3) + ((lane >> 3) & 1);
          uint32_t a0, a1, a2, a3;
          asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
                       : "=r"(a0), "=r"(a1), "=r"(a2), "=r"(a3)
                       : "r"(sbase + (col0 >> 6) * 8192 + roff + ((ch ^ rsw) << 4)));
          mma16816(acc[j], a0, a1, a2, a3, pb0[ks], pb1[ks]);
        }
      }
    }
    lap(4);
  }
  uses += nst;
  if (prof) for (int k = 0; k < 5; ++k) a.tstamp[120 + k] = (unsigned long long)pc[k];
  // ---- normalised partial context of this split + its (max, sum): acc[j] = (col0 + g | heads 2t, 2t+1), (col0 + g + 8 | same heads)
  {
    const float inv = l_run > 0.f ? 1.f / l_run : 0.f;
    const float ie = __shfl_sync(0xffffffffu, inv, 8 * t), io = __shfl_sync(0xffffffffu, inv, 8 * t + 4);
    bf16* de = a.pctx + (((int64_t)b * a.S + z) * H + 2 * t) * D;
    bf16* d_odd = de + D;
#pragma unroll
    for (int j = 0; j < NMT; ++j) {
      const int c = warp * CW + j * 16 + g;
      if (2 * t < H) { de[c] = __float2bfloat16_rn(acc[j][0] * ie); de[c + 8] = __float2bfloat16_rn(acc[j][2] * ie); }
      if (2 * t + 1 < H) { d_odd[c] = __float2bfloat16_rn(acc[j][1] * io); d_odd[c + 8] = __float2bfloat16_rn(acc[j][3] * io); }
    }
    if (warp == 0 && t == 0 && g < H) {
      float* ml = a.pml + (((int64_t)b * a.S + z) * H + g) * 2;
      ml[0] = m_run; ml[1] = l_run;
    }
  }
  cbar();      // the score tiles and qsm are reused by this CTA's next item
}

// ------------------------------------------------------------------------------------------------------------------------------ the step
template <int MT, int D>
__global__ void __launch_bounds__(DP_THREADS, 1) decode_persist_kernel(const DpArgs a) {
  extern __shared__ __align__(128) unsigned char dp_smem[];
  __shared__ uint64_t full[DP_NST];
  __shared__ float lnred[DP_CWARPS];
  const int tid = threadIdx.x, warp = tid >> 5;
  const int G = gridDim.x, cta = blockIdx.x;
  const int d = a.d, B = a.B, H = a.H;
  Smem sm;
  sm.base = dp_smem;
  sm.pitchA = (d > a.KW ? d : a.KW) * 2 + 64;
  sm.A = dp_smem;
  sm.W = sm.A + (size_t)MT * 16 * sm.pitchA;
  sm.A2 = sm.W + (size_t)64 * sm.pitchA;
  sm.W2 = sm.A2 + (size_t)MT * 16 * (DP_DH * 2 + 64);
  sm.red = reinterpret_cast<float*>(sm.W2 + (size_t)128 * (DP_DH * 2 + 64));
  // attention view of the same memory: ring stages (1024-byte aligned: swizzle atoms), then the score tiles and the staged qt rows
  unsigned char* stage0 = dp_smem + ((1024u - (smem_u32(dp_smem) & 1023u)) & 1023u);
  float* ssm = reinterpret_cast<float*>(stage0 + (size_t)DP_NST * (D / 64) * 8192);
  unsigned char* qsm = reinterpret_cast<unsigned char*>(ssm + 2 * 2 * 8 * DP_SPITCH);
  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < DP_NST; ++s) mbar_init(&full[s], 1);
    fence_barrier_init();
  }
  __syncthreads();
  uint32_t uses = 0;                                  // ring-stage uses of this CTA so far
  const int p = *a.pos;
  const int cur = p % a.ML;
  // barrier numbering continues over the steps of one generation (the counter is never reset): 7 per layer
  unsigned long long kbar = (unsigned long long)p * (7ull * a.L);
  int nstamp = 0;
  auto stamp = [&]() {
    if (a.tstamp && cta == 0 && tid == 0) {
      unsigned long long tt;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(tt));
      a.tstamp[nstamp++] = tt;
    }
  };
  stamp();
  const int nC = d / 64;                              // 64-column chunks of a row
  const int nQT = H * nC, nD = (a.ML + 1 + 127) / 128, nBD = H * nD;
  const int nFF1 = (a.di + 15) / 16, nbF2 = (d + 15) / 16, nFF2 = nbF2 * a.KSL, nLM = (a.Vx + 15) / 16;
  // Weight slices do not depend on other CTAs: those of a stage's FIRST item are requested before the grid barrier in front of it (and even
  // across a LayerNorm stage, which leaves the operand memory alone), so after the barrier only the activations are still to come.
  auto qtbd_weights = [&](const DpLayer& ly, int item) {
    if (item < nQT) chain_weights<0>(a, ly, sm, item / nC, item % nC);
    else chain_weights<1>(a, ly, sm, (item - nQT) / nD, (item - nQT) % nD);
  };
  for (int l = 0; l < a.L; ++l) {
    const DpLayer ly = a.layers[l];
    // ---- QTBD (layer 0 also stores the embedding row into its ring slot; later layers got theirs from the previous LN2)
    if (l == 0) {
      for (int e = cta * DP_THREADS + tid; e < B * (d / 8); e += G * DP_THREADS) {
        const int b = e / (d / 8), c = (e % (d / 8)) * 8;
        *reinterpret_cast<uint4*>(ly.ring + ((int64_t)b * a.ML + cur) * d + c) = __ldcg(reinterpret_cast<const uint4*>(a.x + (int64_t)b * d + c));
      }
      fence_proxy_async_all();
    }
    for (int item = cta, first = 1; item < nQT + nBD; item += G, first = 0) {
      const bool wr = first && l > 0;
      if (item < nQT) chain_item<MT, 0>(a, ly, sm, item / nC, item % nC, wr);
      else chain_item<MT, 1>(a, ly, sm, (item - nQT) / nD, (item - nQT) % nD, wr);
    }
    grid_barrier(a.bar, kbar++);
    stamp();
    // ---- ATT (the operand memory changes hands: generic writes so far, TMA writes from here)
    fence_proxy_async_all();
    __syncthreads();
    if (a.tstamp && tid == 0 && l == a.L - 1) { unsigned long long tt; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(tt)); a.tstamp[200 + 2 * cta] = tt; }
    for (int item = cta; item < B * a.S; item += G) att_item<D>(a, &a.tmaps[l], stage0, ssm, qsm, full, uses, item / a.S, item % a.S, cur);
    if (a.tstamp && tid == 0 && l == a.L - 1) { unsigned long long tt; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(tt)); a.tstamp[201 + 2 * cta] = tt; }
    // every ring copy this CTA issued has landed and been consumed (each stage's full barrier was waited on) before the memory is reused
    if (cta < nQT) chain_weights<2>(a, ly, sm, cta / nC, cta % nC);
    grid_barrier(a.bar, kbar++);
    stamp();
    // ---- VAON
    for (int item = cta, first = 1; item < nQT; item += G, first = 0) chain_item<MT, 2>(a, ly, sm, item / nC, item % nC, first != 0);
    if (cta < nFF1) lin_weights(sm, ly.w1, d, a.di, d, cta * 16, 0);
    grid_barrier(a.bar, kbar++);
    stamp();
    // ---- LN1
    ln_rows<D>(a, a.x, H, nullptr, ly.ln1w, ly.ln1b, a.y1, nullptr, 0, lnred);
    grid_barrier(a.bar, kbar++);
    stamp();
    // ---- FF1
    for (int item = cta, first = 1; item < nFF1; item += G, first = 0)
      lin_item<MT, 0>(a, sm, a.y1, d, ly.w1, d, ly.b1, a.di, d, item * 16, 0, a.h1, a.di, first != 0);
    if (cta < nFF2) lin_weights(sm, ly.w2, a.di, d, a.KW, (cta % nbF2) * 16, (cta / nbF2) * a.KW);
    grid_barrier(a.bar, kbar++);
    stamp();
    // ---- FF2
    for (int item = cta, first = 1; item < nFF2; item += G, first = 0) {
      const int ks = item / nbF2, n0 = (item % nbF2) * 16;
      lin_item<MT, 1>(a, sm, a.h1, a.di, ly.w2, a.di, nullptr, d, a.KW, n0, ks * a.KW, a.planes + (int64_t)ks * B * d, d, first != 0);
    }
    if (l + 1 < a.L) { if (cta < nQT + nBD) qtbd_weights(a.layers[l + 1], cta); }
    else if (cta < nLM) lin_weights(sm, a.E, d, a.Vx, d, cta * 16, 0);
    grid_barrier(a.bar, kbar++);
    stamp();
    // ---- LN2 (+ next layer's ring slot)
    ln_rows<D>(a, a.y1, a.KSL, ly.b2, ly.ln2w, ly.ln2b, a.x, l + 1 < a.L ? a.layers[l + 1].ring : nullptr, cur, lnred);
    grid_barrier(a.bar, kbar++);
    stamp();
  }
  // ---- LM head
  for (int item = cta, first = 1; item < nLM; item += G, first = 0)
    lin_item<MT, 2>(a, sm, a.x, d, a.E, d, a.out_bias, a.Vx, d, item * 16, 0, a.logits, a.ldl, first != 0);
  stamp();
  (void)warp;
}

struct DpHostTable {          // what txl_decode_persist_step(build = 1) leaves at the start of the workspace
  DpLayer layers[64];
  CUtensorMap tm[64];
};

size_t dp_smem_bytes(int MT, int d, int KW) {
  const size_t pitchA = (size_t)(d > KW ? d : KW) * 2 + 64;
  const size_t gemm = (size_t)MT * 16 * pitchA + 64 * pitchA + (size_t)MT * 16 * (DP_DH * 2 + 64) + 128 * (DP_DH * 2 + 64) + (size_t)DP_CWARPS * MT * 16 * 17 * 4;
  const size_t att = 1024 + (size_t)DP_NST * (d / 64) * 8192 + 2 * 2 * 8 * DP_SPITCH * 4 + 8 * ((size_t)d * 2 + 16);
  return (gemm > att ? gemm : att) + 128;
}
}  // namespace

static int dp_geometry(int B, int H, int dh, int d, int di, int ML, int L, int Vx, int* S, int* per, int* KW, int* KSL, int* MLP) {
  if (!(B >= 1 && B <= 64 && dh == DP_DH && H >= 1 && H <= 8 && H * dh == d && (d == 128 || d == 512) && ML >= 1 && L >= 1 && L <= 64 && Vx >= 1)) return 0;
  if (di % 32 || !(di <= 512 || di % 512 == 0)) return 0;
  *KW = di <= 512 ? di : 512;
  *KSL = di / *KW;
  if (*KSL > 8) return 0;                            // planes are shared with the H <= 8 head planes
  int s = txl_num_sms() / B;
  if (s > (ML + DP_KS - 1) / DP_KS) s = (ML + DP_KS - 1) / DP_KS;
  if (s > DP_MAXS) s = DP_MAXS;
  if (s < 1) s = 1;
  int p = (ML + s - 1) / s;
  p = (p + DP_KS - 1) / DP_KS * DP_KS;               // whole stages per split
  s = (ML + p - 1) / p;
  *S = s; *per = p;
  *MLP = (ML + 1 + 3) / 4 * 4;
  return 1;
}

static unsigned long long* g_dp_tstamp = nullptr;
extern "C" int txl_decode_persist_set_timestamps(unsigned long long* dev_buf) { g_dp_tstamp = dev_buf; return TXL_OK; }

extern "C" int txl_decode_persist_supported(int B, int H, int dh, int d, int di, int ML, int L, int Vx) {
  int S, per, KW, KSL, MLP;
  return dp_geometry(B, H, dh, d, di, ML, L, Vx, &S, &per, &KW, &KSL, &MLP);
}

extern "C" int64_t txl_decode_persist_ws_bytes(int B, int H, int dh, int d, int di, int ML, int L, int Vx) {
  int S, per, KW, KSL, MLP;
  if (!dp_geometry(B, H, dh, d, di, ML, L, Vx, &S, &per, &KW, &KSL, &MLP)) return 0;
  int64_t n = sizeof(DpHostTable) + 256;
  n += (int64_t)B * H * d * 2 + 256;                 // qt
  n += (int64_t)B * H * MLP * 4 + 256;               // bd
  n += (int64_t)B * S * H * d * 2 + 256;             // pctx
  n += (int64_t)B * S * H * 2 * 4 + 256;             // pml
  n += (int64_t)8 * B * d * 4 + 256;                 // planes
  n += (int64_t)B * d * 2 + 256;                     // y1
  n += (int64_t)B * di * 2 + 256;                    // h1
  n += 256;                                          // barrier counter
  return n;
}

// One decode step of all layers + the LM-head GEMM.  build_table = 1: upload the per-layer pointer table into `ws`, zero the barrier counter
// (synchronises the stream; no launch) - call once per generation before the first step.
extern "C" int txl_decode_persist_step(const void* const* wqkv, const void* const* wkT, const void* const* wo, const void* const* w1,
                                       const void* const* w2, const void* const* rtab, const float* const* b1, const float* const* b2,
                                       const float* const* rwb, const float* const* rrb, const float* const* ln1w, const float* const* ln1b,
                                       const float* const* ln2w, const float* const* ln2b, void* const* ring, const void* E, const float* out_bias,
                                       void* x, const int32_t* pos, float* logits, int64_t ldl, void* ws, int build_table, int B, int H, int dh, int d,
                                       int di, int ML, int L, int Vx, float eps, void* stream) {
  int S, per, KW, KSL, MLP;
  TXL_CHECK_ARG(dp_geometry(B, H, dh, d, di, ML, L, Vx, &S, &per, &KW, &KSL, &MLP),
                "decode_persist: unsupported geometry (needs B<=64, d_head 64, d_model 128 or 512, H<=8, d_inner <= 512 or a multiple of 512)");
  TXL_CHECK_ARG(ws && ((uintptr_t)ws & 255) == 0, "decode_persist: workspace must be 256-byte aligned");
  cudaStream_t st = (cudaStream_t)stream;
  unsigned char* w = (unsigned char*)ws;
  auto take = [&](int64_t bytes) { unsigned char* q = w; w += (bytes + 255) / 256 * 256; return q; };
  DpHostTable* tab = (DpHostTable*)take(sizeof(DpHostTable));
  DpArgs a;
  a.layers = tab->layers;
  a.tmaps = tab->tm;
  a.qt = (bf16*)take((int64_t)B * H * d * 2);
  a.bd = (float*)take((int64_t)B * H * MLP * 4);
  a.pctx = (bf16*)take((int64_t)B * S * H * d * 2);
  a.pml = (float*)take((int64_t)B * S * H * 2 * 4);
  a.planes = (float*)take((int64_t)8 * B * d * 4);
  a.y1 = (bf16*)take((int64_t)B * d * 2);
  a.h1 = (bf16*)take((int64_t)B * di * 2);
  a.bar = (unsigned long long*)take(8);
  if (build_table) {
    TXL_CHECK_ARG(wqkv && wkT && wo && w1 && w2 && rtab && b1 && b2 && rwb && rrb && ln1w && ln1b && ln2w && ln2b && ring, "decode_persist: null table");
    static DpHostTable h;          // (16 KB: not on the stack)
    memset(&h, 0, sizeof(h));
    for (int l = 0; l < L; ++l) {
      DpLayer& y = h.layers[l];
      y.wq = (const bf16*)wqkv[l]; y.wv = (const bf16*)wqkv[l] + (int64_t)2 * d * d; y.wkT = (const bf16*)wkT[l]; y.wo = (const bf16*)wo[l];
      y.w1 = (const bf16*)w1[l]; y.w2 = (const bf16*)w2[l]; y.r = (const bf16*)rtab[l];
      y.b1 = b1[l]; y.b2 = b2[l]; y.rwb = rwb[l]; y.rrb = rrb[l]; y.ln1w = ln1w[l]; y.ln1b = ln1b[l]; y.ln2w = ln2w[l]; y.ln2b = ln2b[l];
      y.ring = (bf16*)ring[l];
      int rc = txl_make_tmap_2d(&h.tm[l], y.ring, (uint64_t)B * ML, (uint64_t)d, (uint64_t)d, 64, 64);
      if (rc) return rc;
      TXL_CHECK_ARG(y.wq && y.wkT && y.wo && y.w1 && y.w2 && y.r && y.ring && ((uintptr_t)y.ring & 15) == 0 && ((uintptr_t)y.wkT & 15) == 0 && ((uintptr_t)y.wq & 15) == 0,
                    "decode_persist: layer %d pointers (16-byte alignment)", l);
    }
    TXL_CUDA(cudaMemcpyAsync(tab, &h, sizeof(h), cudaMemcpyHostToDevice, st));
    TXL_CUDA(cudaMemsetAsync(a.bar, 0, 8, st));
    TXL_CUDA(cudaStreamSynchronize(st));
    return TXL_OK;
  }
  TXL_CHECK_ARG(E && out_bias && x && pos && logits && ldl >= Vx && ((uintptr_t)x & 15) == 0 && ((uintptr_t)E & 15) == 0, "decode_persist: bad step args");
  a.tstamp = g_dp_tstamp;
  a.E = (const bf16*)E; a.out_bias = out_bias; a.x = (bf16*)x; a.logits = logits; a.pos = pos; a.ldl = ldl;
  a.B = B; a.H = H; a.d = d; a.di = di; a.ML = ML; a.MLP = MLP; a.L = L; a.Vx = Vx; a.S = S; a.per = per; a.KW = KW; a.KSL = KSL;
  a.eps = eps; a.scale_log2 = 1.4426950408889634f / sqrtf((float)dh);
  const int MT = B <= 16 ? 1 : 4;
  const size_t smem = dp_smem_bytes(MT, d, KW);
  void* kargs[] = {(void*)&a};
  const dim3 grid((unsigned)txl_num_sms()), block(DP_THREADS);
#define DP_LAUNCH(MTV, DV)                                                                                                                  \
  {                                                                                                                                         \
    static size_t attr[64] = {0};                                                                                                           \
    int dev = 0;                                                                                                                            \
    TXL_CUDA(cudaGetDevice(&dev));                                                                                                          \
    if (smem > attr[dev & 63]) { TXL_CUDA(cudaFuncSetAttribute(decode_persist_kernel<MTV, DV>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); attr[dev & 63] = smem; } \
    TXL_CUDA(cudaLaunchCooperativeKernel((const void*)decode_persist_kernel<MTV, DV>, grid, block, kargs, smem, st));                      \
  }
  if (d == 512) { if (MT == 1) DP_LAUNCH(1, 512) else DP_LAUNCH(4, 512) }
  else { if (MT == 1) DP_LAUNCH(1, 128) else DP_LAUNCH(4, 128) }
#undef DP_LAUNCH
  TXL_LAUNCH_CHECK();
  return TXL_OK;
}
