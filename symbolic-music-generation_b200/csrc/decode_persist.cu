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
// any CTA are visible after it to L1-bypassing loads and (after the proxy fence) to bulk copies.
__device__ __forceinline__ void grid_barrier(unsigned long long* bar, unsigned long long k) {
  fence_proxy_async_all();
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
    __threadfence();
  }
  __syncthreads();
  fence_proxy_async_all();
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
template <int MT, int KIND>
__device__ void chain_item(const DpArgs& a, const DpLayer& ly, const Smem& sm, int h, int c) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, g = lane >> 2, t = lane & 3;
  const int d = a.d, B = a.B, H = a.H;
  constexpr int P2 = DP_DH * 2 + 64;              // pitch of the K = 64 operands
  constexpr int N2 = KIND == 1 ? 128 : 64, NTW2 = N2 / 64;
  if (tid < DP_CWARPS * 32) {
    // ---- operand loads (all cp.async, one group)
    if (KIND == 2) {
      // A1 = merge of the ring splits of head h:  ctx = sum_z coef_z pctx[b, z, h, :],  coef_z = exp2(m_z - M) l_z / sum(...)
      for (int b = warp; b < MT * 16; b += DP_CWARPS) {
        const int bb = min(b, B - 1);
        float coef[DP_MAXS];
        float M = -INFINITY;
        for (int z = 0; z < a.S; ++z) M = fmaxf(M, __ldcg(a.pml + (((int64_t)bb * a.S + z) * H + h) * 2));
        float Lsum = 0.f;
        for (int z = 0; z < a.S; ++z) {
          const float mz = __ldcg(a.pml + (((int64_t)bb * a.S + z) * H + h) * 2), lz = __ldcg(a.pml + (((int64_t)bb * a.S + z) * H + h) * 2 + 1);
          coef[z] = (mz == -INFINITY) ? 0.f : exp2f(mz - M) * lz;
          Lsum += coef[z];
        }
        const float inv = Lsum > 0.f ? 1.f / Lsum : 0.f;
        for (int c0 = lane * 8; c0 < d; c0 += 256) {
          float acc8[8];
#pragma unroll
          for (int k = 0; k < 8; ++k) acc8[k] = 0.f;
          for (int z = 0; z < a.S; ++z) {
            const uint4 u = __ldcg(reinterpret_cast<const uint4*>(a.pctx + (((int64_t)bb * a.S + z) * H + h) * d + c0));
            const bf16* e = reinterpret_cast<const bf16*>(&u);
            const float cz = coef[z] * inv;
#pragma unroll
            for (int k = 0; k < 8; ++k) acc8[k] = fmaf(cz, __bfloat162float(e[k]), acc8[k]);
          }
          uint4 o;
          o.x = pack2(acc8[0], acc8[1]); o.y = pack2(acc8[2], acc8[3]); o.z = pack2(acc8[4], acc8[5]); o.w = pack2(acc8[6], acc8[7]);
          *reinterpret_cast<uint4*>(sm.A + (size_t)b * sm.pitchA + c0 * 2) = o;
        }
      }
    } else {
      load_tile(sm.A, sm.pitchA, a.x, d, MT * 16, B, d, tid, DP_CWARPS * 32);
    }
    const bf16* w1 = (KIND == 2 ? ly.wv : ly.wq) + (int64_t)h * DP_DH * d;
    load_tile(sm.W, sm.pitchA, w1, d, DP_DH, DP_DH, d, tid, DP_CWARPS * 32);
    if (KIND == 0) load_tile(sm.W2, P2, ly.wkT + ((int64_t)h * d + c * 64) * DP_DH, DP_DH, 64, 64, DP_DH, tid, DP_CWARPS * 32);
    if (KIND == 1) load_tile(sm.W2, P2, ly.r + (int64_t)c * 128 * d + h * DP_DH, d, 128, a.ML + 1 - c * 128, DP_DH, tid, DP_CWARPS * 32);
    if (KIND == 2) load_tile(sm.W2, P2, ly.wo + (int64_t)c * 64 * d + h * DP_DH, d, 64, 64, DP_DH, tid, DP_CWARPS * 32);
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
    float b0 = 0.f, b1 = 0.f;
    if (KIND != 2) {
      const float* bias = KIND == 0 ? ly.rwb : ly.rrb;
      b0 = bias[h * DP_DH + n]; b1 = bias[h * DP_DH + n + 1];
    }
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
template <int MT, int MODE>
__device__ void lin_item(const DpArgs& a, const Smem& sm, const bf16* A, int64_t lda, const bf16* W, int64_t ldw, const float* bias, int N, int K, int n0,
                         int k0, void* out, int64_t ldo) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, g = lane >> 2, t = lane & 3;
  const int B = a.B;
  const int pitch = K * 2 + 64;
  if (tid < DP_CWARPS * 32) {
    load_tile(sm.A, pitch, A + k0, lda, MT * 16, B, K, tid, DP_CWARPS * 32);
    load_tile(sm.W, pitch, W + (int64_t)n0 * ldw + k0, ldw, 16, N - n0, K, tid, DP_CWARPS * 32);
    cpa_commit();
    cpa_wait_all();
  }
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
        v = fmaxf(v + bias[n], 0.f);
        reinterpret_cast<bf16*>(out)[(int64_t)m * ldo + n] = __float2bfloat16_rn(v);
      } else if (MODE == 1) {
        reinterpret_cast<float*>(out)[(int64_t)m * ldo + n] = v;
      } else {
        reinterpret_cast<float*>(out)[(int64_t)m * ldo + n] = v + bias[n];
      }
    }
  }
  __syncthreads();
}

// ------------------------------------------------------------------------------------------------------------------------------ LayerNorm rows
// y[b] = LayerNorm(res[b] + sum_p planes[p][b] (+ bias)) * gamma + beta ; one warp per row, fp32 statistics (two-pass variance).  Optionally
// also stored to `ring_row(b)` (the next layer's cache slot of this step).
__device__ void ln_rows(const DpArgs& a, const bf16* res, int nplanes, const float* bias, const float* gamma, const float* beta, bf16* y, bf16* ring,
                        int cur) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int d = a.d, B = a.B;
  for (int b = blockIdx.x * DP_CWARPS + warp; b < B; b += gridDim.x * DP_CWARPS) {
    float v[4][8];      // d <= 1024: lane owns columns 8 (32 e + lane) ..
    float s = 0.f;
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int c = (e * 32 + lane) * 8;
      if (c < d) {
        const uint4 u = __ldcg(reinterpret_cast<const uint4*>(res + (int64_t)b * d + c));
        const bf16* eb = reinterpret_cast<const bf16*>(&u);
#pragma unroll
        for (int k = 0; k < 8; ++k) v[e][k] = __bfloat162float(eb[k]);
        for (int p = 0; p < nplanes; ++p) {
          const float4 p0 = __ldcg(reinterpret_cast<const float4*>(a.planes + ((int64_t)p * B + b) * d + c));
          const float4 p1 = __ldcg(reinterpret_cast<const float4*>(a.planes + ((int64_t)p * B + b) * d + c + 4));
          v[e][0] += p0.x; v[e][1] += p0.y; v[e][2] += p0.z; v[e][3] += p0.w; v[e][4] += p1.x; v[e][5] += p1.y; v[e][6] += p1.z; v[e][7] += p1.w;
        }
        if (bias) {
#pragma unroll
          for (int k = 0; k < 8; ++k) v[e][k] += bias[c + k];
        }
#pragma unroll
        for (int k = 0; k < 8; ++k) s += v[e][k];
      }
    }
    const float mu = warp_sum(s) / d;
    float q = 0.f;
#pragma unroll
    for (int e = 0; e < 4; ++e)
      if ((e * 32 + lane) * 8 < d) {
#pragma unroll
        for (int k = 0; k < 8; ++k) { const float tt = v[e][k] - mu; q += tt * tt; }
      }
    const float rs = rsqrtf(warp_sum(q) / d + a.eps);
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int c = (e * 32 + lane) * 8;
      if (c < d) {
        uint4 o;
        o.x = pack2((v[e][0] - mu) * rs * gamma[c] + beta[c], (v[e][1] - mu) * rs * gamma[c + 1] + beta[c + 1]);
        o.y = pack2((v[e][2] - mu) * rs * gamma[c + 2] + beta[c + 2], (v[e][3] - mu) * rs * gamma[c + 3] + beta[c + 3]);
        o.z = pack2((v[e][4] - mu) * rs * gamma[c + 4] + beta[c + 4], (v[e][5] - mu) * rs * gamma[c + 5] + beta[c + 5]);
        o.w = pack2((v[e][6] - mu) * rs * gamma[c + 6] + beta[c + 6], (v[e][7] - mu) * rs * gamma[c + 7] + beta[c + 7]);
        *reinterpret_cast<uint4*>(y + (int64_t)b * d + c) = o;
        if (ring) *reinterpret_cast<uint4*>(ring + ((int64_t)b * a.ML + cur) * d + c) = o;
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------------------------------------ attention
// D = d_model (compile time: the qt fragments of a warp live in registers).  item = (sequence b, split z): keys [z per, min((z+1) per, ML)).
template <int D>
__device__ void att_item(const DpArgs& a, const DpLayer& ly, unsigned char* stage0, float* ssm, unsigned char* qsm, uint64_t* full, uint32_t& uses, int b,
                         int z, int cur) {
  constexpr int PITCH = D * 2 + 16;                 // 16 mod 128: the 8 rows of an ldmatrix tile fall into distinct 16-byte bank groups
  constexpr int NKS = D / 16;                       // k16 steps of a score row
  constexpr int CW = D / DP_CWARPS;                 // context columns owned by a warp
  constexpr int NTC = CW / 8;                       // n-tiles of those
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, g = lane >> 2, t = lane & 3;
  const int H = a.H, ML = a.ML;
  const int s_begin = z * a.per, s_end = min(s_begin + a.per, ML);
  const int nst = (s_end - s_begin + DP_KS - 1) / DP_KS;
  const bf16* ring = ly.ring + (int64_t)b * ML * D;
  // warp 0 streams the ring: lane j copies rows j and j + 32 of a stage.  A stage buffer is refilled right after the consumer barrier that
  // follows its last reader (see the loop), so no empty barriers are needed.
  auto issue = [&](int i) {
    if (warp == 0 && i < nst) {
      const uint32_t u = uses + i, st = u % DP_NST;
      const int s0 = s_begin + i * DP_KS, n = min(DP_KS, s_end - s0);
      if (lane == 0) mbar_expect_tx(&full[st], (uint32_t)n * D * 2);
      __syncwarp();
      unsigned char* dst = stage0 + (size_t)st * DP_KS * PITCH;
      for (int j = lane; j < n; j += 32) bulk_row(dst + (size_t)j * PITCH, ring + (int64_t)(s0 + j) * D, D * 2, &full[st]);
    }
  };
#pragma unroll
  for (int i = 0; i < DP_NST; ++i) issue(i);
  // A operand of the scores: qt[b] (rows = heads, zero rows past H) staged in shared memory; a thread's fragment words are read per k16 step
  // (64 fragment registers would not fit beside the context accumulators)
  constexpr int QPITCH = D * 2 + 16;
  for (int e = tid; e < 8 * (D / 8); e += DP_THREADS) {
    const int hh = e / (D / 8), v = e % (D / 8);
    uint4 u = make_uint4(0, 0, 0, 0);
    if (hh < H) u = __ldcg(reinterpret_cast<const uint4*>(a.qt + ((int64_t)b * H + hh) * D + v * 8));
    *reinterpret_cast<uint4*>(qsm + (size_t)hh * QPITCH + v * 16) = u;
  }
  __syncthreads();
  const unsigned char* qrow = qsm + (size_t)g * QPITCH + t * 4;
  float acc[NTC][4];
#pragma unroll
  for (int j = 0; j < NTC; ++j)
#pragma unroll
    for (int q = 0; q < 4; ++q) acc[j][q] = 0.f;
  float m_run = -INFINITY, l_run = 0.f;
  const float* bdrow = a.bd + ((int64_t)b * H + min(g, H - 1)) * a.MLP;
  for (int i = 0; i < nst; ++i) {
    const uint32_t u = uses + i, st = u % DP_NST;
    const int s0 = s_begin + i * DP_KS;
    float* S = ssm + (size_t)(i & 1) * 8 * DP_SPITCH;
    // position term of this thread's two keys (stage-local 8 warp + 2t, +1): r row x = ML - ((cur - s) mod ML)
    const int sa = s0 + warp * 8 + 2 * t, sb = sa + 1;
    float bda = 0.f, bdb = 0.f;
    if (g < H) {
      if (sa < s_end) bda = __ldcg(bdrow + (sa <= cur ? ML - cur + sa : sa - cur));
      if (sb < s_end) bdb = __ldcg(bdrow + (sb <= cur ? ML - cur + sb : sb - cur));
    }
    mbar_wait(&full[st], (u / DP_NST) & 1);
    const uint32_t sbase = smem_u32(stage0 + (size_t)st * DP_KS * PITCH);
    {
      // a short last stage: rows past its keys still feed the P.hid MMAs (with p = 0), so they must be finite - zero them.  Every warp has
      // finished the previous use of this buffer (the producer refilled it only after all eight arrived), the copies cover rows < n only.
      const int n = min(DP_KS, s_end - s0);
      if (n < DP_KS) {
        unsigned char* rows = stage0 + (size_t)st * DP_KS * PITCH + (size_t)n * PITCH;
        for (int e = tid; e < (DP_KS - n) * (PITCH / 16); e += DP_THREADS) reinterpret_cast<uint4*>(rows)[e] = make_uint4(0, 0, 0, 0);
      }
    }
    {
      // phase A: scores of keys 8 warp .. 8 warp + 7 over the whole d
      float c[4] = {0.f, 0.f, 0.f, 0.f};
      const uint32_t rowaddr = sbase + (uint32_t)(warp * 8 + (lane & 7)) * PITCH + ((lane >> 3) & 1) * 16;
#pragma unroll 8
      for (int ks = 0; ks < NKS; ++ks) {
        uint32_t b0, b1;
        ldsm_x2(b0, b1, rowaddr + ks * 32);
        const uint32_t qa0 = *reinterpret_cast<const uint32_t*>(qrow + ks * 32), qa2 = *reinterpret_cast<const uint32_t*>(qrow + ks * 32 + 16);
        mma16816(c, qa0, 0u, qa2, 0u, b0, b1);
      }
      const float v0 = sa < s_end ? (c[0] + bda) * a.scale_log2 : -INFINITY;
      const float v1 = sb < s_end ? (c[1] + bdb) * a.scale_log2 : -INFINITY;
      *reinterpret_cast<float2*>(S + g * DP_SPITCH + warp * 8 + 2 * t) = make_float2(v0, v1);
    }
    cbar();
    // every warp is past phase A of this stage, hence done with the previous stage's buffer: refill it with stage i - 1 + NST
    if (i >= 1) issue(i - 1 + DP_NST);
    {
      // phase B: every warp needs head g's probabilities of all 64 keys as A fragments (k16 step ks: keys 16 ks + 2t, +1, +8, +9)
      float sc[4][4];
      float mx = -INFINITY;
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) {
        const float2 lo = *reinterpret_cast<const float2*>(S + g * DP_SPITCH + ks * 16 + 2 * t);
        const float2 hi = *reinterpret_cast<const float2*>(S + g * DP_SPITCH + ks * 16 + 8 + 2 * t);
        sc[ks][0] = lo.x; sc[ks][1] = lo.y; sc[ks][2] = hi.x; sc[ks][3] = hi.y;
        mx = fmaxf(fmaxf(mx, fmaxf(lo.x, lo.y)), fmaxf(hi.x, hi.y));
      }
      mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 1));
      mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 2));
      const float m_new = fmaxf(m_run, mx);
      const float m_safe = (m_new == -INFINITY) ? 0.f : m_new;
      const float corr = exp2f(m_run - m_safe);         // m_run = -inf -> 0
      uint32_t pa0[4], pa2[4];
      float ls = 0.f;
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) {
        const float p0 = exp2f(sc[ks][0] - m_safe), p1 = exp2f(sc[ks][1] - m_safe), p2 = exp2f(sc[ks][2] - m_safe), p3 = exp2f(sc[ks][3] - m_safe);
        ls += (p0 + p1) + (p2 + p3);
        pa0[ks] = pack2(p0, p1); pa2[ks] = pack2(p2, p3);
      }
      ls += __shfl_xor_sync(0xffffffffu, ls, 1);
      ls += __shfl_xor_sync(0xffffffffu, ls, 2);
      l_run = l_run * corr + ls;
      m_run = m_new;
#pragma unroll
      for (int j = 0; j < NTC; ++j) { acc[j][0] *= corr; acc[j][1] *= corr; }
      // context += P . hid over this warp's CW columns: B fragments through ldmatrix.trans (rows = keys, 16-byte pieces = 8 columns)
      const uint32_t rowaddr = sbase + (uint32_t)((lane & 7) + ((lane >> 3) & 1) * 8) * PITCH + (warp * CW + (lane >> 4) * 8) * 2;
#pragma unroll
      for (int ks = 0; ks < 4; ++ks)
#pragma unroll
        for (int j = 0; j < NTC; j += 2) {
          uint32_t b0, b1, b2, b3;
          ldsm_x4_t(b0, b1, b2, b3, rowaddr + ks * 16 * PITCH + j * 16);
          mma16816(acc[j], pa0[ks], 0u, pa2[ks], 0u, b0, b1);
          mma16816(acc[j + 1], pa0[ks], 0u, pa2[ks], 0u, b2, b3);
        }
    }
  }
  uses += nst;
  // ---- normalised partial context of this split + its (max, sum)
  if (g < H) {
    const float inv = l_run > 0.f ? 1.f / l_run : 0.f;
    bf16* dst = a.pctx + (((int64_t)b * a.S + z) * H + g) * D + warp * CW + 2 * t;
#pragma unroll
    for (int j = 0; j < NTC; ++j) *reinterpret_cast<uint32_t*>(dst + j * 8) = pack2(acc[j][0] * inv, acc[j][1] * inv);
    if (warp == 0 && t == 0) {
      float* ml = a.pml + (((int64_t)b * a.S + z) * H + g) * 2;
      ml[0] = m_run; ml[1] = l_run;
    }
  }
  cbar();      // the score tiles are reused by this CTA's next item
}

// ------------------------------------------------------------------------------------------------------------------------------ the step
template <int MT, int D>
__global__ void __launch_bounds__(DP_THREADS, 1) decode_persist_kernel(const DpArgs a) {
  extern __shared__ __align__(128) unsigned char dp_smem[];
  __shared__ uint64_t full[DP_NST];
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
  // attention view of the same memory: ring stages, then the score tiles
  unsigned char* stage0 = dp_smem;
  float* ssm = reinterpret_cast<float*>(dp_smem + (size_t)DP_NST * DP_KS * (D * 2 + 16));
  unsigned char* qsm = reinterpret_cast<unsigned char*>(ssm + 2 * 8 * DP_SPITCH);
  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < DP_NST; ++s) mbar_init(&full[s], 1);
    fence_barrier_init();
  }
  __syncthreads();
  uint32_t uses = 0;                                  // ring-stage uses of this CTA so far (same count in the producer and the consumers)
  const int p = *a.pos;
  const int cur = p % a.ML;
  // barrier numbering continues over the steps of one generation (the counter is never reset): 7 per layer
  unsigned long long kbar = (unsigned long long)p * (7ull * a.L);
  const int nQT = H * (d / 64), nBD = H * ((a.ML + 1 + 127) / 128);
  for (int l = 0; l < a.L; ++l) {
    const DpLayer ly = a.layers[l];
    // ---- QTBD (layer 0 also stores the embedding row into its ring slot; later layers got theirs from the previous LN2)
    if (l == 0)
      for (int e = cta * DP_THREADS + tid; e < B * (d / 8); e += G * DP_THREADS) {
        const int b = e / (d / 8), c = (e % (d / 8)) * 8;
        *reinterpret_cast<uint4*>(ly.ring + ((int64_t)b * a.ML + cur) * d + c) = __ldcg(reinterpret_cast<const uint4*>(a.x + (int64_t)b * d + c));
      }
    for (int item = cta; item < nQT + nBD; item += G) {
      if (item < nQT) chain_item<MT, 0>(a, ly, sm, item / (d / 64), item % (d / 64));
      else { const int q = item - nQT, nd = nBD / H; chain_item<MT, 1>(a, ly, sm, q / nd, q % nd); }
    }
    grid_barrier(a.bar, kbar++);
    // ---- ATT
    for (int item = cta; item < B * a.S; item += G) att_item<D>(a, ly, stage0, ssm, qsm, full, uses, item / a.S, item % a.S, cur);
    // every ring copy this CTA issued has landed and been consumed (each stage's full barrier was waited on) before the memory is reused
    grid_barrier(a.bar, kbar++);
    // ---- VAON
    for (int item = cta; item < nQT; item += G) chain_item<MT, 2>(a, ly, sm, item / (d / 64), item % (d / 64));
    grid_barrier(a.bar, kbar++);
    // ---- LN1
    ln_rows(a, a.x, H, nullptr, ly.ln1w, ly.ln1b, a.y1, nullptr, 0);
    grid_barrier(a.bar, kbar++);
    // ---- FF1
    for (int item = cta; item < (a.di + 15) / 16; item += G) lin_item<MT, 0>(a, sm, a.y1, d, ly.w1, d, ly.b1, a.di, d, item * 16, 0, a.h1, a.di);
    grid_barrier(a.bar, kbar++);
    // ---- FF2
    {
      const int nb = (d + 15) / 16;
      for (int item = cta; item < nb * a.KSL; item += G) {
        const int ks = item / nb, n0 = (item % nb) * 16;
        lin_item<MT, 1>(a, sm, a.h1, a.di, ly.w2, a.di, nullptr, d, a.KW, n0, ks * a.KW, a.planes + (int64_t)ks * B * d, d);
      }
    }
    grid_barrier(a.bar, kbar++);
    // ---- LN2 (+ next layer's ring slot)
    ln_rows(a, a.y1, a.KSL, ly.b2, ly.ln2w, ly.ln2b, a.x, l + 1 < a.L ? a.layers[l + 1].ring : nullptr, cur);
    grid_barrier(a.bar, kbar++);
  }
  // ---- LM head
  for (int item = cta; item < (a.Vx + 15) / 16; item += G) lin_item<MT, 2>(a, sm, a.x, d, a.E, d, a.out_bias, a.Vx, d, item * 16, 0, a.logits, a.ldl);
  (void)warp;
}

struct DpHostTable {          // what txl_decode_persist_step(build = 1) leaves at the start of the workspace
  DpLayer layers[64];
};

size_t dp_smem_bytes(int MT, int d, int KW) {
  const size_t pitchA = (size_t)(d > KW ? d : KW) * 2 + 64;
  const size_t gemm = (size_t)MT * 16 * pitchA + 64 * pitchA + (size_t)MT * 16 * (DP_DH * 2 + 64) + 128 * (DP_DH * 2 + 64) + (size_t)DP_CWARPS * MT * 16 * 17 * 4;
  const size_t att = (size_t)DP_NST * DP_KS * ((size_t)d * 2 + 16) + 2 * 8 * DP_SPITCH * 4 + 8 * ((size_t)d * 2 + 16);
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
    DpHostTable h;
    memset(&h, 0, sizeof(h));
    for (int l = 0; l < L; ++l) {
      DpLayer& y = h.layers[l];
      y.wq = (const bf16*)wqkv[l]; y.wv = (const bf16*)wqkv[l] + (int64_t)2 * d * d; y.wkT = (const bf16*)wkT[l]; y.wo = (const bf16*)wo[l];
      y.w1 = (const bf16*)w1[l]; y.w2 = (const bf16*)w2[l]; y.r = (const bf16*)rtab[l];
      y.b1 = b1[l]; y.b2 = b2[l]; y.rwb = rwb[l]; y.rrb = rrb[l]; y.ln1w = ln1w[l]; y.ln1b = ln1b[l]; y.ln2w = ln2w[l]; y.ln2b = ln2b[l];
      y.ring = (bf16*)ring[l];
      TXL_CHECK_ARG(y.wq && y.wkT && y.wo && y.w1 && y.w2 && y.r && y.ring && ((uintptr_t)y.ring & 15) == 0 && ((uintptr_t)y.wkT & 15) == 0 && ((uintptr_t)y.wq & 15) == 0,
                    "decode_persist: layer %d pointers (16-byte alignment)", l);
    }
    TXL_CUDA(cudaMemcpyAsync(tab, &h, sizeof(h), cudaMemcpyHostToDevice, st));
    TXL_CUDA(cudaMemsetAsync(a.bar, 0, 8, st));
    TXL_CUDA(cudaStreamSynchronize(st));
    return TXL_OK;
  }
  TXL_CHECK_ARG(E && out_bias && x && pos && logits && ldl >= Vx && ((uintptr_t)x & 15) == 0 && ((uintptr_t)E & 15) == 0, "decode_persist: bad step args");
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
