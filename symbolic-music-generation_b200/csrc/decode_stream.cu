// decode_stream.cu — the kernels of the bf16 decode step, second generation (the first generation lives on in decode.cu for the fp32 parity mode).
//
// (1) dec_linear_kernel: y[M<=64, N] = x[M,K] W[N,K]^T (+bias)(ReLU) for the per-step Linears.  The step's Linears are weight-streaming
//     problems of 0.5-2 MB with <= 64 rows: a 128-row tcgen05 tile leaves 4-16 CTAs to pull the whole matrix (measured 12.3 us per launch).
//     Here a CTA owns 8 or 16 output features (96-150 CTAs), x and its weight rows arrive through a cp.async ring, its 8 warps split every
//     256-column chunk in 32-wide blocks and read 16-byte pieces that already ARE mma.sync fragments (the k index inside an MMA is only a
//     label: slot pair (2t,2t+1) / (2t+8,2t+9) of step j is fed from elements 8t+4j .. 8t+4j+3 of the 32-block on both operands), partial sums
//     meet in shared memory.  gridDim.y > 1 = split-K: fp32 planes, summed by (2).                          [A.3, A.6 at T=1; A.7]
// (2) dec_add_ln_kernel: LayerNorm(x + sum of the split-K planes + bias), the residual blocks after o_net and CoreNet.3.        [A.3-8, A.6]
// (3) decode_attn_pipe_kernel: the T=1 relative-position attention over the projected k|v ring.  The first-generation kernel kept only
//     three 16-byte loads per thread in flight (ptxas does not hoist loads over the shuffle / branch of the previous key) and reached
//     2.8 TB/s.  Here a producer lane streams the k|v rows (interleaved per key: one contiguous run per stage) and the r rows of a stage with
//     cp.async.bulk into a shared-memory ring (full/empty mbarriers; default 128 keys x 2 stages, 2 CTAs per SM = 192 KB per SM in flight),
//     8 consumer warps run the one-pass (max, sum, output) update.  The r table is stored head-major [H, mem_len+1, d_head] so that a stage's
//     rows are one (two at the ring's wrap point) contiguous run.  The current token never goes through the ring copy: its k/v come from
//     the qkv row, so no generic->async proxy ordering is needed and the ring streams before the previous kernel has finished.
// All three run under programmatic dependent launch (common.cuh): whatever does not depend on the previous kernel of the step (weights, ring)
// is requested before griddepcontrol.wait.
#include "tc_common.cuh"
#include <stdlib.h>

namespace {

__device__ __forceinline__ void mma_bf16_16816(float* c, uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t word(const uint4& v, int i) { return i == 0 ? v.x : i == 1 ? v.y : i == 2 ? v.z : v.w; }
__device__ __forceinline__ void cp_async16(void* dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

__device__ __forceinline__ void lds8(const bf16* p, float* f) {   // 8 bf16 (one 16-byte load, shared or global) -> fp32
  const uint4 u = *reinterpret_cast<const uint4*>(p);
  const bf16* e = reinterpret_cast<const bf16*>(&u);
#pragma unroll
  for (int k = 0; k < 8; ++k) f[k] = __bfloat162float(e[k]);
}

// L2 prefetch of a slice of the NEXT attention kernel's ring, issued by the latency-bound kernels between two attention kernels (HBM is idle
// while they run): CTA i of the grid asks for bytes [i chunk, (i+1) chunk) of [base, base + bytes) with cp.async.bulk.prefetch.L2 — fire and
// forget, no registers, no completion to wait for.  The ring was written by earlier steps (and L2 is the coherence point), so no ordering
// against the step's other kernels is needed.
__device__ __forceinline__ void l2_prefetch_slice(const void* base, int64_t bytes) {
  if (!base || bytes <= 0) return;
  const int64_t nblk = (int64_t)gridDim.x * gridDim.y, blk = (int64_t)blockIdx.y * gridDim.x + blockIdx.x;
  const int64_t chunk = ((bytes + nblk - 1) / nblk + 127) & ~(int64_t)127;
  int64_t off = blk * chunk;
  const int64_t end = off + chunk < bytes ? off + chunk : bytes;
  const char* p = reinterpret_cast<const char*>(base);
  for (; off < end; off += 32768) {
    const uint32_t n = (uint32_t)((end - off < 32768 ? end - off : 32768) & ~(int64_t)15);
    if (n) asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p + off), "r"(n) : "memory");
  }
}

constexpr int DL_WARPS = 8, DL_THREADS = DL_WARPS * 32;
constexpr int DL_KC = DL_WARPS * 32;                // K columns of one stage: one 32-wide block per warp
constexpr int DL_PITCH = DL_KC * 2 + 64;            // bytes per staged row: +64 shifts consecutive rows by 16 banks -> conflict-free LDS.128 fragments
constexpr size_t dl_smem_bytes(int MT, int NT, int STAGES) {
  size_t ring = (size_t)STAGES * (MT * 16 + NT * 8) * DL_PITCH, red = (size_t)DL_WARPS * MT * 16 * (NT * 8 + 1) * 4;
  return ring > red ? ring : red;
}

// MT = 16-row tiles of x (1, 2 or 4), NT = 8-feature tiles per CTA (1 or 2).  x and the CTA's weight rows stream through a DL_STAGES-deep
// cp.async ring in 256-column chunks (every copy of the first chunks is in flight before the first MMA: ptxas sinks plain global loads between
// the MMAs, and in-order issue then serialises their latencies); warp w multiplies the w-th 32-wide block of each chunk.  Two stages (the
// whole K of a <= 512-column problem, 83-92 KB) leave room for a second CTA or an attention CTA of another sequence group on the SM.
template <int MT, int NT, bool OUT_F32, int DL_STAGES>
__global__ void __launch_bounds__(DL_THREADS) dec_linear_kernel(const bf16* __restrict__ A, int64_t lda, const bf16* __restrict__ W, int64_t ldw,
                                                                const float* __restrict__ bias, void* __restrict__ Cout, int64_t ldc, int M, int N,
                                                                int Ktot, int relu, const void* __restrict__ pf, int64_t pf_bytes) {
  extern __shared__ __align__(128) unsigned char dl_smem[];
  constexpr int ROWS = MT * 16 + NT * 8;             // staged rows: x rows first, then the weight rows
  constexpr int VPR = DL_KC / 8;                     // 16-byte vectors per staged row
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, g = lane >> 2, t = lane & 3;
  const int n0 = blockIdx.x * (NT * 8);
  // split-K: gridDim.y CTAs share an output tile, CTA y multiplies columns [y K, (y+1) K) and writes its fp32 partial sums to plane y of Cout
  // (no bias / ReLU: dec_add_ln_kernel adds the planes, the bias and the residual)
  const int K = Ktot / (int)gridDim.y;
  A += (int64_t)blockIdx.y * K;
  W += (int64_t)blockIdx.y * K;
  const int nch = (K + DL_KC - 1) / DL_KC;
  // rows past M / N re-read the last valid row; their results are never stored
  auto issue_w = [&](int c) {
    if (c < nch) {
      unsigned char* dst = dl_smem + ((size_t)(c % DL_STAGES) * ROWS + MT * 16) * DL_PITCH;
      for (int e = threadIdx.x; e < NT * 8 * VPR; e += DL_THREADS) {
        const int r = e / VPR, v = e % VPR, k = c * DL_KC + v * 8;
        if (k < K) cp_async16(dst + (size_t)r * DL_PITCH + v * 16, W + (int64_t)min(n0 + r, N - 1) * ldw + k);
      }
    }
  };
  auto issue_a = [&](int c) {
    if (c < nch) {
      unsigned char* dst = dl_smem + (size_t)(c % DL_STAGES) * ROWS * DL_PITCH;
      for (int e = threadIdx.x; e < MT * 16 * VPR; e += DL_THREADS) {
        const int r = e / VPR, v = e % VPR, k = c * DL_KC + v * 8;
        if (k < K) cp_async16(dst + (size_t)r * DL_PITCH + v * 16, A + (int64_t)min(r, M - 1) * lda + k);
      }
    }
    cp_async_commit();                               // one group per chunk, empty past the end: keeps wait_group counting uniform
  };
  auto issue = [&](int c) { issue_w(c); issue_a(c); };
  float acc[MT][NT][4];
#pragma unroll
  for (int i = 0; i < MT; ++i)
#pragma unroll
    for (int j = 0; j < NT; ++j)
#pragma unroll
      for (int q = 0; q < 4; ++q) acc[i][j][q] = 0.f;
  // weights do not depend on the previous kernel of the step: their copies start before the programmatic-dependency wait
  pdl_launch_dependents();
  if (threadIdx.x == 0) l2_prefetch_slice(pf, pf_bytes);
#pragma unroll
  for (int c = 0; c < DL_STAGES - 1; ++c) issue_w(c);
  pdl_wait();
#pragma unroll
  for (int c = 0; c < DL_STAGES - 1; ++c) issue_a(c);
  for (int c = 0; c < nch; ++c) {
    issue(c + DL_STAGES - 1);
    cp_async_wait<DL_STAGES - 1>();
    __syncthreads();
    if (c * DL_KC + warp * 32 < K) {
      const unsigned char* base = dl_smem + (size_t)(c % DL_STAGES) * ROWS * DL_PITCH + warp * 64 + t * 16;
      uint4 av[MT][2], wv[NT];
#pragma unroll
      for (int i = 0; i < MT; ++i) {
        av[i][0] = *reinterpret_cast<const uint4*>(base + (size_t)(i * 16 + g) * DL_PITCH);
        av[i][1] = *reinterpret_cast<const uint4*>(base + (size_t)(i * 16 + g + 8) * DL_PITCH);
      }
#pragma unroll
      for (int j = 0; j < NT; ++j) wv[j] = *reinterpret_cast<const uint4*>(base + (size_t)(MT * 16 + j * 8 + g) * DL_PITCH);
      // the k index inside an MMA is only a label: slot pairs (2t,2t+1) / (2t+8,2t+9) of step s are fed from elements 8t+4s .. 8t+4s+3 of
      // the 32-block on BOTH operands, so one 16-byte piece per row is two k16 steps' worth of fragment registers
#pragma unroll
      for (int s = 0; s < 2; ++s)
#pragma unroll
        for (int i = 0; i < MT; ++i)
#pragma unroll
          for (int j = 0; j < NT; ++j)
            mma_bf16_16816(acc[i][j], word(av[i][0], 2 * s), word(av[i][1], 2 * s), word(av[i][0], 2 * s + 1), word(av[i][1], 2 * s + 1),
                           word(wv[j], 2 * s), word(wv[j], 2 * s + 1));
    }
    __syncthreads();
  }
  cp_async_wait<0>();
  // C fragment: c0,c1 -> (row g, cols 2t, 2t+1); c2,c3 -> (row g+8, same cols).  The ring is free now: reuse it for the cross-warp sum.
  float (*red)[MT * 16][NT * 8 + 1] = reinterpret_cast<float (*)[MT * 16][NT * 8 + 1]>(dl_smem);
#pragma unroll
  for (int i = 0; i < MT; ++i)
#pragma unroll
    for (int j = 0; j < NT; ++j) {
      red[warp][i * 16 + g][j * 8 + 2 * t] = acc[i][j][0];
      red[warp][i * 16 + g][j * 8 + 2 * t + 1] = acc[i][j][1];
      red[warp][i * 16 + g + 8][j * 8 + 2 * t] = acc[i][j][2];
      red[warp][i * 16 + g + 8][j * 8 + 2 * t + 1] = acc[i][j][3];
    }
  __syncthreads();
  const bool partial = gridDim.y > 1;
  for (int e = threadIdx.x; e < MT * 16 * NT * 8; e += DL_THREADS) {
    const int m = e / (NT * 8), nl = e % (NT * 8), n = n0 + nl;
    if (m < M && n < N) {
      float x = 0.f;
#pragma unroll
      for (int w = 0; w < DL_WARPS; ++w) x += red[w][m][nl];
      if (partial) { reinterpret_cast<float*>(Cout)[((int64_t)blockIdx.y * M + m) * ldc + n] = x; continue; }
      if (bias) x += bias[n];
      if (relu) x = fmaxf(x, 0.f);
      if constexpr (OUT_F32) reinterpret_cast<float*>(Cout)[(int64_t)m * ldc + n] = x;
      else reinterpret_cast<bf16*>(Cout)[(int64_t)m * ldc + n] = __float2bfloat16_rn(x);
    }
  }
}

// y[m, :] = LayerNorm(x[m, :] + sum_s part[s][m, :] + bias) * gamma + beta     (the residual LayerNorm after o_net / CoreNet.3 at T=1 [A.3-8, A.6])
// part = the fp32 split-K planes of dec_linear_kernel.  One warp per row, 16-byte vectors, fp32 statistics (two-pass variance).
template <int STEPS>
__global__ void __launch_bounds__(64) dec_add_ln_kernel(const bf16* __restrict__ x, const float* __restrict__ part, int nparts, const float* __restrict__ bias,
                                                        const float* __restrict__ gamma, const float* __restrict__ beta, bf16* __restrict__ y, int M, int d,
                                                        float eps, const void* __restrict__ pf, int64_t pf_bytes) {
  pdl_launch_dependents();
  if (threadIdx.x == 0) l2_prefetch_slice(pf, pf_bytes);
  pdl_wait();
  const int lane = threadIdx.x & 31, row = blockIdx.x * 2 + (threadIdx.x >> 5);
  if (row >= M) return;
  float v[STEPS][8];
  float s = 0.f;
#pragma unroll
  for (int e = 0; e < STEPS; ++e) {
    const int c = (e * 32 + lane) * 8;
    if (c < d) {
      lds8(x + (int64_t)row * d + c, v[e]);             // (generic 16-byte load; the helper only converts)
      for (int p = 0; p < nparts; ++p) {
        const float4 a = *reinterpret_cast<const float4*>(part + ((int64_t)p * M + row) * d + c);
        const float4 b = *reinterpret_cast<const float4*>(part + ((int64_t)p * M + row) * d + c + 4);
        v[e][0] += a.x; v[e][1] += a.y; v[e][2] += a.z; v[e][3] += a.w; v[e][4] += b.x; v[e][5] += b.y; v[e][6] += b.z; v[e][7] += b.w;
      }
      if (bias) {
        const float4 a = *reinterpret_cast<const float4*>(bias + c), b = *reinterpret_cast<const float4*>(bias + c + 4);
        v[e][0] += a.x; v[e][1] += a.y; v[e][2] += a.z; v[e][3] += a.w; v[e][4] += b.x; v[e][5] += b.y; v[e][6] += b.z; v[e][7] += b.w;
      }
#pragma unroll
      for (int k = 0; k < 8; ++k) s += v[e][k];
    }
  }
  const float mu = warp_sum(s) / d;
  float q = 0.f;
#pragma unroll
  for (int e = 0; e < STEPS; ++e)
    if ((e * 32 + lane) * 8 < d) {
#pragma unroll
      for (int k = 0; k < 8; ++k) { const float t = v[e][k] - mu; q += t * t; }
    }
  const float rs = rsqrtf(warp_sum(q) / d + eps);
#pragma unroll
  for (int e = 0; e < STEPS; ++e) {
    const int c = (e * 32 + lane) * 8;
    if (c < d) {
      const float4 g0 = *reinterpret_cast<const float4*>(gamma + c), g1 = *reinterpret_cast<const float4*>(gamma + c + 4);
      const float4 b0 = *reinterpret_cast<const float4*>(beta + c), b1 = *reinterpret_cast<const float4*>(beta + c + 4);
      const float gm[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w}, bt[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
      uint4 o;
      bf16* ob = reinterpret_cast<bf16*>(&o);
#pragma unroll
      for (int k = 0; k < 8; ++k) ob[k] = __float2bfloat16_rn((v[e][k] - mu) * rs * gm[k] + bt[k]);
      *reinterpret_cast<uint4*>(y + (int64_t)row * d + c) = o;
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------- attention
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)), "l"(src), "r"(bytes),
               "r"(smem_u32(bar))
               : "memory");
}
constexpr int DA_CWARPS = 8, DA_THREADS = (DA_CWARPS + 1) * 32;
constexpr int DA_TILE = 2048;                       // bf16 elements of one d_head-wide operand for the keys of one pass: (8 warps * keys per warp) * DH
constexpr size_t da_smem_bytes(int passes, int stages) { return (size_t)stages * passes * 3 * DA_TILE * sizeof(bf16); }
constexpr int da_min_ctas(int passes, int stages) {
  return (int)(da_smem_bytes(passes, stages) <= 48 * 1024 ? 4 : da_smem_bytes(passes, stages) <= 72 * 1024 ? 3 : da_smem_bytes(passes, stages) <= 110 * 1024 ? 2 : 1);
}

// Ring cache layout: kv [B, H, mem_len, 2*DH], a key's k row followed by its v row, so a stage of CKS keys is ONE contiguous run of
// CKS * 4 * DH bytes (8-16 KB) instead of two 4 KB runs from two streams: half as many concurrent DRAM streams, twice the burst length.
// A stage = PASSES passes of 32 keys (d_head 64); STAGES stages per CTA.
// gridDim.z = S splits of the ring: split z streams chunks [z per, (z+1) per); with S > 1 every CTA leaves its un-normalised (max, sum,
// output) in `ws` and the LAST CTA of a (b, h) to arrive (atomic counter, reset for the next launch) merges them.  Splits only pay when
// B * H is far below the 592 CTA slots of the chip (8 sequences per GPU): at 64 sequences 1 split 667 us/step, 2 splits 703, 4 splits 720.
template <int DH, int PASSES, int STAGES>
__global__ void __launch_bounds__(DA_THREADS, da_min_ctas(PASSES, STAGES))
decode_attn_pipe_kernel(const bf16* __restrict__ qkv, bf16* __restrict__ kvc, const bf16* __restrict__ rhm, const float* __restrict__ rwb,
                        const float* __restrict__ rrb, bf16* __restrict__ out, const int32_t* __restrict__ pos, int H, int ML, float* __restrict__ ws,
                        int* __restrict__ cnt) {
  constexpr int LPK = DH / 8;                       // lanes per key (16-byte pieces of a d_head row)
  constexpr int KPW = 32 / LPK;                     // keys per warp per pass
  constexpr int CK = DA_CWARPS * KPW;               // keys per pass
  constexpr int CKS = CK * PASSES;                  // keys per stage
  static_assert(CK * DH == DA_TILE, "stage geometry");
  constexpr int STAGE_ELEMS = PASSES * 3 * DA_TILE; // [CKS][k | v] then [CKS] r rows
  extern __shared__ __align__(128) unsigned char da_smem[];
  bf16* ring = reinterpret_cast<bf16*>(da_smem);
  __shared__ uint64_t full[STAGES], empty[STAGES];
  __shared__ float qw[DH], qr[DH], red[DA_CWARPS][DH], wm[DA_CWARPS], wl[DA_CWARPS];
  __shared__ int s_last;
  const int h = blockIdx.x, b = blockIdx.y, z = blockIdx.z, S = gridDim.z, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int HD = H * DH;
  // `pos` is advanced by the step's last kernel only, and every kernel of a step waits for its predecessor: safe to read before pdl_wait
  const int cur = (*pos) % ML;
  bf16* KV = kvc + ((int64_t)b * H + h) * ML * 2 * DH;
  const bf16* R = rhm + (int64_t)h * (ML + 1) * DH;
  const bf16* row = qkv + (int64_t)b * 3 * HD;
  const int nchunks = (ML + CKS - 1) / CKS, per = (nchunks + S - 1) / S;
  const int c_begin = min(z * per, nchunks), c_end = min(c_begin + per, nchunks);
  pdl_launch_dependents();
  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], DA_CWARPS); }
    fence_barrier_init();
  }
  __syncthreads();
  if (warp == DA_CWARPS) {
    // ------------------------------------------------ producer: one lane streams the ring through the stages.  The ring holds what earlier
    // STEPS wrote (slot `cur`, rewritten below, is never taken from the copy), so it does not wait for the previous kernel.
    if (lane == 0) {
      for (int c = c_begin; c < c_end; ++c) {
        const int i = c - c_begin, st = i % STAGES;
        if (i >= STAGES) mbar_wait(&empty[st], ((i / STAGES) & 1) ^ 1);
        const int s0 = c * CKS, n = min(CKS, ML - s0);
        bf16* dst = ring + (size_t)st * STAGE_ELEMS;
        const uint32_t rowb = DH * (uint32_t)sizeof(bf16);
        mbar_expect_tx(&full[st], 3u * n * rowb);
        bulk_g2s(dst, KV + (int64_t)s0 * 2 * DH, 2u * n * rowb, &full[st]);
        // r row of ring slot s: distance (cur - s) mod ML, row ML - distance  =>  s <= cur: ML - cur + s;  s > cur: s - cur
        bf16* rdst = dst + CKS * 2 * DH;
        const int n1 = max(0, min(n, cur + 1 - s0));
        if (n1 > 0) bulk_g2s(rdst, R + (int64_t)(ML - cur + s0) * DH, n1 * rowb, &full[st]);
        if (n1 < n) bulk_g2s(rdst + n1 * DH, R + (int64_t)(s0 + n1 - cur) * DH, (n - n1) * rowb, &full[st]);
      }
    }
    pdl_wait();                                      // a CTA must not retire before its predecessors have (transitive ordering of the chain)
    return;
  }
  // -------------------------------------------------- consumers: lane group `sub` of a warp owns one key of every pass
  pdl_wait();                                        // qkv comes from the previous kernel
  if (tid < DH) {
    const float q = __bfloat162float(row[h * DH + tid]);
    qw[tid] = q + rwb[h * DH + tid];
    qr[tid] = q + rrb[h * DH + tid];
    if (z == 0) {                                    // append for the NEXT steps: this step scores the new token from the qkv row
      KV[(int64_t)cur * 2 * DH + tid] = row[HD + h * DH + tid];
      KV[(int64_t)cur * 2 * DH + DH + tid] = row[2 * HD + h * DH + tid];
    }
  }
  asm volatile("bar.sync 1, %0;" ::"n"(DA_CWARPS * 32) : "memory");   // consumer warps only
  const int sub = lane / LPK, ch = lane % LPK;
  const float scale = rsqrtf((float)DH);
  float qwv[8], qrv[8], acc[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) { qwv[k] = qw[ch * 8 + k]; qrv[k] = qr[ch * 8 + k]; acc[k] = 0.f; }
  float m = -INFINITY, l = 0.f;
  if (z == 0 && warp == 0 && sub == 0) {
    // the current token: distance 0 -> r row ML; k, v straight from the qkv row
    float kf[8], rf[8], a = 0.f;
    lds8(row + HD + h * DH + ch * 8, kf);
    lds8(R + (int64_t)ML * DH + ch * 8, rf);
    lds8(row + 2 * HD + h * DH + ch * 8, acc);
#pragma unroll
    for (int k = 0; k < 8; ++k) a = fmaf(qwv[k], kf[k], fmaf(qrv[k], rf[k], a));
#pragma unroll
    for (int o = LPK / 2; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu >> (32 - LPK), a, o);
    m = a * scale; l = 1.f;
  }
  const int kk0 = warp * KPW + sub;                  // this lane group's key inside a pass
  for (int c = c_begin; c < c_end; ++c) {
    const int i = c - c_begin, st = i % STAGES;
    mbar_wait(&full[st], (i / STAGES) & 1);
    const bf16* stage = ring + (size_t)st * STAGE_ELEMS;
#pragma unroll
    for (int ps = 0; ps < PASSES; ++ps) {
      const int kk = ps * CK + kk0;
      float kf[8], rf[8], vf[8], a = 0.f;
      lds8(stage + kk * 2 * DH + ch * 8, kf);
      lds8(stage + CKS * 2 * DH + kk * DH + ch * 8, rf);
      lds8(stage + kk * 2 * DH + DH + ch * 8, vf);
#pragma unroll
      for (int k = 0; k < 8; ++k) a = fmaf(qwv[k], kf[k], fmaf(qrv[k], rf[k], a));
#pragma unroll
      for (int o = LPK / 2; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
      if (ps == PASSES - 1) {
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty[st]);      // every lane's shared-memory reads of this stage have returned
      }
      const int s = c * CKS + kk;
      const bool ok = s < ML && s != cur;
      a = ok ? a * scale : -INFINITY;
      const float mn = fmaxf(m, a);
      const float cf = (m == -INFINITY) ? 0.f : __expf(m - mn);
      const float p = ok ? __expf(a - mn) : 0.f;
      l = l * cf + p;
#pragma unroll
      for (int k = 0; k < 8; ++k) acc[k] = fmaf(acc[k], cf, ok ? p * vf[k] : 0.f);
      m = mn;
    }
  }
  // merge the KPW key groups of the warp, then the warps through shared memory
#pragma unroll
  for (int o = LPK; o < 32; o <<= 1) {
    const float m2 = __shfl_xor_sync(0xffffffffu, m, o), l2 = __shfl_xor_sync(0xffffffffu, l, o);
    const float mn = fmaxf(m, m2);
    const float c1 = (m == -INFINITY) ? 0.f : __expf(m - mn), c2 = (m2 == -INFINITY) ? 0.f : __expf(m2 - mn);
    l = l * c1 + l2 * c2;
#pragma unroll
    for (int k = 0; k < 8; ++k) acc[k] = acc[k] * c1 + __shfl_xor_sync(0xffffffffu, acc[k], o) * c2;
    m = mn;
  }
  if (sub == 0) {
#pragma unroll
    for (int k = 0; k < 8; ++k) red[warp][ch * 8 + k] = acc[k];
    if (ch == 0) { wm[warp] = m; wl[warp] = l; }
  }
  asm volatile("bar.sync 1, %0;" ::"n"(DA_CWARPS * 32) : "memory");
  float Mx = -INFINITY, L = 0.f, o = 0.f;
  if (tid < DH) {
#pragma unroll
    for (int w = 0; w < DA_CWARPS; ++w) Mx = fmaxf(Mx, wm[w]);
#pragma unroll
    for (int w = 0; w < DA_CWARPS; ++w) {
      const float c = (wm[w] == -INFINITY) ? 0.f : __expf(wm[w] - Mx);
      L += wl[w] * c; o += red[w][tid] * c;
    }
  }
  if (S == 1) {
    if (tid < DH) out[(int64_t)b * HD + h * DH + tid] = __float2bfloat16_rn(o / L);
    return;
  }
  const int bh = b * H + h;
  float* mine = ws + ((int64_t)bh * S + z) * (DH + 2);
  if (tid < DH) {
    mine[tid] = o;
    if (tid == 0) { mine[DH] = Mx; mine[DH + 1] = L; }
    __threadfence();
  }
  asm volatile("bar.sync 1, %0;" ::"n"(DA_CWARPS * 32) : "memory");
  if (tid == 0) s_last = (atomicAdd(&cnt[bh], 1) == S - 1);
  asm volatile("bar.sync 1, %0;" ::"n"(DA_CWARPS * 32) : "memory");
  if (!s_last) return;
  __threadfence();
  if (tid < DH) {
    const float* all = ws + (int64_t)bh * S * (DH + 2);
    float Mt = -INFINITY;
    for (int q = 0; q < S; ++q) Mt = fmaxf(Mt, __ldcg(all + q * (DH + 2) + DH));
    float Lt = 0.f, ot = 0.f;
    for (int q = 0; q < S; ++q) {
      const float mq = __ldcg(all + q * (DH + 2) + DH);
      const float c = (mq == -INFINITY) ? 0.f : __expf(mq - Mt);
      Lt += __ldcg(all + q * (DH + 2) + DH + 1) * c;
      ot += __ldcg(all + q * (DH + 2) + tid) * c;
    }
    out[(int64_t)b * HD + h * DH + tid] = __float2bfloat16_rn(ot / Lt);
    if (tid == 0) cnt[bh] = 0;                       // ready for the next launch (no other CTA of this launch touches it any more)
  }
}

// kv_mem [B*ML, 2*H*DH] (k | v per row, from the GEMM over the hidden-state mems) -> ring cache [B, H, ML, 2*DH] (k row then v row per key)
__global__ void cache_init_kv_kernel(const bf16* __restrict__ kv, int64_t ld, bf16* __restrict__ kvc, int B, int H, int ML, int DH) {
  const int64_t total = (int64_t)B * H * ML * 2 * DH;
  const int HD = H * DH;
  for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(idx % (2 * DH)); int64_t t = idx / (2 * DH); const int j = (int)(t % ML); t /= ML; const int h = (int)(t % H); const int b = (int)(t / H);
    const bf16* src = kv + ((int64_t)b * ML + j) * ld;
    kvc[idx] = c < DH ? src[h * DH + c] : src[HD + h * DH + (c - DH)];
  }
}

// r [ML+1, H*DH] (HF's r_head_k, row = relative-distance index) -> head-major [H, ML+1, DH]
__global__ void rtab_head_major_kernel(const bf16* __restrict__ r, bf16* __restrict__ out, int rows, int H, int DH) {
  const int64_t total = (int64_t)rows * H * DH;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % DH); int64_t q = i / DH; const int x = (int)(q % rows); const int h = (int)(q / rows);
    out[i] = r[((int64_t)x * H + h) * DH + c];
  }
}
}  // namespace

extern "C" int txl_dec_linear(const void* A, int64_t lda, const void* W, int64_t ldw, const float* bias, void* C, int64_t ldc, int M, int N, int K,
                              int relu, int out_f32, int splits, const void* prefetch, int64_t prefetch_bytes, void* stream) {
  TXL_CHECK_ARG(A && W && C && M > 0 && M <= 64 && N > 0 && K > 0 && K % 32 == 0 && lda % 8 == 0 && ldw % 8 == 0, "dec_linear: needs M<=64, K%%32==0, lda,ldw%%8==0");
  TXL_CHECK_ARG(((uintptr_t)A & 15) == 0 && ((uintptr_t)W & 15) == 0, "dec_linear: 16-byte alignment");
  TXL_CHECK_ARG(splits >= 1 && K % (32 * splits) == 0, "dec_linear: K=%d must split into %d multiples of 32", K, splits);
  TXL_CHECK_ARG(splits == 1 || (out_f32 && !bias && !relu), "dec_linear: split-K writes fp32 partial planes (no bias / ReLU)");
  cudaStream_t st = (cudaStream_t)stream;
  const bf16* a = (const bf16*)A; const bf16* w = (const bf16*)W;
  const int mt = M <= 16 ? 1 : M <= 32 ? 2 : 4;
  const int nt = cdiv64(N, 8) * splits <= txl_num_sms() ? 1 : 2;      // 8 features per CTA while that fits one wave, else 16
  dim3 grid((unsigned)cdiv64(N, nt * 8), (unsigned)splits);
  // two ring stages (83-92 KB instead of 166-184 KB) for problems of <= 512 columns per CTA: measured neutral with four sequence groups
  // (546.0 vs 546.2 us/step) and slower with one (662.7 vs 613.2), so it stays an A/B switch
  static const bool allow2 = [] { const char* e = getenv("TXL_DEC_LINEAR_2STAGE"); return e && e[0] == '1'; }();
  const bool two_stage = allow2 && K / splits <= 2 * DL_KC;
#define DL_GO2(MT, NT, F32, STG)                                                                                                                     \
  {                                                                                                                                            \
    static bool attr_set = false;                                                                                                              \
    constexpr size_t smem = dl_smem_bytes(MT, NT, STG);                                                                                        \
    if (!attr_set) { TXL_CUDA(cudaFuncSetAttribute(dec_linear_kernel<MT, NT, F32, STG>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); attr_set = true; } \
    TXL_CUDA(txl_launch_pdl(dec_linear_kernel<MT, NT, F32, STG>, grid, dim3(DL_THREADS), smem, st, a, lda, w, ldw, bias, C, ldc, M, N, K, relu, prefetch, prefetch_bytes));   \
  }
#define DL_GO(MT, NT, F32) { if (two_stage) DL_GO2(MT, NT, F32, 2) else DL_GO2(MT, NT, F32, 4) }
#define DL_F(MT, NT) { if (out_f32) DL_GO(MT, NT, true) else DL_GO(MT, NT, false) }
#define DL_MT(NT) { if (mt == 1) DL_F(1, NT) else if (mt == 2) DL_F(2, NT) else DL_F(4, NT) }
  if (nt == 2) DL_MT(2) else DL_MT(1)
#undef DL_MT
#undef DL_F
#undef DL_GO
#undef DL_GO2
  TXL_LAUNCH_CHECK();
  return TXL_OK;
}

extern "C" int txl_dec_add_ln(const void* x, const float* part, int nparts, const float* bias, const float* gamma, const float* beta, void* y, int M,
                              int d, float eps, const void* prefetch, int64_t prefetch_bytes, void* stream) {
  TXL_CHECK_ARG(x && gamma && beta && y && M > 0 && d > 0 && d % 8 == 0 && d <= 1024 && nparts >= 0 && (nparts == 0 || part), "dec_add_ln: bad args (d %% 8 == 0, d <= 1024)");
  TXL_CHECK_ARG(((uintptr_t)x & 15) == 0 && ((uintptr_t)y & 15) == 0 && ((uintptr_t)part & 15) == 0 && ((uintptr_t)bias & 15) == 0 &&
                ((uintptr_t)gamma & 15) == 0 && ((uintptr_t)beta & 15) == 0, "dec_add_ln: 16-byte alignment");
  cudaStream_t st = (cudaStream_t)stream;
  dim3 grid((unsigned)cdiv64(M, 2));
  const bf16* xp = (const bf16*)x; bf16* yp = (bf16*)y;
  if (d <= 256) { TXL_CUDA(txl_launch_pdl(dec_add_ln_kernel<1>, grid, dim3(64), 0, st, xp, part, nparts, bias, gamma, beta, yp, M, d, eps, prefetch, prefetch_bytes)); }
  else if (d <= 512) { TXL_CUDA(txl_launch_pdl(dec_add_ln_kernel<2>, grid, dim3(64), 0, st, xp, part, nparts, bias, gamma, beta, yp, M, d, eps, prefetch, prefetch_bytes)); }
  else { TXL_CUDA(txl_launch_pdl(dec_add_ln_kernel<4>, grid, dim3(64), 0, st, xp, part, nparts, bias, gamma, beta, yp, M, d, eps, prefetch, prefetch_bytes)); }
  TXL_LAUNCH_CHECK();
  return TXL_OK;
}

extern "C" int txl_decode_rtab_head_major(const void* r, void* out, int rows, int H, int dh, void* stream) {
  TXL_CHECK_ARG(r && out && rows > 0 && H > 0 && dh > 0, "decode_rtab_head_major: bad args");
  const int grid = (int)imin64(cdiv64((int64_t)rows * H * dh, 256), (int64_t)txl_num_sms() * 8);
  rtab_head_major_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>((const bf16*)r, (bf16*)out, rows, H, dh);
  TXL_LAUNCH_CHECK();
  return TXL_OK;
}

static int g_da_cfg = -1;   // -1: take TXL_DECODE_ATTN_CFG at first use
extern "C" int txl_decode_attn_pipe_config(int cfg) { const int old = g_da_cfg; g_da_cfg = cfg; return old; }

extern "C" int64_t txl_decode_attn_pipe_ws_bytes(int B, int H, int dh, int splits) {
  return splits <= 1 ? 0 : (int64_t)B * H * splits * (dh + 2) * (int64_t)sizeof(float);
}

extern "C" int txl_decode_cache_init_kv(const void* kv_mem, int64_t ld, void* kvc, int B, int H, int ML, int dh, void* stream) {
  TXL_CHECK_ARG(kv_mem && kvc && B > 0 && H > 0 && ML > 0 && dh > 0 && ld >= 2 * H * dh, "decode_cache_init_kv: bad args");
  const int grid = (int)imin64(cdiv64((int64_t)B * H * ML * 2 * dh, 256), (int64_t)txl_num_sms() * 16);
  cache_init_kv_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>((const bf16*)kv_mem, ld, (bf16*)kvc, B, H, ML, dh);
  TXL_LAUNCH_CHECK();
  return TXL_OK;
}

extern "C" int txl_decode_attn_pipe(const void* qkv, void* kvc, const void* r_head_major, const float* rwb, const float* rrb, void* out,
                                    const int32_t* pos, int B, int H, int ML, int dh, int splits, void* ws, int* counters, void* stream) {
  TXL_CHECK_ARG(qkv && kvc && r_head_major && rwb && rrb && out && pos && B > 0 && H > 0 && ML > 0, "decode_attn_pipe: bad args");
  TXL_CHECK_ARG(dh == 32 || dh == 64 || dh == 128, "decode_attn_pipe: d_head %d not in {32,64,128}", dh);
  TXL_CHECK_ARG(((uintptr_t)kvc & 15) == 0 && ((uintptr_t)r_head_major & 15) == 0 && ((uintptr_t)qkv & 15) == 0, "decode_attn_pipe: 16-byte alignment");
  TXL_CHECK_ARG(splits >= 1 && splits <= 64 && (splits == 1 || (ws && counters)), "decode_attn_pipe: splits > 1 need ws (txl_decode_attn_pipe_ws_bytes) and zeroed int counters[B*H]");
  dim3 grid(H, B, splits);
  cudaStream_t st = (cudaStream_t)stream;
  // stage geometry (A/B switch TXL_DECODE_ATTN_CFG / txl_decode_attn_pipe_config), keys per stage x stages, us per cfg4 step of 12 layers:
  //   0 = 32 x 4 (48 KB, 4 CTAs/SM) 671    5 = 64 x 2 (48 KB, 4/SM) 635    1 = 64 x 3 (72 KB, 3/SM: 512 CTAs on 444 slots, two waves) 1035
  //   2 = 64 x 4 (96 KB, 2/SM) 631         3 = 128 x 2 (96 KB, 2/SM) 618 <- default         4 = 128 x 4 (192 KB, 1/SM) 723
  // Longer contiguous runs per bulk copy (32 KB of k|v rows at 128 keys) and fewer concurrent streams beat more resident CTAs.
  if (g_da_cfg < 0) { const char* e = getenv("TXL_DECODE_ATTN_CFG"); g_da_cfg = e ? atoi(e) : 3; }
  const int cfg = g_da_cfg;
#define DA_LAUNCH2(DHV, PS, STG)                                                                                                              \
  {                                                                                                                                           \
    static bool attr_set = false;                                                                                                             \
    constexpr size_t smem = da_smem_bytes(PS, STG);                                                                                           \
    if (!attr_set) { TXL_CUDA(cudaFuncSetAttribute(decode_attn_pipe_kernel<DHV, PS, STG>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); attr_set = true; } \
    TXL_CUDA(txl_launch_pdl(decode_attn_pipe_kernel<DHV, PS, STG>, grid, dim3(DA_THREADS), smem, st, (const bf16*)qkv, (bf16*)kvc,            \
                            (const bf16*)r_head_major, rwb, rrb, (bf16*)out, pos, H, ML, (float*)ws, counters));                              \
  }
#define DA_LAUNCH(DHV) { if (cfg == 1) DA_LAUNCH2(DHV, 2, 3) else if (cfg == 2) DA_LAUNCH2(DHV, 2, 4) else if (cfg == 3) DA_LAUNCH2(DHV, 4, 2) \
                         else if (cfg == 4) DA_LAUNCH2(DHV, 4, 4) else if (cfg == 5) DA_LAUNCH2(DHV, 2, 2) else if (cfg == 6) DA_LAUNCH2(DHV, 4, 3) else DA_LAUNCH2(DHV, 1, 4) }
  if (dh == 32) DA_LAUNCH(32) else if (dh == 64) DA_LAUNCH(64) else DA_LAUNCH(128)
#undef DA_LAUNCH
#undef DA_LAUNCH2
  TXL_LAUNCH_CHECK();
  return TXL_OK;
}
