// tc_gemm.cu — bf16 GEMM on the 5th-gen tensor cores: TMA (SWIZZLE_128B) -> shared-memory ring -> tcgen05.mma (one elected
// thread) -> fp32 accumulators in TMEM (double-buffered) -> tcgen05.ld epilogue (bias / ReLU / mask / dropout / accumulate /
// column sums) -> global.  Persistent, warp-specialised: warp 0 = TMA producer, warp 1 = MMA issuer + TMEM owner,
// warps 2..5 = epilogue (one TMEM lane quadrant each).  All four operand layouts (K-major / MN-major for A and B) are taken
// straight from row-major tensors, so forward (x W^T), dgrad (dy W) and wgrad (dy^T x, split-K with fp32 reductions) need no
// transposed copies.   [A.3 qkv_net / r_net / o_net, A.6 CoreNet.0 / CoreNet.3, crit.out_layers.0]
#include "tc_common.cuh"
#include <stdlib.h>
#include <mutex>

// ------------------------------------------------------------------ host: tensor maps through the driver entry point
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)p;
  });
  return fn;
}
// Encoded maps are cached per (base, shape, pitch, box): a training step re-encodes the same few hundred descriptors (weights, activation
// buffers the caching allocator hands back at the same address) every step, 5-12 per attention call.  A map is a pure function of the key
// (no device state), so entries never go stale; the table is dropped when it reaches 8192 entries.
#include <unordered_map>
namespace {
struct TmapKey {
  uint64_t v[7];
  bool operator==(const TmapKey& o) const { for (int i = 0; i < 7; ++i) if (v[i] != o.v[i]) return false; return true; }
};
struct TmapKeyHash {
  size_t operator()(const TmapKey& k) const {
    uint64_t h = 0x9E3779B97F4A7C15ull;
    for (int i = 0; i < 7; ++i) { h ^= k.v[i] + 0x9E3779B97F4A7C15ull + (h << 6) + (h >> 2); }
    return (size_t)h;
  }
};
std::mutex g_tmap_mu;
std::unordered_map<TmapKey, CUtensorMap, TmapKeyHash> g_tmap_cache;
bool tmap_lookup(const TmapKey& k, CUtensorMap* out) {
  std::lock_guard<std::mutex> lk(g_tmap_mu);
  auto it = g_tmap_cache.find(k);
  if (it == g_tmap_cache.end()) return false;
  *out = it->second;
  return true;
}
void tmap_store(const TmapKey& k, const CUtensorMap& m) {
  std::lock_guard<std::mutex> lk(g_tmap_mu);
  if (g_tmap_cache.size() >= 8192) g_tmap_cache.clear();
  g_tmap_cache.emplace(k, m);
}
}  // namespace

int txl_make_tmap_2d(CUtensorMap* out, const void* base, uint64_t rows, uint64_t cols, uint64_t ld, uint32_t box_rows, uint32_t box_cols) {
  const TmapKey key = {{(uint64_t)(uintptr_t)base, rows, cols, ld, ((uint64_t)box_rows << 32) | box_cols, 0, 2}};
  if (tmap_lookup(key, out)) return TXL_OK;
  EncodeTiledFn enc = get_encode();
  if (!enc) { txl_set_error("cuTensorMapEncodeTiled not available from the driver"); return TXL_ECUDA; }
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {ld * 2};
  cuuint32_t box[2] = {box_cols, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { txl_set_error("cuTensorMapEncodeTiled(2d rows=%llu cols=%llu ld=%llu box=%ux%u) failed: %d", (unsigned long long)rows, (unsigned long long)cols, (unsigned long long)ld, box_rows, box_cols, (int)r); return TXL_ECUDA; }
  tmap_store(key, *out);
  return TXL_OK;
}
int txl_make_tmap_3d(CUtensorMap* out, const void* base, uint64_t d2, uint64_t rows, uint64_t cols, uint64_t ld2, uint64_t ld, uint32_t box_rows,
                     uint32_t box_cols) {
  const TmapKey key = {{(uint64_t)(uintptr_t)base, rows, cols, ld, ((uint64_t)box_rows << 32) | box_cols, (d2 << 32) ^ ld2, 3}};
  if (tmap_lookup(key, out)) return TXL_OK;
  EncodeTiledFn enc = get_encode();
  if (!enc) { txl_set_error("cuTensorMapEncodeTiled not available from the driver"); return TXL_ECUDA; }
  cuuint64_t dims[3] = {cols, rows, d2};
  cuuint64_t strides[2] = {ld * 2, ld2 * 2};
  cuuint32_t box[3] = {box_cols, box_rows, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { txl_set_error("cuTensorMapEncodeTiled(3d) failed: %d", (int)r); return TXL_ECUDA; }
  tmap_store(key, *out);
  return TXL_OK;
}

namespace {
constexpr int BM = 128, BK = 64;
constexpr int threads_for(int eg) { return 64 + eg * 128; }   // warp 0 TMA, warp 1 MMA, then EG epilogue groups of four lane-quadrant warps
constexpr int A_BYTES = BM * BK * 2;  // 16 KB
constexpr int STAGE_C_BYTES = BM * 128;  // one 64-column bf16 block of the output tile, SWIZZLE_128B

struct GemmParams {
  int64_t M, N, K, ldc;
  void* C;
  int dtype_c;
  TxlEpilogue epi;
  float inv_keep;
  int tiles_m, tiles_n, ksplits, nkb, kb_per_split;
  int a_mn, b_mn;
  int tma_store;
  // LN > 0 (fused dropout + residual + LayerNorm epilogue): y = LN(resid + dropout(A W^T + bias)), z = the LayerNorm input
  const bf16* resid;
  const float *gamma, *beta;
  float *mean, *rstd;
  float eps;
  int store_z;
};

// CG = 2: a CTA pair (cluster of two CTAs on one TPC) computes one 256 x BN tile with tcgen05.mma.cta_group::2 — each CTA stages its
// own 128 rows of A and HALF of the B tile, so every byte pulled from L2 feeds twice the MMA work; the leader CTA (rank 0) issues.
template <int BN, int STAGES, int CG = 1, int EG = 2, int LN = 0>
struct Cfg {
  static constexpr int B_BYTES = (BN / CG) * BK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int CSTAGE_OFF = STAGES * STAGE_BYTES;
  // output staging buffers per epilogue group: a TMA store takes ~1400-2400 cycles to finish READING its shared-memory source, so with
  // one buffer every 64-column block waits out a full store; the pair kernel has the room for two
  static constexpr int NCST = (CG == 2 && EG == 2) ? 2 : 1;
  static constexpr int BAR_OFF = CSTAGE_OFF + EG * NCST * STAGE_C_BYTES;
  static constexpr int PART_OFF = BAR_OFF + 256;                     // LN: per-row partial sums of the epilogue groups, [2][EG][128] fp32
  static constexpr int SMEM = LN ? PART_OFF + 2 * EG * BM * 4 + 1024 : BAR_OFF + (2 * STAGES + 4) * 8 + 16 + 1024;
  static constexpr int TMEM_COLS = 2 * BN;
  static_assert((2 * STAGES + 4) * 8 + 16 <= 256, "barrier block overlaps the LN partial sums");
};

// butterfly transpose-reduce: on return v[0] of lane l holds sum over the warp's 32 rows of column l (31 shuffles)
__device__ __forceinline__ float warp_colsum32(float* v, int lane) {
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) {
#pragma unroll
    for (int i = 0; i < off; ++i) {
      bool up = lane & off;
      float send = up ? v[i] : v[i + off];
      float keep = up ? v[i + off] : v[i];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
    }
  }
  return v[0];
}
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory"); }
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* m, const void* src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(src)),
               "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int PENDING>
__device__ __forceinline__ void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(PENDING) : "memory"); }
__device__ __forceinline__ void tma_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

// ---- cta_group::2 primitives (PTX forms as in cute/arch/copy_sm100_tma.hpp, cutlass/arch/barrier.h)
__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t mapa_rank(uint32_t smem_addr, uint32_t rank) {
  uint32_t r; asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_addr), "r"(rank)); return r;
}
// tile load into THIS CTA's shared memory; the transaction bytes are credited to `bar_cluster_addr`, a barrier of the pair's leader CTA
__device__ __forceinline__ void tma_load_2d_2sm(void* dst, const CUtensorMap* m, uint32_t bar_cluster_addr, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(smem_u32(dst)),
               "l"(reinterpret_cast<uint64_t>(m)), "r"(bar_cluster_addr), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t bar_cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(bar_cluster_addr) : "memory");
}
__device__ __forceinline__ void umma_bf16_2sm(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrives on the barrier at the same shared-memory offset in BOTH CTAs of the pair once all prior MMAs of this thread are done
__device__ __forceinline__ void umma_commit_2sm(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(bar)), "h"((uint16_t)3)
               : "memory");
}
template <int NCOLS>
__device__ __forceinline__ void tmem_alloc_2sm(uint32_t* slot_in_smem) {  // the same warp of BOTH CTAs
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot_in_smem)), "n"(NCOLS) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
template <int NCOLS>
__device__ __forceinline__ void tmem_dealloc_2sm(uint32_t taddr) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(NCOLS) : "memory");
}

// ---- fused dropout + residual + LayerNorm epilogue (LN column tiles of BN columns = the whole output row).
// Thread (group g, lane quadrant q, lane) owns row 32q + lane of the CTA's 128 rows and, in every column tile, the 64-column block g.
// Pass 1 (per column tile, as soon as its accumulator is complete — the MMAs of the next tile run meanwhile):
//   z = round_bf16(resid + dropout(acc + bias)), row sum.  The bf16 z of the row has to wait for the row statistics, and 96 registers per
//   thread (18 warps) cannot hold it: the first of two tiles parks its z block in the group's staging buffer (where its TMA store reads
//   it anyway) and returns its accumulator at once; the LAST tile writes its packed z back into the first 32 of the 64 TMEM columns it has
//   just read (tcgen05.st) and returns the accumulator after pass 2 — by then the MMAs of the next row block are still in the other buffer.
// The EG groups meet in shared memory for the mean and once more for the centred second moment (the two-pass statistics of
// add_ln_fwd_kernel); pass 2 turns z into y = (z - mean) rstd gamma + beta block by block through the staging buffer -> TMA stores.
template <int BN, int CG, int EG, int LN, typename C_>
__device__ __forceinline__ void ln_epilogue(const GemmParams& p, const CUtensorMap& tmZ, const CUtensorMap& tmY, uint8_t* sm, uint64_t* tfull, uint64_t* tempty,
                                            uint32_t tmem_base, uint32_t cta_rank, int unit0, int unit_step, int q, int grp, int lane, bool leader, uint8_t* cst) {
  static_assert(BN == 64 * EG && (LN == 1 || LN == 2), "one 64-column block per epilogue group and column tile; one tile may wait in the staging buffer");
  const int r_in_tile = q * 32 + lane;
  const TxlEpilogue& e = p.epi;
  float* part = reinterpret_cast<float*>(sm + C_::PART_OFF);
  const uint32_t dkey = dropout_key(e.seed, e.site);
  const uint32_t dthr = dropout_threshold(e.drop_p);
  const bool drop = (e.flags & TXL_EPI_DROPOUT) != 0;
  const float inv_n = 1.f / (float)p.N;
  uint8_t* const crow = cst + r_in_tile * 128;                      // this thread's 128-byte row of the staging block (16-byte chunks swizzled by row)
  const int sw = r_in_tile & 7;
  const uint32_t tm_mine = tmem_base + grp * 64 + ((uint32_t)(q * 32) << 16);      // this thread's lane, its group's 64 columns of accumulator 0
  auto stage_wait = [&]() { if (leader) tma_store_wait_read<0>(); named_bar_sync(1 + grp, 128); };          // the last TMA store has read the staging buffer
  auto stage_store = [&](const CUtensorMap* m, int col, int row) {                                            // staging buffer -> global (rows >= M clipped)
    fence_proxy_async_smem();
    named_bar_sync(1 + grp, 128);
    if (leader) { tma_store_2d(m, cst, col, row); tma_store_commit(); }
  };
  auto release = [&](int acc) {        // whole warp: its reads (and writes) of this accumulator are done
    tc_fence_before();
    __syncwarp();
    if (lane == 0) { if (CG == 2) mbar_arrive_cluster(mapa_rank(smem_u32(&tempty[acc]), 0)); else mbar_arrive(&tempty[acc]); }
  };
  int it = 0;
  for (int tm = unit0; tm < p.tiles_m; tm += unit_step) {
    const int64_t row0 = (int64_t)tm * (BM * CG) + (int64_t)cta_rank * BM;
    const int64_t row = row0 + r_in_tile;
    const bool row_ok = row < p.M;
    float s = 0.f;
    int acc_last = 0;
#pragma unroll
    for (int tn = 0; tn < LN; ++tn, ++it) {
      const bool in_tmem = tn == LN - 1;
      const int acc = it & 1; const uint32_t acc_ph = (it >> 1) & 1;
      const int blk_col0 = tn * BN + grp * 64;
      acc_last = acc;
      // this thread's 64 residual values of the block: requested before the accumulator is waited for, so the global-load latency
      // hides behind the MMAs instead of sitting in front of every 16-column step
      uint4 rxa[8];
#pragma unroll
      for (int c = 0; c < 8; ++c) rxa[c] = make_uint4(0, 0, 0, 0);
      if (row_ok) {
        const uint4* xr = reinterpret_cast<const uint4*>(p.resid + row * p.N + blk_col0);
#pragma unroll
        for (int c = 0; c < 8; ++c) rxa[c] = __ldg(xr + c);
      }
      mbar_wait(&tfull[acc], acc_ph);
      tc_fence_after();
      if (!in_tmem) stage_wait();
#pragma unroll
      for (int qt = 0; qt < 4; ++qt) {                  // 16 columns at a time
        const int col0 = blk_col0 + qt * 16;
        const uint4* rx = rxa + 2 * qt;
        float v[16];
        tmem_ld_32x16(tm_mine + acc * BN + qt * 16, v);
        tmem_ld_wait();
        if (!in_tmem && qt == 3) release(acc);
        if (e.bias) {
#pragma unroll
          for (int g4 = 0; g4 < 4; ++g4) {
            const float4 b4 = __ldg(reinterpret_cast<const float4*>(e.bias + col0 + g4 * 4));
            v[g4 * 4] += b4.x; v[g4 * 4 + 1] += b4.y; v[g4 * 4 + 2] += b4.z; v[g4 * 4 + 3] += b4.w;
          }
        }
        if (drop) {
          const uint64_t base_idx = (uint64_t)row * (uint64_t)p.N + (uint64_t)col0;
#pragma unroll
          for (int i = 0; i < 16; i += 2) {
            const uint32_t hsh = dropout_hash(dkey, (base_idx + i) >> 1);
            v[i] *= (hsh & 0xFFFFu) >= dthr ? p.inv_keep : 0.f;
            v[i + 1] *= (hsh >> 16) >= dthr ? p.inv_keep : 0.f;
          }
        }
        uint32_t zw[8];
        const uint32_t xw[8] = {rx[0].x, rx[0].y, rx[0].z, rx[0].w, rx[1].x, rx[1].y, rx[1].z, rx[1].w};
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const float z0 = __uint_as_float(xw[k] << 16) + v[2 * k];
          const float z1 = __uint_as_float(xw[k] & 0xFFFF0000u) + v[2 * k + 1];
          __nv_bfloat162 zz = __floats2bfloat162_rn(z0, z1);
          const uint32_t w = row_ok ? *reinterpret_cast<uint32_t*>(&zz) : 0u;
          zw[k] = w;
          s += __uint_as_float(w << 16) + __uint_as_float(w & 0xFFFF0000u);
        }
        if (in_tmem) {
          tmem_st_32x8(tm_mine + acc * BN + qt * 8, zw);          // columns 8 qt .. 8 qt + 7: their accumulator values were read in steps <= qt
        } else {
          *reinterpret_cast<uint4*>(crow + (((qt * 2) ^ sw) << 4)) = make_uint4(zw[0], zw[1], zw[2], zw[3]);
          *reinterpret_cast<uint4*>(crow + (((qt * 2 + 1) ^ sw) << 4)) = make_uint4(zw[4], zw[5], zw[6], zw[7]);
        }
      }
      if (in_tmem) tmem_st_wait();
      else if (p.store_z) stage_store(&tmZ, blk_col0, (int)row0);
    }
    // ---- row statistics across the EG groups
    part[grp * BM + r_in_tile] = s;
    named_bar_sync(6, EG * 128);
    float mu = 0.f;
#pragma unroll
    for (int g = 0; g < EG; ++g) mu += part[g * BM + r_in_tile];
    mu *= inv_n;
    float qs = 0.f;
    const uint32_t tm_z = tm_mine + acc_last * BN;        // the last tile's packed z: 32 TMEM columns
#pragma unroll
    for (int hb = 0; hb < 2; ++hb) {
      float zf[16];
      tmem_ld_32x16(tm_z + hb * 16, zf);
      tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const uint32_t w = __float_as_uint(zf[i]);
        const float a = __uint_as_float(w << 16) - mu, b = __uint_as_float(w & 0xFFFF0000u) - mu;
        qs = fmaf(a, a, qs); qs = fmaf(b, b, qs);
      }
    }
    if (LN == 2) {
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        const uint4 u = *reinterpret_cast<const uint4*>(crow + ((c ^ sw) << 4));
        const uint32_t w4[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const float a = __uint_as_float(w4[k] << 16) - mu, b = __uint_as_float(w4[k] & 0xFFFF0000u) - mu;
          qs = fmaf(a, a, qs); qs = fmaf(b, b, qs);
        }
      }
    }
    part[(EG + grp) * BM + r_in_tile] = qs;
    named_bar_sync(6, EG * 128);
    float var = 0.f;
#pragma unroll
    for (int g = 0; g < EG; ++g) var += part[(EG + g) * BM + r_in_tile];
    const float rs = rsqrtf(var * inv_n + p.eps);
    if (grp == 0 && row_ok && p.mean) { p.mean[row] = mu; p.rstd[row] = rs; }
    // ---- pass 2: y = (z - mu) rstd gamma + beta, one 64-column block at a time through the staging buffer
    auto norm8 = [&](const uint32_t* zw, int col, uint32_t* yw) {
      const float4 g0 = __ldg(reinterpret_cast<const float4*>(p.gamma + col)), g1 = __ldg(reinterpret_cast<const float4*>(p.gamma + col + 4));
      const float4 b0 = __ldg(reinterpret_cast<const float4*>(p.beta + col)), b1 = __ldg(reinterpret_cast<const float4*>(p.beta + col + 4));
      const float gm[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w}, bt[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float y0 = (__uint_as_float(zw[k] << 16) - mu) * rs * gm[2 * k] + bt[2 * k];
        const float y1 = (__uint_as_float(zw[k] & 0xFFFF0000u) - mu) * rs * gm[2 * k + 1] + bt[2 * k + 1];
        __nv_bfloat162 yy = __floats2bfloat162_rn(y0, y1);
        yw[k] = *reinterpret_cast<uint32_t*>(&yy);
      }
    };
    // last tile: packed z back from TMEM, then the accumulator goes back to the MMA warp (before the y stores: the next row block's
    // second tile is waiting for it)
    uint32_t zp[32];
    {
      float* zf = reinterpret_cast<float*>(zp);
      tmem_ld_32x16(tm_z, zf);
      tmem_ld_32x16(tm_z + 16, zf + 16);
      tmem_ld_wait();
      release(acc_last);
    }
    // (y straight from registers with 16-byte st.global instead of staging + TMA store measured slower: o_net 52.7 vs 46.3 us)
    if (LN == 2) {      // first tile: z block sits in the staging buffer; normalise it in place once its own store (if any) has read it
      if (p.store_z) stage_wait();
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        uint4* slot = reinterpret_cast<uint4*>(crow + ((c ^ sw) << 4));
        const uint4 u = *slot;
        const uint32_t zw[4] = {u.x, u.y, u.z, u.w};
        uint32_t yw[4];
        norm8(zw, grp * 64 + c * 8, yw);
        *slot = make_uint4(yw[0], yw[1], yw[2], yw[3]);
        asm volatile("" ::: "memory");
      }
      stage_store(&tmY, grp * 64, (int)row0);
    }
    const int last_col0 = (LN - 1) * BN + grp * 64;
    if (p.store_z) {    // last tile's z block: staging -> global
      stage_wait();
#pragma unroll
      for (int c = 0; c < 8; ++c) *reinterpret_cast<uint4*>(crow + ((c ^ sw) << 4)) = make_uint4(zp[c * 4], zp[c * 4 + 1], zp[c * 4 + 2], zp[c * 4 + 3]);
      stage_store(&tmZ, last_col0, (int)row0);
    }
    stage_wait();
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      uint32_t yw[4];
      norm8(zp + c * 4, last_col0 + c * 8, yw);
      *reinterpret_cast<uint4*>(crow + ((c ^ sw) << 4)) = make_uint4(yw[0], yw[1], yw[2], yw[3]);
      asm volatile("" ::: "memory");
    }
    stage_store(&tmY, last_col0, (int)row0);
  }
  if (leader) tma_store_wait_all();
}

// LN > 0: the output row is LN x BN columns wide and every row block's LN column tiles run back to back on the same CTA (pair), so the
// epilogue warps see whole rows: they add bias, dropout and the residual, keep the bf16 LayerNorm input packed in registers, meet once per
// row block for the row statistics (shared-memory partial sums) and write z (optional) and y = LayerNorm(z) through the staging buffers.
template <int BN, int STAGES, typename TC, int CG, int EG, int LN = 0>
__global__ void __launch_bounds__(threads_for(EG), 1)
tc_gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const __grid_constant__ CUtensorMap tmC,
               const __grid_constant__ CUtensorMap tmY, const GemmParams p) {
  using C_ = Cfg<BN, STAGES, CG, EG, LN>;
  const uint32_t cta_rank = CG == 2 ? cluster_ctarank() : 0u;      // 0 = leader of the pair (issues the MMAs)
  const int unit0 = (int)blockIdx.x / CG, unit_step = (int)gridDim.x / CG;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  uint8_t* sm = smem_raw + ((1024u - (raw & 1023u)) & 1023u);
  uint64_t* full = reinterpret_cast<uint64_t*>(sm + C_::BAR_OFF);
  uint64_t* empty = full + STAGES;
  uint64_t* tfull = empty + STAGES;
  uint64_t* tempty = tfull + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(&tfull[a], 1); mbar_init(&tempty[a], 4 * EG * CG); }   // leader's tempty: epilogue warps of both CTAs
    fence_barrier_init();
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
  }
  if (warp == 1) { if (CG == 2) tmem_alloc_2sm<C_::TMEM_COLS>(tmem_slot); else tmem_alloc<C_::TMEM_COLS>(tmem_slot); }
  tc_fence_before();
  __syncthreads();
  if (CG == 2) cluster_sync_all();      // the peer's barriers exist before anything is signalled across the pair
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int total_units = p.tiles_m * p.tiles_n * p.ksplits;
  // work item vi of this CTA (pair): plain GEMM = unit0 + vi * step over (ks, tn, tm); LN = column tile vi % LN of row block unit0 + (vi / LN) * step
  auto unit_of = [&](int vi, int& tm, int& tn, int& ks) -> bool {
    if constexpr (LN > 0) {
      tm = unit0 + (vi / LN) * unit_step; tn = vi % LN; ks = 0;
      return tm < p.tiles_m;
    } else {
      const int unit = unit0 + vi * unit_step;
      if (unit >= total_units) return false;
      tm = unit % p.tiles_m; tn = (unit / p.tiles_m) % p.tiles_n; ks = unit / (p.tiles_m * p.tiles_n);
      return true;
    }
  };

  if (warp == 0) {
    if (lane == 0) {
      // ================= TMA producer
      int s = 0; uint32_t ph = 0;
      int tm, tn, ks;
      for (int vi = 0; unit_of(vi, tm, tn, ks); ++vi) {
        const int kb0 = ks * p.kb_per_split, kb1 = min(kb0 + p.kb_per_split, p.nkb);
        const int m0 = tm * (BM * CG) + (int)cta_rank * BM, n0 = tn * BN + (int)cta_rank * (BN / CG) * (CG - 1);
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&empty[s], ph ^ 1);
          uint8_t* a_s = sm + s * C_::STAGE_BYTES;
          uint8_t* b_s = a_s + A_BYTES;
          const int k0 = kb * BK;
          if (CG == 1) {
            mbar_expect_tx(&full[s], C_::STAGE_BYTES);
            if (!p.a_mn) {
              tma_load_2d(a_s, &tmA, &full[s], k0, m0);
            } else {
#pragma unroll
              for (int i = 0; i < BM / 64; ++i) tma_load_2d(a_s + i * 8192, &tmA, &full[s], m0 + 64 * i, k0);
            }
            if (!p.b_mn) {
              tma_load_2d(b_s, &tmB, &full[s], k0, n0);
            } else {
#pragma unroll
              for (int i = 0; i < BN / 64; ++i) tma_load_2d(b_s + i * 8192, &tmB, &full[s], n0 + 64 * i, k0);
            }
          } else {
            // both CTAs load into their own shared memory; all bytes are credited to the LEADER's full barrier, which expects both halves
            const uint32_t lead_full = mapa_rank(smem_u32(&full[s]), 0);
            if (cta_rank == 0) mbar_expect_tx(&full[s], 2 * C_::STAGE_BYTES);
            if (!p.a_mn) {
              tma_load_2d_2sm(a_s, &tmA, lead_full, k0, m0);
            } else {
#pragma unroll
              for (int i = 0; i < BM / 64; ++i) tma_load_2d_2sm(a_s + i * 8192, &tmA, lead_full, m0 + 64 * i, k0);
            }
            if (!p.b_mn) {
              tma_load_2d_2sm(b_s, &tmB, lead_full, k0, n0);
            } else {
#pragma unroll
              for (int i = 0; i < BN / CG / 64; ++i) tma_load_2d_2sm(b_s + i * 8192, &tmB, lead_full, n0 + 64 * i, k0);
            }
          }
          if (++s == STAGES) { s = 0; ph ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0 && cta_rank == 0) {
      // ================= MMA issuer (the pair's leader for CG = 2: M = 256 rows, 128 per CTA)
      const uint32_t idesc = umma_idesc_bf16(BM * CG, BN, p.a_mn, p.b_mn);
      int s = 0; uint32_t ph = 0; int it = 0;
      int tm, tn, ks;
      for (int vi = 0; unit_of(vi, tm, tn, ks); ++vi, ++it) {
        const int kb0 = ks * p.kb_per_split, kb1 = min(kb0 + p.kb_per_split, p.nkb);
        const int acc = it & 1; const uint32_t acc_ph = (it >> 1) & 1;
        mbar_wait(&tempty[acc], acc_ph ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BN;
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&full[s], ph);
          tc_fence_after();
          const uint32_t a_addr = smem_u32(sm + s * C_::STAGE_BYTES);
          const uint32_t b_addr = a_addr + A_BYTES;
#pragma unroll
          for (int kk = 0; kk < BK / 16; ++kk) {
            const uint64_t ad = p.a_mn ? umma_smem_desc(a_addr + kk * 2048, 8192, 1024) : umma_smem_desc(a_addr + kk * 32, 16, 1024);
            const uint64_t bd = p.b_mn ? umma_smem_desc(b_addr + kk * 2048, 8192, 1024) : umma_smem_desc(b_addr + kk * 32, 16, 1024);
            if (CG == 2) umma_bf16_2sm(d_tmem, ad, bd, idesc, (kb > kb0 || kk > 0) ? 1u : 0u);
            else umma_bf16(d_tmem, ad, bd, idesc, (kb > kb0 || kk > 0) ? 1u : 0u);
          }
          if (CG == 2) umma_commit_2sm(&empty[s]); else umma_commit(&empty[s]);
          if (++s == STAGES) { s = 0; ph ^= 1; }
        }
        if (CG == 2) umma_commit_2sm(&tfull[acc]); else umma_commit(&tfull[acc]);
      }
    }
  } else {
    // ================= epilogue: group g = (warp-2)/4 owns 64-column blocks g, g+2, ...; TMEM lane quadrant = warp % 4
    const int q = warp & 3, grp = (warp - 2) >> 2;
    const int r_in_tile = q * 32 + lane;
    const bool leader = (warp - 2) % 4 == 0 && lane == 0;      // issues this group's TMA stores
    const TxlEpilogue& e = p.epi;
    TC* __restrict__ C = reinterpret_cast<TC*>(p.C);
    uint8_t* const cst0 = sm + C_::CSTAGE_OFF + grp * C_::NCST * STAGE_C_BYTES;
    int nblk = 0;                                 // 64-column blocks this group has staged so far (selects the staging buffer)
    const uint32_t dkey = dropout_key(e.seed, e.site);
    const uint32_t dthr = dropout_threshold(e.drop_p);
    const bool pair_hash = (p.N & 1) == 0;
    int it = 0;
    int tm, tn, ks;
    if constexpr (LN > 0) {
      ln_epilogue<BN, CG, EG, LN, C_>(p, tmC, tmY, sm, tfull, tempty, tmem_base, cta_rank, unit0, unit_step, q, grp, lane, leader, cst0);
    } else
    for (int vi = 0; unit_of(vi, tm, tn, ks); ++vi, ++it) {
      const int acc = it & 1; const uint32_t acc_ph = (it >> 1) & 1;
      const int64_t row0 = (int64_t)tm * (BM * CG) + (int64_t)cta_rank * BM;      // first output row of this CTA's half of the tile
      const int64_t row = row0 + r_in_tile;
      const bool row_ok = row < p.M;
      mbar_wait(&tfull[acc], acc_ph);
      tc_fence_after();
      constexpr int NBLK = BN / 64;
#pragma unroll 1
      for (int bi = grp; bi < NBLK; bi += EG) {
        const int64_t blk_col0 = (int64_t)tn * BN + bi * 64;
        if (p.tma_store) {
          if (leader) tma_store_wait_read<C_::NCST - 1>();   // the store that last used this staging buffer has read it
          named_bar_sync(1 + grp, 128);
        }
        uint8_t* const cst = cst0 + (nblk % C_::NCST) * STAGE_C_BYTES;
        ++nblk;
#pragma unroll 1
        for (int half = 0; half < 2; ++half) {
          const int c = bi * 2 + half;
          float v[32];
          // MASK_LIVE: this thread's 32 mask bits are one word of the [N/32][M] bit plane (lanes = consecutive rows: one coalesced 128-byte
          // read per warp instead of 32 x 64 bytes of the activation); requested before the accumulator is read
          uint32_t live_word = 0xFFFFFFFFu;
          if ((e.flags & TXL_EPI_MASK_LIVE) && row_ok && blk_col0 + half * 32 < p.N) live_word = __ldg(e.live_bits + ((blk_col0 + half * 32) >> 5) * p.M + row);
          tmem_ld_32x32(tmem_base + acc * BN + c * 32 + ((uint32_t)(q * 32) << 16), v);
          tmem_ld_wait();
          if (bi + EG >= NBLK && half == 1) {   // this warp's last read of the accumulator: hand TMEM back to the MMA warp
            tc_fence_before();
            __syncwarp();
            if (lane == 0) { if (CG == 2) mbar_arrive_cluster(mapa_rank(smem_u32(&tempty[acc]), 0)); else mbar_arrive(&tempty[acc]); }
          }
          const int64_t col0 = blk_col0 + half * 32;
          const bool any_col = col0 < p.N;
          const bool full_cols = col0 + 32 <= p.N;
          if (any_col) {
            if (e.flags & TXL_EPI_MASK_POS) {
              if (row_ok && full_cols) {
                const TC* arow = reinterpret_cast<const TC*>(e.aux) + row * p.ldc + col0;
                if constexpr (sizeof(TC) == 2) {
#pragma unroll
                  for (int g4 = 0; g4 < 4; ++g4) {
                    uint4 u = reinterpret_cast<const uint4*>(arow)[g4];
                    const bf16* ab = reinterpret_cast<const bf16*>(&u);
#pragma unroll
                    for (int k = 0; k < 8; ++k) v[g4 * 8 + k] = __bfloat162float(ab[k]) > 0.f ? v[g4 * 8 + k] : 0.f;
                  }
                } else {
#pragma unroll
                  for (int i = 0; i < 32; ++i) v[i] = to_f32(arow[i]) > 0.f ? v[i] : 0.f;
                }
              } else if (row_ok) {
#pragma unroll
                for (int i = 0; i < 32; ++i)
                  if (col0 + i < p.N) v[i] = to_f32(reinterpret_cast<const TC*>(e.aux)[row * p.ldc + col0 + i]) > 0.f ? v[i] : 0.f;
              }
            }
            if (e.flags & TXL_EPI_MASK_LIVE) {
#pragma unroll
              for (int i = 0; i < 32; ++i) v[i] = ((live_word >> i) & 1u) ? v[i] : 0.f;
            }
            const bool col_bias = e.bias && !(e.flags & TXL_EPI_BIAS_ROW);
            const float brow = (e.bias && (e.flags & TXL_EPI_BIAS_ROW) && row_ok) ? e.bias[row] : 0.f;
            if (col_bias) {
              if (full_cols) {
#pragma unroll
                for (int g4 = 0; g4 < 8; ++g4) {
                  const float4 b4 = *reinterpret_cast<const float4*>(e.bias + col0 + g4 * 4);
                  v[g4 * 4] += b4.x; v[g4 * 4 + 1] += b4.y; v[g4 * 4 + 2] += b4.z; v[g4 * 4 + 3] += b4.w;
                }
              } else {
#pragma unroll
                for (int i = 0; i < 32; ++i) if (col0 + i < p.N) v[i] += e.bias[col0 + i];
              }
            }
            if (e.flags & TXL_EPI_RELU) {
#pragma unroll
              for (int i = 0; i < 32; ++i) v[i] = fmaxf(v[i] + brow, 0.f);
            } else if (brow != 0.f) {
#pragma unroll
              for (int i = 0; i < 32; ++i) v[i] += brow;
            }
            if (e.flags & TXL_EPI_MASK_SCALE) {
              // aux is a post-dropout activation: aux > 0 already is (ReLU live) AND (kept); only the 1/(1-p) scale is left
#pragma unroll
              for (int i = 0; i < 32; ++i) v[i] *= p.inv_keep;
            }
            if (e.flags & TXL_EPI_DROPOUT) {
              const uint64_t base_idx = (uint64_t)row * (uint64_t)p.N + (uint64_t)col0;
              if (pair_hash) {
#pragma unroll
                for (int i = 0; i < 32; i += 2) {
                  const uint32_t hsh = dropout_hash(dkey, (base_idx + i) >> 1);
                  v[i] *= (hsh & 0xFFFFu) >= dthr ? p.inv_keep : 0.f;
                  v[i + 1] *= (hsh >> 16) >= dthr ? p.inv_keep : 0.f;
                }
              } else {
#pragma unroll
                for (int i = 0; i < 32; ++i) v[i] *= dropout_scale(e.seed, e.site, base_idx + i, e.drop_p, p.inv_keep);
              }
            }
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] = (row_ok && col0 + i < p.N) ? v[i] : 0.f;
            if ((e.flags & TXL_EPI_EMIT_LIVE) && row_ok) {      // one word per (row, 32 columns); columns >= N are zero bits
              uint32_t w = 0;
#pragma unroll
              for (int i = 0; i < 32; ++i) w |= (v[i] > 0.f ? 1u : 0u) << i;
              e.live_bits[(col0 >> 5) * p.M + row] = w;
            }
            if (e.colsum) {
              float t[32];
#pragma unroll
              for (int i = 0; i < 32; ++i) t[i] = v[i];
              float cs = warp_colsum32(t, lane);
              if (col0 + lane < p.N) atomicAdd(&e.colsum[col0 + lane], cs);
            }
          }
          if (p.tma_store) {
            if constexpr (sizeof(TC) == 2) {
#pragma unroll
              for (int g4 = 0; g4 < 4; ++g4) {
                uint4 o;
                __nv_bfloat162 p0 = __floats2bfloat162_rn(v[g4 * 8], v[g4 * 8 + 1]), p1 = __floats2bfloat162_rn(v[g4 * 8 + 2], v[g4 * 8 + 3]);
                __nv_bfloat162 p2 = __floats2bfloat162_rn(v[g4 * 8 + 4], v[g4 * 8 + 5]), p3 = __floats2bfloat162_rn(v[g4 * 8 + 6], v[g4 * 8 + 7]);
                o.x = *reinterpret_cast<uint32_t*>(&p0); o.y = *reinterpret_cast<uint32_t*>(&p1);
                o.z = *reinterpret_cast<uint32_t*>(&p2); o.w = *reinterpret_cast<uint32_t*>(&p3);
                *reinterpret_cast<uint4*>(cst + r_in_tile * 128 + (((half * 4 + g4) ^ (r_in_tile & 7)) << 4)) = o;
              }
            }
            continue;
          }
          if (!row_ok || !any_col) continue;
          if (e.flags & TXL_EPI_TRANSPOSE) {          // lanes hold consecutive rows: each store instruction writes 32 contiguous elements
#pragma unroll
            for (int i = 0; i < 32; ++i)
              if (col0 + i < p.N) C[(col0 + i) * p.ldc + row] = from_f32<TC>(v[i]);
            continue;
          }
          TC* crow = C + row * p.ldc + col0;
          if (p.ksplits > 1) {
            if (full_cols) {      // split-K partial sums: 16-byte vector reductions (red.global.add.v4.f32), 8 per thread instead of 32 scalar ones
              float4* dst = reinterpret_cast<float4*>(crow);
#pragma unroll
              for (int g4 = 0; g4 < 8; ++g4) atomicAdd(dst + g4, make_float4(v[g4 * 4], v[g4 * 4 + 1], v[g4 * 4 + 2], v[g4 * 4 + 3]));
            } else {
#pragma unroll
              for (int i = 0; i < 32; ++i)
                if (col0 + i < p.N) atomicAdd(reinterpret_cast<float*>(crow) + i, v[i]);
            }
          } else if (full_cols) {
            if constexpr (sizeof(TC) == 2) {
              uint4* dst = reinterpret_cast<uint4*>(crow);
#pragma unroll
              for (int g4 = 0; g4 < 4; ++g4) {
                float a8[8];
                if (e.flags & TXL_EPI_ACCUM) {
                  uint4 old = dst[g4];
                  const bf16* ob = reinterpret_cast<const bf16*>(&old);
#pragma unroll
                  for (int i = 0; i < 8; ++i) a8[i] = v[g4 * 8 + i] + __bfloat162float(ob[i]);
                } else {
#pragma unroll
                  for (int i = 0; i < 8; ++i) a8[i] = v[g4 * 8 + i];
                }
                uint4 o;
                __nv_bfloat162 p0 = __floats2bfloat162_rn(a8[0], a8[1]), p1 = __floats2bfloat162_rn(a8[2], a8[3]);
                __nv_bfloat162 p2 = __floats2bfloat162_rn(a8[4], a8[5]), p3 = __floats2bfloat162_rn(a8[6], a8[7]);
                o.x = *reinterpret_cast<uint32_t*>(&p0); o.y = *reinterpret_cast<uint32_t*>(&p1);
                o.z = *reinterpret_cast<uint32_t*>(&p2); o.w = *reinterpret_cast<uint32_t*>(&p3);
                dst[g4] = o;
              }
            } else {
              float4* dst = reinterpret_cast<float4*>(crow);
#pragma unroll
              for (int g4 = 0; g4 < 8; ++g4) {
                float4 o = make_float4(v[g4 * 4], v[g4 * 4 + 1], v[g4 * 4 + 2], v[g4 * 4 + 3]);
                if (e.flags & TXL_EPI_ACCUM) { float4 old = dst[g4]; o.x += old.x; o.y += old.y; o.z += old.z; o.w += old.w; }
                dst[g4] = o;
              }
            }
          } else {
#pragma unroll
            for (int i = 0; i < 32; ++i) {
              if (col0 + i < p.N) {
                float x = v[i];
                if (e.flags & TXL_EPI_ACCUM) x += to_f32(crow[i]);
                crow[i] = from_f32<TC>(x);
              }
            }
          }
        }
        if (p.tma_store) {
          fence_proxy_async_smem();
          named_bar_sync(1 + grp, 128);
          if (leader && blk_col0 < p.N) {               // TMA clips rows >= M and columns >= N
            tma_store_2d(&tmC, cst, (int)blk_col0, (int)row0);
            tma_store_commit();
          }
        }
      }
    }
    if (p.tma_store && leader) tma_store_wait_all();
  }
  tc_fence_before();
  __syncthreads();
  if (CG == 2) cluster_sync_all();      // neither CTA may leave while its peer can still read its operands or signal its barriers
  if (warp == 1) { if (CG == 2) tmem_dealloc_2sm<C_::TMEM_COLS>(tmem_base); else tmem_dealloc<C_::TMEM_COLS>(tmem_base); }
}

template <int BN, int STAGES, typename TC, int CG = 1, int EG = 2, int LN = 0>
int launch(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmC, const GemmParams& p, int grid, cudaStream_t st, const CUtensorMap* tmY = nullptr) {
  using C_ = Cfg<BN, STAGES, CG, EG, LN>;
  static_assert(C_::SMEM <= 232448, "GEMM shared-memory plan exceeds 227 KB");
  static bool attr_set = false;
  if (!attr_set) {
    TXL_CUDA(cudaFuncSetAttribute(tc_gemm_kernel<BN, STAGES, TC, CG, EG, LN>, cudaFuncAttributeMaxDynamicSharedMemorySize, C_::SMEM));
    attr_set = true;
  }
  const CUtensorMap& y = tmY ? *tmY : tmC;
  if (CG == 1) {
    tc_gemm_kernel<BN, STAGES, TC, CG, EG, LN><<<grid, threads_for(EG), C_::SMEM, st>>>(tmA, tmB, tmC, y, p);
  } else {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(grid); cfg.blockDim = dim3(threads_for(EG)); cfg.dynamicSmemBytes = C_::SMEM; cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    ++g_txl_launches;
    TXL_CUDA(cudaLaunchKernelEx(&cfg, tc_gemm_kernel<BN, STAGES, TC, CG, EG, LN>, tmA, tmB, tmC, y, p));
    return TXL_OK;
  }
  TXL_LAUNCH_CHECK();
  return TXL_OK;
}
}  // namespace

// handled=1 when the tensor-core kernel took the call; 0 => caller uses the SIMT kernel.
int txl_gemm_tc(const void* A, const void* B, void* C, int64_t M, int64_t N, int64_t K, int64_t lda, int64_t ldb, int64_t ldc, int transA,
                int transB, int dtype_c, const TxlEpilogue* epi, void* stream, int* handled) {
  *handled = 0;
  static int disabled = -1;
  if (disabled < 0) { const char* e = getenv("TXL_DISABLE_TC"); disabled = (e && e[0] == '1') ? 1 : 0; }
  if (disabled) return TXL_OK;
  // TMA needs 16-byte aligned bases and row pitches; C vector stores need the same
  auto al16 = [](const void* p) { return ((uintptr_t)p & 15) == 0; };
  const int csz = dtype_c == TXL_F32 ? 4 : 2;
  if (!al16(A) || !al16(B) || !al16(C) || (lda % 8) || (ldb % 8) || ((ldc * csz) % 16) || K < 16 || N < 8) return TXL_OK;
  if ((epi->flags & TXL_EPI_MASK_POS) && !epi->aux) return TXL_OK;
  if ((epi->flags & (TXL_EPI_EMIT_LIVE | TXL_EPI_MASK_LIVE)) && (!epi->live_bits || (epi->flags & (TXL_EPI_ACCUM | TXL_EPI_TRANSPOSE)))) return TXL_OK;
  if (M >= (1ll << 31) || N >= (1ll << 31) || K >= (1ll << 31)) return TXL_OK;

  GemmParams p;
  p.M = M; p.N = N; p.K = K; p.ldc = ldc; p.C = C; p.dtype_c = dtype_c; p.epi = *epi;
  p.resid = nullptr; p.gamma = p.beta = nullptr; p.mean = p.rstd = nullptr; p.eps = 0.f; p.store_z = 0;
  p.inv_keep = (epi->flags & (TXL_EPI_DROPOUT | TXL_EPI_MASK_SCALE)) ? 1.f / (1.f - epi->drop_p) : 1.f;
  p.a_mn = transA ? 1 : 0;   // transA: A stored [K, M] => M contiguous
  p.b_mn = transB ? 0 : 1;   // transB: B stored [N, K] => K contiguous
  const int BN = N >= 256 ? 256 : 128;
  // CTA pairs (cta_group::2, 256 x 256 tiles) for the big training shapes; TXL_GEMM_2SM=0 switches them off
  static int use_2sm = -1;
  if (use_2sm < 0) { const char* e = getenv("TXL_GEMM_2SM"); use_2sm = (e && e[0] == '0') ? 0 : 1; }
  // Short-K (K < 1024) bf16-output shapes are bound by the epilogue (bias / ReLU / mask / dropout / column sums per element, staging, store):
  // they run on the pair kernel with FOUR epilogue groups (16 epilogue warps, one 64-column block each): +16..28 % on the cfg2 shapes.
  // Long-K and split-K launches measured 2-4 % slower that way and keep two groups.  TXL_GEMM_EG4=0 switches the variant off.
  static int eg4 = -1;
  if (eg4 < 0) { const char* e = getenv("TXL_GEMM_EG4"); eg4 = (e && e[0] == '0') ? 0 : 1; }
  const bool pair_ok = use_2sm && BN == 256 && M >= 512 && !(epi->flags & TXL_EPI_TRANSPOSE);
  const bool short_k_eg4 = pair_ok && eg4 && K < 1024 && dtype_c == TXL_BF16;
  const int CG = (pair_ok && (K >= 1024 || short_k_eg4)) ? 2 : 1;
  p.tiles_m = (int)cdiv64(M, BM * CG); p.tiles_n = (int)cdiv64(N, BN);
  p.nkb = (int)cdiv64(K, BK);
  p.ksplits = 1;
  const int sms = txl_num_sms();
  const bool pure_accum = dtype_c == TXL_F32 && epi->flags == TXL_EPI_ACCUM && !epi->bias && !epi->colsum;
  const int tiles = p.tiles_m * p.tiles_n;
  const int workers = sms / CG;      // CTAs, or CTA pairs
  if (pure_accum && tiles < workers && p.nkb >= 8) {
    int want = workers / tiles;
    int maxs = p.nkb / 4;
    p.ksplits = want < maxs ? want : maxs;
    if (p.ksplits < 1) p.ksplits = 1;
  }
  p.kb_per_split = (int)cdiv64(p.nkb, p.ksplits);
  p.ksplits = (int)cdiv64(p.nkb, p.kb_per_split);   // no empty split

  p.tma_store = (dtype_c == TXL_BF16 && !(epi->flags & (TXL_EPI_ACCUM | TXL_EPI_TRANSPOSE)) && p.ksplits == 1) ? 1 : 0;
  CUtensorMap tmA, tmB, tmC;
  int rc;
  if (p.tma_store) { if ((rc = txl_make_tmap_2d(&tmC, C, (uint64_t)M, (uint64_t)N, (uint64_t)ldc, BM, 64))) return rc; }
  else tmC = CUtensorMap{};
  if (!p.a_mn) rc = txl_make_tmap_2d(&tmA, A, (uint64_t)M, (uint64_t)K, (uint64_t)lda, BM, BK);
  else rc = txl_make_tmap_2d(&tmA, A, (uint64_t)K, (uint64_t)M, (uint64_t)lda, BK, 64);
  if (rc) return rc;
  if (!p.b_mn) rc = txl_make_tmap_2d(&tmB, B, (uint64_t)N, (uint64_t)K, (uint64_t)ldb, BN / CG, BK);
  else rc = txl_make_tmap_2d(&tmB, B, (uint64_t)K, (uint64_t)N, (uint64_t)ldb, BK, 64);
  if (rc) return rc;

  const int total = tiles * p.ksplits;
  const int grid = (total < workers ? total : workers) * CG;
  cudaStream_t st = (cudaStream_t)stream;
  if (short_k_eg4) {
    rc = launch<256, 5, bf16, 2, 4>(tmA, tmB, tmC, p, grid, st);
  } else if (CG == 2) {
    if (dtype_c == TXL_F32) rc = launch<256, 5, float, 2>(tmA, tmB, tmC, p, grid, st); else rc = launch<256, 5, bf16, 2>(tmA, tmB, tmC, p, grid, st);
  } else if (BN == 256) {
    if (dtype_c == TXL_F32) rc = launch<256, 4, float>(tmA, tmB, tmC, p, grid, st); else rc = launch<256, 4, bf16>(tmA, tmB, tmC, p, grid, st);
  } else {
    if (dtype_c == TXL_F32) rc = launch<128, 6, float>(tmA, tmB, tmC, p, grid, st); else rc = launch<128, 6, bf16>(tmA, tmB, tmC, p, grid, st);
  }
  if (rc) return rc;
  *handled = 1;
  return TXL_OK;
}


// y = LayerNorm(resid + dropout(A W^T + bias)) in ONE kernel: the CTA-pair GEMM with the fused row-statistics epilogue (LN template
// parameter above).  A [M, K] and W [N, K] bf16 row-major, resid / y / z [M, N] bf16 contiguous, N = 256 or 512 (the whole row lives in
// the pair's TMEM: 2 column tiles of 256).  z (the LayerNorm input), mean and rstd are what add_ln_bwd reads; pass NULL for all three in
// evaluation.  handled = 0: shape not covered, the caller runs txl_gemm + txl_add_ln_fwd.   [A.3 o_net + layer_norm, A.6 CoreNet.3 + layer_norm]
extern "C" int txl_gemm_add_ln_fwd(const void* A, const void* W, const float* bias, const void* resid, const float* gamma, const float* beta, void* Y, void* Z,
                                   float* mean, float* rstd, int64_t M, int64_t N, int64_t K, int64_t lda, int64_t ldw, float eps, float drop_p, uint64_t seed,
                                   uint32_t site, void* stream, int* handled) {
  TXL_CHECK_ARG(handled != nullptr, "gemm_add_ln_fwd: handled must not be NULL");
  *handled = 0;
  TXL_CHECK_ARG(A && W && resid && gamma && beta && Y && M > 0 && N > 0 && K > 0, "gemm_add_ln_fwd: null operand or empty shape (M=%lld N=%lld K=%lld)", (long long)M, (long long)N, (long long)K);
  TXL_CHECK_ARG((Z == nullptr) == (mean == nullptr) && (Z == nullptr) == (rstd == nullptr), "gemm_add_ln_fwd: z, mean and rstd are saved together or not at all");
  TXL_CHECK_ARG(drop_p >= 0.f && drop_p < 1.f, "gemm_add_ln_fwd: drop_p=%f outside [0, 1)", drop_p);
  static int off = -1;
  if (off < 0) { const char* e = getenv("TXL_DISABLE_TC"); const char* e2 = getenv("TXL_GEMM_LN"); off = ((e && e[0] == '1') || (e2 && e2[0] == '0')) ? 1 : 0; }
  if (off) return TXL_OK;
  auto al16 = [](const void* p) { return ((uintptr_t)p & 15) == 0; };
  if (!(N == 256 || N == 512) || M < 256 || K < 64 || (K % 8) || (lda % 8) || (ldw % 8) || M >= (1ll << 31)) return TXL_OK;
  if (!al16(A) || !al16(W) || !al16(resid) || !al16(Y) || !al16(gamma) || !al16(beta) || (bias && !al16(bias)) || (Z && !al16(Z))) return TXL_OK;
  GemmParams p;
  p.M = M; p.N = N; p.K = K; p.ldc = N; p.C = Y; p.dtype_c = TXL_BF16;
  p.epi = TxlEpilogue{};
  p.epi.bias = bias; p.epi.drop_p = drop_p; p.epi.seed = seed; p.epi.site = site; p.epi.flags = drop_p > 0.f ? TXL_EPI_DROPOUT : 0;
  p.inv_keep = drop_p > 0.f ? 1.f / (1.f - drop_p) : 1.f;
  p.a_mn = 0; p.b_mn = 0;
  p.tiles_m = (int)cdiv64(M, 2 * BM); p.tiles_n = (int)(N / 256);
  p.nkb = (int)cdiv64(K, BK); p.ksplits = 1; p.kb_per_split = p.nkb;
  p.tma_store = 1;
  p.resid = (const bf16*)resid; p.gamma = gamma; p.beta = beta; p.mean = mean; p.rstd = rstd; p.eps = eps; p.store_z = Z ? 1 : 0;
  CUtensorMap tmA, tmB, tmZ, tmY;
  int rc;
  if ((rc = txl_make_tmap_2d(&tmA, A, (uint64_t)M, (uint64_t)K, (uint64_t)lda, BM, BK))) return rc;
  if ((rc = txl_make_tmap_2d(&tmB, W, (uint64_t)N, (uint64_t)K, (uint64_t)ldw, 128, BK))) return rc;
  if ((rc = txl_make_tmap_2d(&tmY, Y, (uint64_t)M, (uint64_t)N, (uint64_t)N, BM, 64))) return rc;
  if (Z) { if ((rc = txl_make_tmap_2d(&tmZ, Z, (uint64_t)M, (uint64_t)N, (uint64_t)N, BM, 64))) return rc; } else tmZ = tmY;
  const int workers = txl_num_sms() / 2;
  const int grid = (p.tiles_m < workers ? p.tiles_m : workers) * 2;
  if (N == 512) rc = launch<256, 4, bf16, 2, 4, 2>(tmA, tmB, tmZ, p, grid, (cudaStream_t)stream, &tmY);
  else rc = launch<256, 4, bf16, 2, 4, 1>(tmA, tmB, tmZ, p, grid, (cudaStream_t)stream, &tmY);
  if (rc) return rc;
  *handled = 1;
  return TXL_OK;
}
