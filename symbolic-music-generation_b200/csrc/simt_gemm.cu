// simt_gemm.cu — fp32-FMA GEMM with the shared epilogue.  This is the *fp32 parity mode* of every Linear on the
// path (true fp32 multiply-accumulate, no TF32) and the catch-all for shapes the tcgen05 kernel does not take.
// [A.3 qkv_net/r_net/o_net, A.6 CoreNet.0/3, crit.out_layers.0]
#include "common.cuh"

int txl_gemm_tc(const void* A, const void* B, void* C, int64_t M, int64_t N, int64_t K, int64_t lda, int64_t ldb, int64_t ldc,
                int transA, int transB, int dtype_c, const TxlEpilogue* epi, void* stream, int* handled);

namespace {
constexpr int BM = 64, BN = 64, BK = 16, TM = 4, TN = 4;  // 256 threads

template <typename TA, typename TC>
__global__ void __launch_bounds__(256) gemm_simt_kernel(const TA* __restrict__ A, const TA* __restrict__ B, TC* __restrict__ C,
                                                        int64_t M, int64_t N, int64_t K, int64_t lda, int64_t ldb, int64_t ldc,
                                                        int transA, int transB, TxlEpilogue epi, float inv_keep) {
  __shared__ float As[BK][BM + 4];
  __shared__ float Bs[BK][BN + 4];
  const int tid = threadIdx.x;
  const int64_t m0 = (int64_t)blockIdx.y * BM, n0 = (int64_t)blockIdx.x * BN;
  const int tx = tid % 16, ty = tid / 16;
  float acc[TM][TN];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

  for (int64_t k0 = 0; k0 < K; k0 += BK) {
    // ---- stage A tile: As[k][m]
    for (int e = tid; e < BM * BK; e += 256) {
      int m, k;
      if (transA) { m = e % BM; k = e / BM; } else { k = e % BK; m = e / BK; }
      int64_t gm = m0 + m, gk = k0 + k;
      float v = 0.f;
      if (gm < M && gk < K) v = to_f32(transA ? A[gk * lda + gm] : A[gm * lda + gk]);
      As[k][m] = v;
    }
    for (int e = tid; e < BN * BK; e += 256) {
      int n, k;
      if (transB) { k = e % BK; n = e / BK; } else { n = e % BN; k = e / BN; }
      int64_t gn = n0 + n, gk = k0 + k;
      float v = 0.f;
      if (gn < N && gk < K) v = to_f32(transB ? B[gn * ldb + gk] : B[gk * ldb + gn]);
      Bs[k][n] = v;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      float a[TM], b[TN];
#pragma unroll
      for (int i = 0; i < TM; ++i) a[i] = As[k][ty * TM + i];
#pragma unroll
      for (int j = 0; j < TN; ++j) b[j] = Bs[k][tx * TN + j];
#pragma unroll
      for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
  // ---- epilogue
  float cs[TN];
#pragma unroll
  for (int j = 0; j < TN; ++j) cs[j] = 0.f;
#pragma unroll
  for (int i = 0; i < TM; ++i) {
    int64_t gm = m0 + ty * TM + i;
    if (gm >= M) continue;
#pragma unroll
    for (int j = 0; j < TN; ++j) {
      int64_t gn = n0 + tx * TN + j;
      if (gn >= N) continue;
      float v = acc[i][j];
      if (epi.bias) v += (epi.flags & TXL_EPI_BIAS_ROW) ? epi.bias[gm] : epi.bias[gn];
      if (epi.flags & TXL_EPI_RELU) v = fmaxf(v, 0.f);
      if (epi.flags & TXL_EPI_MASK_POS) v = to_f32(((const TC*)epi.aux)[gm * ldc + gn]) > 0.f ? v : 0.f;
      if (epi.flags & TXL_EPI_MASK_LIVE) v = ((epi.live_bits[(gn >> 5) * M + gm] >> (gn & 31)) & 1u) ? v : 0.f;
      if (epi.flags & TXL_EPI_MASK_SCALE) v *= inv_keep;
      if (epi.flags & TXL_EPI_DROPOUT) v *= dropout_scale(epi.seed, epi.site, (uint64_t)(gm * N + gn), epi.drop_p, inv_keep);
      if ((epi.flags & TXL_EPI_EMIT_LIVE) && v > 0.f) atomicOr(&epi.live_bits[(gn >> 5) * M + gm], 1u << (gn & 31));   // plane zeroed by the host wrapper
      cs[j] += v;
      const int64_t ci = (epi.flags & TXL_EPI_TRANSPOSE) ? gn * ldc + gm : gm * ldc + gn;
      if (epi.flags & TXL_EPI_ACCUM) v += to_f32(C[ci]);
      C[ci] = from_f32<TC>(v);
    }
  }
  if (epi.colsum) {
    // reduce the 16 row-groups of this block through shared memory, one atomic per column per block
    __shared__ float red[BN];
    if (tid < BN) red[tid] = 0.f;
    __syncthreads();
#pragma unroll
    for (int j = 0; j < TN; ++j) atomicAdd(&red[tx * TN + j], cs[j]);
    __syncthreads();
    if (tid < BN && n0 + tid < N) atomicAdd(&epi.colsum[n0 + tid], red[tid]);
  }
}
}  // namespace

extern "C" int txl_gemm(const void* A, const void* B, void* C, int64_t M, int64_t N, int64_t K, int64_t lda, int64_t ldb,
                        int64_t ldc, int transA, int transB, int dtype_ab, int dtype_c, const TxlEpilogue* epi_in, void* stream) {
  TXL_CHECK_ARG(A && B && C && M > 0 && N > 0 && K > 0, "gemm: null pointer or empty shape (M=%ld N=%ld K=%ld)", (long)M, (long)N, (long)K);
  TXL_CHECK_ARG(lda >= (transA ? M : K) && ldb >= (transB ? K : N) && ldc >= ((epi_in && (epi_in->flags & TXL_EPI_TRANSPOSE)) ? M : N), "gemm: leading dimension too small");
  TxlEpilogue epi;
  if (epi_in) epi = *epi_in; else { epi.bias = nullptr; epi.aux = nullptr; epi.colsum = nullptr; epi.drop_p = 0.f; epi.seed = 0; epi.site = 0; epi.flags = 0; epi.live_bits = nullptr; }
  if ((epi.flags & TXL_EPI_DROPOUT) && !(epi.drop_p > 0.f)) epi.flags &= ~TXL_EPI_DROPOUT;
  TXL_CHECK_ARG(!(epi.flags & TXL_EPI_MASK_POS) || epi.aux, "gemm: MASK_POS needs aux");
  TXL_CHECK_ARG(!(epi.flags & (TXL_EPI_EMIT_LIVE | TXL_EPI_MASK_LIVE)) || epi.live_bits, "gemm: EMIT_LIVE / MASK_LIVE need live_bits");
  TXL_CHECK_ARG(!((epi.flags & TXL_EPI_EMIT_LIVE) && (epi.flags & (TXL_EPI_MASK_LIVE | TXL_EPI_ACCUM))), "gemm: EMIT_LIVE cannot be combined with MASK_LIVE or ACCUM");
  TXL_CHECK_ARG(!(epi.flags & TXL_EPI_MASK_SCALE) || ((epi.flags & (TXL_EPI_MASK_POS | TXL_EPI_MASK_LIVE)) && epi.drop_p >= 0.f && epi.drop_p < 1.f),
                "gemm: MASK_SCALE needs MASK_POS or MASK_LIVE and 0 <= drop_p < 1");
  if (dtype_ab == TXL_BF16) {
    int handled = 0;
    int rc = txl_gemm_tc(A, B, C, M, N, K, lda, ldb, ldc, transA, transB, dtype_c, &epi, stream, &handled);
    if (rc != TXL_OK) return rc;
    if (handled) return TXL_OK;
  }
  if (epi.flags & TXL_EPI_EMIT_LIVE) TXL_CUDA(cudaMemsetAsync(epi.live_bits, 0, (size_t)cdiv64(N, 32) * M * sizeof(uint32_t), (cudaStream_t)stream));   // set with atomicOr below
  dim3 grid((unsigned)cdiv64(N, BN), (unsigned)cdiv64(M, BM));
  float ik = (epi.flags & (TXL_EPI_DROPOUT | TXL_EPI_MASK_SCALE)) ? 1.f / (1.f - epi.drop_p) : 1.f;
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype_ab == TXL_F32 && dtype_c == TXL_F32)
    gemm_simt_kernel<float, float><<<grid, 256, 0, st>>>((const float*)A, (const float*)B, (float*)C, M, N, K, lda, ldb, ldc, transA, transB, epi, ik);
  else if (dtype_ab == TXL_BF16 && dtype_c == TXL_BF16)
    gemm_simt_kernel<bf16, bf16><<<grid, 256, 0, st>>>((const bf16*)A, (const bf16*)B, (bf16*)C, M, N, K, lda, ldb, ldc, transA, transB, epi, ik);
  else if (dtype_ab == TXL_BF16 && dtype_c == TXL_F32)
    gemm_simt_kernel<bf16, float><<<grid, 256, 0, st>>>((const bf16*)A, (const bf16*)B, (float*)C, M, N, K, lda, ldb, ldc, transA, transB, epi, ik);
  else { txl_set_error("gemm: unsupported dtype combination ab=%d c=%d", dtype_ab, dtype_c); return TXL_EINVAL; }
  TXL_LAUNCH_CHECK();
  return TXL_OK;
}
