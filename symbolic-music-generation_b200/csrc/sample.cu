// sample.cu — logits warpers + draw for batched generation: Temperature -> TopK (ties kept) -> TopP (first token
// crossing p kept) -> renormalise -> inverse-CDF draw / argmax.  One CTA per sequence, the whole (small) music
// vocabulary lives in shared memory; sort, scans and reductions are warp-shuffle based.  [A.7; eval.py:277-333]
#include "common.cuh"

namespace {
constexpr int NT = 1024;

__device__ __forceinline__ bool before(float va, int ia, float vb, int ib) {  // total order: value desc, index asc
  return va > vb || (va == vb && ia < ib);
}

__device__ float block_sum(float v, float* red) {
  v = warp_sum(v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  float t = 0.f;
  for (int w = 0; w < NT / 32; ++w) t += red[w];
  return t;
}

__global__ void __launch_bounds__(NT) sample_kernel(const float* __restrict__ scores, int V, int NP, int do_sample, float temperature,
                                                    int top_k, float top_p, const float* __restrict__ u, int64_t* __restrict__ next,
                                                    uint8_t* __restrict__ keep, float* __restrict__ warped) {
  extern __shared__ float sm[];
  float* val = sm;                    // [NP] sorted values
  int* idx = (int*)(sm + NP);         // [NP] sorted indices
  float* ex = sm + 2 * NP;            // [NP] exp / cumulative
  __shared__ float red[NT / 32];
  __shared__ int s_n1, s_n2, s_pick;
  __shared__ float s_total;
  const int b = blockIdx.x, tid = threadIdx.x;
  const float* s = scores + (int64_t)b * V;
  const float inv_t = (do_sample && temperature != 1.0f) ? 1.0f / temperature : 1.0f;
  for (int v = tid; v < NP; v += NT) {
    val[v] = v < V ? s[v] * inv_t : -INFINITY;
    idx[v] = v;
  }
  __syncthreads();
  // bitonic sort, descending under `before`.  Compare-exchange c of a stage pairs t = (c with a 0 inserted at bit log2 j) and t | j; with
  // c = tid (+ NT i) a warp's 32 exchanges of every stage with j <= 32 stay inside one 64-element block that only this warp touches, so
  // those stages (51 of the 66 at 2048 elements) need __syncwarp() only; block-wide barriers remain around the j >= 64 stages.
  for (int k = 2; k <= NP; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int c = tid; c < NP / 2; c += NT) {
        const int t = ((c & ~(j - 1)) << 1) | (c & (j - 1)), p = t | j;
        const bool up = (t & k) == 0;
        const float va = val[t], vb = val[p]; const int ia = idx[t], ib = idx[p];
        const bool swap = up ? before(vb, ib, va, ia) : before(va, ia, vb, ib);
        if (swap) { val[t] = vb; val[p] = va; idx[t] = ib; idx[p] = ia; }
      }
      const int next_j = j > 1 ? (j >> 1) : k;      // first stride of the next merge size is k
      if (j >= 64 || next_j >= 64) __syncthreads(); else __syncwarp();
    }
  }
  __syncthreads();
  if (!do_sample) {  // greedy: first maximal index
    if (tid == 0) next[b] = idx[0];
    if (keep) for (int v = tid; v < V; v += NT) keep[(int64_t)b * V + v] = (v == idx[0]);
    if (warped) for (int v = tid; v < V; v += NT) warped[(int64_t)b * V + v] = s[v];
    return;
  }
  // ---- top-k: keep every value >= the k-th largest
  if (tid == 0) s_n1 = V;
  __syncthreads();
  if (top_k > 0) {
    int k = min(max(top_k, 1), V);
    float kth = val[k - 1];
    int cnt = 0;
    for (int t = tid; t < V; t += NT) cnt += (val[t] >= kth) ? 1 : 0;
    float c = block_sum((float)cnt, red);
    if (tid == 0) s_n1 = (int)c;
    __syncthreads();
  }
  // -inf scores are never kept
  {
    int cnt = 0;
    for (int t = tid; t < s_n1; t += NT) cnt += (val[t] > -INFINITY) ? 1 : 0;
    float c = block_sum((float)cnt, red);
    __syncthreads();
    if (tid == 0) s_n1 = max((int)c, 1);
    __syncthreads();
  }
  const int n1 = s_n1;
  const float m = val[0];
  float part = 0.f;
  for (int t = tid; t < n1; t += NT) { float e = __expf(val[t] - m); ex[t] = e; part += e; }
  float total1 = block_sum(part, red);
  // ---- inclusive scan of ex[0..n1) by warp 0 (chunks of 32 with carry)
  if (tid < 32) {
    float carry = 0.f;
    for (int base = 0; base < n1; base += 32) {
      int t = base + tid;
      float x = t < n1 ? ex[t] : 0.f;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) { float y = __shfl_up_sync(0xffffffffu, x, o); if (tid >= o) x += y; }
      x += carry;
      if (t < n1) ex[t] = x;   // cumulative un-normalised mass
      carry = __shfl_sync(0xffffffffu, x, 31);
    }
  }
  __syncthreads();
  // ---- top-p: position t kept iff t == 0 or cum[t-1]/total <= top_p
  if (tid == 0) s_n2 = n1;
  __syncthreads();
  if (top_p < 1.0f) {
    int cnt = 0;
    for (int t = tid; t < n1; t += NT) cnt += (ex[t] / total1 <= top_p) ? 1 : 0;
    float c = block_sum((float)cnt, red);
    if (tid == 0) s_n2 = min(n1, 1 + (int)c);
    __syncthreads();
  }
  const int n2 = s_n2;
  const float total2 = ex[n2 - 1];
  const float lse = m + __logf(total2);
  if (keep) {
    for (int v = tid; v < V; v += NT) keep[(int64_t)b * V + v] = 0;
  }
  if (warped) {
    for (int v = tid; v < V; v += NT) warped[(int64_t)b * V + v] = -INFINITY;
  }
  __syncthreads();
  for (int t = tid; t < n2; t += NT) {
    if (keep) keep[(int64_t)b * V + idx[t]] = 1;
    if (warped) warped[(int64_t)b * V + idx[t]] = val[t] - lse;
  }
  // ---- inverse-CDF draw in descending-probability order
  const float target = u[b] * total2;
  if (tid == 0) s_pick = n2 - 1;
  __syncthreads();
  for (int t = tid; t < n2; t += NT) {
    float prev = t > 0 ? ex[t - 1] : 0.f;
    if (target >= prev && target < ex[t]) s_pick = t;   // exactly one t satisfies this
  }
  __syncthreads();
  if (tid == 0) next[b] = idx[s_pick];
  (void)s_total;
}
}  // namespace

extern "C" int txl_sample(const float* scores, int B, int V, int do_sample, float temperature, int top_k, float top_p,
                          const float* u, int64_t* next, uint8_t* keep, float* warped, void* stream) {
  TXL_CHECK_ARG(scores && next && B > 0 && V > 0, "sample: bad args");
  TXL_CHECK_ARG(!do_sample || (u && temperature > 0.f && top_p > 0.f), "sample: sampling needs u, temperature>0, top_p>0");
  int NP = 32;
  while (NP < V) NP <<= 1;
  TXL_CHECK_ARG(NP <= 16384, "sample: vocab %d too large for the shared-memory sampler", V);
  size_t smem = (size_t)NP * 12;
  static size_t attr_smem = 0;
  if (smem > 48 * 1024 && smem > attr_smem) { TXL_CUDA(cudaFuncSetAttribute(sample_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); attr_smem = smem; }
  sample_kernel<<<B, NT, smem, (cudaStream_t)stream>>>(scores, V, NP, do_sample, temperature, top_k, top_p, u, next, keep, warped);
  TXL_LAUNCH_CHECK();
  return TXL_OK;
}
