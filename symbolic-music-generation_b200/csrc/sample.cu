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

constexpr int SEL_K = 64, SEL_MAX = 128;   // selection path: top_k <= SEL_K, at most SEL_MAX kept entries (ties with the k-th value included)

// One CTA = one sequence.  s = the row of scores (global or shared), sm = NP*12 bytes of dynamic shared memory.  Returns the drawn token in
// every thread.  HF order: Temperature -> TopK (ties with the k-th value kept) -> TopP (first token crossing p kept) -> renormalise -> draw.
// With 0 < top_k <= 64 (and for greedy) the kept set is built by repeated block-wide arg-max in the total order (value desc, index asc) —
// k+1 rounds of two barriers instead of the 66 compare-exchange stages of a 2048-element sort; anything else takes the bitonic sort.
__device__ int sample_block(const float* __restrict__ s, int V, int NP, int do_sample, float temperature, int top_k, float top_p, float u01,
                            uint8_t* __restrict__ keep_row, float* __restrict__ warped_row, float* sm) {
  float* val = sm;                    // [NP] sorted values
  int* idx = (int*)(sm + NP);         // [NP] sorted indices
  float* ex = sm + 2 * NP;            // [NP] exp / cumulative (selection path: taken flags first)
  __shared__ float red[NT / 32];
  __shared__ int s_n1, s_n2, s_pick;
  __shared__ float sel_v[SEL_MAX], wred_v[NT / 32], s_wv;
  __shared__ int sel_i[SEL_MAX], wred_i[NT / 32], s_wi;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const float inv_t = (do_sample && temperature != 1.0f) ? 1.0f / temperature : 1.0f;
  for (int v = tid; v < NP; v += NT) {
    val[v] = v < V ? s[v] * inv_t : -INFINITY;
    idx[v] = v;
    ex[v] = 0.f;
  }
  __syncthreads();
  const int kmax = !do_sample ? 1 : (top_k > 0 ? min(top_k, V) : 0);
  bool selected = false;
  int n_sel = 0;
  if (kmax > 0 && kmax <= SEL_K) {
    float kth = 0.f;
    bool overflow = false;
    for (int r = 0;; ++r) {
      float bv = -INFINITY; int bi = -1;                       // best un-taken element of this thread (it owns v = tid + NT i)
      for (int v = tid; v < V; v += NT)
        if (ex[v] == 0.f) { const float x = val[v]; if (bi < 0 || before(x, v, bv, bi)) { bv = x; bi = v; } }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const float ov = __shfl_xor_sync(0xffffffffu, bv, o); const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (oi >= 0 && (bi < 0 || before(ov, oi, bv, bi))) { bv = ov; bi = oi; }
      }
      if (lane == 0) { wred_v[warp] = bv; wred_i[warp] = bi; }
      __syncthreads();
      if (warp == 0) {
        bv = wred_v[lane]; bi = wred_i[lane];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          const float ov = __shfl_xor_sync(0xffffffffu, bv, o); const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
          if (oi >= 0 && (bi < 0 || before(ov, oi, bv, bi))) { bv = ov; bi = oi; }
        }
        if (lane == 0) { s_wv = bv; s_wi = bi; }
      }
      __syncthreads();
      const float wv = s_wv; const int wi = s_wi;
      if (wi < 0) break;                                       // everything taken
      if (r >= kmax && !(wv == kth)) break;                    // no further tie with the k-th value
      if (r >= SEL_MAX) { overflow = true; break; }
      if (tid == 0) { sel_v[r] = wv; sel_i[r] = wi; }
      if (r == kmax - 1) kth = wv;
      if ((wi % NT) == tid) ex[wi] = 1.f;                      // only the owner reads this flag again: no barrier needed
      n_sel = r + 1;
      if (!do_sample) break;                                   // greedy: the first maximal index is all that is needed
    }
    __syncthreads();
    if (!overflow) {
      for (int t = tid; t < n_sel; t += NT) { val[t] = sel_v[t]; idx[t] = sel_i[t]; }
      selected = true;
    }
    __syncthreads();
  }
  if (!selected) {
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
  }
  if (!do_sample) {  // greedy: first maximal index
    const int pick = idx[0];
    if (keep_row) for (int v = tid; v < V; v += NT) keep_row[v] = (v == pick);
    if (warped_row) for (int v = tid; v < V; v += NT) warped_row[v] = s[v];
    return pick;
  }
  // ---- top-k: keep every value >= the k-th largest
  if (tid == 0) s_n1 = selected ? n_sel : V;
  __syncthreads();
  if (top_k > 0 && !selected) {
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
  if (keep_row) {
    for (int v = tid; v < V; v += NT) keep_row[v] = 0;
  }
  if (warped_row) {
    for (int v = tid; v < V; v += NT) warped_row[v] = -INFINITY;
  }
  __syncthreads();
  for (int t = tid; t < n2; t += NT) {
    if (keep_row) keep_row[idx[t]] = 1;
    if (warped_row) warped_row[idx[t]] = val[t] - lse;
  }
  // ---- inverse-CDF draw in descending-probability order
  const float target = u01 * total2;
  if (tid == 0) s_pick = n2 - 1;
  __syncthreads();
  for (int t = tid; t < n2; t += NT) {
    float prev = t > 0 ? ex[t - 1] : 0.f;
    if (target >= prev && target < ex[t]) s_pick = t;   // exactly one t satisfies this
  }
  __syncthreads();
  return idx[s_pick];
}

__global__ void __launch_bounds__(NT) sample_kernel(const float* __restrict__ scores, int V, int NP, int do_sample, float temperature,
                                                    int top_k, float top_p, const float* __restrict__ u, int64_t* __restrict__ next,
                                                    uint8_t* __restrict__ keep, float* __restrict__ warped) {
  extern __shared__ float sm[];
  const int b = blockIdx.x;
  const int pick = sample_block(scores + (int64_t)b * V, V, NP, do_sample, temperature, top_k, top_p, do_sample ? u[b] : 0.f,
                                keep ? keep + (int64_t)b * V : nullptr, warped ? warped + (int64_t)b * V : nullptr, sm);
  if (threadIdx.x == 0) next[b] = pick;
}

// The tail of a decode step in ONE kernel (one CTA per sequence): log-softmax of the LM-head logits (warp 0, in the exact operation order of
// lsm_nll_fwd_kernel, so the scores are bit-identical to the forward path's), the keyed uniform of decode_uniform_kernel, warpers + draw
// (sample_block), HF's eos / pad bookkeeping (decode_commit_kernel), the embedding row of the chosen token for the next step, and the
// step counter (advanced by the last CTA to arrive).                                                                        [A.6, A.7, A.2]
__global__ void __launch_bounds__(NT) decode_tail_kernel(const float* __restrict__ logits, int64_t ldl, float* __restrict__ scores, int V, int NP,
                                                         int do_sample, float temperature, int top_k, float top_p, uint64_t seed, int64_t seq_offset,
                                                         int64_t* __restrict__ tok, int64_t* __restrict__ unfinished, int64_t* __restrict__ out_ids,
                                                         int64_t ld_out, int col0, int32_t* pos, int* __restrict__ arrive, int64_t eos, int64_t pad,
                                                         int use_eos, const bf16* __restrict__ E, bf16* __restrict__ x0, int d, float emb_scale) {
  extern __shared__ float sm[];
  float* srow = sm + 3 * NP;          // [NP] log-probs of this sequence
  __shared__ float s_lse;
  __shared__ int64_t s_tok;
  __shared__ int s_is_last;
  pdl_launch_dependents();
  pdl_wait();
  const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31;
  const int p = *pos;
  const float* l = logits + (int64_t)b * ldl;
  if (tid < 32) {                     // lsm_nll_fwd_kernel's arithmetic, one warp per row
    float m = -INFINITY;
    for (int v = lane; v < V; v += 32) m = fmaxf(m, l[v]);
    const float mm = warp_max(m);
    float s = 0.f;
    for (int v = lane; v < V; v += 32) s += expf(l[v] - mm);
    const float lse = mm + logf(warp_sum(s));
    if (lane == 0) s_lse = lse;
  }
  __syncthreads();
  const float lse = s_lse;
  for (int v = tid; v < V; v += NT) {
    const float x = l[v] - lse;
    srow[v] = x;
    if (scores) scores[(int64_t)b * V + v] = x;
  }
  __syncthreads();
  float u01 = 0.f;
  if (do_sample) {                    // decode_uniform_kernel's hash: keyed on (seed, global sequence index, step)
    uint64_t x = seed ^ (0x9E3779B97F4A7C15ull * (uint64_t)(seq_offset + b + 1)) ^ (0xC2B2AE3D27D4EB4Full * (uint64_t)(p + 1));
    x ^= x >> 33; x *= 0xff51afd7ed558ccdull; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ull; x ^= x >> 33;
    u01 = (float)(uint32_t)(x >> 40) * (1.0f / 16777216.0f);
  }
  const int pick = sample_block(srow, V, NP, do_sample, temperature, top_k, top_p, u01, nullptr, nullptr, sm);
  if (tid == 0) {
    int64_t t = pick;
    if (use_eos) {
      const int64_t un = unfinished[b];
      t = un ? t : pad;
      unfinished[b] = un && (t != eos);
    }
    tok[b] = t;
    out_ids[(int64_t)b * ld_out + col0 + p] = t;
    s_tok = t;
  }
  __syncthreads();
  const int64_t t = s_tok;
  for (int c = tid; c < d; c += NT)
    x0[(int64_t)b * d + c] = (t >= 0 && t < V) ? __float2bfloat16_rn(__bfloat162float(E[t * d + c]) * emb_scale) : __float2bfloat16_rn(0.f);
  // every CTA has read *pos long before the last one arrives here
  if (tid == 0) {
    __threadfence();
    s_is_last = (atomicAdd(arrive, 1) == (int)gridDim.x - 1);
    if (s_is_last) { *arrive = 0; *pos = p + 1; }
  }
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
  static size_t attr_smem[64] = {0};      // per device: the attribute belongs to the (function, device) pair
  int dev = 0;
  TXL_CUDA(cudaGetDevice(&dev));
  if (smem > 48 * 1024 && smem > attr_smem[dev & 63]) { TXL_CUDA(cudaFuncSetAttribute(sample_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); attr_smem[dev & 63] = smem; }
  sample_kernel<<<B, NT, smem, (cudaStream_t)stream>>>(scores, V, NP, do_sample, temperature, top_k, top_p, u, next, keep, warped);
  TXL_LAUNCH_CHECK();
  return TXL_OK;
}

extern "C" int txl_decode_tail(const float* logits, int64_t ldl, float* scores, int B, int V, int do_sample, float temperature, int top_k, float top_p,
                               uint64_t seed, int64_t seq_offset, int64_t* tok, int64_t* unfinished, int64_t* out_ids, int64_t ld_out, int col0,
                               int32_t* pos, int* arrive, int64_t eos, int64_t pad, int use_eos, const void* E, void* x0, int d, float emb_scale,
                               void* stream) {
  TXL_CHECK_ARG(logits && tok && out_ids && pos && arrive && E && x0 && B > 0 && V > 0 && d > 0 && ldl >= V && (!use_eos || unfinished), "decode_tail: bad args");
  TXL_CHECK_ARG(!do_sample || (temperature > 0.f && top_p > 0.f), "decode_tail: sampling needs temperature>0, top_p>0");
  int NP = 32;
  while (NP < V) NP <<= 1;
  TXL_CHECK_ARG(NP <= 8192, "decode_tail: vocab %d too large for the shared-memory sampler", V);
  const size_t smem = (size_t)NP * 16;
  static size_t attr_smem[64] = {0};
  int dev = 0;
  TXL_CUDA(cudaGetDevice(&dev));
  if (smem > 48 * 1024 && smem > attr_smem[dev & 63]) { TXL_CUDA(cudaFuncSetAttribute(decode_tail_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); attr_smem[dev & 63] = smem; }
  TXL_CUDA(txl_launch_pdl(decode_tail_kernel, dim3(B), dim3(NT), smem, (cudaStream_t)stream, logits, ldl, scores, V, NP, do_sample, temperature, top_k, top_p,
                          seed, seq_offset, tok, unfinished, out_ids, ld_out, col0, pos, arrive, eos, pad, use_eos, (const bf16*)E, (bf16*)x0, d, emb_scale));
  TXL_LAUNCH_CHECK();
  return TXL_OK;
}
