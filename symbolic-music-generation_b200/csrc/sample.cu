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

// ------------------------------------------------------------------ large vocabularies (row does not fit the shared-memory sort)
// Same warpers, no sort: the k-th largest score and the top-p boundary are found by radix selection over the order-preserving 32-bit
// keys of the row (4 passes of 8 bits each, 256-bin histograms of counts and probability mass); ties at the boundary are kept in index
// order, which is the total order of the small-vocabulary path (value descending, index ascending).  Mass is summed as 40-bit fixed point
// (integer adds commute: the kept set and the draw are reproducible whatever order the atomics land in).  The inverse-CDF draw walks the
// kept tokens in INDEX order (any fixed order samples the same distribution; the small path walks them by descending probability).
// WordPiece 32k / the 267,735-token default TransfoXL vocabulary: reference musicnlp/models/transformer_xl.py:56-63.        [A.7]
constexpr float MASS_SCALE = 1099511627776.f;      // 2^40
__device__ __forceinline__ uint32_t fkey(float x) { const uint32_t u = __float_as_uint(x); return (u & 0x80000000u) ? ~u : (u | 0x80000000u); }
__device__ __forceinline__ unsigned long long mass_of(float x, float m) { return __float2ull_rn(__expf(x - m) * MASS_SCALE); }

struct BigShared {
  unsigned long long hmass[256];
  int hcnt[256];
  unsigned long long wmass[NT / 32];
  int wcnt[NT / 32];
  uint32_t prefix;
  int found, base_cnt, bin_cnt;
  unsigned long long base_mass, bin_mass;
  float red[NT / 32];
  int pick;
};

// Radix descent to the element where the running (count | mass) of the row, taken in descending key order, first reaches the target:
//   by_mass == 0: the element of rank `k_target` (1-based)            -> sh.prefix = its key
//   by_mass == 1: the first element whose inclusive mass exceeds `p_target` -> sh.prefix = its key; sh.found = 0 when the whole row stays below
// On return base_cnt / base_mass describe the elements with a LARGER key, bin_cnt / bin_mass those with exactly this key.
// Only elements with key >= min_key and a finite score take part.
__device__ void radix_select(const float* __restrict__ s, int V, float inv_t, float m, uint32_t min_key, int by_mass, int k_target,
                             unsigned long long p_target, BigShared& sh) {
  const int tid = threadIdx.x;
  if (tid == 0) { sh.prefix = 0; sh.found = 1; sh.base_cnt = 0; sh.base_mass = 0; }
  __syncthreads();
  for (int pass = 0; pass < 4; ++pass) {
    const int shift = 24 - 8 * pass;
    const uint32_t hi_mask = pass == 0 ? 0u : (0xFFFFFFFFu << (shift + 8));
    if (tid < 256) { sh.hmass[tid] = 0; sh.hcnt[tid] = 0; }
    __syncthreads();
    const uint32_t prefix = sh.prefix;
    for (int v = tid; v < V; v += NT) {
      const float x = s[v] * inv_t;
      const uint32_t key = fkey(x);
      if (key < min_key || !(x > -INFINITY) || (key & hi_mask) != prefix) continue;
      const int bin = (key >> shift) & 255;
      atomicAdd(&sh.hcnt[bin], 1);
      atomicAdd(&sh.hmass[bin], mass_of(x, m));
    }
    __syncthreads();
    if (tid == 0 && sh.found) {
      int cnt = sh.base_cnt; unsigned long long mass = sh.base_mass;
      int bin = 255;
      bool hit = false;
      for (; bin >= 0; --bin) {
        const int c = sh.hcnt[bin]; const unsigned long long w = sh.hmass[bin];
        if (c > 0 && (by_mass ? (mass + w > p_target) : (cnt + c >= k_target))) { hit = true; break; }
        cnt += c; mass += w;
      }
      if (!hit) { sh.found = 0; }
      else {
        sh.prefix = prefix | ((uint32_t)bin << shift);
        sh.base_cnt = cnt; sh.base_mass = mass; sh.bin_cnt = sh.hcnt[bin]; sh.bin_mass = sh.hmass[bin];
      }
    }
    __syncthreads();
    if (!sh.found) break;
  }
}

// exclusive prefix over the block of one (count, mass) pair per thread; totals in tot_cnt / tot_mass
__device__ void block_scan_pair(int cnt, unsigned long long mass, int& ex_cnt, unsigned long long& ex_mass, int& tot_cnt, unsigned long long& tot_mass, BigShared& sh) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  int c = cnt; unsigned long long w = mass;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int oc = __shfl_up_sync(0xffffffffu, c, o); const unsigned long long ow = __shfl_up_sync(0xffffffffu, w, o);
    if (lane >= o) { c += oc; w += ow; }
  }
  __syncthreads();
  if (lane == 31) { sh.wcnt[warp] = c; sh.wmass[warp] = w; }
  __syncthreads();
  int bc = 0; unsigned long long bw = 0;
  tot_cnt = 0; tot_mass = 0;
  for (int i = 0; i < NT / 32; ++i) {
    if (i < warp) { bc += sh.wcnt[i]; bw += sh.wmass[i]; }
    tot_cnt += sh.wcnt[i]; tot_mass += sh.wmass[i];
  }
  ex_cnt = bc + c - cnt; ex_mass = bw + w - mass;
}

__global__ void __launch_bounds__(NT) sample_large_kernel(const float* __restrict__ scores, int V, int do_sample, float temperature, int top_k, float top_p,
                                                          const float* __restrict__ u, int64_t* __restrict__ next, uint8_t* __restrict__ keep,
                                                          float* __restrict__ warped) {
  __shared__ BigShared sh;
  const int b = blockIdx.x, tid = threadIdx.x;
  const float* s = scores + (int64_t)b * V;
  uint8_t* keep_row = keep ? keep + (int64_t)b * V : nullptr;
  float* warped_row = warped ? warped + (int64_t)b * V : nullptr;
  const float inv_t = (do_sample && temperature != 1.0f) ? 1.0f / temperature : 1.0f;
  // ---- row maximum and its first index (greedy answer; softmax shift)
  float bv = -INFINITY; int bi = -1;
  for (int v = tid; v < V; v += NT) { const float x = s[v] * inv_t; if (bi < 0 || before(x, v, bv, bi)) { bv = x; bi = v; } }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ov = __shfl_xor_sync(0xffffffffu, bv, o); const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
    if (oi >= 0 && (bi < 0 || before(ov, oi, bv, bi))) { bv = ov; bi = oi; }
  }
  if ((tid & 31) == 0) { sh.red[tid >> 5] = bv; sh.wcnt[tid >> 5] = bi; }
  __syncthreads();
  if (tid == 0) {
    for (int w = 1; w < NT / 32; ++w) { const float ov = sh.red[w]; const int oi = sh.wcnt[w]; if (oi >= 0 && (bi < 0 || before(ov, oi, bv, bi))) { bv = ov; bi = oi; } }
    sh.red[0] = bv; sh.pick = bi;
  }
  __syncthreads();
  const float m = sh.red[0];
  const int arg = sh.pick;
  __syncthreads();
  if (!do_sample) {
    if (keep_row) for (int v = tid; v < V; v += NT) keep_row[v] = (v == arg);
    if (warped_row) for (int v = tid; v < V; v += NT) warped_row[v] = s[v];
    if (tid == 0) next[b] = arg;
    return;
  }
  // ---- top-k: everything >= the k-th largest score stays (ties kept, as HF's `scores < kth` removal)
  uint32_t min_key = 0;
  if (top_k > 0 && top_k < V) {
    radix_select(s, V, inv_t, m, 0u, 0, top_k, 0ull, sh);
    if (sh.found) min_key = sh.prefix;
    __syncthreads();
  }
  // ---- mass of the top-k survivors (finite scores only), this thread's share over a CONTIGUOUS index range (index-order scans below)
  const int chunk = (V + NT - 1) / NT, v0 = min(tid * chunk, V), v1 = min(v0 + chunk, V);
  unsigned long long my_mass = 0; int my_cnt = 0;
  for (int v = v0; v < v1; ++v) {
    const float x = s[v] * inv_t;
    if (fkey(x) >= min_key && x > -INFINITY) { my_mass += mass_of(x, m); ++my_cnt; }
  }
  int ex_c, tot_c; unsigned long long ex_m, total1;
  block_scan_pair(my_cnt, my_mass, ex_c, ex_m, tot_c, total1, sh);
  __syncthreads();
  // ---- top-p: tokens in descending order stay while the mass BEFORE them is <= top_p (the first one past the boundary is kept)
  uint32_t tau = min_key; int ties_kept = 0x7fffffff; bool bounded = false;
  if (top_p < 1.0f) {
    const unsigned long long P = __double2ull_rd((double)top_p * (double)total1);
    radix_select(s, V, inv_t, m, min_key, 1, 0, P, sh);
    if (sh.found) {
      bounded = true;
      tau = sh.prefix;
      const unsigned long long each = sh.bin_cnt > 0 ? sh.bin_mass / (unsigned long long)sh.bin_cnt : 0ull;      // all ties carry the same mass
      const unsigned long long room = P - sh.base_mass;                                                          // >= 0: the elements above tau stayed at or below P
      ties_kept = each > 0 ? (int)min((unsigned long long)sh.bin_cnt, room / each + 1ull) : sh.bin_cnt;
    }
    __syncthreads();
  }
  // ---- kept set in index order: score above tau, or one of the first `ties_kept` scores equal to tau
  int my_ties = 0;
  if (bounded) for (int v = v0; v < v1; ++v) { const float x = s[v] * inv_t; if (fkey(x) == tau && x > -INFINITY) ++my_ties; }
  int tie_rank0, tot_t; unsigned long long dummy0, dummy1;
  block_scan_pair(my_ties, 0ull, tie_rank0, dummy0, tot_t, dummy1, sh);
  __syncthreads();
  auto kept = [&](float x, int& tie_rank) -> bool {
    if (!(x > -INFINITY)) return false;
    const uint32_t key = fkey(x);
    if (!bounded) return key >= min_key;
    if (key > tau) return true;
    if (key == tau) return tie_rank++ < ties_kept;
    return false;
  };
  unsigned long long kept_mass = 0; int kept_cnt = 0;
  { int tr = tie_rank0; for (int v = v0; v < v1; ++v) { const float x = s[v] * inv_t; if (kept(x, tr)) { kept_mass += mass_of(x, m); ++kept_cnt; } } }
  int ex_kc, tot_kc; unsigned long long ex_km, total2;
  block_scan_pair(kept_cnt, kept_mass, ex_kc, ex_km, tot_kc, total2, sh);
  __syncthreads();
  if (total2 == 0) {      // nothing representable survived (cannot happen with a finite maximum: its mass is 2^40): fall back to the arg-max
    if (tid == 0) next[b] = arg;
    return;
  }
  const float lse = m + __logf((float)((double)total2 / (double)MASS_SCALE));
  if (keep_row || warped_row) {
    int tr = tie_rank0;
    for (int v = v0; v < v1; ++v) {
      const float x = s[v] * inv_t;
      const bool k = kept(x, tr);
      if (keep_row) keep_row[v] = k;
      if (warped_row) warped_row[v] = k ? x - lse : -INFINITY;
    }
  }
  // ---- inverse-CDF draw over the kept tokens in index order
  const unsigned long long target = min(__double2ull_rd((double)u[b] * (double)total2), total2 - 1);
  if (tid == 0) sh.pick = -1;
  __syncthreads();
  if (kept_cnt > 0 && target >= ex_km && target < ex_km + kept_mass) {
    unsigned long long acc = ex_km; int tr = tie_rank0; int choice = -1;
    for (int v = v0; v < v1 && choice < 0; ++v) {
      const float x = s[v] * inv_t;
      if (kept(x, tr)) { acc += mass_of(x, m); if (target < acc) choice = v; }
    }
    sh.pick = choice;
  }
  __syncthreads();
  if (tid == 0) next[b] = sh.pick >= 0 ? sh.pick : arg;
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
  if (NP > 16384) {      // the row does not fit the shared-memory sort: radix-selection sampler
    sample_large_kernel<<<B, NT, 0, (cudaStream_t)stream>>>(scores, V, do_sample, temperature, top_k, top_p, u, next, keep, warped);
    TXL_LAUNCH_CHECK();
    return TXL_OK;
  }
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
