// decode.cu — autoregressive decode step kernels (batched generation, one new token per sequence per step).
//
// HF re-projects K and V from all `mem_len` cached hidden states at every step (~1.3e10 FLOP per token per sequence);
// here generate() keeps a private ring cache of the PROJECTED keys/values, [B, H, mem_len, d_head], so a step only streams the
// cache once with coalesced 16-byte loads.  The API-level `mems` remain hidden states (A.8').  With T = 1 and mlen == mem_len the
// live band is keys 1..mem_len of cat(mems, current): exactly "the ring after the new token has overwritten the oldest slot".
// Relative distance of ring slot s when the current token sits in slot c is (c - s) mod mem_len; the r row is x = mem_len - dist
// (r = HF's r_head_k for klen = mem_len + 1, cached per layer — it depends on weights only).      [A.3-A.5, A.7, A.8']
#include "common.cuh"

namespace {
template <typename T> struct V8 {};
template <typename T> __device__ __forceinline__ void load8(const T* p, float* f) {
  if constexpr (sizeof(T) == 2) {
    uint4 u = *reinterpret_cast<const uint4*>(p);
    const bf16* e = reinterpret_cast<const bf16*>(&u);
#pragma unroll
    for (int k = 0; k < 8; ++k) f[k] = __bfloat162float(e[k]);
  } else {
    float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
    f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w; f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
  }
}

// kv_mem [B*ML, 2*H*dh] (k | v, row = b*ML + j) -> kc, vc [B, H, ML, dh]
template <typename T>
__global__ void cache_init_kernel(const T* __restrict__ kv, int64_t ld, T* __restrict__ kc, T* __restrict__ vc, int B, int H, int ML, int dh) {
  const int64_t total = (int64_t)B * H * ML * dh;
  const int HD = H * dh;
  for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    int c = (int)(idx % dh); int64_t t = idx / dh; int j = (int)(t % ML); t /= ML; int h = (int)(t % H); int b = (int)(t / H);
    const T* row = kv + ((int64_t)b * ML + j) * ld;
    kc[idx] = row[h * dh + c];
    vc[idx] = row[HD + h * dh + c];
  }
}

constexpr int DEC_THREADS = 512;
template <typename T, int DH>
__global__ void __launch_bounds__(DEC_THREADS) decode_attn_kernel(const T* __restrict__ qkv, T* __restrict__ kc, T* __restrict__ vc, const T* __restrict__ r,
                                                                  const float* __restrict__ rwb, const float* __restrict__ rrb, T* __restrict__ out,
                                                                  const int32_t* __restrict__ pos, int H, int ML) {
  extern __shared__ float sm[];
  float* sc = sm;                 // [ML] scores -> probabilities
  float* qw = sm + ML;            // [DH]
  float* qr = qw + DH;            // [DH]
  float* red = qr + DH;           // [DEC_THREADS/32 * DH] partial outputs, also scratch for reductions
  const int h = blockIdx.x, b = blockIdx.y, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int HD = H * DH;
  const int cur = (*pos) % ML;
  T* K = kc + ((int64_t)b * H + h) * ML * DH;
  T* V = vc + ((int64_t)b * H + h) * ML * DH;
  const T* row = qkv + (int64_t)b * 3 * HD;
  if (tid < DH) {
    const float q = to_f32(row[h * DH + tid]);
    qw[tid] = q + rwb[h * DH + tid];
    qr[tid] = q + rrb[h * DH + tid];
    K[(int64_t)cur * DH + tid] = row[HD + h * DH + tid];        // append: overwrite the oldest slot
    V[(int64_t)cur * DH + tid] = row[2 * HD + h * DH + tid];
  }
  __syncthreads();
  // ---- scores: groups of (DH/8) lanes share one key, 16-byte loads along d_head
  constexpr int LPK = DH / 8;                 // lanes per key
  constexpr int KPW = 32 / LPK;               // keys per warp per pass
  const int sub = lane / LPK, ch = lane % LPK;
  const float scale = rsqrtf((float)DH);
  float qwv[8], qrv[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) { qwv[k] = qw[ch * 8 + k]; qrv[k] = qr[ch * 8 + k]; }
  if constexpr (sizeof(T) == 2) {
    // bf16: ONE pass over the ring (flash-decoding): K, R and V rows of a key are requested together, every lane group keeps a running
    // (max, sum, partial output) and the groups are merged at the end — no second sweep that has to wait for the soft-max, so all
    // loads of the step are independent and in flight together.  (fp32 keeps the two-pass form below: bit-for-bit soft-max order.)
    float m = -INFINITY, l = 0.f, acc[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) acc[k] = 0.f;
#pragma unroll 4
    for (int s0 = warp * KPW; s0 < ML; s0 += (DEC_THREADS / 32) * KPW) {
      const int s = s0 + sub;
      const bool ok = s < ML;
      float a = 0.f, vf[8];
      if (ok) {
        int dist = cur - s; if (dist < 0) dist += ML;
        float kf[8], rf[8];
        load8(K + (int64_t)s * DH + ch * 8, kf);
        load8(r + (int64_t)(ML - dist) * HD + h * DH + ch * 8, rf);
        load8(V + (int64_t)s * DH + ch * 8, vf);
#pragma unroll
        for (int k = 0; k < 8; ++k) a = fmaf(qwv[k], kf[k], fmaf(qrv[k], rf[k], a));
      }
#pragma unroll
      for (int o = LPK / 2; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
      if (ok) {
        a *= scale;
        const float mn = fmaxf(m, a), c = __expf(m - mn), p = __expf(a - mn);
        l = l * c + p;
#pragma unroll
        for (int k = 0; k < 8; ++k) acc[k] = fmaf(acc[k], c, p * vf[k]);
        m = mn;
      }
    }
    // merge the KPW key groups of the warp ...
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
    // ... then the warps through shared memory
    float* wm = red + (DEC_THREADS / 32) * DH; float* wl = wm + DEC_THREADS / 32;
    if (sub == 0) {
#pragma unroll
      for (int k = 0; k < 8; ++k) red[warp * DH + ch * 8 + k] = acc[k];
      if (ch == 0) { wm[warp] = m; wl[warp] = l; }
    }
    __syncthreads();
    if (tid < DH) {
      float M = -INFINITY;
      for (int w = 0; w < DEC_THREADS / 32; ++w) M = fmaxf(M, wm[w]);
      float L = 0.f, o = 0.f;
      for (int w = 0; w < DEC_THREADS / 32; ++w) {
        const float c = (wm[w] == -INFINITY) ? 0.f : __expf(wm[w] - M);
        L += wl[w] * c; o += red[w * DH + tid] * c;
      }
      out[(int64_t)b * HD + h * DH + tid] = from_f32<T>(o / L);
    }
    return;
  }
#pragma unroll 4
  for (int s0 = warp * KPW; s0 < ML; s0 += (DEC_THREADS / 32) * KPW) {
    const int s = s0 + sub;
    float a = 0.f;
    if (s < ML) {
      int dist = cur - s; if (dist < 0) dist += ML;
      float kf[8], rf[8];
      load8(K + (int64_t)s * DH + ch * 8, kf);
      load8(r + (int64_t)(ML - dist) * HD + h * DH + ch * 8, rf);
#pragma unroll
      for (int k = 0; k < 8; ++k) a = fmaf(qwv[k], kf[k], fmaf(qrv[k], rf[k], a));
    }
#pragma unroll
    for (int o = LPK / 2; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
    if (ch == 0 && s < ML) sc[s] = a * scale;
  }
  __syncthreads();
  // ---- softmax over the ML ring entries
  float m = -INFINITY;
  for (int s = tid; s < ML; s += DEC_THREADS) m = fmaxf(m, sc[s]);
  m = warp_max(m);
  if (lane == 0) red[warp] = m;
  __syncthreads();
  m = red[0];
  for (int w = 1; w < DEC_THREADS / 32; ++w) m = fmaxf(m, red[w]);
  __syncthreads();
  float sum = 0.f;
  for (int s = tid; s < ML; s += DEC_THREADS) { float p = __expf(sc[s] - m); sc[s] = p; sum += p; }
  sum = warp_sum(sum);
  if (lane == 0) red[warp] = sum;
  __syncthreads();
  sum = 0.f;
  for (int w = 0; w < DEC_THREADS / 32; ++w) sum += red[w];
  __syncthreads();
  // ---- P.V: same lane grouping, each lane accumulates its 8 head dims over the keys of its group
  float acc[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) acc[k] = 0.f;
#pragma unroll 4
  for (int s0 = warp * KPW; s0 < ML; s0 += (DEC_THREADS / 32) * KPW) {
    const int s = s0 + sub;
    if (s < ML) {
      const float p = sc[s];
      float vf[8];
      load8(V + (int64_t)s * DH + ch * 8, vf);
#pragma unroll
      for (int k = 0; k < 8; ++k) acc[k] = fmaf(p, vf[k], acc[k]);
    }
  }
  // reduce over the KPW key groups of the warp, then over warps through shared memory
#pragma unroll
  for (int o = LPK; o < 32; o <<= 1)
#pragma unroll
    for (int k = 0; k < 8; ++k) acc[k] += __shfl_xor_sync(0xffffffffu, acc[k], o);
  if (sub == 0) {
#pragma unroll
    for (int k = 0; k < 8; ++k) red[warp * DH + ch * 8 + k] = acc[k];
  }
  __syncthreads();
  if (tid < DH) {
    float o = 0.f;
    for (int w = 0; w < DEC_THREADS / 32; ++w) o += red[w * DH + tid];
    out[(int64_t)b * HD + h * DH + tid] = from_f32<T>(o / sum);
  }
}

// C[M<=64, N] = A[M,K] . W[N,K]^T (+bias)(ReLU) for the per-step Linears (M = sequences on this GPU).  Weight-streaming:
// a block owns 16 output columns; the activations are staged once per block in shared memory (cp.async, double-buffered 256-wide K
// chunks); warp (rh, cg) computes rows 32*rh.. for the 4 columns of group cg, lanes split K in 16-byte pieces, so every activation
// fetched from shared memory is used for 4 columns and every weight element is read from HBM exactly once; the per-lane partial sums
// are reduced across the warp by butterfly transposes.
constexpr int SK_MAXM = 64, SK_KC = 256, SK_WARPS = 8, SK_CPW = 4, SK_COLS = (SK_WARPS / 2) * SK_CPW;   // 16 columns per block
__device__ __forceinline__ void cp_async16(void* dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

template <typename T>
__global__ void __launch_bounds__(SK_WARPS * 32) skinny_gemm_kernel(const T* __restrict__ A, int64_t lda, const T* __restrict__ W, int64_t ldw,
                                                                    const float* __restrict__ bias, T* __restrict__ C, int64_t ldc, int M, int N, int K,
                                                                    int relu) {
  extern __shared__ __align__(16) unsigned char sk_smem[];
  T* As = reinterpret_cast<T*>(sk_smem);                       // [2][SK_MAXM][SK_KC]
  constexpr int VPR = SK_KC * (int)sizeof(T) / 16;             // 16-byte vectors per staged row
  constexpr int EPV = 16 / (int)sizeof(T);                     // elements per vector
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int rh = warp & 1, cg = warp >> 1;
  const int n0 = blockIdx.x * SK_COLS + cg * SK_CPW;
  const int nchunks = (K + SK_KC - 1) / SK_KC;
  auto stage = [&](int ck, int buf) {
    T* dst = As + (size_t)buf * SK_MAXM * SK_KC;
    for (int e = threadIdx.x; e < SK_MAXM * VPR; e += SK_WARPS * 32) {
      const int m = e / VPR, v = e % VPR, k = ck * SK_KC + v * EPV;
      if (m < M && k < K) cp_async16(dst + m * SK_KC + v * EPV, A + (int64_t)m * lda + k);
    }
    cp_async_commit();
  };
  float acc[SK_CPW][32];
#pragma unroll
  for (int c = 0; c < SK_CPW; ++c)
#pragma unroll
    for (int m = 0; m < 32; ++m) acc[c][m] = 0.f;
  stage(0, 0);
  for (int ck = 0; ck < nchunks; ++ck) {
    if (ck + 1 < nchunks) { stage(ck + 1, (ck + 1) & 1); cp_async_wait<1>(); } else { cp_async_wait<0>(); }
    __syncthreads();
    const int k = ck * SK_KC + lane * 8;
    if (k < K) {
      float w[SK_CPW][8];
#pragma unroll
      for (int c = 0; c < SK_CPW; ++c) {
        if (n0 + c < N) load8(W + (int64_t)(n0 + c) * ldw + k, w[c]);
        else {
#pragma unroll
          for (int j = 0; j < 8; ++j) w[c][j] = 0.f;
        }
      }
      const T* abuf = As + (size_t)(ck & 1) * SK_MAXM * SK_KC + (size_t)(rh * 32) * SK_KC + lane * 8;
#pragma unroll
      for (int m = 0; m < 32; ++m) {
        float a[8];
        load8(abuf + m * SK_KC, a);
#pragma unroll
        for (int c = 0; c < SK_CPW; ++c)
#pragma unroll
          for (int j = 0; j < 8; ++j) acc[c][m] = fmaf(a[j], w[c][j], acc[c][m]);
      }
    }
    __syncthreads();
  }
#pragma unroll
  for (int c = 0; c < SK_CPW; ++c) {
    float* v = acc[c];
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) {
#pragma unroll
      for (int i = 0; i < off; ++i) {
        const bool up = lane & off;
        const float send = up ? v[i] : v[i + off];
        const float keep = up ? v[i + off] : v[i];
        v[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
      }
    }
    const int n = n0 + c, m = rh * 32 + lane;
    if (n < N && m < M) {
      float x = v[0] + (bias ? bias[n] : 0.f);
      if (relu) x = fmaxf(x, 0.f);
      C[(int64_t)m * ldc + n] = from_f32<T>(x);
    }
  }
}

// end-of-step bookkeeping of HF's sample()/greedy_search() loop, on the device: finished rows emit pad, eos finishes a row,
// the token is stored at column col0 + *pos of out_ids and fed back as next input; then the step counter advances.
__global__ void decode_commit_kernel(const int64_t* __restrict__ next, int64_t* __restrict__ tok, int64_t* __restrict__ unfinished, int64_t* __restrict__ out_ids,
                                     int64_t ld_out, int col0, int32_t* pos, int B, int64_t eos, int64_t pad, int use_eos) {
  const int p = *pos;
  __syncthreads();
  for (int b = threadIdx.x; b < B; b += blockDim.x) {
    int64_t t = next[b];
    if (use_eos) {
      const int64_t u = unfinished[b];
      t = u ? t : pad;
      unfinished[b] = u && (t != eos);
    }
    tok[b] = t;
    out_ids[(int64_t)b * ld_out + col0 + p] = t;
  }
  __syncthreads();
  if (threadIdx.x == 0) *pos = p + 1;
}

// u[b] = uniform(0,1) from (seed, global sequence index, step): reproducible across any sharding of the sequences over GPUs
__global__ void decode_uniform_kernel(float* __restrict__ u, int B, uint64_t seed, int64_t seq_offset, const int32_t* __restrict__ pos) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  uint64_t x = seed ^ (0x9E3779B97F4A7C15ull * (uint64_t)(seq_offset + b + 1)) ^ (0xC2B2AE3D27D4EB4Full * (uint64_t)(*pos + 1));
  x ^= x >> 33; x *= 0xff51afd7ed558ccdull; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ull; x ^= x >> 33;
  u[b] = (float)(uint32_t)(x >> 40) * (1.0f / 16777216.0f);
}
}  // namespace

#define DEC_DISPATCH(dtype, ...)                                   \
  if ((dtype) == TXL_F32) { typedef float T; __VA_ARGS__; }        \
  else if ((dtype) == TXL_BF16) { typedef bf16 T; __VA_ARGS__; }   \
  else { txl_set_error("decode: bad dtype"); return TXL_EINVAL; }

extern "C" int txl_decode_cache_init(const void* kv_mem, int64_t ld, void* kc, void* vc, int B, int H, int ML, int dh, int dtype, void* stream) {
  TXL_CHECK_ARG(kv_mem && kc && vc && B > 0 && H > 0 && ML > 0 && dh > 0 && ld >= 2 * H * dh, "decode_cache_init: bad args");
  int grid = (int)imin64(cdiv64((int64_t)B * H * ML * dh, 256), (int64_t)txl_num_sms() * 16);
  TXL_CHECK_ARG(dtype == TXL_F32, "decode_cache_init: fp32 parity mode only (bf16 decodes over the interleaved ring of txl_decode_cache_init_kv)");
  typedef float T;
  cache_init_kernel<T><<<grid, 256, 0, (cudaStream_t)stream>>>((const T*)kv_mem, ld, (T*)kc, (T*)vc, B, H, ML, dh);
  TXL_LAUNCH_CHECK();
  return TXL_OK;
}

extern "C" int txl_decode_attn(const void* qkv, void* kc, void* vc, const void* r, const float* rwb, const float* rrb, void* out, const int32_t* pos,
                               int B, int H, int ML, int dh, int dtype, void* stream) {
  TXL_CHECK_ARG(qkv && kc && vc && r && rwb && rrb && out && pos && B > 0 && H > 0 && ML > 0, "decode_attn: bad args");
  TXL_CHECK_ARG(dh == 32 || dh == 64 || dh == 128, "decode_attn: d_head %d not in {32,64,128}", dh);
  dim3 grid(H, B);
  size_t smem = ((size_t)ML + 2 * dh + (DEC_THREADS / 32) * dh + 2 * (DEC_THREADS / 32)) * sizeof(float);
  cudaStream_t st = (cudaStream_t)stream;
#define DEC_LAUNCH(DHV)                                                                                                                   \
  {                                                                                                                                       \
    static size_t attr_smem = 0;                                                                                                          \
    if (smem > 48 * 1024 && smem > attr_smem) { TXL_CUDA(cudaFuncSetAttribute(decode_attn_kernel<T, DHV>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); attr_smem = smem; } \
    decode_attn_kernel<T, DHV><<<grid, DEC_THREADS, smem, st>>>((const T*)qkv, (T*)kc, (T*)vc, (const T*)r, rwb, rrb, (T*)out, pos, H, ML);  \
  }
  TXL_CHECK_ARG(dtype == TXL_F32, "decode_attn: fp32 parity mode only (bf16: txl_decode_attn_pipe)");
  typedef float T;
  if (dh == 32) DEC_LAUNCH(32) else if (dh == 64) DEC_LAUNCH(64) else DEC_LAUNCH(128)
#undef DEC_LAUNCH
  TXL_LAUNCH_CHECK();
  return TXL_OK;
}

extern "C" int txl_skinny_gemm(const void* A, int64_t lda, const void* W, int64_t ldw, const float* bias, void* C, int64_t ldc, int M, int N, int K,
                               int relu, int dtype, void* stream) {
  TXL_CHECK_ARG(A && W && C && M > 0 && M <= SK_MAXM && N > 0 && K > 0 && K % 8 == 0 && lda % 8 == 0 && ldw % 8 == 0, "skinny_gemm: needs M<=64, K,lda,ldw multiples of 8");
  TXL_CHECK_ARG(((uintptr_t)A & 15) == 0 && ((uintptr_t)W & 15) == 0, "skinny_gemm: 16-byte alignment");
  dim3 grid((unsigned)cdiv64(N, SK_COLS));
  DEC_DISPATCH(dtype, {
    const size_t smem = 2 * sizeof(T) * SK_MAXM * SK_KC;
    static bool attr_set = false;
    if (smem > 48 * 1024 && !attr_set) { TXL_CUDA(cudaFuncSetAttribute(skinny_gemm_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); attr_set = true; }
    skinny_gemm_kernel<T><<<grid, SK_WARPS * 32, smem, (cudaStream_t)stream>>>((const T*)A, lda, (const T*)W, ldw, bias, (T*)C, ldc, M, N, K, relu);
  });
  TXL_LAUNCH_CHECK();
  return TXL_OK;
}

extern "C" int txl_decode_commit(const int64_t* next, int64_t* tok, int64_t* unfinished, int64_t* out_ids, int64_t ld_out, int col0, int32_t* pos, int B,
                                 int64_t eos, int64_t pad, int use_eos, void* stream) {
  TXL_CHECK_ARG(next && tok && out_ids && pos && B > 0 && (!use_eos || unfinished), "decode_commit: bad args");
  decode_commit_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(next, tok, unfinished, out_ids, ld_out, col0, pos, B, eos, pad, use_eos);
  TXL_LAUNCH_CHECK();
  return TXL_OK;
}

extern "C" int txl_decode_uniform(float* u, int B, uint64_t seed, int64_t seq_offset, const int32_t* pos, void* stream) {
  TXL_CHECK_ARG(u && pos && B > 0, "decode_uniform: bad args");
  decode_uniform_kernel<<<(B + 127) / 128, 128, 0, (cudaStream_t)stream>>>(u, B, seed, seq_offset, pos);
  TXL_LAUNCH_CHECK();
  return TXL_OK;
}
