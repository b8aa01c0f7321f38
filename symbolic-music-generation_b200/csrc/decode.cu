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

constexpr int DEC_THREADS = 256;
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

// C[M<=64, N] = A[M,K] . W[N,K]^T (+bias)(ReLU) for the per-step Linears (M = sequences on this GPU).  Weight-streaming: the block's
// 8 warps cover CPB output columns x KSPLIT slices of K, every weight element is read once with 16-byte loads, the activations come
// straight from L1/L2 (all warps of a block, and all blocks, read the same 64 rows); the 64 per-row partial sums of a lane are
// reduced across the warp by a butterfly transpose (62 shuffles) and across K slices through shared memory.
constexpr int SK_MAXM = 64, SK_WARPS = 8;
template <typename T>
__global__ void __launch_bounds__(SK_WARPS * 32) skinny_gemm_kernel(const T* __restrict__ A, int64_t lda, const T* __restrict__ W, int64_t ldw,
                                                                    const float* __restrict__ bias, T* __restrict__ C, int64_t ldc, int M, int N, int K,
                                                                    int relu, int ksplit) {
  __shared__ float part[SK_WARPS][SK_MAXM];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int cpb = SK_WARPS / ksplit;
  const int col_in_blk = warp / ksplit, ks = warp % ksplit;
  const int n = blockIdx.x * cpb + col_in_blk;
  const int kper = ((K + ksplit - 1) / ksplit + 7) / 8 * 8;
  const int kbeg = ks * kper, kend = min(K, kbeg + kper);
  float acc[SK_MAXM];
#pragma unroll
  for (int m = 0; m < SK_MAXM; ++m) acc[m] = 0.f;
  if (n < N) {
    for (int k0 = kbeg + lane * 8; k0 < kend; k0 += 256) {
      float w[8];
      load8(W + (int64_t)n * ldw + k0, w);
#pragma unroll
      for (int m = 0; m < SK_MAXM; ++m) {
        if (m < M) {
          float a[8];
          load8(A + (int64_t)m * lda + k0, a);
#pragma unroll
          for (int k = 0; k < 8; ++k) acc[m] = fmaf(a[k], w[k], acc[m]);
        }
      }
    }
  }
  // transpose-reduce: lane l ends with the totals of rows l and 32 + l
#pragma unroll
  for (int half = 0; half < 2; ++half) {
    float* v = acc + 32 * half;
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
  }
  part[warp][lane] = acc[0];
  part[warp][32 + lane] = acc[32];
  __syncthreads();
  if (ks == 0 && n < N) {
    const float bb = bias ? bias[n] : 0.f;
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      const int m = 32 * half + lane;
      if (m < M) {
        float x = bb;
        for (int j = 0; j < ksplit; ++j) x += part[warp + j][m];
        if (relu) x = fmaxf(x, 0.f);
        C[(int64_t)m * ldc + n] = from_f32<T>(x);
      }
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
  DEC_DISPATCH(dtype, (cache_init_kernel<T><<<grid, 256, 0, (cudaStream_t)stream>>>((const T*)kv_mem, ld, (T*)kc, (T*)vc, B, H, ML, dh)));
  TXL_LAUNCH_CHECK();
  return TXL_OK;
}

extern "C" int txl_decode_attn(const void* qkv, void* kc, void* vc, const void* r, const float* rwb, const float* rrb, void* out, const int32_t* pos,
                               int B, int H, int ML, int dh, int dtype, void* stream) {
  TXL_CHECK_ARG(qkv && kc && vc && r && rwb && rrb && out && pos && B > 0 && H > 0 && ML > 0, "decode_attn: bad args");
  TXL_CHECK_ARG(dh == 32 || dh == 64 || dh == 128, "decode_attn: d_head %d not in {32,64,128}", dh);
  dim3 grid(H, B);
  size_t smem = ((size_t)ML + 2 * dh + (DEC_THREADS / 32) * dh) * sizeof(float);
  cudaStream_t st = (cudaStream_t)stream;
#define DEC_LAUNCH(DHV)                                                                                                                   \
  {                                                                                                                                       \
    static size_t attr_smem = 0;                                                                                                          \
    if (smem > 48 * 1024 && smem > attr_smem) { TXL_CUDA(cudaFuncSetAttribute(decode_attn_kernel<T, DHV>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); attr_smem = smem; } \
    decode_attn_kernel<T, DHV><<<grid, DEC_THREADS, smem, st>>>((const T*)qkv, (T*)kc, (T*)vc, (const T*)r, rwb, rrb, (T*)out, pos, H, ML);  \
  }
  DEC_DISPATCH(dtype, { if (dh == 32) DEC_LAUNCH(32) else if (dh == 64) DEC_LAUNCH(64) else DEC_LAUNCH(128) });
#undef DEC_LAUNCH
  TXL_LAUNCH_CHECK();
  return TXL_OK;
}

extern "C" int txl_skinny_gemm(const void* A, int64_t lda, const void* W, int64_t ldw, const float* bias, void* C, int64_t ldc, int M, int N, int K,
                               int relu, int dtype, void* stream) {
  TXL_CHECK_ARG(A && W && C && M > 0 && M <= SK_MAXM && N > 0 && K > 0 && K % 8 == 0 && lda % 8 == 0 && ldw % 8 == 0, "skinny_gemm: needs M<=64, K,lda,ldw multiples of 8");
  TXL_CHECK_ARG(((uintptr_t)A & 15) == 0 && ((uintptr_t)W & 15) == 0, "skinny_gemm: 16-byte alignment");
  int ksplit = 1;
  while (ksplit < SK_WARPS && K / (ksplit * 2) >= 256) ksplit *= 2;     // each warp gets >= 256 of K (one 16-byte piece per lane)
  const int cpb = SK_WARPS / ksplit;
  dim3 grid((unsigned)cdiv64(N, cpb));
  DEC_DISPATCH(dtype, (skinny_gemm_kernel<T><<<grid, SK_WARPS * 32, 0, (cudaStream_t)stream>>>((const T*)A, lda, (const T*)W, ldw, bias, (T*)C, ldc, M, N, K, relu, ksplit)));
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
