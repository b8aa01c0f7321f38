// decode_fused.cu — the whole transformer stack of one decode step as ONE persistent cooperative kernel.
//
// A decode step at 64 sequences is ~90 dependent micro-ops (5 Linears, attention and 2 LayerNorms per layer); launched separately
// each costs ~10 us of fixed latency against ~1 us of work.  Here one CTA per SM stays resident for the whole step and the stages
// are separated by grid-wide barriers (cooperative launch):
//   per layer:  qkv = x Wqkv^T | ring append + band attention over the projected-K/V cache | ao = vec Wo^T | y1 = LN(x + ao) |
//               h = y1 W1^T | f = relu(h + b1) W2^T | x' = LN(y1 + f + b2)
//   then the LM-head GEMM; log-softmax, sampling and bookkeeping stay separate (tiny) kernels.
// Every Linear is split over (16 output columns) x (256-wide K slices) work items so that all SMs stream weights; each item writes
// an fp32 partial [ks][B][N] and the consumer stage sums the K slices while staging its operand (deterministic, no atomics), applying
// bias / ReLU on the fly.  Weights are read from HBM exactly once per step, with 16-byte loads.   [A.2, A.3-A.5 at T=1, A.6, A.7]
#include "common.cuh"
#include <cooperative_groups.h>
namespace cg = cooperative_groups;

namespace {
constexpr int FD_THREADS = 256, FD_MAXB = 64, FD_KC = 256, FD_COLS = 16, FD_MAXKS = 8;

template <typename T>
struct FusedLayer {
  const T *wqkv, *wr_unused, *wo, *w1, *w2, *r;      // r: [ML+1, d] cached r_head_k table
  const float *b1, *b2, *rwb, *rrb, *ln1w, *ln1b, *ln2w, *ln2b;
  T *kc, *vc;                                         // [B, H, ML, dh] rings
};
template <typename T>
struct FusedArgs {
  const FusedLayer<T>* layers;
  const T* E;                 // [V, d]
  const int64_t* tok;         // [B]
  const int32_t* pos;
  T *x, *y1, *vec;            // [B, d] activations in the compute dtype
  float *part;                // fp32 partial sums, max over stages of ks * B * N
  float* logits;              // [B, Vp] fp32, written without bias (bias added by the log-softmax stage... see host)
  const float* out_bias;
  int B, H, dh, d, di, ML, L, V, Vp;
  float eps, emb_scale;
  unsigned long long* tstamp;   // optional [1 + 7 L + 2] globaltimer stamps taken by block 0 after every stage (profiling only)
};

template <typename T> __device__ __forceinline__ void ldv8(const T* p, float* f) {
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

// source of a GEMM operand A[B, K]: either a tensor in the compute dtype, or the sum of `nks` fp32 partials (+bias)(ReLU)
template <typename T>
struct ASrc {
  const T* t;            // if non-null: A = t[b*K + k]
  const float* part;     // else: A = f(sum_s part[(s*B + b)*K + k])
  int nks;
  const float* bias;
  int relu;
};
template <typename T>
__device__ __forceinline__ float a_elem(const ASrc<T>& s, int B, int K, int b, int k) {
  if (s.t) return to_f32(s.t[(int64_t)b * K + k]);
  float v = 0.f;
  for (int j = 0; j < s.nks; ++j) v += s.part[((int64_t)j * B + b) * K + k];
  if (s.bias) v += s.bias[k];
  if (s.relu) v = fmaxf(v, 0.f);
  return v;
}

// out_part[ks][b][n] = sum_{k in slice ks} A[b,k] W[n,k]   for all (column group, K slice) items, grid-strided over CTAs
template <typename T>
__device__ void gemm_stage(const ASrc<T>& A, const T* __restrict__ W, float* __restrict__ out_part, int B, int N, int K, T* As /* smem [FD_MAXB][FD_KC] */) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int rh = warp & 1, cgp = warp >> 1;                 // rows 32*rh.., columns 4*cgp.. of the 16-column item
  const int ncg = (N + FD_COLS - 1) / FD_COLS, nks = (K + FD_KC - 1) / FD_KC;
  for (int item = blockIdx.x; item < ncg * nks; item += gridDim.x) {
    const int cgi = item % ncg, ks = item / ncg;
    const int k0 = ks * FD_KC;
    __syncthreads();
    // stage A[:, k0 : k0+256] into shared memory, 8 elements per thread per step, all loads of a step issued before use
    constexpr int VPR = FD_KC / 8;                       // 8-element vectors per row
#pragma unroll 4
    for (int e = threadIdx.x; e < FD_MAXB * VPR; e += FD_THREADS) {
      const int b = e / VPR, k = k0 + (e % VPR) * 8;
      float f[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) f[j] = 0.f;
      if (b < B && k < K) {
        if (A.t) {
          ldv8(A.t + (int64_t)b * K + k, f);
        } else {
          for (int sidx = 0; sidx < A.nks; ++sidx) {
            const float4 p0 = *reinterpret_cast<const float4*>(A.part + ((int64_t)sidx * B + b) * K + k);
            const float4 p1 = *reinterpret_cast<const float4*>(A.part + ((int64_t)sidx * B + b) * K + k + 4);
            f[0] += p0.x; f[1] += p0.y; f[2] += p0.z; f[3] += p0.w; f[4] += p1.x; f[5] += p1.y; f[6] += p1.z; f[7] += p1.w;
          }
          if (A.bias) {
#pragma unroll
            for (int j = 0; j < 8; ++j) f[j] += A.bias[k + j];
          }
          if (A.relu) {
#pragma unroll
            for (int j = 0; j < 8; ++j) f[j] = fmaxf(f[j], 0.f);
          }
        }
      }
      T* dst = As + (size_t)b * FD_KC + (e % VPR) * 8;
#pragma unroll
      for (int j = 0; j < 8; ++j) dst[j] = from_f32<T>(f[j]);
    }
    __syncthreads();
    const int n0 = cgi * FD_COLS + cgp * 4;
    const int k = k0 + lane * 8;
    float acc[4][32];
#pragma unroll
    for (int c = 0; c < 4; ++c)
#pragma unroll
      for (int m = 0; m < 32; ++m) acc[c][m] = 0.f;
    if (k < K) {
      float w[4][8];
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        if (n0 + c < N) ldv8(W + (int64_t)(n0 + c) * K + k, w[c]);
        else {
#pragma unroll
          for (int j = 0; j < 8; ++j) w[c][j] = 0.f;
        }
      }
      const T* ab = As + (size_t)(rh * 32) * FD_KC + lane * 8;
#pragma unroll
      for (int m = 0; m < 32; ++m) {
        float a[8];
        ldv8(ab + m * FD_KC, a);
#pragma unroll
        for (int c = 0; c < 4; ++c)
#pragma unroll
          for (int j = 0; j < 8; ++j) acc[c][m] = fmaf(a[j], w[c][j], acc[c][m]);
      }
    }
#pragma unroll
    for (int c = 0; c < 4; ++c) {
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
      if (n < N && m < B) out_part[((int64_t)ks * B + m) * N + n] = v[0];
    }
  }
}

// y[b,:] = LN(resid[b,:] + sum_s part[s][b,:] (+bias)) * gamma + beta, one warp per row, the row held in registers (d <= 1024)
template <typename T>
__device__ void ln_stage(const T* __restrict__ resid, const float* __restrict__ part, int nks, const float* __restrict__ bias, const float* __restrict__ gamma,
                         const float* __restrict__ beta, T* __restrict__ y, int B, int d, float eps) {
  // rows are spread over CTAs first (row b -> CTA b, warp 0): a CTA-local burst of 8 rows would leave 140 CTAs spinning on the barrier
  const int lane = threadIdx.x & 31, gw = (threadIdx.x >> 5) * gridDim.x + blockIdx.x, nw = gridDim.x * (FD_THREADS / 32);
  constexpr int MAXV = 4;                                  // 8-element vectors per lane
  for (int b = gw; b < B; b += nw) {
    float v[MAXV][8];
    float s = 0.f;
#pragma unroll
    for (int e = 0; e < MAXV; ++e) {
      const int c = (e * 32 + lane) * 8;
      if (c < d) {
        ldv8(resid + (int64_t)b * d + c, v[e]);
        for (int j = 0; j < nks; ++j) {
          const float4 p0 = *reinterpret_cast<const float4*>(part + ((int64_t)j * B + b) * d + c);
          const float4 p1 = *reinterpret_cast<const float4*>(part + ((int64_t)j * B + b) * d + c + 4);
          v[e][0] += p0.x; v[e][1] += p0.y; v[e][2] += p0.z; v[e][3] += p0.w; v[e][4] += p1.x; v[e][5] += p1.y; v[e][6] += p1.z; v[e][7] += p1.w;
        }
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          if (bias) v[e][k] += bias[c + k];
          v[e][k] = to_f32(from_f32<T>(v[e][k]));
          s += v[e][k];
        }
      }
    }
    const float mu = warp_sum(s) / d;
    float q = 0.f;
#pragma unroll
    for (int e = 0; e < MAXV; ++e) {
      const int c = (e * 32 + lane) * 8;
      if (c < d) {
#pragma unroll
        for (int k = 0; k < 8; ++k) { const float t = v[e][k] - mu; q += t * t; }
      }
    }
    const float rs = rsqrtf(warp_sum(q) / d + eps);
#pragma unroll
    for (int e = 0; e < MAXV; ++e) {
      const int c = (e * 32 + lane) * 8;
      if (c < d) {
#pragma unroll
        for (int k = 0; k < 8; ++k) y[(int64_t)b * d + c + k] = from_f32<T>((v[e][k] - mu) * rs * gamma[c + k] + beta[c + k]);
      }
    }
  }
}

// ring append + single-query band attention for every (b, h), reading q,k,v as sums of the qkv partials
template <typename T, int DH>
__device__ void attn_stage(const FusedArgs<T>& a, const FusedLayer<T>& Lw, const float* __restrict__ qkv_part, int nks, float* sm) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int H = a.H, ML = a.ML, HD = H * DH, B = a.B;
  float* sc = sm;
  float* qw = sm + ML;
  float* qr = qw + DH;
  float* red = qr + DH;
  const int cur = (*a.pos) % ML;
  constexpr int LPK = DH / 8, KPW = 32 / LPK;
  const int sub = lane / LPK, ch = lane % LPK;
  const float scale = rsqrtf((float)DH);
  for (int item = blockIdx.x; item < B * H; item += gridDim.x) {
    const int b = item / H, h = item % H;
    T* K = Lw.kc + ((int64_t)b * H + h) * ML * DH;
    T* V = Lw.vc + ((int64_t)b * H + h) * ML * DH;
    __syncthreads();
    if (tid < DH) {
      float q = 0.f, k = 0.f, v = 0.f;
      for (int j = 0; j < nks; ++j) {
        const float* row = qkv_part + ((int64_t)j * B + b) * 3 * HD;
        q += row[h * DH + tid]; k += row[HD + h * DH + tid]; v += row[2 * HD + h * DH + tid];
      }
      q = to_f32(from_f32<T>(q));                       // the unfused path stores qkv in the compute dtype
      qw[tid] = q + Lw.rwb[h * DH + tid];
      qr[tid] = q + Lw.rrb[h * DH + tid];
      K[(int64_t)cur * DH + tid] = from_f32<T>(k);
      V[(int64_t)cur * DH + tid] = from_f32<T>(v);
    }
    __syncthreads();
    float qwv[8], qrv[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) { qwv[k] = qw[ch * 8 + k]; qrv[k] = qr[ch * 8 + k]; }
#pragma unroll 4
    for (int s0 = warp * KPW; s0 < ML; s0 += (FD_THREADS / 32) * KPW) {
      const int s = s0 + sub;
      float acc = 0.f;
      if (s < ML) {
        int dist = cur - s; if (dist < 0) dist += ML;
        float kf[8], rf[8];
        ldv8(K + (int64_t)s * DH + ch * 8, kf);
        ldv8(Lw.r + (int64_t)(ML - dist) * HD + h * DH + ch * 8, rf);
#pragma unroll
        for (int k = 0; k < 8; ++k) acc = fmaf(qwv[k], kf[k], fmaf(qrv[k], rf[k], acc));
      }
#pragma unroll
      for (int o = LPK / 2; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
      if (ch == 0 && s < ML) sc[s] = acc * scale;
    }
    __syncthreads();
    float m = -INFINITY;
    for (int s = tid; s < ML; s += FD_THREADS) m = fmaxf(m, sc[s]);
    m = warp_max(m);
    if (lane == 0) red[warp] = m;
    __syncthreads();
    m = red[0];
    for (int w = 1; w < FD_THREADS / 32; ++w) m = fmaxf(m, red[w]);
    __syncthreads();
    float sum = 0.f;
    for (int s = tid; s < ML; s += FD_THREADS) { float p = __expf(sc[s] - m); sc[s] = p; sum += p; }
    sum = warp_sum(sum);
    if (lane == 0) red[warp] = sum;
    __syncthreads();
    sum = 0.f;
    for (int w = 0; w < FD_THREADS / 32; ++w) sum += red[w];
    __syncthreads();
    float acc8[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) acc8[k] = 0.f;
#pragma unroll 4
    for (int s0 = warp * KPW; s0 < ML; s0 += (FD_THREADS / 32) * KPW) {
      const int s = s0 + sub;
      if (s < ML) {
        const float p = sc[s];
        float vf[8];
        ldv8(V + (int64_t)s * DH + ch * 8, vf);
#pragma unroll
        for (int k = 0; k < 8; ++k) acc8[k] = fmaf(p, vf[k], acc8[k]);
      }
    }
#pragma unroll
    for (int o = LPK; o < 32; o <<= 1)
#pragma unroll
      for (int k = 0; k < 8; ++k) acc8[k] += __shfl_xor_sync(0xffffffffu, acc8[k], o);
    if (sub == 0) {
#pragma unroll
      for (int k = 0; k < 8; ++k) red[warp * DH + ch * 8 + k] = acc8[k];
    }
    __syncthreads();
    if (tid < DH) {
      float o = 0.f;
      for (int w = 0; w < FD_THREADS / 32; ++w) o += red[w * DH + tid];
      a.vec[(int64_t)b * HD + h * DH + tid] = from_f32<T>(o / sum);
    }
  }
}

template <typename T, int DH>
__global__ void __launch_bounds__(FD_THREADS, 1) decode_fused_kernel(const FusedArgs<T> a) {
  cg::grid_group grid = cg::this_grid();
  extern __shared__ __align__(16) unsigned char fd_smem[];
  T* As = reinterpret_cast<T*>(fd_smem);
  float* fsm = reinterpret_cast<float*>(fd_smem);
  const int B = a.B, d = a.d, di = a.di;
  const int ks_d = (d + FD_KC - 1) / FD_KC, ks_di = (di + FD_KC - 1) / FD_KC;
  int ts_i = 0;
  auto stamp = [&]() {
    if (a.tstamp && blockIdx.x == 0 && threadIdx.x == 0) { unsigned long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); a.tstamp[ts_i] = t; }
    ++ts_i;
  };
  stamp();
  // ---- embedding
  for (int e = blockIdx.x * FD_THREADS + threadIdx.x; e < B * d; e += gridDim.x * FD_THREADS) {
    const int b = e / d, c = e % d;
    const int64_t id = a.tok[b];
    a.x[e] = from_f32<T>((id >= 0 && id < a.V) ? to_f32(a.E[id * d + c]) * a.emb_scale : 0.f);
  }
  grid.sync();
  for (int l = 0; l < a.L; ++l) {
    const FusedLayer<T>& Lw = a.layers[l];
    ASrc<T> sx{a.x, nullptr, 0, nullptr, 0};
    gemm_stage<T>(sx, Lw.wqkv, a.part, B, 3 * d, d, As);                       // qkv partials [ks_d][B][3d]
    grid.sync();
    stamp();
    attn_stage<T, DH>(a, Lw, a.part, ks_d, fsm);                               // -> vec
    grid.sync();
    stamp();
    ASrc<T> sv{a.vec, nullptr, 0, nullptr, 0};
    gemm_stage<T>(sv, Lw.wo, a.part, B, d, d, As);                             // ao partials [ks_d][B][d]
    grid.sync();
    stamp();
    ln_stage<T>(a.x, a.part, ks_d, nullptr, Lw.ln1w, Lw.ln1b, a.y1, B, d, a.eps);
    grid.sync();
    stamp();
    ASrc<T> sy{a.y1, nullptr, 0, nullptr, 0};
    gemm_stage<T>(sy, Lw.w1, a.part, B, di, d, As);                            // h partials [ks_d][B][di]
    grid.sync();
    stamp();
    // FF2 reads h = relu(sum partials + b1) while staging; its own partials go behind the h partials
    float* fpart = a.part + (size_t)ks_d * B * di;
    ASrc<T> sh{nullptr, a.part, ks_d, Lw.b1, 1};
    gemm_stage<T>(sh, Lw.w2, fpart, B, d, di, As);                             // f partials [ks_di][B][d]
    grid.sync();
    stamp();
    ln_stage<T>(a.y1, fpart, ks_di, Lw.b2, Lw.ln2w, Lw.ln2b, a.x, B, d, a.eps);
    grid.sync();
    stamp();
  }
  // ---- LM head: logits partials, summed (+bias) into a.logits by a final pass
  ASrc<T> sx{a.x, nullptr, 0, nullptr, 0};
  gemm_stage<T>(sx, a.E, a.part, B, a.V, d, As);
  grid.sync();
  stamp();
  if (a.tstamp) { grid.sync(); stamp(); }     // calibration: an empty stage = the cost of one grid barrier
  for (int e = blockIdx.x * FD_THREADS + threadIdx.x; e < B * a.V; e += gridDim.x * FD_THREADS) {
    const int b = e / a.V, v = e % a.V;
    float s = a.out_bias[v];
    for (int j = 0; j < ks_d; ++j) s += a.part[((int64_t)j * B + b) * a.V + v];
    a.logits[(int64_t)b * a.Vp + v] = s;
  }
}

template <typename T>
int launch_fused(const FusedArgs<T>& a, cudaStream_t st) {
  const size_t smem_gemm = sizeof(T) * FD_MAXB * FD_KC;
  const size_t smem_attn = sizeof(float) * ((size_t)a.ML + 2 * a.dh + (FD_THREADS / 32) * a.dh);
  const size_t smem = smem_gemm > smem_attn ? smem_gemm : smem_attn;
  void* fn = nullptr;
  if (a.dh == 32) fn = (void*)decode_fused_kernel<T, 32>;
  else if (a.dh == 64) fn = (void*)decode_fused_kernel<T, 64>;
  else fn = (void*)decode_fused_kernel<T, 128>;
  static size_t attr[2][3] = {{0, 0, 0}, {0, 0, 0}};
  size_t& cur = attr[sizeof(T) == 2][a.dh == 32 ? 0 : (a.dh == 64 ? 1 : 2)];
  if (smem > 48 * 1024 && smem > cur) { TXL_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); cur = smem; }
  int per_sm = 0;
  TXL_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fn, FD_THREADS, smem));
  TXL_CHECK_ARG(per_sm >= 1, "decode_fused: kernel does not fit on an SM");
  const int grid = txl_num_sms();          // one resident CTA per SM: the grid barrier needs co-residency
  void* args[] = {(void*)&a};
  TXL_CUDA(cudaLaunchCooperativeKernel(fn, dim3(grid), dim3(FD_THREADS), args, smem, st));
  ++g_txl_launches;
  return TXL_OK;
}
}  // namespace

static unsigned long long* g_fd_tstamp = nullptr;
/* profiling hook: device buffer of >= 7 L + 4 uint64 that receives %globaltimer stamps after every stage (NULL disables) */
extern "C" int txl_decode_fused_set_timestamps(unsigned long long* dev_buf) { g_fd_tstamp = dev_buf; return TXL_OK; }

// C-ABI: pointers arrive in flat arrays (one entry per layer) so that no struct layout crosses the language boundary.
extern "C" int64_t txl_decode_fused_workspace(int B, int d, int di, int V, int L, int dtype) {
  const int64_t ks_d = (d + FD_KC - 1) / FD_KC, ks_di = (di + FD_KC - 1) / FD_KC;
  int64_t part = ks_d * B * (int64_t)di + ks_di * B * (int64_t)d;
  const int64_t alt1 = ks_d * B * 3ll * d, alt2 = ks_d * B * (int64_t)V;
  if (alt1 > part) part = alt1;
  if (alt2 > part) part = alt2;
  part = (part + 63) / 64 * 64;
  const int64_t layer_bytes = (dtype == TXL_F32 ? sizeof(FusedLayer<float>) : sizeof(FusedLayer<bf16>)) * (int64_t)L;
  const int64_t esz = dtype == TXL_F32 ? 4 : 2;
  return part * 4 + 3 * (((int64_t)B * d * esz + 255) / 256 * 256) + layer_bytes + 4096;
}

extern "C" int txl_decode_fused_step(const void* const* wqkv, const void* const* wo, const void* const* w1, const void* const* w2, const void* const* rtab,
                                     const float* const* b1, const float* const* b2, const float* const* rwb, const float* const* rrb,
                                     const float* const* ln1w, const float* const* ln1b, const float* const* ln2w, const float* const* ln2b,
                                     void* const* kc, void* const* vc, const void* E, const float* out_bias, const int64_t* tok, const int32_t* pos,
                                     float* logits, void* ws, int build_layer_table, int B, int H, int dh, int d, int di, int ML, int L, int V, int Vp,
                                     float eps, int dtype, void* stream) {
  TXL_CHECK_ARG(B > 0 && B <= FD_MAXB && L > 0 && d % 8 == 0 && di % 8 == 0 && H * dh == d && d <= 1024, "decode_fused: needs B<=64, d<=1024, d and d_inner multiples of 8");
  TXL_CHECK_ARG(dh == 32 || dh == 64 || dh == 128, "decode_fused: d_head %d not in {32,64,128}", dh);
  TXL_CHECK_ARG(E && out_bias && tok && pos && logits && ws, "decode_fused: null pointer");
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t ks_d = (d + FD_KC - 1) / FD_KC, ks_di = (di + FD_KC - 1) / FD_KC;
  int64_t part = ks_d * B * (int64_t)di + ks_di * B * (int64_t)d;
  const int64_t alt1 = ks_d * B * 3ll * d, alt2 = ks_d * B * (int64_t)V;
  if (alt1 > part) part = alt1;
  if (alt2 > part) part = alt2;
  part = (part + 63) / 64 * 64;          // keeps the activation buffers behind the partials 256-byte aligned
#define FD_RUN(TT)                                                                                                         \
  {                                                                                                                        \
    char* p = (char*)ws;                                                                                                   \
    float* partp = (float*)p; p += part * 4;                                                                               \
    const int64_t act = ((int64_t)B * d * sizeof(TT) + 255) / 256 * 256;                                                   \
    TT* x = (TT*)p; p += act;                                                                                              \
    TT* y1 = (TT*)p; p += act;                                                                                             \
    TT* vec = (TT*)p; p += act;                                                                                            \
    p = (char*)(((uintptr_t)p + 255) & ~(uintptr_t)255);                                                                   \
    FusedLayer<TT>* table = (FusedLayer<TT>*)p;                                                                            \
    if (build_layer_table) {                                                                                               \
      TXL_CHECK_ARG(L <= 64, "decode_fused: more than 64 layers");                                                         \
      FusedLayer<TT> host[64];                                                                                             \
      for (int l = 0; l < L; ++l) {                                                                                        \
        host[l].wqkv = (const TT*)wqkv[l]; host[l].wr_unused = nullptr; host[l].wo = (const TT*)wo[l]; host[l].w1 = (const TT*)w1[l];   \
        host[l].w2 = (const TT*)w2[l]; host[l].r = (const TT*)rtab[l]; host[l].b1 = b1[l]; host[l].b2 = b2[l]; host[l].rwb = rwb[l];     \
        host[l].rrb = rrb[l]; host[l].ln1w = ln1w[l]; host[l].ln1b = ln1b[l]; host[l].ln2w = ln2w[l]; host[l].ln2b = ln2b[l];             \
        host[l].kc = (TT*)kc[l]; host[l].vc = (TT*)vc[l];                                                                  \
      }                                                                                                                    \
      TXL_CUDA(cudaMemcpyAsync(table, host, sizeof(FusedLayer<TT>) * L, cudaMemcpyHostToDevice, st));                      \
      TXL_CUDA(cudaStreamSynchronize(st));                                                                                 \
      return TXL_OK;                                                                                                       \
    }                                                                                                                      \
    FusedArgs<TT> a;                                                                                                       \
    a.layers = table; a.E = (const TT*)E; a.tok = tok; a.pos = pos; a.x = x; a.y1 = y1; a.vec = vec; a.part = partp; a.logits = logits;   \
    a.out_bias = out_bias; a.B = B; a.H = H; a.dh = dh; a.d = d; a.di = di; a.ML = ML; a.L = L; a.V = V; a.Vp = Vp; a.eps = eps;         \
    a.emb_scale = sqrtf((float)d); a.tstamp = g_fd_tstamp;                                                                 \
    return launch_fused<TT>(a, st);                                                                                        \
  }
  if (dtype == TXL_F32) FD_RUN(float) else if (dtype == TXL_BF16) FD_RUN(bf16) else { txl_set_error("decode_fused: bad dtype"); return TXL_EINVAL; }
#undef FD_RUN
}
