// umma_probe.cu — micro-benchmarks behind the attention-kernel design decisions (profiles/README.md):
//   (1) cycles per tcgen05.mma (M=128, K=16, bf16) for the N and operand-major combinations the attention kernels issue, and for three
//       ways of issuing it (divergent single thread / converged warp + lane predicate / converged warp + elect.sync),
//   (2) tcgen05.ld throughput with 4 / 8 / 16 warps reading TMEM,
//   (3) cost of st.shared + fence.proxy.async + mbarrier.arrive by every thread vs. named barrier + one fence.
// Build (from the repo root):
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -Iinclude -Isymbolic-music-generation_b200/csrc profiles/umma_probe.cu -o profiles/umma_probe.bin
// Run on the GPU box: ./profiles/umma_probe.bin
#include "tc_common.cuh"
#include <cstdio>
#include <vector>

int g_txl_launches_dummy;
namespace {
constexpr int SMEM = 200 * 1024;

struct MmaCfg { int N, a_mn, b_mn, nk; const char* name; };

__device__ __forceinline__ void umma_elect(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p, q;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "elect.sync _|q, 0xffffffff;\n\t"
      "@q tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// chains of nk accumulating MMAs (the k-steps of one tile), issued by a converged warp through elect.sync; shape and operand layouts are
// RUN-TIME arguments here, so the descriptors are rebuilt per MMA: the table shows what that scalar work costs (94 cycles per issue for
// the small shapes) next to the tight `issue` variants below (48)
__global__ void __launch_bounds__(128, 1) mma_probe(int N, int a_mn, int b_mn, int nk, int reps, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  uint8_t* sm = smem_raw + ((1024u - (raw & 1023u)) & 1023u);
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  for (int e = threadIdx.x; e < 160 * 1024 / 16; e += blockDim.x) reinterpret_cast<uint4*>(sm)[e] = make_uint4(0, 0, 0, 0);
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
  if (threadIdx.x < 32) tmem_alloc<512>(&slot);
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = slot;
  if (threadIdx.x < 32) {
    const uint32_t idesc = umma_idesc_bf16(128, N, a_mn, b_mn);
    const uint32_t a0 = smem_u32(sm), b0 = smem_u32(sm) + 64 * 1024;
    const long long t0 = clock64();
    for (int r = 0; r < reps; ++r) {
      for (int k = 0; k < nk; ++k) {
        // K-major: +32 B per 16-element k-step inside a 64-wide atom; MN-major: +2048 B per 16 k-rows
        const uint64_t ad = a_mn ? umma_smem_desc(a0 + k * 2048, 16384, 1024) : umma_smem_desc(a0 + (k >> 2) * 16384 + (k & 3) * 32, 16, 1024);
        const uint64_t bd = b_mn ? umma_smem_desc(b0 + k * 2048, 8192, 1024) : umma_smem_desc(b0 + (k >> 2) * 24576 + (k & 3) * 32, 16, 1024);
        umma_elect(tm + (r & 1) * 256, ad, bd, idesc, k > 0);
      }
    }
    if (threadIdx.x == 0) umma_commit(&bar);
    mbar_wait(&bar, 0);
    const long long t1 = clock64();
    if (blockIdx.x == 0 && threadIdx.x == 0) out[0] = t1 - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc<512>(tm);
}

// the N = 64 chain, unrolled with compile-time shape, issued from divergent `if (threadIdx.x == 0)` code
__global__ void __launch_bounds__(128, 1) solo_issue_probe(int reps, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  uint8_t* sm = smem_raw + ((1024u - (raw & 1023u)) & 1023u);
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  for (int e = threadIdx.x; e < 160 * 1024 / 16; e += blockDim.x) reinterpret_cast<uint4*>(sm)[e] = make_uint4(0, 0, 0, 0);
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
  if (threadIdx.x < 32) tmem_alloc<512>(&slot);
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = slot;
  if (threadIdx.x == 0) {
    constexpr uint32_t idesc = umma_idesc_bf16(128, 64, 0, 0);
    const uint32_t a0 = smem_u32(sm), b0 = smem_u32(sm) + 64 * 1024;
    const long long t0 = clock64();
    for (int r = 0; r < reps; ++r) {
#pragma unroll
      for (int k = 0; k < 4; ++k) umma_bf16(tm + (r & 1) * 256, umma_smem_desc(a0 + k * 32, 16, 1024), umma_smem_desc(b0 + k * 32, 16, 1024), idesc, k > 0);
    }
    umma_commit(&bar);
    mbar_wait(&bar, 0);
    const long long t1 = clock64();
    if (blockIdx.x == 0) out[0] = t1 - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc<512>(tm);
}

// issue styles: the same N=64 K-major MMA chain issued (1) by a whole converged warp with the instruction predicated on lane 0,
// (2) predicated on elect.sync; descriptors are warp-uniform so they can live in uniform registers
__device__ __forceinline__ void umma_bf16_pred(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate, uint32_t issue) {
  asm volatile(
      "{\n\t.reg .pred p, q;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "setp.ne.b32 q, %5, 0;\n\t"
      "@q tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate), "r"(issue)
      : "memory");
}
__device__ __forceinline__ void umma_bf16_elect(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p, q;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "elect.sync _|q, 0xffffffff;\n\t"
      "@q tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
template <int STYLE, int N>
__global__ void __launch_bounds__(128, 1) issue_probe(int reps, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  uint8_t* sm = smem_raw + ((1024u - (raw & 1023u)) & 1023u);
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  for (int e = threadIdx.x; e < 160 * 1024 / 16; e += blockDim.x) reinterpret_cast<uint4*>(sm)[e] = make_uint4(0, 0, 0, 0);
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
  if (threadIdx.x < 32) tmem_alloc<512>(&slot);
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = slot;
  if (threadIdx.x < 32) {
    const uint32_t lane0 = (threadIdx.x == 0);
    constexpr uint32_t idesc = umma_idesc_bf16(128, N, 0, 0);
    const uint32_t a0 = smem_u32(sm), b0 = smem_u32(sm) + 64 * 1024;
    const long long t0 = clock64();
    for (int r = 0; r < reps; ++r) {
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const uint64_t ad = umma_smem_desc(a0 + k * 32, 16, 1024), bd = umma_smem_desc(b0 + k * 32, 16, 1024);
        if (STYLE == 1) umma_bf16_pred(tm + (r & 1) * 256, ad, bd, idesc, k > 0, lane0);
        else umma_bf16_elect(tm + (r & 1) * 256, ad, bd, idesc, k > 0);
      }
    }
    if (lane0) umma_commit(&bar);
    mbar_wait(&bar, 0);
    const long long t1 = clock64();
    if (blockIdx.x == 0 && lane0) out[0] = t1 - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc<512>(tm);
}

template <int X32>
__global__ void __launch_bounds__(512, 1) tmem_ld_probe(int reps, long long* out, float* sink) {
  __shared__ uint32_t slot;
  if (threadIdx.x < 32) tmem_alloc<512>(&slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = slot;
  const int warp = threadIdx.x >> 5;
  const uint32_t lane_base = (uint32_t)(32 * (warp & 3)) << 16;
  float acc = 0.f;
  __syncthreads();
  const long long t0 = clock64();
  for (int r = 0; r < reps; ++r) {
    float v[32];
    if (X32) tmem_ld_32x32(tm + lane_base + ((r * 32) & 255), v);
    else { tmem_ld_32x16(tm + lane_base + ((r * 32) & 255), v); tmem_ld_32x16(tm + lane_base + ((r * 32 + 16) & 255), v + 16); }
    tmem_ld_wait();
#pragma unroll
    for (int c = 0; c < 32; ++c) acc += v[c];
  }
  __syncthreads();
  const long long t1 = clock64();
  if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = t1 - t0;
  if (acc == 123.456f) sink[0] = acc;
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc<512>(tm);
}

// publish pattern: every thread stores 64 B to smem, then (mode 0) fence.proxy.async + arrive per thread, (mode 1) syncwarp, lane 0 fence + arrive,
// (mode 2) bar.sync + one thread fence + arrive.  One waiter thread (the last warp) consumes the barrier each round.
__global__ void __launch_bounds__(544, 1) publish_probe(int mode, int reps, long long* out) {
  __shared__ uint64_t bar, back;
  __shared__ __align__(16) uint8_t tile[512 * 64];
  const int tid = threadIdx.x, lane = tid & 31;
  if (tid == 0) { mbar_init(&bar, mode == 0 ? 512 : (mode == 1 ? 16 : 1)); mbar_init(&back, 1); fence_barrier_init(); }
  __syncthreads();
  const long long t0 = clock64();
  if (tid >= 512) {
    if (lane == 0)
      for (int r = 0; r < reps; ++r) { mbar_wait(&bar, r & 1); mbar_arrive(&back); }
  } else {
    for (int r = 0; r < reps; ++r) {
      uint4 v = make_uint4(r, tid, r, tid);
#pragma unroll
      for (int c = 0; c < 4; ++c) *reinterpret_cast<uint4*>(tile + tid * 64 + ((c ^ (tid & 3)) << 4)) = v;
      if (mode == 0) { fence_proxy_async_smem(); mbar_arrive(&bar); }
      else if (mode == 1) { fence_proxy_async_smem(); __syncwarp(); if (lane == 0) mbar_arrive(&bar); }
      else { asm volatile("bar.sync 1, 512;" ::: "memory"); if (tid == 0) { fence_proxy_async_smem(); mbar_arrive(&bar); } }
      mbar_wait(&back, r & 1);
    }
  }
  __syncthreads();
  if (tid == 0 && blockIdx.x == 0) out[0] = clock64() - t0;
}
}  // namespace

int main() {
  long long* d_out; float* d_sink;
  cudaMalloc(&d_out, 64); cudaMalloc(&d_sink, 64);
  cudaFuncSetAttribute(mma_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM);
  const MmaCfg cfgs[] = {
      {64, 0, 0, 4, "N=64  A K-major  B K-major  (S = Qw.K^T, dP = dO.V^T)"}, {64, 0, 1, 4, "N=64  A K-major  B MN-major (O += P.V, dQw += dS.K)"},
      {64, 0, 1, 12, "N=64  A K-major  B MN-major, 12 k-steps (dQr += dBD0.Rwin)"}, {64, 1, 1, 8, "N=64  A MN-major B MN-major (dK += dS^T.Qw, dV += P^T.dO, dR)"},
      {192, 0, 0, 4, "N=192 A K-major  B K-major  (BD0 = Qr.Rwin^T)"}, {128, 0, 0, 4, "N=128 A K-major  B K-major"}, {256, 0, 0, 4, "N=256 A K-major  B K-major (GEMM tile)"},
      {128, 0, 1, 4, "N=128 A K-major  B MN-major"}, {128, 1, 1, 8, "N=128 A MN-major B MN-major"}, {32, 0, 0, 4, "N=32  A K-major  B K-major"}, {16, 0, 0, 4, "N=16  A K-major  B K-major"}};
  const int reps = 2000;
  for (const auto& c : cfgs) {
    for (int grid : {148}) {
      mma_probe<<<grid, 128, SMEM>>>(c.N, c.a_mn, c.b_mn, c.nk, reps, d_out);
      long long cyc = 0;
      cudaError_t e = cudaDeviceSynchronize();
      cudaMemcpy(&cyc, d_out, 8, cudaMemcpyDeviceToHost);
      printf("mma  %-66s grid %3d: %7.1f cycles per MMA (K=16)   [%s]\n", c.name, grid, (double)cyc / (reps * c.nk), cudaGetErrorString(e));
    }
  }
  {
    auto run = [&](const char* name, auto kern) {
      cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM);
      kern<<<148, 128, SMEM>>>(reps, d_out);
      long long cyc = 0;
      cudaError_t e = cudaDeviceSynchronize();
      cudaMemcpy(&cyc, d_out, 8, cudaMemcpyDeviceToHost);
      printf("issue %-60s: %7.1f cycles per MMA   [%s]\n", name, (double)cyc / (reps * 4), cudaGetErrorString(e));
    };
    run("divergent single thread (if (threadIdx.x == 0)), N=64", solo_issue_probe);
    run("converged warp, lane-0 predicate, N=64", issue_probe<1, 64>);
    run("converged warp, elect.sync predicate, N=64", issue_probe<2, 64>);
    run("converged warp, lane-0 predicate, N=128", issue_probe<1, 128>);
    run("converged warp, lane-0 predicate, N=256", issue_probe<1, 256>);
    run("converged warp, lane-0 predicate, N=16", issue_probe<1, 16>);
  }
  for (int threads : {128, 256, 512}) {
    for (int x32 = 0; x32 < 2; ++x32) {
      if (x32) tmem_ld_probe<1><<<148, threads>>>(4000, d_out, d_sink); else tmem_ld_probe<0><<<148, threads>>>(4000, d_out, d_sink);
      long long cyc = 0;
      cudaError_t e = cudaDeviceSynchronize();
      cudaMemcpy(&cyc, d_out, 8, cudaMemcpyDeviceToHost);
      const double bytes = 4000.0 * threads * 32 * 4;
      printf("tmem_ld %s  %3d threads: %6.1f B/cycle/SM  (%5.1f cycles per warp-level 32-column read)   [%s]\n", x32 ? "32x32b.x32  " : "2 x 32x32b.x16", threads, bytes / cyc,
             (double)cyc / 4000.0, cudaGetErrorString(e));
    }
  }
  for (int mode = 0; mode < 3; ++mode) {
    publish_probe<<<148, 544>>>(mode, 2000, d_out);
    long long cyc = 0;
    cudaError_t e = cudaDeviceSynchronize();
    cudaMemcpy(&cyc, d_out, 8, cudaMemcpyDeviceToHost);
    const char* names[] = {"per-thread fence.proxy.async + 512 arrives", "per-thread fence, 16 warp-elected arrives", "bar.sync + one fence + one arrive"};
    printf("publish  %-48s: %7.1f cycles per round trip   [%s]\n", names[mode], (double)cyc / 2000, cudaGetErrorString(e));
  }
  return 0;
}
