// l2_stream_probe.cu — what can one SM pull from L2 with N bytes of cp.async in flight?  (sizing of the decode weight streams, round 2)
// Every warp streams 16-row x 256-byte units (the decode_cluster.cu unit) of a bf16 matrix through a private ring of STAGES x 5 KB.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o profiles/l2_stream_probe.bin profiles/l2_stream_probe.cu && profiles/l2_stream_probe.bin
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdint.h>
__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void cpa16(void* d, const void* s) { asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s32(d)), "l"(s) : "memory"); }
__device__ __forceinline__ void commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void waitg() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

template <int STAGES, int MODE>   // MODE 0: cp.async ring, MODE 1: direct 16-byte register loads (8 in flight per lane)
__global__ void __launch_bounds__(256, 1) stream(const unsigned char* base, size_t bytes_per_cta, int same, float* sink, unsigned long long* cyc) {
  extern __shared__ __align__(128) unsigned char sm[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const unsigned char* src = base + (same ? 0 : (size_t)blockIdx.x * bytes_per_cta) + (size_t)warp * (bytes_per_cta / 8);
  const int nunits = (int)(bytes_per_cta / 8 / 4096);
  unsigned char* ring = sm + warp * STAGES * 5120;
  float acc = 0.f;
  const long long t0 = clock64();
  if (MODE == 0) {
    auto issue = [&](int u) {
      if (u < nunits) {
        unsigned char* dst = ring + (u % STAGES) * 5120;
        for (int e = lane; e < 256; e += 32) cpa16(dst + (e >> 4) * 320 + (e & 15) * 16, src + (size_t)u * 4096 + e * 16);
      }
      commit();
    };
    for (int u = 0; u < STAGES - 1; ++u) issue(u);
    for (int u = 0; u < nunits; ++u) {
      issue(u + STAGES - 1);
      waitg<STAGES - 1>();
      __syncwarp();
      const unsigned char* st = ring + (u % STAGES) * 5120;
      for (int k = 0; k < 8; ++k) acc += *reinterpret_cast<const float*>(st + (lane >> 2) * 320 + k * 32 + (lane & 3) * 4);
      __syncwarp();
    }
  } else {
    for (int u = 0; u < nunits; ++u) {
      uint4 v[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) v[k] = __ldg(reinterpret_cast<const uint4*>(src + (size_t)u * 4096 + (k * 32 + lane) * 16));
#pragma unroll
      for (int k = 0; k < 8; ++k) acc += __uint_as_float(v[k].x ^ v[k].y ^ v[k].z ^ v[k].w);
    }
  }
  const long long t1 = clock64();
  if (threadIdx.x == 0) cyc[blockIdx.x] = (unsigned long long)(t1 - t0);
  if (acc == 123.456f) sink[0] = acc;
}

template <int STAGES, int MODE>
static void run(const unsigned char* buf, int ctas, size_t per_cta, int same, const char* what) {
  float* sink; unsigned long long* cyc;
  cudaMalloc(&sink, 4); cudaMalloc(&cyc, 8 * 256);
  const size_t smem = 8 * STAGES * 5120;
  cudaFuncSetAttribute(stream<STAGES, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int it = 0; it < 3; ++it) stream<STAGES, MODE><<<ctas, 256, smem>>>(buf, per_cta, same, sink, cyc);
  cudaEventRecord(e0);
  const int reps = 20;
  for (int it = 0; it < reps; ++it) stream<STAGES, MODE><<<ctas, 256, smem>>>(buf, per_cta, same, sink, cyc);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  const double us = ms * 1e3 / reps;
  printf("%-44s ctas %3d  %6.1f KB in flight/SM  %7.1f us  %6.1f GB/s per SM  %6.2f TB/s total  (%s)\n", what, ctas, MODE == 0 ? 8.0 * (STAGES - 1) * 4 : 8.0 * 4, us,
         per_cta / us / 1e3, per_cta * (double)ctas / us / 1e6, cudaGetErrorString(cudaGetLastError()));
}

int main() {
  const size_t total = 64ull << 20;       // 64 MB: L2-resident after the warm-up passes
  unsigned char* buf; cudaMalloc(&buf, total); cudaMemset(buf, 1, total);
  for (int ctas : {8, 64, 128}) {
    const size_t per = 512 << 10;         // 512 KB per CTA (a layer's FF slices)
    run<3, 0>(buf, ctas, per, 0, "cp.async ring 3 stages, private regions");
    run<3, 0>(buf, ctas, per, 1, "cp.async ring 3 stages, all read the same");
    run<6, 0>(buf, ctas, per, 0, "cp.async ring 6 stages, private regions");
    run<9, 0>(buf, ctas, per, 0, "cp.async ring 9 stages, private regions");
    run<3, 1>(buf, ctas, per, 0, "direct 16-byte loads x 8, private regions");
    run<3, 1>(buf, ctas, per, 1, "direct 16-byte loads x 8, all read the same");
  }
  return 0;
}
