// Development microbenchmark: issue rate / execution time of small tcgen05.mma kind::f16 instructions
// (M=128, K=16) as a function of N, operand source of A (TMEM vs shared memory) and the shared-memory layout.
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo, int swz) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)swz << 61;          // 0 none, 2 = 128B, 4 = 64B, 6 = 32B
  return d;
}
__device__ __forceinline__ uint32_t make_idesc(int M, int N) {
  return (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// MODE 0: TS (A in TMEM), 1: SS.  NDST independent accumulators used round robin; 60 MMAs per commit, fully unrolled
// with precomputed descriptors so that the issuing thread is not the limit.
template <int N, int MODE, int NDST>
__global__ void __launch_bounds__(128) rate_kernel(int reps, long long* out) {
  extern __shared__ __align__(1024) uint8_t sm[];
  __shared__ __align__(8) uint64_t mbar;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < 160 * 1024 / 4; i += 128) reinterpret_cast<uint32_t*>(sm)[i] = 0x3c003c00u;   // 1.0h
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&mbar)));
    asm volatile("fence.mbarrier_init.release.cluster;");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tmem_base_s)));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("fence.proxy.async.shared::cta;");
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  const uint32_t tbase = tmem_base_s;
  if (tid == 0) {
    const uint32_t idesc = make_idesc(128, N);
    const uint64_t a0 = make_desc(smem_u32(sm), 128, 256, 0), b0 = make_desc(smem_u32(sm) + 64 * 1024, 128, 256, 0);
    constexpr int DW = N < 128 ? N : 128;
    uint32_t parity = 0;
    long long t_issue = 0, t_total = 0;
    for (int r = 0; r < reps; ++r) {
      const long long t0 = clock64();
#pragma unroll
      for (int i = 0; i < 60; ++i) {
        const uint64_t bdesc = b0 + (uint64_t)((i % 6) * 8192 >> 4);
        const uint32_t d = tbase + 256 + (i % NDST) * DW;
        if (MODE == 1) {
          const uint64_t adesc = a0 + (uint64_t)((i % 6) * 4096 >> 4);
          asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d),
                       "l"(adesc), "l"(bdesc), "r"(idesc), "r"(1u));
        } else {
          asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d),
                       "r"(tbase + (i % 16) * 8), "l"(bdesc), "r"(idesc), "r"(1u));
        }
      }
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&mbar)));
      const long long t1 = clock64();
      uint32_t done = 0;
      while (!done)
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(smem_u32(&mbar)), "r"(parity));
      parity ^= 1;
      const long long t2 = clock64();
      if (r > 0) { t_issue += t1 - t0; t_total += t2 - t0; }
    }
    if (blockIdx.x == 0) { out[0] = t_issue; out[1] = t_total; }
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tbase));
}

template <int N, int MODE, int NDST>
void run(long long* d_out) {
  const int reps = 21;
  CK(cudaFuncSetAttribute(rate_kernel<N, MODE, NDST>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
  rate_kernel<N, MODE, NDST><<<1, 128, 160 * 1024>>>(reps, d_out);
  CK(cudaDeviceSynchronize());
  long long h[2]; CK(cudaMemcpy(h, d_out, 16, cudaMemcpyDeviceToHost));
  printf("%4d %4s %5d | %10.1f %10.1f\n", N, MODE ? "SS" : "TS", NDST, (double)h[0] / ((reps - 1) * 60), (double)h[1] / ((reps - 1) * 60));
}

int main() {
  long long* d_out; CK(cudaMalloc(&d_out, 16));
  printf("%4s %4s %5s | %10s %10s   (cycles per MMA, M=128 K=16 f16)\n", "N", "mode", "ndst", "issue", "total");
  run<16, 0, 1>(d_out); run<32, 0, 1>(d_out); run<64, 0, 1>(d_out); run<80, 0, 1>(d_out); run<128, 0, 1>(d_out); run<256, 0, 1>(d_out);
  run<16, 1, 1>(d_out); run<32, 1, 1>(d_out); run<64, 1, 1>(d_out); run<80, 1, 1>(d_out); run<128, 1, 1>(d_out); run<256, 1, 1>(d_out);
  run<16, 0, 2>(d_out); run<16, 0, 4>(d_out); run<80, 0, 2>(d_out); run<80, 0, 3>(d_out);
  run<16, 1, 2>(d_out); run<16, 1, 4>(d_out); run<80, 1, 2>(d_out); run<80, 1, 3>(d_out);
  return 0;
}
