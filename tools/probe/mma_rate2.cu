// Development microbenchmark (round 2): what bounds small-N tcgen05.mma kind::f16 (M=128, K=16)?
//  * N sweep with one issuing thread (TS mode: A in TMEM)
//  * A-operand collector reuse (collector::a::fill / ::lastuse) for the hi*hi, hi*lo pair of the 3-term split
//  * several issuing warps of one CTA at once (do their MMAs overlap or share one queue?)
//  * M = 64
// Prints cycles per MMA measured by the issuing thread(s) over 60-MMA batches closed by one commit.
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
  d |= (uint64_t)swz << 61;
  return d;
}
__device__ __forceinline__ uint32_t make_idesc(int M, int N) {
  return (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// REUSE 0: plain; 1: triples (A0 fill, A0 lastuse, A1 plain) = the 3-term split order hi*Bh, hi*Bl, lo*Bh
// NISSUE: number of warps (one thread each) that issue concurrently, each into its own accumulator columns
template <int M, int N, int REUSE, int NISSUE>
__global__ void __launch_bounds__(128) rate_kernel(int reps, long long* out) {
  extern __shared__ __align__(1024) uint8_t sm[];
  __shared__ __align__(8) uint64_t mbar[4];
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < 96 * 1024 / 4; i += 128) reinterpret_cast<uint32_t*>(sm)[i] = 0x3c003c00u;   // 1.0h
  if (tid == 0) {
    for (int i = 0; i < 4; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&mbar[i])));
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
  if ((tid & 31) == 0 && warp < NISSUE) {
    const uint32_t idesc = make_idesc(M, N);
    const uint64_t b0 = make_desc(smem_u32(sm), 128, 256, 0);
    constexpr int DW = N < 96 ? N : 96;          // accumulator columns per issuer (N > 96: issuers share, timing only)
    uint32_t parity = 0;
    long long t_issue = 0, t_total = 0;
    const uint32_t d = tbase + 128 + warp * DW;
    for (int r = 0; r < reps; ++r) {
      const long long t0 = clock64();
#pragma unroll
      for (int i = 0; i < 60; ++i) {
        const uint64_t bdesc = b0 + (uint64_t)((i % 6) * 8192 >> 4);
        const uint32_t a = tbase + ((i / 3) % 8) * 16 + ((i % 3) == 2 ? 8 : 0);
        if (REUSE == 1 && (i % 3) == 0) {
          asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16.collector::a::fill [%0], [%1], %2, %3, p;\n\t}" ::"r"(d),
                       "r"(a), "l"(bdesc), "r"(idesc), "r"(1u));
        } else if (REUSE == 1 && (i % 3) == 1) {
          asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16.collector::a::lastuse [%0], [%1], %2, %3, p;\n\t}" ::"r"(d),
                       "r"(a), "l"(bdesc), "r"(idesc), "r"(1u));
        } else {
          asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d),
                       "r"(a), "l"(bdesc), "r"(idesc), "r"(1u));
        }
      }
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&mbar[warp])));
      const long long t1 = clock64();
      uint32_t done = 0;
      while (!done)
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(smem_u32(&mbar[warp])), "r"(parity));
      parity ^= 1;
      const long long t2 = clock64();
      if (r > 0) { t_issue += t1 - t0; t_total += t2 - t0; }
    }
    if (blockIdx.x == 0) { out[2 * warp] = t_issue; out[2 * warp + 1] = t_total; }
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tbase));
}

template <int M, int N, int REUSE, int NISSUE>
void run(long long* d_out) {
  const int reps = 21;
  CK(cudaFuncSetAttribute(rate_kernel<M, N, REUSE, NISSUE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
  rate_kernel<M, N, REUSE, NISSUE><<<1, 128, 96 * 1024>>>(reps, d_out);
  CK(cudaDeviceSynchronize());
  long long h[8]; CK(cudaMemcpy(h, d_out, 64, cudaMemcpyDeviceToHost));
  printf("%4d %4d %5d %6d |", M, N, REUSE, NISSUE);
  for (int w = 0; w < NISSUE; ++w) printf(" %7.1f %7.1f", (double)h[2 * w] / ((reps - 1) * 60), (double)h[2 * w + 1] / ((reps - 1) * 60));
  printf("\n");
}

int main() {
  long long* d_out; CK(cudaMalloc(&d_out, 64));
  printf("%4s %4s %5s %6s | per issuer: issue total (cycles per MMA, K=16 f16, A in TMEM)\n", "M", "N", "reuse", "nissue");
  run<128, 16, 0, 1>(d_out); run<128, 32, 0, 1>(d_out); run<128, 48, 0, 1>(d_out); run<128, 64, 0, 1>(d_out);
  run<128, 80, 0, 1>(d_out); run<128, 96, 0, 1>(d_out); run<128, 112, 0, 1>(d_out); run<128, 128, 0, 1>(d_out);
  run<128, 144, 0, 1>(d_out); run<128, 160, 0, 1>(d_out); run<128, 192, 0, 1>(d_out); run<128, 256, 0, 1>(d_out);
  run<128, 16, 1, 1>(d_out); run<128, 32, 1, 1>(d_out); run<128, 64, 1, 1>(d_out); run<128, 80, 1, 1>(d_out); run<128, 144, 1, 1>(d_out);
  run<128, 16, 0, 2>(d_out); run<128, 16, 0, 3>(d_out); run<128, 64, 0, 3>(d_out); run<128, 80, 0, 3>(d_out);
  run<128, 16, 1, 3>(d_out); run<128, 80, 1, 3>(d_out);
  run<64, 16, 0, 1>(d_out); run<64, 32, 0, 1>(d_out); run<64, 64, 0, 1>(d_out); run<64, 80, 0, 1>(d_out); run<64, 128, 0, 1>(d_out);
  run<64, 16, 0, 3>(d_out); run<64, 80, 0, 3>(d_out);
  return 0;
}
