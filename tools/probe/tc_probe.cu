// Development probe: one CTA, tcgen05.mma kind::f16 (fp16 in, fp32 accumulate), A from TMEM
// (written with tcgen05.st, lane = row), B from shared memory (K-major, no swizzle).
// Checks D[128 x N] = A[128 x K] * B[N x K]^T against the host.
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint64_t make_b_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;            // version = 1 (sm100)
  return d;                          // layout_type 0 = no swizzle
}

__device__ __forceinline__ uint32_t make_idesc(int M, int N) {
  uint32_t d = 0;
  d |= 1u << 4;                      // D format f32
  d |= 0u << 7;                      // A format f16
  d |= 0u << 10;                     // B format f16
  d |= (uint32_t)(N >> 3) << 17;
  d |= (uint32_t)(M >> 4) << 24;
  return d;
}

template <int N, int K, bool SS>
__global__ void __launch_bounds__(128) probe_kernel(const __half* __restrict__ A, const __half* __restrict__ B,
                                                    float* __restrict__ D) {
  __shared__ __align__(128) __half sB[N * K];            // per k-step slab: [n/8][khalf][8][8]
  __shared__ __align__(128) __half sA[SS ? 128 * K : 8];  // SS mode: A in the same canonical layout
  __shared__ __align__(8) uint64_t mbar;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5;
  // stage B in canonical K-major no-swizzle layout: slab ks at ks*N*16 halfs
  for (int idx = tid; idx < N * K; idx += 128) {
    const int n = idx / K, k = idx % K;
    const int ks = k / 16, kk = k % 16;
    const int off = ks * N * 16 + (n / 8) * 128 + (kk / 8) * 64 + (n % 8) * 8 + (kk % 8);
    sB[off] = B[idx];
  }
  if (SS) {
    for (int idx = tid; idx < 128 * K; idx += 128) {
      const int m = idx / K, k = idx % K;
      const int ks = k / 16, kk = k % 16;
      sA[ks * 128 * 16 + (m / 8) * 128 + (kk / 8) * 64 + (m % 8) * 8 + (kk % 8)] = A[idx];
    }
  }
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&mbar)));
    asm volatile("fence.mbarrier_init.release.cluster;");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 256;" ::"r"(smem_u32(&tmem_base_s)));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("fence.proxy.async.shared::cta;");        // generic smem writes -> visible to the MMA (async proxy)
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  const uint32_t tbase = tmem_base_s;
  const uint32_t lane_base = (uint32_t)(warp * 32) << 16;
  // A row of this thread -> TMEM columns [0, K/2): column c holds (k=2c, k=2c+1)
  const uint32_t* arow = reinterpret_cast<const uint32_t*>(A + (size_t)tid * K);
#pragma unroll
  for (int ks = 0; ks < K / 16; ++ks) {
    uint32_t r[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) r[j] = arow[ks * 8 + j];
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(tbase + lane_base + ks * 8),
                 "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]));
  }
  asm volatile("tcgen05.wait::st.sync.aligned;");
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  const uint32_t d_col = 128;                            // accumulator at columns [128, 128+N)
  if (tid == 0) {
    const uint32_t idesc = make_idesc(128, N);
#pragma unroll
    for (int ks = 0; ks < K / 16; ++ks) {
      const uint64_t bdesc = make_b_desc(smem_u32(sB) + ks * N * 32, 128, 256);
      const uint32_t acc = ks > 0 ? 1u : 0u;
      if (SS) {
        const uint64_t adesc = make_b_desc(smem_u32(sA) + ks * 128 * 32, 128, 256);
        asm volatile(
            "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
            "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tbase + d_col),
            "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc));
      } else {
      asm volatile(
          "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
          "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(tbase + d_col),
          "r"(tbase + ks * 8), "l"(bdesc), "r"(idesc), "r"(acc));
      }
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&mbar)));
  }
  // wait for the MMAs
  {
    uint32_t done = 0;
    while (!done) {
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                   : "=r"(done) : "r"(smem_u32(&mbar)), "r"(0u));
    }
  }
  asm volatile("tcgen05.fence::after_thread_sync;");
#pragma unroll
  for (int c0 = 0; c0 < N; c0 += 16) {
    uint32_t v[16];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                   "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                 : "r"(tbase + lane_base + d_col + c0));
    asm volatile("tcgen05.wait::ld.sync.aligned;");
#pragma unroll
    for (int j = 0; j < 16; ++j) D[(size_t)tid * N + c0 + j] = __uint_as_float(v[j]);
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 256;" ::"r"(tbase));
}

template <int N, int K, bool SS>
int run() {
  std::vector<__half> hA(128 * K), hB(N * K);
  std::vector<float> fA(128 * K), fB(N * K), ref(128 * N), out(128 * N);
  srand(1234 + N * 7 + K);
  for (size_t i = 0; i < hA.size(); ++i) { float v = (rand() % 2001 - 1000) / 512.0f; hA[i] = __float2half(v); fA[i] = __half2float(hA[i]); }
  for (size_t i = 0; i < hB.size(); ++i) { float v = (rand() % 2001 - 1000) / 1024.0f; hB[i] = __float2half(v); fB[i] = __half2float(hB[i]); }
  for (int m = 0; m < 128; ++m)
    for (int n = 0; n < N; ++n) {
      double s = 0;
      for (int k = 0; k < K; ++k) s += (double)fA[m * K + k] * fB[n * K + k];
      ref[m * N + n] = (float)s;
    }
  __half *dA, *dB; float* dD;
  CK(cudaMalloc(&dA, hA.size() * 2)); CK(cudaMalloc(&dB, hB.size() * 2)); CK(cudaMalloc(&dD, out.size() * 4));
  CK(cudaMemcpy(dA, hA.data(), hA.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dB, hB.data(), hB.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemset(dD, 0xff, out.size() * 4));
  probe_kernel<N, K, SS><<<1, 128>>>(dA, dB, dD);
  CK(cudaDeviceSynchronize());
  CK(cudaMemcpy(out.data(), dD, out.size() * 4, cudaMemcpyDeviceToHost));
  double maxerr = 0; int bad = 0;
  for (size_t i = 0; i < out.size(); ++i) { double e = fabs(out[i] - ref[i]); if (e > maxerr) maxerr = e; if (!(e <= 1e-3)) ++bad; }
  printf("%s N=%d K=%d: max err %.3e, bad %d / %zu  (out[0]=%f ref[0]=%f out[last]=%f ref[last]=%f)\n", SS ? "SS" : "TS", N, K, maxerr, bad,
         out.size(), out[0], ref[0], out.back(), ref.back());
  cudaFree(dA); cudaFree(dB); cudaFree(dD);
  return bad;
}

int main() {
  int bad = 0;
  bad += run<16, 16, false>();
  bad += run<80, 96, false>();
  bad += run<16, 16, true>();
  bad += run<80, 96, true>();
  bad += run<64, 64, true>();
  printf(bad ? "PROBE FAILED\n" : "PROBE OK\n");
  return bad != 0;
}
