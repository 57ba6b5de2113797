// Development probe: TMA tile::gather4 (row gather by index) into a 128B-swizzled K-major shared-memory tile and
// a tcgen05.mma (SS mode, SWIZZLE_128B descriptors) that consumes it.  D[128 x N] = X[idx[m], :] * B[N x 64]^T.
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc_sw128(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)1 << 16;                    // LBO (unused for swizzled K-major)
  d |= (uint64_t)(1024 >> 4) << 32;          // SBO: 8 rows x 128 B
  d |= (uint64_t)1 << 46;                    // version
  d |= (uint64_t)2 << 61;                    // SWIZZLE_128B
  return d;
}
__device__ __forceinline__ uint32_t make_idesc(int M, int N) {
  return (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

constexpr int N = 80;
__global__ void __launch_bounds__(128) probe(const __grid_constant__ CUtensorMap tmap, const int* __restrict__ idx,
                                             const __half* __restrict__ B, float* __restrict__ D, uint4* __restrict__ dump) {
  extern __shared__ __align__(1024) uint8_t sm[];
  uint8_t* sA = sm;                 // 128 rows x 128 B, SW128
  uint8_t* sB = sm + 16384;         // N rows x 128 B, SW128 (written by threads)
  __shared__ __align__(8) uint64_t mbar_tma, mbar_mma;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < N * 8; i += 128) {           // 16-B chunks of B rows, swizzled by hand
    const int n = i >> 3, c = i & 7;
    *reinterpret_cast<uint4*>(sB + n * 128 + ((c ^ (n & 7)) << 4)) = *reinterpret_cast<const uint4*>(B + n * 64 + c * 8);
  }
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&mbar_tma)));
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&mbar_mma)));
    asm volatile("fence.mbarrier_init.release.cluster;");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 128;" ::"r"(smem_u32(&tmem_base_s)));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("fence.proxy.async.shared::cta;");
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  const uint32_t tbase = tmem_base_s;
  if (warp == 0) {
    // lane l issues the gather of rows 4l..4l+3
    if (tid == 0) asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&mbar_tma)), "r"(128 * 128));
    __syncwarp();
    const int r0 = idx[4 * tid], r1 = idx[4 * tid + 1], r2 = idx[4 * tid + 2], r3 = idx[4 * tid + 3];
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.tile::gather4.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
        ::"r"(smem_u32(sA + tid * 512)), "l"(&tmap), "r"(smem_u32(&mbar_tma)), "r"(0), "r"(r0), "r"(r1), "r"(r2), "r"(r3)
        : "memory");
  }
  {
    uint32_t done = 0;
    while (!done)
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                   : "=r"(done) : "r"(smem_u32(&mbar_tma)), "r"(0u));
  }
  for (int i = tid; i < 1024; i += 128) dump[i] = reinterpret_cast<const uint4*>(sA)[i];
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  if (tid == 0) {
    const uint32_t idesc = make_idesc(128, N);
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
      const uint64_t adesc = make_desc_sw128(smem_u32(sA) + ks * 32);
      const uint64_t bdesc = make_desc_sw128(smem_u32(sB) + ks * 32);
      asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tbase),
                   "l"(adesc), "l"(bdesc), "r"(idesc), "r"(ks > 0 ? 1u : 0u));
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&mbar_mma)));
  }
  {
    uint32_t done = 0;
    while (!done)
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                   : "=r"(done) : "r"(smem_u32(&mbar_mma)), "r"(0u));
  }
  asm volatile("tcgen05.fence::after_thread_sync;");
  const uint32_t lane_base = (uint32_t)(warp * 32) << 16;
#pragma unroll
  for (int c0 = 0; c0 < N; c0 += 16) {
    uint32_t v[16];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                   "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                 : "r"(tbase + lane_base + c0));
    asm volatile("tcgen05.wait::ld.sync.aligned;");
#pragma unroll
    for (int j = 0; j < 16; ++j) D[(size_t)tid * N + c0 + j] = __uint_as_float(v[j]);
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 128;" ::"r"(tbase));
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main(int argc, char** argv) {
  const int box1 = argc > 1 ? atoi(argv[1]) : 1;
  const int NR = 4096;
  std::vector<__half> hX((size_t)NR * 64), hB(N * 64);
  std::vector<int> hidx(128);
  srand(7);
  for (auto& v : hX) v = __float2half((rand() % 2001 - 1000) / 512.0f);
  for (auto& v : hB) v = __float2half((rand() % 2001 - 1000) / 1024.0f);
  for (auto& v : hidx) v = rand() % NR;
  __half *dX, *dB; int* didx; float* dD; uint4* ddump;
  CK(cudaMalloc(&dX, hX.size() * 2)); CK(cudaMalloc(&dB, hB.size() * 2)); CK(cudaMalloc(&didx, 512));
  CK(cudaMalloc(&dD, 128 * N * 4)); CK(cudaMalloc(&ddump, 16384));
  CK(cudaMemcpy(dX, hX.data(), hX.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dB, hB.data(), hB.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(didx, hidx.data(), 512, cudaMemcpyHostToDevice));
  EncodeFn encode = nullptr;
  cudaDriverEntryPointQueryResult qres;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", (void**)&encode, cudaEnableDefault, &qres));
  if (!encode) { printf("no cuTensorMapEncodeTiled\n"); return 1; }
  CUtensorMap tmap;
  const cuuint64_t dims[2] = {64, (cuuint64_t)NR};
  const cuuint64_t strides[1] = {128};
  const cuuint32_t box[2] = {64, (cuuint32_t)box1};
  const cuuint32_t estr[2] = {1, 1};
  CUresult r = encode(&tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, dX, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                      CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  printf("encode (box1=%d) -> %d\n", box1, (int)r);
  if (r != CUDA_SUCCESS) return 1;
  CK(cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 32768));
  probe<<<1, 128, 32768>>>(tmap, didx, dB, dD, ddump);
  CK(cudaDeviceSynchronize());
  std::vector<uint4> dump(1024); std::vector<float> out(128 * N);
  CK(cudaMemcpy(dump.data(), ddump, 16384, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(out.data(), dD, out.size() * 4, cudaMemcpyDeviceToHost));
  // layout check: chunk c of tile row m expected at m*128 + ((c ^ (m & 7)) << 4)
  int bad_layout = 0;
  for (int m = 0; m < 128; ++m)
    for (int c = 0; c < 8; ++c) {
      const uint4 got = dump[(m * 128 + ((c ^ (m & 7)) << 4)) / 16];
      const uint4 exp = *reinterpret_cast<const uint4*>(&hX[(size_t)hidx[m] * 64 + c * 8]);
      if (memcmp(&got, &exp, 16) != 0) ++bad_layout;
    }
  double maxerr = 0; int bad = 0;
  for (int m = 0; m < 128; ++m)
    for (int n = 0; n < N; ++n) {
      double s = 0;
      for (int k = 0; k < 64; ++k) s += (double)__half2float(hX[(size_t)hidx[m] * 64 + k]) * __half2float(hB[n * 64 + k]);
      const double e = fabs(out[m * N + n] - s);
      if (e > maxerr) maxerr = e;
      if (!(e <= 2e-3)) ++bad;
    }
  printf("gather layout mismatches: %d / 1024 chunks; MMA max err %.3e bad %d / %d\n", bad_layout, maxerr, bad, 128 * N);
  printf((bad_layout || bad) ? "PROBE FAILED\n" : "PROBE OK\n");
  return 0;
}
