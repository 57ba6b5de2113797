// Thread-private small dense layers: activations in registers, weights broadcast from
// shared memory as 128-bit loads.  Weights are staged TRANSPOSED, Wt[in][OUTP] with
// OUTP = OUT rounded up to 4 (zero padded), so one LDS.128 feeds four FMAs.
#pragma once
#include <cuda_runtime.h>

namespace mpn {

template <int OUT>
struct Pad4 { static constexpr int value = (OUT + 3) / 4 * 4; };

// Cooperative copy of nn.Linear weight W[out][in] (global) into Wt[in][OUTP] (shared).
template <int IN, int OUT>
__device__ __forceinline__ void stage_weight_t(const float* __restrict__ w, float* __restrict__ wt) {
  constexpr int OUTP = Pad4<OUT>::value;
  for (int idx = threadIdx.x; idx < IN * OUTP; idx += blockDim.x) {
    const int i = idx / OUTP, o = idx - i * OUTP;
    wt[idx] = o < OUT ? w[o * IN + i] : 0.f;
  }
}

template <int OUT>
__device__ __forceinline__ void stage_bias(const float* __restrict__ b, float* __restrict__ sb) {
  constexpr int OUTP = Pad4<OUT>::value;
  for (int o = threadIdx.x; o < OUTP; o += blockDim.x) sb[o] = o < OUT ? b[o] : 0.f;
}

template <int OUTP>
__device__ __forceinline__ void load_bias(float (&acc)[OUTP], const float* __restrict__ sb) {
#pragma unroll
  for (int o = 0; o < OUTP; o += 4) {
    const float4 b = *reinterpret_cast<const float4*>(sb + o);
    acc[o] = b.x; acc[o + 1] = b.y; acc[o + 2] = b.z; acc[o + 3] = b.w;
  }
}

// acc[:] += a * Wt[row][:]
template <int OUTP>
__device__ __forceinline__ void axpy_row(float (&acc)[OUTP], float a, const float* __restrict__ wrow) {
#pragma unroll
  for (int o = 0; o < OUTP; o += 4) {
    const float4 w = *reinterpret_cast<const float4*>(wrow + o);
    acc[o] = fmaf(a, w.x, acc[o]);
    acc[o + 1] = fmaf(a, w.y, acc[o + 1]);
    acc[o + 2] = fmaf(a, w.z, acc[o + 2]);
    acc[o + 3] = fmaf(a, w.w, acc[o + 3]);
  }
}

// acc += in[0..IN) @ Wt[row0 .. row0+IN)
template <int IN, int OUTP, int INP>
__device__ __forceinline__ void dense_acc(float (&acc)[OUTP], const float (&in)[INP],
                                          const float* __restrict__ wt, int row0 = 0) {
  static_assert(IN <= INP, "input array too small");
#pragma unroll
  for (int i = 0; i < IN; ++i) axpy_row<OUTP>(acc, in[i], wt + (row0 + i) * OUTP);
}

template <int N>
__device__ __forceinline__ void relu_inplace(float (&v)[N]) {
#pragma unroll
  for (int i = 0; i < N; ++i) v[i] = fmaxf(v[i], 0.f);
}

}  // namespace mpn
