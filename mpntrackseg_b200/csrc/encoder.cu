// Encoders: global average pool, generic Linear(+ReLU) layer (fp32 SIMT tiled GEMM),
// row gather, and the fused edge-feature MLP in slot order.
#include "common.cuh"
#include "mlp_regs.cuh"

namespace mpn {

constexpr unsigned kFull = 0xffffffffu;

// ------------------------------------------------------------------ average pool
// x[n*c][hw] -> out[n*c].  GROUP = hw/4 lanes cooperate on one channel (float4 each).
template <int GROUP>
__global__ void avgpool_vec_kernel(const float4* __restrict__ x, int64_t rows, float inv_hw_unused,
                                   int hw, float* __restrict__ out) {
  constexpr int ROWS_PER_WARP = 32 / GROUP;
  const int lane = threadIdx.x & 31;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  // UNROLL independent 16-byte loads per thread are issued before any of them is consumed: the kernel is a pure
  // HBM stream and needs the bytes in flight (one load per thread left it at 87 % of the measured peak)
  constexpr int UNROLL = 4;
  const int64_t stride = nwarps * ROWS_PER_WARP;
  for (int64_t r0 = warp * ROWS_PER_WARP; r0 < rows; r0 += UNROLL * stride) {
    float4 v[UNROLL];
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) {
      const int64_t ru = r0 + u * stride;
      v[u] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (ru + lane / GROUP < rows) v[u] = __ldcs(x + ru * GROUP + lane);      // streaming: read once
    }
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) {
      const int64_t r = r0 + u * stride + lane / GROUP;
      float s = (v[u].x + v[u].y) + (v[u].z + v[u].w);
#pragma unroll
      for (int d = GROUP / 2; d > 0; d >>= 1) s += __shfl_xor_sync(kFull, s, d);
      if (r < rows && (lane % GROUP) == 0) out[r] = s / (float)hw;
    }
  }
}

__global__ void avgpool_scalar_kernel(const float* __restrict__ x, int64_t rows, int hw,
                                      float* __restrict__ out) {
  for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < rows;
       r += (int64_t)gridDim.x * blockDim.x) {
    const float* p = x + r * hw;
    float s = 0.f;
    for (int i = 0; i < hw; ++i) s += p[i];
    out[r] = s / (float)hw;
  }
}

// ------------------------------------------------------------------ generic linear
// out[m][o] = act(sum_k in[m][k] * w[o][k] + b[o]);  64x64 tile, 16-deep k slab, 4x4 per thread.
constexpr int LT = 64, LK = 16;
__global__ void __launch_bounds__(256) linear_kernel(const float* __restrict__ in, int64_t m, int64_t k,
                                                     const float* __restrict__ w,
                                                     const float* __restrict__ b, int64_t o, int relu,
                                                     float* __restrict__ out) {
  __shared__ float As[LK][LT + 4];
  __shared__ float Bs[LK][LT + 4];
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const int64_t m0 = (int64_t)blockIdx.y * LT, o0 = (int64_t)blockIdx.x * LT;
  float acc[4][4] = {};
  for (int64_t k0 = 0; k0 < k; k0 += LK) {
    for (int idx = threadIdx.x; idx < LT * LK; idx += 256) {
      const int r = idx / LK, kk = idx % LK;
      const int64_t gm = m0 + r, go = o0 + r, gk = k0 + kk;
      As[kk][r] = (gm < m && gk < k) ? in[gm * k + gk] : 0.f;
      Bs[kk][r] = (go < o && gk < k) ? w[go * k + gk] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < LK; ++kk) {
      const float4 a = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
      const float4 bb = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
      const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {bb.x, bb.y, bb.z, bb.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int64_t gm = m0 + ty * 4 + i;
    if (gm >= m) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int64_t go = o0 + tx * 4 + j;
      if (go >= o) continue;
      float v = acc[i][j] + (b != nullptr ? b[go] : 0.f);
      out[gm * o + go] = relu ? fmaxf(v, 0.f) : v;
    }
  }
}

// ------------------------------------------------------------------ fused edge encoder
template <int D0, int D1, int D2, int D3>
__global__ void __launch_bounds__(256) edge_encoder_kernel(
    const float* __restrict__ attr, const int32_t* __restrict__ slot_edge, int64_t e,
    const float* __restrict__ w0, const float* __restrict__ b0, const float* __restrict__ w1,
    const float* __restrict__ b1, const float* __restrict__ w2, const float* __restrict__ b2,
    float* __restrict__ e_init) {
  constexpr int P1 = Pad4<D1>::value, P2 = Pad4<D2>::value, P3 = Pad4<D3>::value;
  static_assert(D3 % 4 == 0, "edge latent width must be a multiple of 4");
  __shared__ __align__(16) float s_w0[D0 * P1], s_w1[D1 * P2], s_w2[D2 * P3];
  __shared__ __align__(16) float s_b0[P1], s_b1[P2], s_b2[P3];
  stage_weight_t<D0, D1>(w0, s_w0);
  stage_weight_t<D1, D2>(w1, s_w1);
  stage_weight_t<D2, D3>(w2, s_w2);
  stage_bias<D1>(b0, s_b0);
  stage_bias<D2>(b1, s_b1);
  stage_bias<D3>(b2, s_b2);
  __syncthreads();
  for (int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; s < e;
       s += (int64_t)gridDim.x * blockDim.x) {
    const float* a = attr + (int64_t)(slot_edge != nullptr ? slot_edge[s] : s) * D0;
    float in[D0];
#pragma unroll
    for (int i = 0; i < D0; ++i) in[i] = a[i];
    float h1[P1], h2[P2], h3[P3];
    load_bias<P1>(h1, s_b0);
    dense_acc<D0, P1>(h1, in, s_w0);
    relu_inplace(h1);
    load_bias<P2>(h2, s_b1);
    dense_acc<D1, P2>(h2, h1, s_w1);
    relu_inplace(h2);
    load_bias<P3>(h3, s_b2);
    dense_acc<D2, P3>(h3, h2, s_w2);
    relu_inplace(h3);
    float4* dst = reinterpret_cast<float4*>(e_init + s * D3);
#pragma unroll
    for (int q = 0; q < D3 / 4; ++q) dst[q] = make_float4(h3[4 * q], h3[4 * q + 1], h3[4 * q + 2], h3[4 * q + 3]);
  }
}

__global__ void gather_rows_kernel(const float* __restrict__ in, const int32_t* __restrict__ idx,
                                   int64_t rows, int64_t width, float* __restrict__ out) {
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < rows * width;
       t += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = t / width, c = t - r * width;
    out[t] = in[(int64_t)idx[r] * width + c];
  }
}

static inline unsigned grid_for(int64_t n, int threads, int per_sm = 8) {
  int64_t want = ceil_div(n > 0 ? n : 1, threads);
  int64_t cap = (int64_t)sm_count() * per_sm;
  return (unsigned)(want < cap ? want : cap);
}

}  // namespace mpn

using namespace mpn;

extern "C" {

int mpn_avgpool(const float* x, int64_t n, int64_t c, int64_t hw, float* out, void* stream) {
  MPN_CHECK_ARG(n >= 0 && c > 0 && hw > 0, "avgpool: bad sizes");
  const int64_t rows = n * c;
  if (rows == 0) return MPN_OK;
  MPN_CHECK_ARG(x && out, "avgpool: null pointer");
  cudaStream_t s = as_stream(stream);
  const int64_t vec = hw / 4;
  const bool vec_ok = (hw % 4 == 0) && vec >= 1 && vec <= 32 && (vec & (vec - 1)) == 0 &&
                      (reinterpret_cast<uintptr_t>(x) % 16 == 0);
  if (vec_ok) {
    const unsigned grid = grid_for(rows * vec, 256, 16);
    const float4* x4 = reinterpret_cast<const float4*>(x);
    switch (vec) {
      case 1: avgpool_vec_kernel<1><<<grid, 256, 0, s>>>(x4, rows, 0.f, (int)hw, out); count_launch(); break;
      case 2: avgpool_vec_kernel<2><<<grid, 256, 0, s>>>(x4, rows, 0.f, (int)hw, out); count_launch(); break;
      case 4: avgpool_vec_kernel<4><<<grid, 256, 0, s>>>(x4, rows, 0.f, (int)hw, out); count_launch(); break;
      case 8: avgpool_vec_kernel<8><<<grid, 256, 0, s>>>(x4, rows, 0.f, (int)hw, out); count_launch(); break;
      case 16: avgpool_vec_kernel<16><<<grid, 256, 0, s>>>(x4, rows, 0.f, (int)hw, out); count_launch(); break;
      default: avgpool_vec_kernel<32><<<grid, 256, 0, s>>>(x4, rows, 0.f, (int)hw, out); count_launch(); break;
    }
  } else {
    avgpool_scalar_kernel<<<grid_for(rows, 256), 256, 0, s>>>(x, rows, (int)hw, out); count_launch();
  }
  MPN_LAUNCH_CHECK();
  return MPN_OK;
}

int mpn_linear(const float* in, int64_t m, int64_t k, const float* w, const float* b, int64_t o,
               int relu, float* out, void* stream) {
  MPN_CHECK_ARG(m >= 0 && k > 0 && o > 0, "linear: bad sizes");
  if (m == 0) return MPN_OK;
  MPN_CHECK_ARG(in && w && out, "linear: null pointer");
  dim3 grid((unsigned)ceil_div(o, LT), (unsigned)ceil_div(m, LT));
  MPN_CHECK_ARG(grid.y <= 65535u, "linear: at most %d rows per call (gridDim.y limit)", 65535 * LT);
  linear_kernel<<<grid, 256, 0, as_stream(stream)>>>(in, m, k, w, b, o, relu, out); count_launch();
  MPN_LAUNCH_CHECK();
  return MPN_OK;
}

int mpn_gather_rows(const float* in, const int32_t* idx, int64_t rows, int64_t width, float* out,
                    void* stream) {
  if (rows == 0) return MPN_OK;
  MPN_CHECK_ARG(in && idx && out && width > 0, "gather_rows: bad args");
  gather_rows_kernel<<<grid_for(rows * width, 256), 256, 0, as_stream(stream)>>>(in, idx, rows, width, out); count_launch();
  MPN_LAUNCH_CHECK();
  return MPN_OK;
}

int mpn_edge_encoder(const float* edge_attr, const int32_t* slot_edge, int64_t e, const int32_t* dims,
                     int32_t n_layers, const float* const* w, const float* const* b, float* e_init,
                     void* stream) {
  MPN_CHECK_ARG(e >= 0 && dims && w && b, "edge_encoder: bad args");
  const bool shipped = n_layers == 3 && dims[0] == 6 && dims[1] == 18 && dims[2] == 18 && dims[3] == 16;
  MPN_CHECK_ARG(shipped, "edge_encoder: fused kernel is built for widths 6-18-18-16 "
                         "(configs/tracking_cfg.yaml:141-144); use mpn_gather_rows + mpn_linear for others");
  if (e == 0) return MPN_OK;
  MPN_CHECK_ARG(edge_attr && e_init, "edge_encoder: null pointer");
  edge_encoder_kernel<6, 18, 18, 16><<<grid_for(e, 256, 4), 256, 0, as_stream(stream)>>>(
      edge_attr, slot_edge, e, w[0], b[0], w[1], b[1], w[2], b[2], e_init); count_launch();
  MPN_LAUNCH_CHECK();
  return MPN_OK;
}

}  // extern "C"
