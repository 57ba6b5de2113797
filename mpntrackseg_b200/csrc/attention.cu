// Attentive aggregation of the node feature maps (models/mpn.py:117-137):
//   w_e  = softmax of the edge logits over each node's future (row<col) resp. past (row>col) neighbours
//          (torch_scatter.composite.scatter_softmax: exp(l - max) / (sum + 1e-12)),
//   flow = sum_e w_e * z[col_e]      (z: [N, C*H*W] feature maps),
// evaluated per (node, direction) over that node's contiguous slot range, neighbours taken in slot
// order (= the reference's scatter_add order on CPU): deterministic, no atomics.
#include <math.h>

#include "common.cuh"

namespace mpn {

constexpr int ATT_THREADS = 256;
constexpr int ATT_MAX_DEG = 1024;   // neighbours whose weights are cached in shared memory per pass

__global__ void __launch_bounds__(ATT_THREADS) attn_aggregate_kernel(
    const float* __restrict__ z, int64_t feat, const int32_t* __restrict__ slot_col,
    const int32_t* __restrict__ slot_edge, const int32_t* __restrict__ out_ptr, const int32_t* __restrict__ in_ptr,
    const float* __restrict__ logits, int64_t num_nodes, float* __restrict__ flow_in, float* __restrict__ flow_out) {
  __shared__ float s_w[ATT_MAX_DEG];
  __shared__ int32_t s_c[ATT_MAX_DEG];
  __shared__ float s_red[32];
  const int64_t node = blockIdx.x >> 1;
  const int dir = blockIdx.x & 1;                       // 0 = flow_in (row>col), 1 = flow_out (row<col)
  const int32_t* ptr = dir == 0 ? in_ptr : out_ptr;
  float* out = (dir == 0 ? flow_in : flow_out) + node * feat;
  const int s0 = ptr[node], s1 = ptr[node + 1];
  const int deg = s1 - s0;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (deg == 0) {                                       // scatter_add zero fill
    for (int64_t i = tid; i < feat; i += ATT_THREADS) out[i] = 0.f;
    return;
  }
  // segment max
  float m = -INFINITY;
  for (int q = tid; q < deg; q += ATT_THREADS) m = fmaxf(m, logits[slot_edge[s0 + q]]);
  for (int d = 16; d > 0; d >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, d));
  if (lane == 0) s_red[warp] = m;
  __syncthreads();
  m = s_red[0];
  for (int w = 1; w < ATT_THREADS / 32; ++w) m = fmaxf(m, s_red[w]);
  __syncthreads();
  // segment sum of exp, sequential order per thread then fixed tree (deterministic)
  float sum = 0.f;
  for (int q = tid; q < deg; q += ATT_THREADS) sum += expf(logits[slot_edge[s0 + q]] - m);
  for (int d = 16; d > 0; d >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, d);
  if (lane == 0) s_red[warp] = sum;
  __syncthreads();
  sum = 0.f;
  for (int w = 0; w < ATT_THREADS / 32; ++w) sum += s_red[w];
  const float inv = 1.f / (sum + 1e-12f);
  // weighted sum of the neighbours' maps
  for (int q0 = 0; q0 < deg; q0 += ATT_MAX_DEG) {
    const int nq = deg - q0 < ATT_MAX_DEG ? deg - q0 : ATT_MAX_DEG;
    __syncthreads();
    for (int q = tid; q < nq; q += ATT_THREADS) {
      s_w[q] = expf(logits[slot_edge[s0 + q0 + q]] - m) * inv;
      s_c[q] = slot_col[s0 + q0 + q];
    }
    __syncthreads();
    for (int64_t i = tid; i < feat; i += ATT_THREADS) {
      float acc = q0 == 0 ? 0.f : out[i];
      for (int q = 0; q < nq; ++q) acc += z[(int64_t)s_c[q] * feat + i] * s_w[q];   // x[col] * w, then add (mpn.py:123-124)
      out[i] = acc;
    }
  }
}

}  // namespace mpn

using namespace mpn;

extern "C" int mpn_attn_aggregate(const float* z, int64_t num_nodes, int64_t feat, const mpn_edge_layout* g,
                                  const float* logits, float* flow_in, float* flow_out, void* stream) {
  MPN_CHECK_ARG(g != nullptr && num_nodes == g->num_nodes, "attn_aggregate: layout / node count mismatch");
  if (num_nodes == 0) return MPN_OK;
  MPN_CHECK_ARG(z && flow_in && flow_out && feat > 0, "attn_aggregate: null pointer");
  MPN_CHECK_ARG(logits != nullptr || g->num_edges == 0, "attn_aggregate: null logits");
  MPN_CHECK_ARG(2 * num_nodes < (1ll << 31), "attn_aggregate: too many nodes");
  attn_aggregate_kernel<<<(unsigned)(2 * num_nodes), ATT_THREADS, 0, as_stream(stream)>>>(
      z, feat, g->slot_col, g->slot_edge, g->out_ptr, g->in_ptr, logits, num_nodes, flow_in, flow_out);
  count_launch();
  MPN_LAUNCH_CHECK();
  return MPN_OK;
}
