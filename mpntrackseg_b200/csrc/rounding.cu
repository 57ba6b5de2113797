// Rounding of the edge predictions and identity assignment on the device: the step right after
// the message-passing path, so that the sequence graph does not have to go to the host as numpy.
//   * flow-conservation statistics          utils/evaluation.py:370-414  compute_constr_satisfaction_rate
//   * greedy rounding                        tracker/projectors.py:11-67   GreedyProjector.project
//   * identities = connected components      tracker/mpn_tracker.py:231-248 _assign_ped_ids (scipy connected_components)
// Everything is integer / comparison work: counts with integer atomics, winners with 64-bit atomicMax on
// (prediction bits, ~edge index) keys -- order independent, so the results are deterministic and bit-exact.
#include "common.cuh"

namespace mpn {
namespace {

// flows of BINARISED edge values (exactly 0 or 1) as integer counts; row / col already ordered (row = earlier node)
// unless `undirected`, in which case every pair is stored in both directions and is sorted here (evaluation.py:391-394).
__global__ void flow_count_kernel(const int64_t* __restrict__ row, const int64_t* __restrict__ col,
                                  const float* __restrict__ edges_out, int64_t num_edges, int undirected,
                                  int32_t* __restrict__ cnt_out, int32_t* __restrict__ cnt_in,
                                  int32_t* __restrict__ has_out, int32_t* __restrict__ has_in, int32_t* __restrict__ bad) {
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < num_edges; e += (int64_t)gridDim.x * blockDim.x) {
    int64_t r = row[e], c = col[e];
    if (undirected && r > c) { const int64_t t = r; r = c; c = t; }
    has_out[r] = 1;                                  // the node owns an outgoing / incoming constraint
    has_in[c] = 1;
    const float v = edges_out[e];
    if (v == 1.f) { atomicAdd(&cnt_out[r], 1); atomicAdd(&cnt_in[c], 1); }
    else if (v != 0.f) *bad = 1;                     // not binarised
  }
}

// per node: float flows (count / div), violation and constraint counters
__global__ void flow_stats_kernel(const int32_t* __restrict__ cnt_out, const int32_t* __restrict__ cnt_in,
                                  const int32_t* __restrict__ has_out, const int32_t* __restrict__ has_in,
                                  int64_t num_nodes, float div, float* __restrict__ flow_in, float* __restrict__ flow_out,
                                  unsigned long long* __restrict__ counts /* [0] violated_in [1] violated_out [2] constraints */) {
  unsigned long long vi = 0, vo = 0, nc = 0;
  for (int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; v < num_nodes; v += (int64_t)gridDim.x * blockDim.x) {
    const float fo = (float)cnt_out[v] / div, fi = (float)cnt_in[v] / div;
    if (flow_out) flow_out[v] = fo;
    if (flow_in) flow_in[v] = fi;
    vo += fo > 1.f;
    vi += fi > 1.f;
    nc += (has_out[v] != 0) + (has_in[v] != 0);
  }
  for (int d = 16; d > 0; d >>= 1) {
    vi += __shfl_xor_sync(0xffffffffu, vi, d);
    vo += __shfl_xor_sync(0xffffffffu, vo, d);
    nc += __shfl_xor_sync(0xffffffffu, nc, d);
  }
  if ((threadIdx.x & 31) == 0) {
    if (vi) atomicAdd(&counts[0], vi);
    if (vo) atomicAdd(&counts[1], vo);
    if (nc) atomicAdd(&counts[2], nc);
  }
}

__global__ void round_kernel(const float* __restrict__ preds, int64_t num_edges, float* __restrict__ rounded) {
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < num_edges; e += (int64_t)gridDim.x * blockDim.x)
    rounded[e] = preds[e] > 0.5f ? 1.f : 0.f;
}

// One phase of the greedy projection over the constraints of one type (side 0: outgoing, keyed by row; side 1:
// incoming, keyed by col).  The constraints of a type own disjoint edge sets, so the order in which the reference walks
// them (projectors.py:41-60) does not matter; outgoing constraints come first (descending sort on the type, :38).
//   pass A: a node whose INITIAL flow violates the constraint and whose CURRENT active-edge count is still > 1 keeps the
//           active edge with the largest prediction (first such edge on ties, projectors.py:56).
__global__ void greedy_count_kernel(const int64_t* __restrict__ key_node, const float* __restrict__ rounded,
                                    int64_t num_edges, int32_t* __restrict__ cur_cnt) {
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < num_edges; e += (int64_t)gridDim.x * blockDim.x)
    if (rounded[e] == 1.f) atomicAdd(&cur_cnt[key_node[e]], 1);
}
__global__ void greedy_winner_kernel(const int64_t* __restrict__ key_node, const float* __restrict__ preds,
                                     const float* __restrict__ rounded, int64_t num_edges,
                                     const int32_t* __restrict__ init_cnt, const int32_t* __restrict__ cur_cnt,
                                     unsigned long long* __restrict__ winner) {
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < num_edges; e += (int64_t)gridDim.x * blockDim.x) {
    const int64_t v = key_node[e];
    if (rounded[e] == 1.f && init_cnt[v] > 1 && cur_cnt[v] > 1) {
      // predictions of active edges are > 0.5: their bit patterns order like the values; ties go to the smaller edge id
      const unsigned long long key = ((unsigned long long)__float_as_uint(preds[e]) << 32) | (0xffffffffu - (uint32_t)e);
      atomicMax(&winner[v], key);
    }
  }
}
__global__ void greedy_apply_kernel(const int64_t* __restrict__ key_node, int64_t num_edges,
                                    const int32_t* __restrict__ init_cnt, const int32_t* __restrict__ cur_cnt,
                                    const unsigned long long* __restrict__ winner, float* __restrict__ rounded) {
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < num_edges; e += (int64_t)gridDim.x * blockDim.x) {
    const int64_t v = key_node[e];
    if (init_cnt[v] > 1 && cur_cnt[v] > 1) {                    // projectors.py:58-59: all edges of the constraint off, winner on
      const uint32_t win = 0xffffffffu - (uint32_t)(winner[v] & 0xffffffffull);
      rounded[e] = (uint32_t)e == win ? 1.f : 0.f;
    }
  }
}

// Connected components of the graph of active edges: parent[v] converges to the smallest node index of v's component
// (hooking of larger roots under smaller ones + pointer jumping).  One CTA, so that the fixed point is detected with a
// block barrier instead of a host round trip per iteration; run once per sequence.
__global__ void __launch_bounds__(1024) cc_parent_kernel(const int64_t* __restrict__ row, const int64_t* __restrict__ col,
                                                         const float* __restrict__ edge_vals, int64_t num_edges,
                                                         int64_t num_nodes, int32_t* __restrict__ parent) {
  __shared__ int changed;
  for (int64_t v = threadIdx.x; v < num_nodes; v += blockDim.x) parent[v] = (int32_t)v;
  __syncthreads();
  while (true) {
    if (threadIdx.x == 0) changed = 0;
    __syncthreads();
    for (int64_t e = threadIdx.x; e < num_edges; e += blockDim.x) {
      if (edge_vals[e] != 1.f) continue;
      int32_t ru = parent[row[e]], rv = parent[col[e]];
      while (ru != parent[ru]) ru = parent[ru];
      while (rv != parent[rv]) rv = parent[rv];
      if (ru != rv) {
        const int32_t hi = ru > rv ? ru : rv, lo = ru > rv ? rv : ru;
        atomicMin(&parent[hi], lo);
        changed = 1;
      }
    }
    __syncthreads();
    for (int64_t v = threadIdx.x; v < num_nodes; v += blockDim.x) {      // full compression
      int32_t r = parent[v];
      while (r != parent[r]) r = parent[r];
      parent[v] = r;
    }
    __syncthreads();
    if (!changed) break;
    __syncthreads();
  }
}
__global__ void cc_root_flag_kernel(const int32_t* __restrict__ parent, int64_t num_nodes, int32_t* __restrict__ flag) {
  for (int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; v < num_nodes; v += (int64_t)gridDim.x * blockDim.x)
    flag[v] = parent[v] == (int32_t)v ? 1 : 0;
}
// scipy numbers the components in the order of their smallest node: label = number of roots below the node's root
__global__ void cc_label_kernel(const int32_t* __restrict__ parent, const int32_t* __restrict__ root_rank, int64_t num_nodes,
                                int64_t* __restrict__ labels) {
  for (int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; v < num_nodes; v += (int64_t)gridDim.x * blockDim.x)
    labels[v] = root_rank[parent[v]];
}

unsigned grid_for_items(int64_t n) {
  const int64_t g = ceil_div(n > 0 ? n : 1, 256);
  const int64_t cap = (int64_t)sm_count() * 8;
  return (unsigned)(g < cap ? g : cap);
}

struct FlowWs { int32_t* cnt_out; int32_t* cnt_in; int32_t* has_out; int32_t* has_in; int32_t* cur; unsigned long long* winner;
                unsigned long long* counts; int32_t* bad; };
int64_t carve_flow(void* ws, int64_t n, FlowWs* out) {
  Carver cv(ws);
  FlowWs w;
  w.cnt_out = cv.take<int32_t>(n); w.cnt_in = cv.take<int32_t>(n);
  w.has_out = cv.take<int32_t>(n); w.has_in = cv.take<int32_t>(n);
  w.cur = cv.take<int32_t>(n);
  w.winner = cv.take<unsigned long long>(n);
  w.counts = cv.take<unsigned long long>(4);
  w.bad = cv.take<int32_t>(4);
  if (out) *out = w;
  return cv.off;
}

int flow_stats(const int64_t* row, const int64_t* col, const float* vals, int64_t e, int64_t n, int undirected, const FlowWs& w,
               float* flow_in, float* flow_out, cudaStream_t s) {
  MPN_CUDA(cudaMemsetAsync(w.cnt_out, 0, (char*)w.cur - (char*)w.cnt_out, s));          // the four count / presence arrays
  MPN_CUDA(cudaMemsetAsync(w.counts, 0, 32, s));
  MPN_CUDA(cudaMemsetAsync(w.bad, 0, 16, s));
  if (e > 0) {
    flow_count_kernel<<<grid_for_items(e), 256, 0, s>>>(row, col, vals, e, undirected, w.cnt_out, w.cnt_in, w.has_out, w.has_in, w.bad);
    count_launch();
  }
  flow_stats_kernel<<<grid_for_items(n), 256, 0, s>>>(w.cnt_out, w.cnt_in, w.has_out, w.has_in, n, undirected ? 2.f : 1.f, flow_in,
                                                     flow_out, w.counts);
  count_launch();
  MPN_LAUNCH_CHECK();
  return MPN_OK;
}

}  // namespace
}  // namespace mpn

using namespace mpn;

extern "C" {

int64_t mpn_rounding_workspace(int64_t num_nodes) { return carve_flow(nullptr, num_nodes > 0 ? num_nodes : 1, nullptr) + 256; }

int mpn_constr_satisfaction(const int64_t* row, const int64_t* col, const float* edges_out, int64_t num_edges,
                            int64_t num_nodes, int undirected_edges, void* workspace, float* flow_in, float* flow_out,
                            int64_t* h_counts, void* stream) {
  MPN_CHECK_ARG(num_edges >= 0 && num_nodes >= 0 && workspace && h_counts, "constr_satisfaction: bad arguments");
  MPN_CHECK_ARG(num_edges == 0 || (row && col && edges_out), "constr_satisfaction: null edge arrays");
  MPN_CHECK_ARG(num_nodes < 2147483647LL, "constr_satisfaction: too many nodes");
  cudaStream_t s = as_stream(stream);
  FlowWs w;
  carve_flow(workspace, num_nodes > 0 ? num_nodes : 1, &w);
  int rc = flow_stats(row, col, edges_out, num_edges, num_nodes, undirected_edges, w, flow_in, flow_out, s);
  if (rc) return rc;
  unsigned long long h[4];
  int32_t bad = 0;
  MPN_CUDA(cudaMemcpyAsync(h, w.counts, 32, cudaMemcpyDeviceToHost, s));
  MPN_CUDA(cudaMemcpyAsync(&bad, w.bad, 4, cudaMemcpyDeviceToHost, s));
  MPN_CUDA(cudaStreamSynchronize(s));
  MPN_CHECK_ARG(bad == 0, "constr_satisfaction: edges_out must be binarised (exactly 0 or 1)");
  h_counts[0] = (int64_t)h[0]; h_counts[1] = (int64_t)h[1]; h_counts[2] = (int64_t)h[2];
  return MPN_OK;
}

int mpn_greedy_project(const int64_t* row, const int64_t* col, const float* edge_preds, int64_t num_edges,
                       int64_t num_nodes, void* workspace, float* round_preds, int64_t* h_counts, void* stream) {
  MPN_CHECK_ARG(num_edges >= 0 && num_nodes >= 0 && workspace && h_counts, "greedy_project: bad arguments");
  MPN_CHECK_ARG(num_edges == 0 || (row && col && edge_preds && round_preds), "greedy_project: null edge arrays");
  MPN_CHECK_ARG(num_nodes < 2147483647LL && num_edges < 4294967295LL, "greedy_project: graph too large");
  cudaStream_t s = as_stream(stream);
  FlowWs w;
  carve_flow(workspace, num_nodes > 0 ? num_nodes : 1, &w);
  const unsigned eg = grid_for_items(num_edges);
  if (num_edges > 0) { round_kernel<<<eg, 256, 0, s>>>(edge_preds, num_edges, round_preds); count_launch(); }
  int rc = flow_stats(row, col, round_preds, num_edges, num_nodes, 0, w, nullptr, nullptr, s);      // projectors.py:22-25
  if (rc) return rc;
  unsigned long long h[4];
  MPN_CUDA(cudaMemcpyAsync(h, w.counts, 32, cudaMemcpyDeviceToHost, s));
  if (num_edges > 0) {
    for (int side = 0; side < 2; ++side) {                                     // outgoing constraints, then incoming
      const int64_t* key = side == 0 ? row : col;
      const int32_t* init = side == 0 ? w.cnt_out : w.cnt_in;
      MPN_CUDA(cudaMemsetAsync(w.cur, 0, sizeof(int32_t) * num_nodes, s));
      MPN_CUDA(cudaMemsetAsync(w.winner, 0, sizeof(unsigned long long) * num_nodes, s));
      greedy_count_kernel<<<eg, 256, 0, s>>>(key, round_preds, num_edges, w.cur); count_launch();
      greedy_winner_kernel<<<eg, 256, 0, s>>>(key, edge_preds, round_preds, num_edges, init, w.cur, w.winner); count_launch();
      greedy_apply_kernel<<<eg, 256, 0, s>>>(key, num_edges, init, w.cur, w.winner, round_preds); count_launch();
    }
    MPN_LAUNCH_CHECK();
  }
  MPN_CUDA(cudaStreamSynchronize(s));
  h_counts[0] = (int64_t)h[0]; h_counts[1] = (int64_t)h[1]; h_counts[2] = (int64_t)h[2];
  return MPN_OK;
}

int64_t mpn_connected_components_workspace(int64_t num_nodes) {
  const int64_t n = num_nodes > 0 ? num_nodes : 1;
  return 3 * align_up((n + 1) * 4, 256) + 256;
}

int mpn_connected_components(const int64_t* row, const int64_t* col, const float* edge_vals, int64_t num_edges,
                             int64_t num_nodes, void* workspace, int64_t* labels, int64_t* h_num_components, void* stream) {
  MPN_CHECK_ARG(num_edges >= 0 && num_nodes >= 0 && workspace && h_num_components, "connected_components: bad arguments");
  MPN_CHECK_ARG(num_edges == 0 || (row && col && edge_vals), "connected_components: null edge arrays");
  MPN_CHECK_ARG(num_nodes < 2147483647LL, "connected_components: too many nodes");
  cudaStream_t s = as_stream(stream);
  *h_num_components = 0;
  if (num_nodes == 0) return MPN_OK;
  MPN_CHECK_ARG(labels != nullptr, "connected_components: null labels");
  Carver cv(workspace);
  int32_t* parent = cv.take<int32_t>(num_nodes + 1);
  int32_t* flag = cv.take<int32_t>(num_nodes + 1);
  int32_t* rank = cv.take<int32_t>(num_nodes + 1);
  cc_parent_kernel<<<1, 1024, 0, s>>>(row, col, edge_vals, num_edges, num_nodes, parent); count_launch();
  cc_root_flag_kernel<<<grid_for_items(num_nodes), 256, 0, s>>>(parent, num_nodes, flag); count_launch();
  MPN_LAUNCH_CHECK();
  int rc = exclusive_scan_i32(flag, rank, num_nodes, s);
  if (rc) return rc;
  cc_label_kernel<<<grid_for_items(num_nodes), 256, 0, s>>>(parent, rank, num_nodes, labels); count_launch();
  MPN_LAUNCH_CHECK();
  int32_t total = 0;
  MPN_CUDA(cudaMemcpyAsync(&total, rank + num_nodes, 4, cudaMemcpyDeviceToHost, s));
  MPN_CUDA(cudaStreamSynchronize(s));
  *h_num_components = total;
  return MPN_OK;
}

}  // extern "C"
