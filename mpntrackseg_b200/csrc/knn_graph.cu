// Fused, batched construction of KNN-pruned window graphs (training-mode semantics of
// MOTGraph._get_edge_ixs, data/mot_graph.py:195-221): for G independent windows at once
//   1. dense ReID distance blocks  D_g[i][j] = || reid_i - reid_j + 1e-6 ||_2  (i < j, mirrored;
//      pairs of the same frame or farther apart than max_frame_dist = +inf),
//   2. per-row k-th smallest (distance, index) threshold                         (utils/graph.py:65-70),
//   3. keep (i<j) iff in_k(i,j) AND/OR in_k(j,i); pairs are emitted sorted by (i, j) (utils/graph.py:73-85),
// with a single host synchronisation (the pair count).  Node ids are batch-global.
#include <math.h>

#include "common.cuh"

namespace mpn {

constexpr unsigned kFullMask = 0xffffffffu;
constexpr int DT = 64;        // distance tile (rows x cols per CTA)
constexpr int DK = 32;        // feature slab

__device__ __forceinline__ uint32_t order_key_f(float f) {
  uint32_t u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

__device__ __forceinline__ bool frames_connect(int64_t fi, int64_t fj, int64_t max_dist) {
  const int64_t d = fi > fj ? fi - fj : fj - fi;
  return d > 0 && (max_dist < 0 || d <= max_dist);
}

// blockIdx.y = window, blockIdx.x = upper-triangular tile id inside the window.
__global__ void __launch_bounds__(256) dist_blocks_kernel(const float* __restrict__ reid, int64_t dim,
                                                          const int64_t* __restrict__ frame,
                                                          const int64_t* __restrict__ gptr,
                                                          const int64_t* __restrict__ doff, int64_t max_dist,
                                                          float* __restrict__ dense) {
  const int g = blockIdx.y;
  const int64_t n0 = gptr[g], n = gptr[g + 1] - n0;
  const int nt = (int)((n + DT - 1) / DT);
  // tile id -> (ti <= tj) over the upper triangle, row-major
  int t = blockIdx.x;
  if (t >= nt * (nt + 1) / 2) return;
  int ti = 0;
  while (t >= nt - ti) { t -= nt - ti; ++ti; }
  const int tj = ti + t;
  float* D = dense + doff[g];
  __shared__ float sa[DK][DT + 1];
  __shared__ float sb[DK][DT + 1];
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;          // 16 x 16 threads, 4 x 4 outputs each
  float acc[4][4] = {};
  const float eps = 1e-6f;
  for (int64_t k0 = 0; k0 < dim; k0 += DK) {
    for (int idx = threadIdx.x; idx < DT * DK; idx += 256) {
      const int r = idx / DK, kk = idx % DK;
      const int64_t ia = (int64_t)ti * DT + r, ib = (int64_t)tj * DT + r, k = k0 + kk;
      sa[kk][r] = (ia < n && k < dim) ? reid[(n0 + ia) * dim + k] : 0.f;
      sb[kk][r] = (ib < n && k < dim) ? reid[(n0 + ib) * dim + k] : 0.f;
    }
    __syncthreads();
#pragma unroll 8
    for (int kk = 0; kk < DK; ++kk) {
      float a[4], b[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) { a[q] = sa[kk][ty * 4 + q]; b[q] = sb[kk][tx * 4 + q]; }
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float d = (a[i] - b[j]) + eps;                      // F.pairwise_distance: x1 - x2 + eps
          acc[i][j] = fmaf(d, d, acc[i][j]);
        }
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int64_t li = (int64_t)ti * DT + ty * 4 + i;
    if (li >= n) continue;
    const int64_t fi = frame[n0 + li];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int64_t lj = (int64_t)tj * DT + tx * 4 + j;
      if (lj >= n || lj <= li) continue;                            // strictly upper triangle; mirrored below
      const float v = frames_connect(fi, frame[n0 + lj], max_dist) ? sqrtf(acc[i][j]) : INFINITY;
      D[li * n + lj] = v;
      D[lj * n + li] = v;                                           // utils/graph.py:58-60 (symmetric fill)
    }
  }
  // diagonal entries (no self edges)
  if (ti == tj && threadIdx.x < DT) {
    const int64_t li = (int64_t)ti * DT + threadIdx.x;
    if (li < n) D[li * n + li] = INFINITY;
  }
}

// One CTA per node (batch-global row): radix-select the k-th smallest (key, local index) of its row.
__global__ void __launch_bounds__(256) batch_row_kth_kernel(const float* __restrict__ dense,
                                                            const int64_t* __restrict__ gptr, int64_t num_graphs,
                                                            const int64_t* __restrict__ doff, int64_t k,
                                                            const int32_t* __restrict__ row_mask,
                                                            uint32_t* __restrict__ thr_key,
                                                            int32_t* __restrict__ thr_idx) {
  __shared__ int hist[256];
  __shared__ uint32_t s_prefix;
  __shared__ int s_remaining, s_count;
  __shared__ int32_t s_found;
  const int64_t i = blockIdx.x;
  if (row_mask != nullptr && row_mask[i] == 0) return;
  int64_t lo = 0, hi = num_graphs;                                  // window of node i
  while (hi - lo > 1) { const int64_t mid = (lo + hi) >> 1; if (gptr[mid] <= i) lo = mid; else hi = mid; }
  const int64_t n0 = gptr[lo], n = gptr[lo + 1] - n0;
  const float* rowp = dense + doff[lo] + (i - n0) * n;
  if (k >= n) {                                                     // every rank < k
    if (threadIdx.x == 0) { thr_key[i] = 0xffffffffu; thr_idx[i] = (int32_t)n; }
    return;
  }
  if (threadIdx.x == 0) { s_prefix = 0u; s_remaining = (int)k; s_found = (int32_t)(n - 1); }
  uint32_t mask = 0u;
  const int lane = threadIdx.x & 31;
  for (int shift = 24; shift >= 0; shift -= 8) {
    hist[threadIdx.x] = 0;
    __syncthreads();
    const uint32_t prefix = s_prefix;
    for (int64_t j = threadIdx.x; j < n; j += blockDim.x) {
      const uint32_t key = order_key_f(rowp[j]);
      if ((key & mask) == prefix) atomicAdd(&hist[(key >> shift) & 255u], 1);
    }
    __syncthreads();
    if (threadIdx.x < 32) {
      // bin holding the `remaining`-th element: lane l owns bins 8l..8l+7, warp scan over the lane totals
      int h[8], tot = 0;
#pragma unroll
      for (int q = 0; q < 8; ++q) { h[q] = hist[8 * lane + q]; tot += h[q]; }
      int incl = tot;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) { const int v = __shfl_up_sync(kFullMask, incl, d); if (lane >= d) incl += v; }
      const int rem0 = s_remaining;
      const bool here = incl >= rem0 && incl - tot < rem0;              // exactly one lane
      if (here) {
        int rem = rem0 - (incl - tot), b = 0;
#pragma unroll
        for (int q = 0; q < 8; ++q) { if (h[q] >= rem) { b = q; break; } rem -= h[q]; }
        s_remaining = rem;
        s_prefix = prefix | ((uint32_t)(8 * lane + b) << shift);
        s_count = h[b];
      }
    }
    mask |= 255u << shift;
    __syncthreads();
  }
  const uint32_t tau = s_prefix;
  const int need = s_remaining;
  if (s_count == need) {
    // no tie straddles the boundary: the k-th element (ties by ascending index) is the LAST element equal to tau
    if (threadIdx.x == 0) s_found = -1;
    __syncthreads();
    int32_t mine = -1;
    for (int64_t j = threadIdx.x; j < n; j += blockDim.x)
      if (order_key_f(rowp[j]) == tau) mine = (int32_t)j;                // ascending j per thread
    if (mine >= 0) atomicMax(&s_found, mine);
    __syncthreads();
    if (threadIdx.x == 0) { thr_key[i] = tau; thr_idx[i] = s_found; }
    return;
  }
  if (threadIdx.x < 32) {                                           // a tie straddles the boundary: first `need` in index order
    int seen = 0;
    int32_t found = (int32_t)(n - 1);
    for (int64_t j0 = 0; j0 < n; j0 += 32) {
      const int64_t j = j0 + lane;
      const bool eq = j < n && order_key_f(rowp[j]) == tau;
      const unsigned m = __ballot_sync(kFullMask, eq);
      const int c = __popc(m);
      if (seen + c >= need) {
        unsigned mm = m;
        for (int q = 1; q < need - seen; ++q) mm &= mm - 1;
        found = (int32_t)(j0 + __ffs(mm) - 1);
        break;
      }
      seen += c;
    }
    if (lane == 0) { thr_key[i] = tau; thr_idx[i] = found; }
  }
}

// warp per node i: kept pairs (i, j>i) of its window; kFill=false counts, true writes at row_start.
template <bool kFill>
__global__ void knn_pairs_kernel(const float* __restrict__ dense, const int64_t* __restrict__ gptr,
                                 int64_t num_graphs, const int64_t* __restrict__ doff,
                                 const int64_t* __restrict__ frame, int64_t num_nodes, int64_t max_dist,
                                 const uint32_t* __restrict__ thr_key, const int32_t* __restrict__ thr_idx,
                                 int prune, int reciprocal, int64_t* __restrict__ row_cnt_or_start,
                                 int64_t* __restrict__ out_row, int64_t* __restrict__ out_col,
                                 float* __restrict__ out_dist) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t i = warp; i < num_nodes; i += nwarps) {
    int64_t lo = 0, hi = num_graphs;
    while (hi - lo > 1) { const int64_t mid = (lo + hi) >> 1; if (gptr[mid] <= i) lo = mid; else hi = mid; }
    const int64_t n0 = gptr[lo], n = gptr[lo + 1] - n0, li = i - n0;
    const float* D = dense + doff[lo];
    const int64_t fi = frame[i];
    const uint32_t ki = prune ? thr_key[i] : 0u;
    const int32_t ti = prune ? thr_idx[i] : 0;
    int64_t cursor = kFill ? row_cnt_or_start[i] : 0;
    for (int64_t j0 = li + 1; j0 < n; j0 += 32) {
      const int64_t lj = j0 + lane;
      bool ok = lj < n && frames_connect(fi, frame[n0 + lj], max_dist);
      float d = 0.f;
      if (ok) {
        d = D[li * n + lj];
        if (prune) {
          const uint32_t key = order_key_f(d);
          const bool a = key < ki || (key == ki && lj <= ti);
          const uint32_t kj = thr_key[n0 + lj];
          const bool b = key < kj || (key == kj && li <= thr_idx[n0 + lj]);   // D is symmetric
          ok = reciprocal ? (a && b) : (a || b);
        }
      }
      const unsigned m = __ballot_sync(kFullMask, ok);
      if (kFill && ok) {
        const int64_t pos = cursor + __popc(m & ((1u << lane) - 1u));
        out_row[pos] = i;
        out_col[pos] = n0 + lj;
        out_dist[pos] = d;
      }
      cursor += __popc(m);
    }
    if (!kFill && lane == 0) row_cnt_or_start[i] = cursor;
  }
}

__global__ void graph_offsets_kernel(const int64_t* __restrict__ gptr, int64_t num_graphs, int64_t* __restrict__ doff) {
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    int64_t acc = 0;
    for (int64_t g = 0; g < num_graphs; ++g) { doff[g] = acc; const int64_t n = gptr[g + 1] - gptr[g]; acc += n * n; }
    doff[num_graphs] = acc;
  }
}

__global__ void graph_pair_ptr_kernel(const int64_t* __restrict__ gptr, int64_t num_graphs, int64_t num_nodes,
                                      const int64_t* __restrict__ row_start, int64_t* __restrict__ out) {
  for (int64_t g = blockIdx.x * blockDim.x + threadIdx.x; g <= num_graphs; g += (int64_t)gridDim.x * blockDim.x) {
    const int64_t node = g < num_graphs ? gptr[g] : num_nodes;
    out[g] = row_start[node];
  }
}

int64_t gram_workspace_bytes(int64_t num_nodes, int64_t total_tiles, int64_t num_graphs, int64_t dim);
int gram_dist_blocks(const float* reid, int64_t dim, const int64_t* frame, const int64_t* gptr, const int64_t* h_gptr,
                     int64_t num_graphs, const int64_t* doff, int64_t max_dist, void* ws, float* dense,
                     int32_t* status, float** norm2_out, int32_t** amb_out, int32_t** amb_count_out, cudaStream_t s);
int gram_fix_ambiguous(const float* reid, int64_t dim, const int64_t* frame, const int64_t* gptr, int64_t num_graphs,
                       int64_t num_nodes, const int64_t* doff, int64_t max_dist, float* dense, const float* norm2,
                       const uint32_t* thr_key, const int32_t* thr_idx, float beta, int32_t* amb, int32_t* amb_count,
                       cudaStream_t s);

}  // namespace mpn

using namespace mpn;

extern "C" {

int64_t mpn_knn_graph_workspace(int64_t num_nodes, int64_t sum_sq_nodes, int64_t num_graphs) {
  const int64_t base = align_up(sum_sq_nodes * 4, 256) + align_up((num_graphs + 1) * 8, 256) +
                       2 * align_up(num_nodes * 4, 256) + align_up((num_nodes + 1) * 8, 256) + 2048;
  // tensor-core path: packed fp16 hi/lo image of the embeddings (dim <= 512 covered) + per-node scalars
  return base + gram_workspace_bytes(num_nodes, num_nodes / 128 + num_graphs, num_graphs, 512) + 256;
}

int mpn_knn_graph_pairs(const int64_t* frame, const int64_t* gptr, const int64_t* h_gptr, int64_t num_graphs,
                        const float* reid, int64_t dim, int64_t top_k, int reciprocal, int64_t max_frame_dist,
                        int use_tensor_cores, void* ws, int64_t capacity, int64_t* out_row, int64_t* out_col,
                        float* out_dist, int64_t* graph_pair_ptr, int64_t* h_graph_pair_ptr, int64_t* h_stats,
                        void* stream) {
  MPN_CHECK_ARG(num_graphs >= 1 && h_gptr && gptr && frame && reid && ws, "knn_graph_pairs: null / empty arguments");
  MPN_CHECK_ARG(out_row && out_col && out_dist && graph_pair_ptr && h_graph_pair_ptr, "knn_graph_pairs: null outputs");
  MPN_CHECK_ARG(num_graphs <= 65535, "knn_graph_pairs: at most 65535 windows per call");
  cudaStream_t s = as_stream(stream);
  const int64_t n = h_gptr[num_graphs];
  int64_t sum_sq = 0, max_n = 0;
  for (int64_t g = 0; g < num_graphs; ++g) {
    const int64_t ng = h_gptr[g + 1] - h_gptr[g];
    MPN_CHECK_ARG(ng >= 0, "knn_graph_pairs: node_graph_ptr must be non-decreasing");
    sum_sq += ng * ng;
    max_n = ng > max_n ? ng : max_n;
  }
  if (n == 0) { for (int64_t g = 0; g <= num_graphs; ++g) h_graph_pair_ptr[g] = 0; return MPN_OK; }
  Carver cv(ws);
  float* dense = cv.take<float>(sum_sq);
  int64_t* doff = cv.take<int64_t>(num_graphs + 1);
  uint32_t* tk = cv.take<uint32_t>(n);
  int32_t* ti = cv.take<int32_t>(n);
  int64_t* row_start = cv.take<int64_t>(n + 1);
  const int prune = top_k >= 0 ? 1 : 0;
  // Gram distances only pre-rank (they are repaired / recomputed exactly below); without pruning every
  // distance is an output, so the exact kernel is used directly.
  const bool tc = use_tensor_cores && prune && dim % 64 == 0 && dim <= 512 && reinterpret_cast<uintptr_t>(reid) % 16 == 0;
  int32_t* status = reinterpret_cast<int32_t*>(cv.take<int32_t>(4));
  float* norm2 = nullptr;
  int32_t* amb = nullptr;
  int32_t* amb_count = nullptr;
  int rc = MPN_OK;
  MPN_CUDA(cudaMemsetAsync(status, 0, 16, s));

  graph_offsets_kernel<<<1, 32, 0, s>>>(gptr, num_graphs, doff); count_launch();
  if (tc) {
    rc = gram_dist_blocks(reid, dim, frame, gptr, h_gptr, num_graphs, doff, max_frame_dist,
                          static_cast<char*>(ws) + cv.off, dense, status, &norm2, &amb, &amb_count, s);
    if (rc) return rc;
  } else {
    const int64_t nt = ceil_div(max_n, DT);
    dim3 grid((unsigned)(nt * (nt + 1) / 2), (unsigned)num_graphs);
    dist_blocks_kernel<<<grid, 256, 0, s>>>(reid, dim, frame, gptr, doff, max_frame_dist, dense); count_launch();
  }
  if (prune) {
    batch_row_kth_kernel<<<(unsigned)n, 256, 0, s>>>(dense, gptr, num_graphs, doff, top_k, nullptr, tk, ti); count_launch();
    if (tc) {                                                       // repair rows whose top-k set is not certain
      rc = gram_fix_ambiguous(reid, dim, frame, gptr, num_graphs, n, doff, max_frame_dist, dense, norm2, tk, ti, 2e-6f,
                              amb, amb_count, s);
      if (rc) return rc;
      batch_row_kth_kernel<<<(unsigned)n, 256, 0, s>>>(dense, gptr, num_graphs, doff, top_k, amb, tk, ti); count_launch();
    }
  }
  const unsigned wgrid = (unsigned)std::min<int64_t>(ceil_div(n * 32, 256), (int64_t)sm_count() * 16);
  knn_pairs_kernel<false><<<wgrid, 256, 0, s>>>(dense, gptr, num_graphs, doff, frame, n, max_frame_dist, tk, ti, prune,
                                               reciprocal, row_start, nullptr, nullptr, nullptr); count_launch();
  MPN_LAUNCH_CHECK();
  rc = exclusive_scan_i64(row_start, row_start, n, s);
  if (rc) return rc;
  graph_pair_ptr_kernel<<<(unsigned)ceil_div(num_graphs + 1, 256), 256, 0, s>>>(gptr, num_graphs, n, row_start, graph_pair_ptr);
  count_launch();
  MPN_CUDA(cudaMemcpyAsync(h_graph_pair_ptr, graph_pair_ptr, 8 * (num_graphs + 1), cudaMemcpyDeviceToHost, s));
  int32_t h_flags[2] = {0, 0};                                      // fp16 overflow in the Gram kernel, #ambiguous rows
  MPN_CUDA(cudaMemcpyAsync(&h_flags[0], status, 4, cudaMemcpyDeviceToHost, s));
  if (tc) MPN_CUDA(cudaMemcpyAsync(&h_flags[1], amb_count, 4, cudaMemcpyDeviceToHost, s));
  MPN_CUDA(cudaStreamSynchronize(s));
  if (tc && h_flags[0] != 0)                                        // embeddings beyond the fp16 range: exact kernel
    return mpn_knn_graph_pairs(frame, gptr, h_gptr, num_graphs, reid, dim, top_k, reciprocal, max_frame_dist, 0, ws,
                               capacity, out_row, out_col, out_dist, graph_pair_ptr, h_graph_pair_ptr, h_stats, stream);
  if (h_stats != nullptr) { h_stats[0] = tc ? 1 : 0; h_stats[1] = h_flags[1]; }
  const int64_t total = h_graph_pair_ptr[num_graphs];
  if (total > capacity) {
    set_error("knn_graph_pairs: %lld pairs exceed the output capacity %lld", (long long)total, (long long)capacity);
    return MPN_ENOSPC;
  }
  knn_pairs_kernel<true><<<wgrid, 256, 0, s>>>(dense, gptr, num_graphs, doff, frame, n, max_frame_dist, tk, ti, prune,
                                              reciprocal, row_start, out_row, out_col, out_dist); count_launch();
  MPN_LAUNCH_CHECK();
  if (tc && total > 0)                                              // the edge feature is always the exact distance
    return mpn_pair_reid_dist(reid, n, dim, out_row, out_col, total, out_dist, stream);
  return MPN_OK;
}

}  // extern "C"
