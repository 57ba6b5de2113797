// Slot layout of the directed edge list for the fused message-passing kernels:
// flow_out group (row<col) then flow_in group (row>col), each sorted by row, stable.
// The sort is cub::DeviceRadixSort (ships with the CUDA toolkit); it runs once per graph.
#include <cub/device/device_radix_sort.cuh>

#include "common.cuh"

namespace mpn {

__global__ void layout_keys_kernel(const int64_t* __restrict__ ei, int64_t e, int64_t n,
                                   uint32_t* __restrict__ keys, int32_t* __restrict__ vals,
                                   unsigned long long* __restrict__ bad) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < e;
       i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = ei[i], c = ei[e + i];
    if (r == c || r < 0 || c < 0 || r >= n || c >= n) atomicAdd(bad, 1ull);
    keys[i] = (uint32_t)((r < c ? 0 : n) + (r < 0 || r >= n ? 0 : r));
    vals[i] = (int32_t)i;
  }
}

__global__ void layout_finish_kernel(const int64_t* __restrict__ ei, int64_t e, int64_t n,
                                     const uint32_t* __restrict__ skeys,
                                     const int32_t* __restrict__ svals, int32_t* __restrict__ srow,
                                     int32_t* __restrict__ scol) {
  for (int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; s < e;
       s += (int64_t)gridDim.x * blockDim.x) {
    const uint32_t k = skeys[s];
    srow[s] = (int32_t)(k >= (uint32_t)n ? k - (uint32_t)n : k);
    scol[s] = (int32_t)ei[e + svals[s]];
  }
}

// ptr[r] = first slot whose key >= base + r  (r in [0, n]); lower bound over sorted keys.
__global__ void layout_ptr_kernel(const uint32_t* __restrict__ skeys, int64_t e, int64_t n,
                                  int32_t* __restrict__ out_ptr, int32_t* __restrict__ in_ptr) {
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < 2 * (n + 1);
       t += (int64_t)gridDim.x * blockDim.x) {
    const bool in = t >= n + 1;
    const int64_t r = in ? t - (n + 1) : t;
    const uint64_t target = (uint64_t)(in ? n : 0) + (uint64_t)r;   // may equal 2n -> e
    int64_t lo = 0, hi = e;
    while (lo < hi) {
      const int64_t mid = (lo + hi) >> 1;
      if ((uint64_t)skeys[mid] < target) lo = mid + 1; else hi = mid;
    }
    (in ? in_ptr : out_ptr)[r] = (int32_t)lo;
  }
}

static int key_bits(int64_t n) {
  int b = 1;
  while (((int64_t)1 << b) < 2 * n + 1 && b < 32) ++b;
  return b;
}

}  // namespace mpn

using namespace mpn;

extern "C" {

int64_t mpn_edge_layout_workspace(int64_t e, int64_t n) {
  size_t temp = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, temp, (const uint32_t*)nullptr, (uint32_t*)nullptr,
                                  (const int32_t*)nullptr, (int32_t*)nullptr, (int)(e > 0 ? e : 1), 0,
                                  key_bits(n));
  return align_up((int64_t)temp, 256) + 3 * align_up((e > 0 ? e : 1) * 4, 256) + 512;
}

int mpn_edge_layout_build(const int64_t* edge_index, int64_t e, int64_t n, void* ws,
                          int32_t* slot_row, int32_t* slot_col, int32_t* slot_edge,
                          int32_t* out_ptr, int32_t* in_ptr, int64_t* h_num_out, void* stream) {
  MPN_CHECK_ARG(e >= 0 && n >= 0 && n < (1ll << 30) && e < (1ll << 31), "edge_layout_build: sizes out of range");
  MPN_CHECK_ARG(out_ptr && in_ptr && h_num_out, "edge_layout_build: null pointer");
  cudaStream_t s = as_stream(stream);
  if (e == 0) {
    MPN_CUDA(cudaMemsetAsync(out_ptr, 0, 4 * (n + 1), s));
    MPN_CUDA(cudaMemsetAsync(in_ptr, 0, 4 * (n + 1), s));
    *h_num_out = 0;
    return MPN_OK;
  }
  MPN_CHECK_ARG(edge_index && ws && slot_row && slot_col && slot_edge, "edge_layout_build: null pointer");
  Carver cv(ws);
  unsigned long long* bad = cv.take<unsigned long long>(1);
  uint32_t* keys = cv.take<uint32_t>(e);
  uint32_t* skeys = cv.take<uint32_t>(e);
  int32_t* vals = cv.take<int32_t>(e);
  size_t temp = 0;
  const int bits = key_bits(n);
  cub::DeviceRadixSort::SortPairs(nullptr, temp, (const uint32_t*)keys, skeys, (const int32_t*)vals,
                                  slot_edge, (int)e, 0, bits, s);
  void* temp_ptr = cv.take<char>((int64_t)temp);
  MPN_CUDA(cudaMemsetAsync(bad, 0, 8, s));
  const unsigned grid = (unsigned)std::min<int64_t>(ceil_div(e, 256), (int64_t)sm_count() * 8);
  layout_keys_kernel<<<grid, 256, 0, s>>>(edge_index, e, n, keys, vals, bad); count_launch();
  MPN_CUDA(cub::DeviceRadixSort::SortPairs(temp_ptr, temp, (const uint32_t*)keys, skeys,
                                           (const int32_t*)vals, slot_edge, (int)e, 0, bits, s));
  layout_finish_kernel<<<grid, 256, 0, s>>>(edge_index, e, n, skeys, slot_edge, slot_row, slot_col); count_launch();
  layout_ptr_kernel<<<(unsigned)ceil_div(2 * (n + 1), 256), 256, 0, s>>>(skeys, e, n, out_ptr, in_ptr); count_launch();
  MPN_LAUNCH_CHECK();
  unsigned long long h_bad = 0;
  int32_t h_out = 0;
  MPN_CUDA(cudaMemcpyAsync(&h_bad, bad, 8, cudaMemcpyDeviceToHost, s));
  MPN_CUDA(cudaMemcpyAsync(&h_out, out_ptr + n, 4, cudaMemcpyDeviceToHost, s));
  MPN_CUDA(cudaStreamSynchronize(s));
  MPN_CHECK_ARG(h_bad == 0, "edge_layout_build: %llu edges are self-loops or out of range [0,%lld)",
                h_bad, (long long)n);
  *h_num_out = h_out;
  return MPN_OK;
}

}  // extern "C"
