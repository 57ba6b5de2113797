// ReID distance blocks on the tcgen05 tensor cores (north-star item (a)):
//   d(i,j)^2 = ||a||^2 + ||b||^2 - 2 a.b + 2 eps (sum a - sum b) + K eps^2 ,   a.b from a split-fp16 Gram MMA
// (hi*hi + hi*lo + lo*hi, fp32 accumulation).  The Gram distances are APPROXIMATE (~1e-6 relative), so they
// only pre-rank: every row whose k-th / (k+1)-th neighbour are closer than an error band is recomputed with
// the exact fp32 formula of the reference (same kernel arithmetic as knn_graph.cu) and re-ranked, and the
// distances attached to the kept pairs are always recomputed exactly.  The kept edge set is therefore the
// one the exact path produces.
#include <math.h>

#include "common.cuh"
#include "tc_ptx.cuh"

namespace mpn {
namespace gram {

using namespace ptx;

constexpr int TS = 128, KC = 64, NTHREADS = 256;
constexpr int SLAB = TS * 32;                               // one K=16 step of a 128-row B tile
constexpr int CHUNK_BYTES = 2 * (KC / 16) * SLAB;           // hi + lo: 32 KB
constexpr int C_A = 0, C_D = 128;
constexpr int SM_W = 0;                                     // [2 groups][2 bufs][CHUNK_BYTES]
constexpr int SM_NB = 4 * CHUNK_BYTES;                      // float [2 groups][3][TS]: ||b||^2, sum b, frame of the column tile
constexpr int SM_BAR = SM_NB + 2 * 3 * TS * 4;
constexpr int SM_TMEM = SM_BAR + 4 * 8;
constexpr int SMEM_BYTES = SM_TMEM + 16;

__device__ __forceinline__ int slab_off(int n, int k16) {
  return (n >> 3) * 256 + (k16 >> 3) * 128 + (n & 7) * 16 + (k16 & 7) * 2;
}

// warp per node: packed fp16 hi/lo image of its row inside its 128-row tile, ||a||^2 and sum(a).
__global__ void pack_reid_kernel(const float* __restrict__ reid, int64_t dim, const int64_t* __restrict__ gptr,
                                 int64_t num_graphs, const int64_t* __restrict__ tile_off, int64_t num_nodes,
                                 uint8_t* __restrict__ img, float* __restrict__ norm2, float* __restrict__ sum1) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  const int nchunks = (int)(dim / KC);
  for (int64_t i = warp; i < num_nodes; i += nwarps) {
    int64_t lo = 0, hi = num_graphs;
    while (hi - lo > 1) { const int64_t mid = (lo + hi) >> 1; if (gptr[mid] <= i) lo = mid; else hi = mid; }
    const int64_t li = i - gptr[lo];
    const int64_t tile = tile_off[lo] + li / TS;
    const int n = (int)(li % TS);
    float s2 = 0.f, s1 = 0.f;
    for (int64_t k = lane; k < dim; k += 32) {
      const float v = reid[i * dim + k];
      s2 = fmaf(v, v, s2);
      s1 += v;
      const __half h = __float2half_rn(v), l = __float2half_rn(v - __half2float(h));
      uint8_t* base = img + (tile * nchunks + k / KC) * CHUNK_BYTES + ((k % KC) >> 4) * SLAB + slab_off(n, (int)(k & 15));
      *reinterpret_cast<__half*>(base) = h;
      *reinterpret_cast<__half*>(base + (KC / 16) * SLAB) = l;
    }
    for (int d = 16; d > 0; d >>= 1) { s2 += __shfl_xor_sync(0xffffffffu, s2, d); s1 += __shfl_xor_sync(0xffffffffu, s1, d); }
    if (lane == 0) { norm2[i] = s2; sum1[i] = s1; }
  }
}

struct GramArgs {
  const float* reid; int64_t dim;
  const int64_t* frame; const int64_t* gptr; const int64_t* doff; const int64_t* tile_off;
  const uint8_t* img; const float* norm2; const float* sum1;
  int64_t max_dist; float* dense; int32_t* status;
};

// blockIdx.y = window, blockIdx.x = pair of upper-triangular tiles (one per group).
__global__ void __launch_bounds__(NTHREADS, 1) gram_blocks_kernel(GramArgs a) {
  extern __shared__ __align__(128) uint8_t smem[];
  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
  const int g = warp >> 2, wq = warp & 3, gt = tid & (TS - 1);
  const int w = blockIdx.y;
  const int64_t n0 = a.gptr[w], n = a.gptr[w + 1] - n0;
  const int nt = (int)((n + TS - 1) / TS);
  const int ntri = nt * (nt + 1) / 2;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + SM_BAR) + 2 * g;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + SM_TMEM);
  if (tid == 0) {
    for (int i = 0; i < 4; ++i) mbar_init(reinterpret_cast<uint64_t*>(smem + SM_BAR) + i, 1);
    mbar_fence_init();
  }
  if (warp == 0) tmem_alloc<512>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tcol = __shfl_sync(0xffffffffu, *tmem_slot, 0) + (uint32_t)g * 256u;
  const uint32_t tlane = tcol + ((uint32_t)(wq * 32) << 16);
  const uint32_t wbuf = smem_u32(smem + SM_W + g * 2 * CHUNK_BYTES);
  float* s_nb = reinterpret_cast<float*>(smem + SM_NB) + g * 3 * TS;
  const uint64_t dbase = smem_desc_kmajor(0, 128, 256);
  const int nchunks = (int)(a.dim / KC);
  float* D = a.dense + a.doff[w];
  __half2 vmax = __floats2half2_rn(0.f, 0.f);
  uint32_t par[2] = {0, 0};
  const float eps = 1e-6f;

  for (int t = blockIdx.x * 2 + g; t < ntri; t += gridDim.x * 2) {
    int ti = 0, rem = t;
    while (rem >= nt - ti) { rem -= nt - ti; ++ti; }
    const int tj = ti + rem;
    int64_t li = (int64_t)ti * TS + gt;
    const bool valid = li < n;
    if (!valid) li = n - 1;
    const float4* xr = reinterpret_cast<const float4*>(a.reid + (n0 + li) * a.dim);
    const uint8_t* bimg = a.img + (a.tile_off[w] + tj) * (int64_t)nchunks * CHUNK_BYTES;
    {                                                           // column tile's norms / sums / frames
      const int64_t lj = (int64_t)tj * TS + gt;
      const bool vj = lj < n;
      s_nb[gt] = vj ? a.norm2[n0 + lj] : 0.f;
      s_nb[TS + gt] = vj ? a.sum1[n0 + lj] : 0.f;
      s_nb[2 * TS + gt] = vj ? __int_as_float((int)a.frame[n0 + lj]) : __int_as_float(INT_MIN);
    }
    auto load_w = [&](int c) {
      const uint8_t* src = bimg + (int64_t)c * CHUNK_BYTES;
      const uint32_t dst = wbuf + (c & 1) * CHUNK_BYTES;
#pragma unroll
      for (int j = 0; j < CHUNK_BYTES / 16 / TS; ++j) cp_async16(dst + (j * TS + gt) * 16, src + (j * TS + gt) * 16);
    };
    float4 xv[KC / 4];
    load_w(0);
#pragma unroll
    for (int j = 0; j < KC / 4; ++j) xv[j] = __ldg(xr + j);
    for (int c = 0; c < nchunks; ++c) {
      const uint32_t abuf = tlane + C_A + (c & 1) * 64;
#pragma unroll
      for (int q = 0; q < KC / 16; ++q) {
        uint32_t hi[8], lo[8];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float4 v = xv[4 * q + j];
          split2(v.x, v.y, hi[2 * j], lo[2 * j]);
          split2(v.z, v.w, hi[2 * j + 1], lo[2 * j + 1]);
          vmax = __hmax2(vmax, __habs2(*reinterpret_cast<const __half2*>(&hi[2 * j])));
          vmax = __hmax2(vmax, __habs2(*reinterpret_cast<const __half2*>(&hi[2 * j + 1])));
        }
        tmem_st8(abuf + 8 * q, hi);
        tmem_st8(abuf + 32 + 8 * q, lo);
      }
      cp_async_wait_all();
      fence_async_smem();
      tc_wait_st();
      tc_fence_before();
      named_barrier(1 + g, TS);
      if (wq == 0) {
        tc_fence_after();
        if (elect_one()) {
          const uint32_t wb = wbuf + (c & 1) * CHUNK_BYTES;
#pragma unroll
          for (int ks = 0; ks < KC / 16; ++ks) {
            const uint64_t dh = dbase + (uint64_t)((wb + ks * SLAB) >> 4);
            const uint64_t dl = dbase + (uint64_t)((wb + (KC / 16 + ks) * SLAB) >> 4);
            const uint32_t ah = tcol + C_A + (c & 1) * 64 + 8 * ks, al = ah + 32;
            mma_ts(tcol + C_D, ah, dh, idesc_f16(128, TS), (c > 0 || ks > 0) ? 1u : 0u);
            mma_ts(tcol + C_D, ah, dl, idesc_f16(128, TS), 1u);
            mma_ts(tcol + C_D, al, dh, idesc_f16(128, TS), 1u);
          }
          mma_commit(&bars[c & 1]);
        }
        __syncwarp();
      }
      if (c + 1 < nchunks) {
        if (c >= 1) { mbar_wait(&bars[(c - 1) & 1], par[(c - 1) & 1]); par[(c - 1) & 1] ^= 1; }
        load_w(c + 1);
#pragma unroll
        for (int j = 0; j < KC / 4; ++j) xv[j] = __ldg(xr + (c + 1) * (KC / 4) + j);
      }
    }
    if (nchunks >= 2) { mbar_wait(&bars[(nchunks - 2) & 1], par[(nchunks - 2) & 1]); par[(nchunks - 2) & 1] ^= 1; }
    mbar_wait(&bars[(nchunks - 1) & 1], par[(nchunks - 1) & 1]); par[(nchunks - 1) & 1] ^= 1;
    tc_fence_after();
    // ---- epilogue: distances of row li against the 128 columns of tile tj
    const float na = a.norm2[n0 + li], sa = a.sum1[n0 + li];
    const int fi = (int)a.frame[n0 + li];
    const float keps = (float)a.dim * eps * eps;
#pragma unroll 1
    for (int ch = 0; ch < TS / 16; ++ch) {
      uint32_t acc[16];
      tmem_ld16(tlane + C_D + 16 * ch, acc);
      tc_wait_ld();
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const int cj = 16 * ch + j;
        const int64_t lj = (int64_t)tj * TS + cj;
        const int fj = __float_as_int(s_nb[2 * TS + cj]);
        if (!valid || lj >= n || lj <= li) continue;               // strictly upper triangle, mirrored below
        int df = fi - fj; df = df < 0 ? -df : df;
        const bool conn = fj != INT_MIN && df > 0 && (a.max_dist < 0 || df <= a.max_dist);
        float d2 = na + s_nb[cj] - 2.f * __uint_as_float(acc[j]) + 2.f * eps * (sa - s_nb[TS + cj]) + keps;
        const float v = conn ? sqrtf(fmaxf(d2, 0.f)) : INFINITY;
        D[li * n + lj] = v;
        D[lj * n + li] = v;
      }
    }
    if (ti == tj && valid) D[li * n + li] = INFINITY;
    tc_fence_before();
    named_barrier(1 + g, TS);
  }
  {
    const uint32_t wv = *reinterpret_cast<const uint32_t*>(&vmax);
    if ((wv & 0x7FFFu) >= 0x7BFFu || ((wv >> 16) & 0x7FFFu) >= 0x7BFFu) atomicOr(a.status, 1);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc<512>(*tmem_slot);
}

__device__ __forceinline__ uint32_t okey(float f) {
  uint32_t u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

// One CTA per row: is the top-k SET of this row certain given the error band of the Gram distances?
// ambiguous <=> an entry outside the top-k lies within the band above the k-th value (in d^2).
__global__ void __launch_bounds__(256) row_ambiguity_kernel(const float* __restrict__ dense, const int64_t* __restrict__ gptr,
                                                            int64_t num_graphs, const int64_t* __restrict__ doff,
                                                            const float* __restrict__ norm2, const uint32_t* __restrict__ thr_key,
                                                            const int32_t* __restrict__ thr_idx, float beta,
                                                            int32_t* __restrict__ amb, int32_t* __restrict__ amb_count) {
  __shared__ float s_min[8];
  __shared__ float s_nmax[8];
  const int64_t i = blockIdx.x;
  int64_t lo = 0, hi = num_graphs;
  while (hi - lo > 1) { const int64_t mid = (lo + hi) >> 1; if (gptr[mid] <= i) lo = mid; else hi = mid; }
  const int64_t n0 = gptr[lo], n = gptr[lo + 1] - n0;
  const float* rowp = dense + doff[lo] + (i - n0) * n;
  const uint32_t tk = thr_key[i];
  const int32_t ti = thr_idx[i];
  if (tk == 0xffffffffu) { if (threadIdx.x == 0) amb[i] = 0; return; }      // k >= row length: everything is in
  float mn = INFINITY, nmax = 0.f;
  for (int64_t j = threadIdx.x; j < n; j += blockDim.x) {
    const float v = rowp[j];
    const uint32_t key = okey(v);
    const bool outside = key > tk || (key == tk && j > ti);
    if (outside) mn = fminf(mn, v);
    nmax = fmaxf(nmax, norm2[n0 + j]);
  }
  for (int d = 16; d > 0; d >>= 1) { mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, d)); nmax = fmaxf(nmax, __shfl_xor_sync(0xffffffffu, nmax, d)); }
  if ((threadIdx.x & 31) == 0) { s_min[threadIdx.x >> 5] = mn; s_nmax[threadIdx.x >> 5] = nmax; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int q = 1; q < 8; ++q) { mn = fminf(mn, s_min[q]); nmax = fmaxf(nmax, s_nmax[q]); }
    // k-th value from its key
    const uint32_t u = (tk & 0x80000000u) ? (tk & 0x7fffffffu) : ~tk;
    const float vk = __uint_as_float(u);
    int flag = 0;
    if (isfinite(vk) && isfinite(mn)) {
      const float band = beta * (norm2[i] + nmax);                           // absolute error bound on d^2
      flag = (mn * mn - vk * vk) <= 2.f * band ? 1 : 0;
    }
    amb[i] = flag;
    if (flag) atomicAdd(amb_count, 1);
  }
}

// Exact fp32 row (reference formula, sequential over the feature dimension) for the flagged rows.
__global__ void __launch_bounds__(256) exact_rows_kernel(const float* __restrict__ reid, int64_t dim,
                                                         const int64_t* __restrict__ frame, const int64_t* __restrict__ gptr,
                                                         int64_t num_graphs, const int64_t* __restrict__ doff, int64_t max_dist,
                                                         const int32_t* __restrict__ amb, float* __restrict__ dense) {
  const int64_t i = blockIdx.x;
  if (!amb[i]) return;
  int64_t lo = 0, hi = num_graphs;
  while (hi - lo > 1) { const int64_t mid = (lo + hi) >> 1; if (gptr[mid] <= i) lo = mid; else hi = mid; }
  const int64_t n0 = gptr[lo], n = gptr[lo + 1] - n0, li = i - n0;
  float* rowp = dense + doff[lo] + li * n;
  const int64_t fi = frame[i];
  const float eps = 1e-6f;
  for (int64_t lj = threadIdx.x; lj < n; lj += blockDim.x) {
    const int64_t fj = frame[n0 + lj];
    const int64_t df = fi > fj ? fi - fj : fj - fi;
    float v = INFINITY;
    if (lj != li && df > 0 && (max_dist < 0 || df <= max_dist)) {
      // the reference evaluates pairs as (row = smaller index, col = larger index)
      const float* pa = reid + (n0 + (li < lj ? li : lj)) * dim;
      const float* pb = reid + (n0 + (li < lj ? lj : li)) * dim;
      float acc = 0.f;
      for (int64_t k = 0; k < dim; ++k) { const float d = (pa[k] - pb[k]) + eps; acc = fmaf(d, d, acc); }
      v = sqrtf(acc);
    }
    rowp[lj] = v;
  }
}

__global__ void tile_offsets_kernel(const int64_t* __restrict__ gptr, int64_t num_graphs, int64_t* __restrict__ tile_off) {
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    int64_t acc = 0;
    for (int64_t g = 0; g < num_graphs; ++g) { tile_off[g] = acc; acc += (gptr[g + 1] - gptr[g] + TS - 1) / TS; }
    tile_off[num_graphs] = acc;
  }
}

}  // namespace gram

int64_t gram_workspace_bytes(int64_t num_nodes, int64_t total_tiles, int64_t num_graphs, int64_t dim) {
  return align_up(total_tiles * (dim / gram::KC) * gram::CHUNK_BYTES, 256) + 2 * align_up(num_nodes * 4, 256) +
         align_up((num_graphs + 1) * 8, 256) + align_up(num_nodes * 4, 256) + 1024;
}

// Fills the dense blocks with Gram distances, ranks, then repairs ambiguous rows exactly.
// `rank_rows(mask)` is provided by knn_graph.cu (batch_row_kth_kernel launcher).
int gram_dist_blocks(const float* reid, int64_t dim, const int64_t* frame, const int64_t* gptr, const int64_t* h_gptr,
                     int64_t num_graphs, const int64_t* doff, int64_t max_dist, void* ws, float* dense,
                     int32_t* status, float** norm2_out, int32_t** amb_out, int32_t** amb_count_out, cudaStream_t s) {
  using namespace gram;
  const int64_t n = h_gptr[num_graphs];
  int64_t total_tiles = 0, max_tiles = 0;
  for (int64_t g = 0; g < num_graphs; ++g) {
    const int64_t t = ceil_div(h_gptr[g + 1] - h_gptr[g], TS);
    total_tiles += t;
    max_tiles = t > max_tiles ? t : max_tiles;
  }
  Carver cv(ws);
  uint8_t* img = cv.take<uint8_t>(total_tiles * (dim / KC) * CHUNK_BYTES);
  float* norm2 = cv.take<float>(n);
  float* sum1 = cv.take<float>(n);
  int64_t* tile_off = cv.take<int64_t>(num_graphs + 1);
  int32_t* amb = cv.take<int32_t>(n + 1);
  MPN_CUDA(cudaMemsetAsync(img, 0, total_tiles * (dim / KC) * CHUNK_BYTES, s));
  MPN_CUDA(cudaMemsetAsync(amb + n, 0, 4, s));
  tile_offsets_kernel<<<1, 32, 0, s>>>(gptr, num_graphs, tile_off); count_launch();
  pack_reid_kernel<<<(unsigned)std::min<int64_t>(ceil_div(n * 32, 256), (int64_t)sm_count() * 16), 256, 0, s>>>(
      reid, dim, gptr, num_graphs, tile_off, n, img, norm2, sum1); count_launch();
  static bool attr_set = false;
  if (!attr_set) {
    MPN_CUDA(cudaFuncSetAttribute(gram_blocks_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    attr_set = true;
  }
  GramArgs a;
  a.reid = reid; a.dim = dim; a.frame = frame; a.gptr = gptr; a.doff = doff; a.tile_off = tile_off;
  a.img = img; a.norm2 = norm2; a.sum1 = sum1; a.max_dist = max_dist; a.dense = dense; a.status = status;
  const int64_t ntri = max_tiles * (max_tiles + 1) / 2;
  int64_t ctas_x = ceil_div(ntri, 2);
  const int64_t want = ceil_div((int64_t)sm_count(), num_graphs);          // ~1 CTA per SM over the whole batch
  if (ctas_x > want) ctas_x = want > 0 ? want : 1;
  dim3 grid((unsigned)ctas_x, (unsigned)num_graphs);
  gram_blocks_kernel<<<grid, NTHREADS, SMEM_BYTES, s>>>(a); count_launch();
  MPN_LAUNCH_CHECK();
  *norm2_out = norm2; *amb_out = amb; *amb_count_out = amb + n;
  return MPN_OK;
}

int gram_fix_ambiguous(const float* reid, int64_t dim, const int64_t* frame, const int64_t* gptr, int64_t num_graphs,
                       int64_t num_nodes, const int64_t* doff, int64_t max_dist, float* dense, const float* norm2,
                       const uint32_t* thr_key, const int32_t* thr_idx, float beta, int32_t* amb, int32_t* amb_count,
                       cudaStream_t s) {
  using namespace gram;
  row_ambiguity_kernel<<<(unsigned)num_nodes, 256, 0, s>>>(dense, gptr, num_graphs, doff, norm2, thr_key, thr_idx, beta,
                                                         amb, amb_count); count_launch();
  exact_rows_kernel<<<(unsigned)num_nodes, 256, 0, s>>>(reid, dim, frame, gptr, num_graphs, doff, max_dist, amb, dense);
  count_launch();
  MPN_LAUNCH_CHECK();
  return MPN_OK;
}

}  // namespace mpn
