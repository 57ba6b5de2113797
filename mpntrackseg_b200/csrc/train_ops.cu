// Building blocks of the training step (forward with stored activations + backward), fp32,
// deterministic (no float atomics; every reduction runs in a fixed order):
//   mpn_gemm          C = act(op(A) op(B) + bias)  with optional transposes / accumulation / ReLU-mask on A
//   mpn_gather_cols   out[e, off:off+w] = src[idx[e], :]
//   mpn_segment_sum   out[n, off:off+w] (+)= sum over the node's contiguous (optionally permuted) segment
//   mpn_relu_mask     g *= (y > 0)
//   mpn_adam_step     fused Adam with L2 weight decay (torch.optim.Adam semantics) on flat buffers
// They serve pl_module/pl_module.py:122-135 (loss.backward + optimizer step) for the core network.
#include <math.h>

#include "common.cuh"

namespace mpn {

constexpr int GT = 64, GK = 16;

// C[m][n] = sum_k A(m,k) B(k,n);  A(m,k) = ta ? a[k*lda + m] : a[m*lda + k],  B(k,n) = tb ? b[n*ldb + k] : b[k*ldb + n]
// optional: A is multiplied elementwise by (mask_a > 0) (ReLU backward of the producer), bias[n], ReLU, C += .
__global__ void __launch_bounds__(256) gemm_kernel(const float* __restrict__ a, int64_t lda, int ta,
                                                   const float* __restrict__ mask_a, int64_t ldm,
                                                   const float* __restrict__ b, int64_t ldb, int tb,
                                                   const float* __restrict__ bias, int relu, int accumulate,
                                                   float* __restrict__ c, int64_t ldc, int64_t m, int64_t n, int64_t k,
                                                   int64_t k_per_split, float* __restrict__ partial) {
  __shared__ float As[GK][GT + 4];
  __shared__ float Bs[GK][GT + 4];
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const int64_t m0 = (int64_t)blockIdx.y * GT, n0 = (int64_t)blockIdx.x * GT;
  const int64_t kbeg = (int64_t)blockIdx.z * k_per_split;
  const int64_t kend = kbeg + k_per_split < k ? kbeg + k_per_split : k;
  float acc[4][4] = {};
  for (int64_t k0 = kbeg; k0 < kend; k0 += GK) {
    for (int idx = threadIdx.x; idx < GT * GK; idx += 256) {
      int r, kk;
      if (ta) { r = idx % GT; kk = idx / GT; } else { r = idx / GK; kk = idx % GK; }   // coalesced along the contiguous dim
      const int64_t gm = m0 + r, gk = k0 + kk;
      float v = 0.f;
      if (gm < m && gk < kend) {
        v = ta ? a[gk * lda + gm] : a[gm * lda + gk];
        if (mask_a != nullptr) v = (ta ? mask_a[gk * ldm + gm] : mask_a[gm * ldm + gk]) > 0.f ? v : 0.f;
      }
      As[kk][r] = v;
    }
    for (int idx = threadIdx.x; idx < GT * GK; idx += 256) {
      int r, kk;
      if (tb) { r = idx / GK; kk = idx % GK; } else { r = idx % GT; kk = idx / GT; }
      const int64_t gn = n0 + r, gk = k0 + kk;
      Bs[kk][r] = (gn < n && gk < kend) ? (tb ? b[gn * ldb + gk] : b[gk * ldb + gn]) : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < GK; ++kk) {
      const float4 av = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
      const float4 bv = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
      const float ar[4] = {av.x, av.y, av.z, av.w}, br[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(ar[i], br[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int64_t gm = m0 + ty * 4 + i;
    if (gm >= m) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int64_t gn = n0 + tx * 4 + j;
      if (gn >= n) continue;
      if (partial != nullptr) { partial[((int64_t)blockIdx.z * m + gm) * n + gn] = acc[i][j]; continue; }
      float v = acc[i][j] + (bias != nullptr ? bias[gn] : 0.f);
      if (accumulate) v += c[gm * ldc + gn];
      c[gm * ldc + gn] = relu ? fmaxf(v, 0.f) : v;
    }
  }
}

// c = act(sum_z partial[z] + bias) (+ c): the split-K slices are added in ascending z (deterministic)
__global__ void gemm_reduce_kernel(const float* __restrict__ partial, int splits, const float* __restrict__ bias, int relu,
                                   int accumulate, float* __restrict__ c, int64_t ldc, int64_t m, int64_t n) {
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < m * n; t += (int64_t)gridDim.x * blockDim.x) {
    const int64_t gm = t / n, gn = t - gm * n;
    float v = 0.f;
    for (int z = 0; z < splits; ++z) v += partial[(int64_t)z * m * n + t];
    v += bias != nullptr ? bias[gn] : 0.f;
    if (accumulate) v += c[gm * ldc + gn];
    c[gm * ldc + gn] = relu ? fmaxf(v, 0.f) : v;
  }
}

// column sums of (A .* (mask > 0)) -> out[n] (+)= ...; one CTA per 32 columns, rows walked in a fixed order
__global__ void __launch_bounds__(256) colsum_kernel(const float* __restrict__ a, int64_t lda,
                                                     const float* __restrict__ mask, int64_t ldm, int64_t m,
                                                     int64_t n, int64_t rows_per_slice, float* __restrict__ partial) {
  __shared__ float part[8][33];
  const int col = blockIdx.x * 32 + (threadIdx.x & 31), rgrp = threadIdx.x >> 5;
  const int64_t r0 = (int64_t)blockIdx.y * rows_per_slice;
  const int64_t r1 = r0 + rows_per_slice < m ? r0 + rows_per_slice : m;
  float s = 0.f;
  if (col < n)
    for (int64_t r = r0 + rgrp; r < r1; r += 8) {
      float v = a[r * lda + col];
      if (mask != nullptr && !(mask[r * ldm + col] > 0.f)) v = 0.f;
      s += v;
    }
  part[rgrp][threadIdx.x & 31] = s;
  __syncthreads();
  if (rgrp == 0 && col < n) {
    float t = 0.f;
    for (int g = 0; g < 8; ++g) t += part[g][threadIdx.x & 31];
    partial[(int64_t)blockIdx.y * n + col] = t;
  }
}

__global__ void colsum_reduce_kernel(const float* __restrict__ partial, int slices, int64_t n, int accumulate,
                                     float* __restrict__ out) {
  for (int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; c < n; c += (int64_t)gridDim.x * blockDim.x) {
    float t = 0.f;
    for (int y = 0; y < slices; ++y) t += partial[(int64_t)y * n + c];     // ascending slices: deterministic
    out[c] = accumulate ? out[c] + t : t;
  }
}

__global__ void gather_cols_kernel(const float* __restrict__ src, int64_t lds, int64_t w, const int32_t* __restrict__ idx,
                                   int64_t rows, float* __restrict__ out, int64_t ldo, int64_t off) {
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < rows * w; t += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = t / w, c = t - r * w;
    const int64_t s = idx != nullptr ? idx[r] : r;
    out[r * ldo + off + c] = src[s * lds + c];
  }
}

// out[node, off + f] (+)= sum_{q in [ptr[node], ptr[node+1])} in[(perm ? perm[q] : q) * ld + in_off + f], in slot order
__global__ void segment_sum_kernel(const float* __restrict__ in, int64_t ld, int64_t in_off, int64_t w,
                                   const int32_t* __restrict__ ptr, const int32_t* __restrict__ perm, int64_t nodes,
                                   int accumulate, float* __restrict__ out, int64_t ldo, int64_t off) {
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < nodes * w; t += (int64_t)gridDim.x * blockDim.x) {
    const int64_t node = t / w, f = t - node * w;
    float s = 0.f;
    for (int q = ptr[node]; q < ptr[node + 1]; ++q) s += in[(int64_t)(perm != nullptr ? perm[q] : q) * ld + in_off + f];
    float* o = out + node * ldo + off + f;
    *o = accumulate ? *o + s : s;
  }
}

__global__ void relu_mask_kernel(float* __restrict__ g, const float* __restrict__ y, int64_t n) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    if (!(y[i] > 0.f)) g[i] = 0.f;
}

__global__ void adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
                            int64_t n, float lr, float beta1, float beta2, float eps, float weight_decay,
                            float bc1, float bc2, float grad_scale) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    float gi = g[i] * grad_scale + weight_decay * p[i];           // torch.optim.Adam: L2 term added to the gradient
    const float mi = beta1 * m[i] + (1.f - beta1) * gi;
    const float vi = beta2 * v[i] + (1.f - beta2) * gi * gi;
    m[i] = mi; v[i] = vi;
    const float denom = sqrtf(vi) / sqrtf(bc2) + eps;
    p[i] -= (lr / bc1) * (mi / denom);
  }
}

// step counter on the device (CUDA-graph replays of a training step must not bake the step number in)
__global__ void bump_step_kernel(int64_t* step) { *step += 1; }
__global__ void adam_dev_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
                                int64_t n, float lr, float beta1, float beta2, float eps, float weight_decay,
                                const int64_t* __restrict__ step, float grad_scale) {
  const float t = (float)*step;
  const float bc1 = 1.f - powf(beta1, t), bc2 = 1.f - powf(beta2, t);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    float gi = g[i] * grad_scale + weight_decay * p[i];
    const float mi = beta1 * m[i] + (1.f - beta1) * gi;
    const float vi = beta2 * v[i] + (1.f - beta2) * gi * gi;
    m[i] = mi; v[i] = vi;
    const float denom = sqrtf(vi) / sqrtf(bc2) + eps;
    p[i] -= (lr / bc1) * (mi / denom);
  }
}

static void keep_async_pool() {
  static bool done = false;
  if (done) return;
  int dev = 0;
  cudaMemPool_t pool;
  if (cudaGetDevice(&dev) == cudaSuccess && cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
    unsigned long long keep = ~0ull;
    cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
  }
  done = true;
}

static inline unsigned flat_grid(int64_t n) {
  return (unsigned)std::min<int64_t>(ceil_div(n > 0 ? n : 1, 256), (int64_t)sm_count() * 8);
}

}  // namespace mpn

using namespace mpn;

extern "C" {

int mpn_gemm(const float* a, int64_t lda, int trans_a, const float* mask_a, int64_t ldm, const float* b, int64_t ldb,
             int trans_b, const float* bias, int relu, int accumulate, float* c, int64_t ldc, int64_t m, int64_t n,
             int64_t k, void* stream) {
  MPN_CHECK_ARG(m >= 0 && n >= 0 && k >= 0, "gemm: negative size");
  if (m == 0 || n == 0) return MPN_OK;
  MPN_CHECK_ARG(c != nullptr && (k == 0 || (a && b)), "gemm: null pointer");
  MPN_CHECK_ARG(ceil_div(m, GT) <= 65535, "gemm: at most %d rows per call (gridDim.y limit)", 65535 * GT);
  cudaStream_t s = as_stream(stream);
  keep_async_pool();
  const int64_t tiles = ceil_div(n, GT) * ceil_div(m, GT);
  // few output tiles but a long contraction (weight gradients: contraction over the edges): split K
  int64_t splits = 1;
  if (tiles < 2 * sm_count() && k >= 2048) {
    splits = std::min<int64_t>(ceil_div(4 * (int64_t)sm_count(), tiles), ceil_div(k, 512));
    if (splits > 256) splits = 256;
  }
  if (splits <= 1) {
    dim3 grid((unsigned)ceil_div(n, GT), (unsigned)ceil_div(m, GT), 1);
    gemm_kernel<<<grid, 256, 0, s>>>(a, lda, trans_a, mask_a, ldm, b, ldb, trans_b, bias, relu, accumulate, c, ldc, m, n, k,
                                     k, nullptr);
    count_launch();
  } else {
    const int64_t kps = align_up(ceil_div(k, splits), GK);
    splits = ceil_div(k, kps);
    float* partial = nullptr;
    MPN_CUDA(cudaMallocAsync(&partial, sizeof(float) * splits * m * n, s));
    dim3 grid((unsigned)ceil_div(n, GT), (unsigned)ceil_div(m, GT), (unsigned)splits);
    gemm_kernel<<<grid, 256, 0, s>>>(a, lda, trans_a, mask_a, ldm, b, ldb, trans_b, nullptr, 0, 0, c, ldc, m, n, k, kps,
                                     partial);
    count_launch();
    gemm_reduce_kernel<<<flat_grid(m * n), 256, 0, s>>>(partial, (int)splits, bias, relu, accumulate, c, ldc, m, n);
    count_launch();
    MPN_CUDA(cudaFreeAsync(partial, s));
  }
  MPN_LAUNCH_CHECK();
  return MPN_OK;
}

int mpn_colsum(const float* a, int64_t lda, const float* mask, int64_t ldm, int64_t m, int64_t n, int accumulate,
               float* out, void* stream) {
  if (n == 0) return MPN_OK;
  MPN_CHECK_ARG(out != nullptr && (m == 0 || a != nullptr), "colsum: null pointer");
  cudaStream_t s = as_stream(stream);
  keep_async_pool();
  const int64_t slices = std::max<int64_t>(1, std::min<int64_t>(ceil_div(m, 256), 128));
  const int64_t rps = ceil_div(m > 0 ? m : 1, slices);
  float* partial = nullptr;
  MPN_CUDA(cudaMallocAsync(&partial, sizeof(float) * slices * n, s));
  dim3 grid((unsigned)ceil_div(n, 32), (unsigned)slices);
  colsum_kernel<<<grid, 256, 0, s>>>(a, lda, mask, ldm, m, n, rps, partial);
  count_launch();
  colsum_reduce_kernel<<<(unsigned)ceil_div(n, 256), 256, 0, s>>>(partial, (int)slices, n, accumulate, out);
  count_launch();
  MPN_CUDA(cudaFreeAsync(partial, s));
  MPN_LAUNCH_CHECK();
  return MPN_OK;
}

int mpn_gather_cols(const float* src, int64_t lds, int64_t width, const int32_t* idx, int64_t rows, float* out,
                    int64_t ldo, int64_t col_off, void* stream) {
  if (rows == 0 || width == 0) return MPN_OK;
  MPN_CHECK_ARG(src && out, "gather_cols: null pointer");
  gather_cols_kernel<<<flat_grid(rows * width), 256, 0, as_stream(stream)>>>(src, lds, width, idx, rows, out, ldo, col_off);
  count_launch();
  MPN_LAUNCH_CHECK();
  return MPN_OK;
}

int mpn_segment_sum(const float* in, int64_t ld, int64_t in_off, int64_t width, const int32_t* ptr, const int32_t* perm,
                    int64_t nodes, int accumulate, float* out, int64_t ldo, int64_t col_off, void* stream) {
  if (nodes == 0 || width == 0) return MPN_OK;
  MPN_CHECK_ARG(ptr && out, "segment_sum: null pointer");
  segment_sum_kernel<<<flat_grid(nodes * width), 256, 0, as_stream(stream)>>>(in, ld, in_off, width, ptr, perm, nodes,
                                                                           accumulate, out, ldo, col_off);
  count_launch();
  MPN_LAUNCH_CHECK();
  return MPN_OK;
}

int mpn_relu_mask(float* g, const float* y, int64_t n, void* stream) {
  if (n == 0) return MPN_OK;
  MPN_CHECK_ARG(g && y, "relu_mask: null pointer");
  relu_mask_kernel<<<flat_grid(n), 256, 0, as_stream(stream)>>>(g, y, n);
  count_launch();
  MPN_LAUNCH_CHECK();
  return MPN_OK;
}

int mpn_adam_step(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, int64_t n, float lr, float beta1,
                  float beta2, float eps, float weight_decay, int64_t step, float grad_scale, void* stream) {
  if (n == 0) return MPN_OK;
  MPN_CHECK_ARG(params && grads && exp_avg && exp_avg_sq && step >= 1, "adam_step: bad arguments");
  const float bc1 = 1.f - powf(beta1, (float)step), bc2 = 1.f - powf(beta2, (float)step);
  adam_kernel<<<flat_grid(n), 256, 0, as_stream(stream)>>>(params, grads, exp_avg, exp_avg_sq, n, lr, beta1, beta2, eps,
                                                          weight_decay, bc1, bc2, grad_scale);
  count_launch();
  MPN_LAUNCH_CHECK();
  return MPN_OK;
}

int mpn_adam_step_dev(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, int64_t n, float lr, float beta1,
                      float beta2, float eps, float weight_decay, int64_t* d_step, float grad_scale, void* stream) {
  MPN_CHECK_ARG(d_step != nullptr, "adam_step_dev: null step counter");
  bump_step_kernel<<<1, 1, 0, as_stream(stream)>>>(d_step); count_launch();
  if (n > 0) {
    MPN_CHECK_ARG(params && grads && exp_avg && exp_avg_sq, "adam_step_dev: bad arguments");
    adam_dev_kernel<<<flat_grid(n), 256, 0, as_stream(stream)>>>(params, grads, exp_avg, exp_avg_sq, n, lr, beta1, beta2, eps,
                                                                weight_decay, d_step, grad_scale);
    count_launch();
  }
  MPN_LAUNCH_CHECK();
  return MPN_OK;
}

}  // extern "C"
