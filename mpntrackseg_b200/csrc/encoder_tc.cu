// Node encoder  x_init = ReLU(W1 ReLU(W0 x + b0) + b1)  (2048 -> 128 -> 32, models/mpn.py:355 with
// configs/tracking_cfg.yaml:146-148) on the tcgen05 tensor cores.
//
// One 128-node tile per group of 4 warps (thread = node row = TMEM lane), two groups per CTA.  The
// K = 2048 contraction is streamed in chunks of 64: each thread loads its row's 64 fp32 values,
// splits them into fp16 hi/lo (22 significant bits, same scheme as mp_step_tc.cu) and writes them to
// a double-buffered TMEM A operand; the matching 32 KB slice of the packed fp16 hi/lo weight image (the K stream of
// the B operand) is brought into shared memory by ONE TMA bulk copy per chunk (cp.async.bulk, completion counted in
// bytes on an mbarrier); 12 MMAs (128 x 128 x 16, hi*hi + hi*lo + lo*hi) per chunk
// accumulate in TMEM while the next chunk is being loaded.  The 128 -> 32 layer reuses the accumulator
// in place as its operand.
#include "common.cuh"
#include "tc_ptx.cuh"

namespace mpn {
namespace enc {

using namespace ptx;

constexpr int TS = 128, H = 128, O = 32, KC = 64;                    // tile rows, hidden, out, K per chunk
constexpr int NTHREADS = 256;
constexpr int SLAB1 = H * 32;                                         // one K=16 step of W0: 128 rows x 32 B
constexpr int CHUNK_BYTES = 2 * (KC / 16) * SLAB1;                    // hi + lo, 4 k-steps = 32 KB
constexpr int SLAB2 = O * 32;                                         // one K=16 step of W1: 32 rows x 32 B
constexpr int W2_BYTES = 2 * (H / 16) * SLAB2;                        // 16 KB
constexpr int F_B0 = 0, F_B1 = H, F_COUNT = H + O;
constexpr int TAIL_BYTES = W2_BYTES + F_COUNT * 4;                    // W1 image + biases

constexpr int C_A = 0;            // A operand double buffer: [buf][hi 32 cols | lo 32 cols]
constexpr int C_D1 = 128;         // 128 accumulator columns, reused in place as the layer-2 operand
constexpr int C_D2 = 0;           // layer-2 accumulator (32 cols) over the dead A buffers

constexpr int SM_W = 0;                                               // [2 groups][2 bufs][CHUNK_BYTES]
constexpr int SM_TAIL = SM_W + 4 * CHUNK_BYTES;
constexpr int SM_BAR = (SM_TAIL + TAIL_BYTES + 127) / 128 * 128;      // u64 bars[2 groups][2] (MMA done), wbars[2 groups][2] (weights landed)
constexpr int SM_TMEM = SM_BAR + 8 * 8;
constexpr int SMEM_BYTES = SM_TMEM + 16;

__device__ __forceinline__ int slab_off(int n, int k16) {
  return (n >> 3) * 256 + (k16 >> 3) * 128 + (n & 7) * 16 + (k16 & 7) * 2;
}

// global image: [K0/64 chunks][hi: 4 slabs | lo: 4 slabs] then the tail (W1 hi slabs, lo slabs, b0, b1)
__global__ void pack_kernel(const float* __restrict__ w0, const float* __restrict__ b0, const float* __restrict__ w1,
                            const float* __restrict__ b1, int64_t k0, uint8_t* __restrict__ img) {
  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x, nt = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = tid; i < (int64_t)H * k0; i += nt) {
    const int n = (int)(i / k0);
    const int64_t k = i % k0;
    const float v = w0[i];
    const __half h = __float2half_rn(v), l = __float2half_rn(v - __half2float(h));
    const int64_t off = (k / KC) * CHUNK_BYTES + ((k % KC) >> 4) * SLAB1 + slab_off(n, (int)(k & 15));
    *reinterpret_cast<__half*>(img + off) = h;
    *reinterpret_cast<__half*>(img + off + (KC / 16) * SLAB1) = l;
  }
  uint8_t* tail = img + (k0 / KC) * CHUNK_BYTES;
  for (int64_t i = tid; i < O * H; i += nt) {
    const int n = (int)(i / H), k = (int)(i % H);
    const float v = w1[i];
    const __half h = __float2half_rn(v), l = __float2half_rn(v - __half2float(h));
    const int off = (k >> 4) * SLAB2 + slab_off(n, k & 15);
    *reinterpret_cast<__half*>(tail + off) = h;
    *reinterpret_cast<__half*>(tail + off + (H / 16) * SLAB2) = l;
  }
  float* ft = reinterpret_cast<float*>(tail + W2_BYTES);
  for (int64_t i = tid; i < F_COUNT; i += nt) ft[i] = i < H ? b0[i] : b1[i - H];
}

__global__ void __launch_bounds__(NTHREADS, 1) node_encoder_tc_kernel(const float* __restrict__ x, int64_t n, int64_t k0,
                                                                     const uint8_t* __restrict__ img,
                                                                     float* __restrict__ out,
                                                                     int32_t* __restrict__ status) {
  extern __shared__ __align__(128) uint8_t smem[];
  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
  const int g = warp >> 2, wq = warp & 3, gt = tid & (TS - 1);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + SM_BAR) + 2 * g;   // [0]/[1]: even / odd chunks
  uint64_t* wbars = reinterpret_cast<uint64_t*>(smem + SM_BAR) + 4 + 2 * g;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + SM_TMEM);
  const int nchunks = (int)(k0 / KC);
  const uint8_t* tail_g = img + (int64_t)nchunks * CHUNK_BYTES;
  for (int i = tid; i < TAIL_BYTES / 16; i += NTHREADS)
    reinterpret_cast<uint4*>(smem + SM_TAIL)[i] = __ldg(reinterpret_cast<const uint4*>(tail_g) + i);
  if (tid == 0) {
    for (int i = 0; i < 8; ++i) mbar_init(reinterpret_cast<uint64_t*>(smem + SM_BAR) + i, 1);
    mbar_fence_init();
  }
  if (warp == 0) tmem_alloc<512>(tmem_slot);
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tcol = __shfl_sync(0xffffffffu, *tmem_slot, 0) + (uint32_t)g * 256u;
  const uint32_t tlane = tcol + ((uint32_t)(wq * 32) << 16);
  const float* s_f = reinterpret_cast<const float*>(smem + SM_TAIL + W2_BYTES);
  const uint32_t wbuf = smem_u32(smem + SM_W + g * 2 * CHUNK_BYTES);
  const uint64_t dbase = smem_desc_kmajor(0, 128, 256);
  __half2 vmax = __floats2half2_rn(0.f, 0.f);
  uint32_t par[2] = {0, 0}, wpar[2] = {0, 0};

  const int64_t tiles = (n + TS - 1) / TS;
  for (int64_t t = (int64_t)blockIdx.x * 2 + g; t < tiles; t += (int64_t)gridDim.x * 2) {
    int64_t row = t * TS + gt;
    const bool valid = row < n;
    if (!valid) row = n - 1;
    const float4* xr = reinterpret_cast<const float4*>(x + row * k0);
    auto load_w = [&](int c) {                                       // 32 KB weight chunk: one TMA bulk copy
      if (gt == 0) {
        mbar_arrive_expect_tx(&wbars[c & 1], CHUNK_BYTES);
        tma_bulk_g2s(wbuf + (c & 1) * CHUNK_BYTES, img + (int64_t)c * CHUNK_BYTES, CHUNK_BYTES, &wbars[c & 1]);
      }
    };
    float4 xv[KC / 4];
    load_w(0);
#pragma unroll
    for (int j = 0; j < KC / 4; ++j) xv[j] = __ldcs(xr + j);
    for (int c = 0; c < nchunks; ++c) {
      // ---- this chunk's A operand: split the 64 values and store them to TMEM buffer c&1
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
      mbar_wait(&wbars[c & 1], wpar[c & 1]); wpar[c & 1] ^= 1;       // this chunk's weights have landed
      tc_wait_st();
      tc_fence_before();
      named_barrier(1 + g, TS);
      if (wq == 0) {
        tc_fence_after();
        if (elect_one()) {
          const uint32_t wb = wbuf + (c & 1) * CHUNK_BYTES;
#pragma unroll
          for (int ks = 0; ks < KC / 16; ++ks) {
            const uint64_t dh = dbase + (uint64_t)((wb + ks * SLAB1) >> 4);
            const uint64_t dl = dbase + (uint64_t)((wb + (KC / 16 + ks) * SLAB1) >> 4);
            const uint32_t ah = tcol + C_A + (c & 1) * 64 + 8 * ks, al = ah + 32;
            mma_ts(tcol + C_D1, ah, dh, idesc_f16(128, H), (c > 0 || ks > 0) ? 1u : 0u);
            mma_ts(tcol + C_D1, ah, dl, idesc_f16(128, H), 1u);
            mma_ts(tcol + C_D1, al, dh, idesc_f16(128, H), 1u);
          }
          mma_commit(&bars[c & 1]);
        }
        __syncwarp();
      }
      // ---- next chunk: its buffers were last read by the MMAs of chunk c-1
      if (c + 1 < nchunks) {
        if (c >= 1) { mbar_wait(&bars[(c - 1) & 1], par[(c - 1) & 1]); par[(c - 1) & 1] ^= 1; }
        load_w(c + 1);
#pragma unroll
        for (int j = 0; j < KC / 4; ++j) xv[j] = __ldcs(xr + (c + 1) * (KC / 4) + j);
      }
    }
    // drain: the last two commits
    if (nchunks >= 2) { mbar_wait(&bars[(nchunks - 2) & 1], par[(nchunks - 2) & 1]); par[(nchunks - 2) & 1] ^= 1; }
    mbar_wait(&bars[(nchunks - 1) & 1], par[(nchunks - 1) & 1]); par[(nchunks - 1) & 1] ^= 1;
    tc_fence_after();
    // ---- epilogue 1: h = ReLU(D1 + b0) -> layer-2 operand in place
#pragma unroll
    for (int ch = 0; ch < H / 16; ++ch) {
      uint32_t acc[16];
      tmem_ld16(tlane + C_D1 + 16 * ch, acc);
      tc_wait_ld();
      uint32_t hi[8], lo[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        split2_relu(__uint_as_float(acc[2 * j]) + s_f[F_B0 + 16 * ch + 2 * j],
                    __uint_as_float(acc[2 * j + 1]) + s_f[F_B0 + 16 * ch + 2 * j + 1], hi[j], lo[j]);
        vmax = __hmax2(vmax, *reinterpret_cast<const __half2*>(&hi[j]));
      }
      tmem_st8(tlane + C_D1 + 16 * ch, hi);
      tmem_st8(tlane + C_D1 + 16 * ch + 8, lo);
    }
    tc_wait_st();
    tc_fence_before();
    named_barrier(1 + g, TS);
    if (wq == 0) {
      tc_fence_after();
      if (elect_one()) {
        const uint32_t w2 = smem_u32(smem + SM_TAIL);
#pragma unroll
        for (int ks = 0; ks < H / 16; ++ks) {
          const uint64_t dh = dbase + (uint64_t)((w2 + ks * SLAB2) >> 4);
          const uint64_t dl = dbase + (uint64_t)((w2 + (H / 16 + ks) * SLAB2) >> 4);
          const uint32_t ah = tcol + C_D1 + 16 * ks, al = ah + 8;
          mma_ts(tcol + C_D2, ah, dh, idesc_f16(128, O), ks > 0 ? 1u : 0u);
          mma_ts(tcol + C_D2, ah, dl, idesc_f16(128, O), 1u);
          mma_ts(tcol + C_D2, al, dh, idesc_f16(128, O), 1u);
        }
        mma_commit(&bars[0]);
      }
      __syncwarp();
    }
    mbar_wait(&bars[0], par[0]); par[0] ^= 1;
    tc_fence_after();
    // ---- epilogue 2: x_init = ReLU(D2 + b1)
#pragma unroll
    for (int ch = 0; ch < O / 16; ++ch) {
      uint32_t acc[16];
      tmem_ld16(tlane + C_D2 + 16 * ch, acc);
      tc_wait_ld();
      if (valid) {
        float4* dst = reinterpret_cast<float4*>(out + (t * TS + gt) * O + 16 * ch);
#pragma unroll
        for (int j = 0; j < 4; ++j)
          dst[j] = make_float4(fmaxf(__uint_as_float(acc[4 * j]) + s_f[F_B1 + 16 * ch + 4 * j], 0.f),
                               fmaxf(__uint_as_float(acc[4 * j + 1]) + s_f[F_B1 + 16 * ch + 4 * j + 1], 0.f),
                               fmaxf(__uint_as_float(acc[4 * j + 2]) + s_f[F_B1 + 16 * ch + 4 * j + 2], 0.f),
                               fmaxf(__uint_as_float(acc[4 * j + 3]) + s_f[F_B1 + 16 * ch + 4 * j + 3], 0.f));
      }
    }
    tc_fence_before();
    named_barrier(1 + g, TS);                                        // D2 / A buffers are free for the next tile
  }
  {
    const uint32_t w = *reinterpret_cast<const uint32_t*>(&vmax);
    if ((w & 0x7FFFu) >= 0x7BFFu || ((w >> 16) & 0x7FFFu) >= 0x7BFFu) atomicOr(status, 1);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc<512>(*tmem_slot);
}

}  // namespace enc
}  // namespace mpn

using namespace mpn;

extern "C" {

int64_t mpn_node_encoder_tc_workspace(int64_t k0) { return (k0 / enc::KC) * enc::CHUNK_BYTES + enc::TAIL_BYTES + 512; }

int mpn_node_encoder_tc(const float* x, int64_t n, int64_t k0, const float* w0, const float* b0, int64_t hidden,
                        const float* w1, const float* b1, int64_t out_dim, void* ws, float* out, int32_t* status,
                        void* stream) {
  MPN_CHECK_ARG(hidden == enc::H && out_dim == enc::O && k0 >= enc::KC && k0 % enc::KC == 0,
                "node_encoder_tc: built for K -> 128 -> 32 with K a multiple of 64 (got %lld -> %lld -> %lld)",
                (long long)k0, (long long)hidden, (long long)out_dim);
  MPN_CHECK_ARG(ws && status, "node_encoder_tc: null workspace / status");
  cudaStream_t s = as_stream(stream);
  MPN_CUDA(cudaMemsetAsync(status, 0, 4, s));
  if (n == 0) return MPN_OK;
  MPN_CHECK_ARG(x && w0 && b0 && w1 && b1 && out, "node_encoder_tc: null pointer");
  MPN_CHECK_ARG(reinterpret_cast<uintptr_t>(x) % 16 == 0, "node_encoder_tc: x must be 16-byte aligned");
  uint8_t* img = static_cast<uint8_t*>(ws);
  MPN_CUDA(cudaFuncSetAttribute(enc::node_encoder_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                enc::SMEM_BYTES));
  enc::pack_kernel<<<64, 256, 0, s>>>(w0, b0, w1, b1, k0, img); count_launch();
  const int64_t tiles = ceil_div(n, enc::TS);
  int grid = (int)std::min<int64_t>(ceil_div(tiles, 2), sm_count());
  enc::node_encoder_tc_kernel<<<grid, enc::NTHREADS, enc::SMEM_BYTES, s>>>(x, n, k0, img, out, status); count_launch();
  MPN_LAUNCH_CHECK();
  return MPN_OK;
}

}  // extern "C"
