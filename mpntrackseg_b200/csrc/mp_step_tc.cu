// Fused message-passing step on the 5th-gen tensor cores (tcgen05, sm_100a).
//
// One 128-edge tile = one M=128 MMA tile: TMEM lane i <-> edge slot i <-> epilogue thread i.
// The four dense layers of a step (edge MLP 160->80->16, flow MLP 80->56->32) run as
// tcgen05.mma kind::f16 with the ACTIVATIONS in TMEM (A operand, written by the epilogue
// threads with tcgen05.st) and the WEIGHTS resident in shared memory (B operand, K-major).
// Precision: every operand is split v ~= hi + lo in fp16 (22 significant bits) and each K step
// issues hi*hi + hi*lo + lo*hi with fp32 accumulation in TMEM -- single-pass bf16/tf32 breaks
// the 1e-3 parity bar after 12 recurrent steps, bf16 hi/lo is 10-20x less accurate than fp16
// hi/lo (tools/emulate_split_bf16.py).  fp16 range overflow (> 65504) is detected and reported
// through `status`; the caller then reruns on the fp32 kernels (mp_step.cu).
//
//   x[row] part of edge layer 0 is hoisted: prow[r] = W0[:, 0:64] [x_init[r] | x_lat[r]] + b0
//   (fp32, computed by the node kernel), added in the first epilogue.
//
// State lives in HBM already split: per node 128 B = [hi(32 halfs) | lo(32 halfs)], per edge
// 64 B = [hi(16) | lo(16)], so loads are plain copies into TMEM and the bytes per edge-update
// stay at 200.
//
// Warp roles (512 threads, 1 CTA / SM): two groups of 8 warps; each group keeps one tile in flight
// and owns 256 of the 512 TMEM columns.  Every edge row is served by two threads (halves A and B of
// the group) that split the columns of each epilogue.  After the group's 256 threads have written a
// layer's operand, an elected lane of the group's first warp issues that layer's MMAs and commits
// them to the group's mbarrier; while one group runs an epilogue the tensor core runs the other
// group's layer.  Operand rows of the next tile are staged with cp.async during the current one,
// the per-row message sums of the previous tile run under the current tile's first layer.
#include <cuda.h>
#include <stdlib.h>

#include "common.cuh"
#include "tc_ptx.cuh"

namespace mpn {
namespace tc {

int variant();   // 2: two tiles in flight, operands staged through TMEM; 3 (default): three tiles, SS-mode layer-1 operands

using namespace ptx;

constexpr int TS = 128;
constexpr int DN = 32, DE = 16, EH = 80, FH = 56, FHP = 64, CH = 8;
constexpr int NTHREADS = 512;

// ---- shared-memory weight image (bytes). Slab = one K=16 step of a B operand: [n/8][k/8][n%8][8 halfs]
constexpr int L1_KS = 6, L1_SLAB = EH * 32;
constexpr int L2_KS = 5, L2_SLAB = DE * 32;
constexpr int L3_KS = 5, L3_SLAB = FHP * 32;
constexpr int L4_KS = 4, L4_SLAB = DN * 32;
constexpr int OFF_L1H = 0, OFF_L1L = OFF_L1H + L1_KS * L1_SLAB;
constexpr int OFF_L2H = OFF_L1L + L1_KS * L1_SLAB, OFF_L2L = OFF_L2H + L2_KS * L2_SLAB;
constexpr int OFF_L3H = OFF_L2L + L2_KS * L2_SLAB, OFF_L3L = OFF_L3H + L3_KS * L3_SLAB;
constexpr int OFF_L4H = OFF_L3L + L3_KS * L3_SLAB, OFF_L4L = OFF_L4H + L4_KS * L4_SLAB;
constexpr int OFF_F32 = OFF_L4L + L4_KS * L4_SLAB;          // fp32 tail
constexpr int F_B1 = 0, F_FB0 = F_B1 + DE, F_FB1 = F_FB0 + FHP, F_CW0 = F_FB1 + DN, F_CB0 = F_CW0 + DE * CH,
              F_CW1 = F_CB0 + CH, F_CB1 = F_CW1 + CH, F_COUNT = F_CB1 + 4;
constexpr int IMG_BYTES = (OFF_F32 + F_COUNT * 4 + 127) / 128 * 128;

// ---- TMEM column map inside a group's 256 columns.  Each epilogue rewrites an accumulator chunk of
//      16 fp32 columns IN PLACE as the next layer's operand: [hi: 8 cols of fp16 pairs | lo: 8 cols].
constexpr int C_XCH = 0, C_XCL = 32;        // x[col] hi / lo            (K = 64)
constexpr int C_EH = 64, C_EL = 80;         // [e_init | e] hi / lo      (K = 32)
constexpr int C_D1 = 96, C_A2 = C_D1;       // layer-1 accumulator (80 cols) -> layer-2 operand (K = 80)
constexpr int C_D2 = 64, C_A3 = 80;         // layer-2 accumulator (16 cols) and e' operand (16 cols), over the dead E region
constexpr int C_D3 = 176, C_A4 = C_D3;      // layer-3 accumulator (64 cols) -> layer-4 operand (K = 64)
constexpr int C_D4 = 0;                     // layer-4 accumulator, 32 cols (over the dead x[col])

// ---- per-row sums are made per warp over its 32 consecutive slots ("chunk"); rows that cross a
//      chunk boundary leave partial sums that the node kernel adds up in chunk order.
constexpr int CHUNK = 32;

// ---- dynamic shared memory map
constexpr int MSG_LD = DN + 1;
constexpr int STAGE_ROW = 400;                                      // 24 x 16 B operands per edge + 16 B pad (bank spread)
constexpr int SM_MSG = IMG_BYTES;                                   // float [8 warps][32*MSG_LD]
constexpr int SM_STAGE = SM_MSG + 8 * CHUNK * MSG_LD * 4;           // [2 groups][TS][STAGE_ROW]  next tile's operands
constexpr int SM_BAR = SM_STAGE + 2 * TS * STAGE_ROW;               // u64 d_ready[2] (+2 spare)
constexpr int SM_TMEM = SM_BAR + 4 * 8;
constexpr int SMEM_BYTES = SM_TMEM + 16;

__device__ __forceinline__ int slab_off(int n, int k16) {           // byte offset inside a slab
  return (n >> 3) * 256 + (k16 >> 3) * 128 + (n & 7) * 16 + (k16 & 7) * 2;
}

// =================================================================== weight packing (once per forward)
// One image per direction (flow_out / flow_in); layers 1-2 and the classifier are shared.
__global__ void pack_weights_kernel(mpn_core_weights w, uint8_t* __restrict__ img_out, uint8_t* __restrict__ img_in,
                                    int cls_in_l3) {
  for (int dir = 0; dir < 2; ++dir) {
    uint8_t* img = dir == 0 ? img_out : img_in;
    const float* f0 = dir == 0 ? w.fout_w0 : w.fin_w0;
    const float* f1 = dir == 0 ? w.fout_w1 : w.fin_w1;
    const float* fb0 = dir == 0 ? w.fout_b0 : w.fin_b0;
    const float* fb1 = dir == 0 ? w.fout_b1 : w.fin_b1;
    auto put = [&](int off_h, int off_l, int slab_bytes, int n, int k, float v) {
      const __half h = __float2half_rn(v);
      const __half l = __float2half_rn(v - __half2float(h));
      const int o = (k >> 4) * slab_bytes + slab_off(n, k & 15);
      *reinterpret_cast<__half*>(img + off_h + o) = h;
      *reinterpret_cast<__half*>(img + off_l + o) = l;
    };
    const int tid = blockIdx.x * blockDim.x + threadIdx.x, nt = gridDim.x * blockDim.x;
    for (int i = tid; i < EH * 96; i += nt) {            // edge layer 0, input columns 64..159
      const int n = i / 96, k = i % 96;
      put(OFF_L1H, OFF_L1L, L1_SLAB, n, k, w.edge_w0[n * 160 + 64 + k]);
    }
    for (int i = tid; i < DE * EH; i += nt) {            // edge layer 1
      const int n = i / EH, k = i % EH;
      put(OFF_L2H, OFF_L2L, L2_SLAB, n, k, w.edge_w1[n * EH + k]);
    }
    for (int i = tid; i < FHP * 80; i += nt) {           // flow layer 0 (rows padded 56 -> 64)
      const int n = i / 80, k = i % 80;
      // rows 56..63 are padding for the flow MLP; variant 3 puts the classifier's first layer there (it reads
      // only the e' K step, columns 64..79)
      float v = n < FH ? f0[n * 80 + k] : 0.f;
      if (cls_in_l3 && n >= FH && k >= 64) v = w.cls_w0[(n - FH) * DE + (k - 64)];
      put(OFF_L3H, OFF_L3L, L3_SLAB, n, k, v);
    }
    for (int i = tid; i < DN * FHP; i += nt) {           // flow layer 1 (K padded 56 -> 64)
      const int n = i / FHP, k = i % FHP;
      put(OFF_L4H, OFF_L4L, L4_SLAB, n, k, k < FH ? f1[n * FH + k] : 0.f);
    }
    float* ft = reinterpret_cast<float*>(img + OFF_F32);
    for (int i = tid; i < F_COUNT; i += nt) {
      float v = 0.f;
      if (i < F_FB0) v = w.edge_b1[i - F_B1];
      else if (i < F_FB1) v = (i - F_FB0) < FH ? fb0[i - F_FB0] : 0.f;
      else if (i < F_CW0) v = fb1[i - F_FB1];
      else if (i < F_CB0) { const int q = i - F_CW0, in = q / CH, o = q % CH; v = w.cls_w0[o * DE + in]; }
      else if (i < F_CW1) v = w.cls_b0[i - F_CB0];
      else if (i < F_CB1) v = w.cls_w1[i - F_CW1];
      else if (i == F_CB1) v = w.cls_b1[0];
      ft[i] = v;
    }
  }
}

// =================================================================== node-side kernels
__device__ __forceinline__ void store_split_row(__half* __restrict__ row64, int lane, float v, int* ovf) {
  const __half h = __float2half_rn(v);
  const __half l = __float2half_rn(v - __half2float(h));
  row64[lane] = h;
  row64[32 + lane] = l;
  if (!(fabsf(v) < 65000.f)) *ovf = 1;
}

// Once per forward: split x_init, and the hoisted row terms
//   pinit[r] = W0[:, 0:32] x_init[r] + b0 ; prow[r] = pinit[r] + W0[:, 32:64] x_init[r]  (x_lat = x_init before step 1)
__global__ void __launch_bounds__(256) prep_nodes_kernel(const float* __restrict__ x_init, int64_t n,
                                                         const float* __restrict__ w0, const float* __restrict__ b0,
                                                         __half* __restrict__ xi, __half* __restrict__ xl0,
                                                         float* __restrict__ pinit, float* __restrict__ prow,
                                                         int32_t* __restrict__ status) {
  __shared__ float s_w[64 * EH];     // [i][o], i over W0 columns 0..63
  __shared__ float s_b[EH];
  for (int idx = threadIdx.x; idx < 64 * EH; idx += blockDim.x) {        // coalesced along a weight row
    const int o = idx / 64, i = idx - o * 64;
    s_w[i * EH + o] = w0[o * 160 + i];
  }
  for (int o = threadIdx.x; o < EH; o += blockDim.x) s_b[o] = b0[o];
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  int ovf = 0;
  for (int64_t r = warp; r < n; r += nwarps) {
    const float v = x_init[r * DN + lane];
    store_split_row(xi + r * 64, lane, v, &ovf);
    store_split_row(xl0 + r * 64, lane, v, &ovf);
    float a0[3], a1[3];
#pragma unroll
    for (int q = 0; q < 3; ++q) { const int o = lane + 32 * q; a0[q] = o < EH ? s_b[o] : 0.f; a1[q] = 0.f; }
#pragma unroll
    for (int i = 0; i < DN; ++i) {
      const float xv = __shfl_sync(0xffffffffu, v, i);
#pragma unroll
      for (int q = 0; q < 3; ++q) {
        const int o = lane + 32 * q;
        if (o < EH) { a0[q] = fmaf(xv, s_w[i * EH + o], a0[q]); a1[q] = fmaf(xv, s_w[(32 + i) * EH + o], a1[q]); }
      }
    }
#pragma unroll
    for (int q = 0; q < 3; ++q) {
      const int o = lane + 32 * q;
      if (o < EH) { pinit[r * EH + o] = a0[q]; prow[r * EH + o] = a0[q] + a1[q]; }
    }
  }
  if (ovf) atomicOr(status, 1);
}

// Per step: x' = ReLU(Wn [flow_in | flow_out] + bn)  (models/mpn.py:97-99), then the split copy of
// x' for the next step's gathers and prow[r] = pinit[r] + W0[:, 32:64] x'.
__global__ void __launch_bounds__(256, 2) node_tc_kernel(const int32_t* __restrict__ out_ptr, const int32_t* __restrict__ in_ptr,
                                                         int64_t num_nodes, int64_t num_out, int32_t chunks_out,
                                                         int chunk_shift, const float* __restrict__ flow, const float* __restrict__ part,
                                                         const float* __restrict__ node_w, const float* __restrict__ node_b,
                                                         const float* __restrict__ w0, const float* __restrict__ pinit,
                                                         __half* __restrict__ xl_next, float* __restrict__ prow,
                                                         float* __restrict__ x_out, int32_t* __restrict__ status) {
  // Warp per node, lane = output feature.  The lane's column of the node Linear lives in REGISTERS (64 values); a
  // node's input vector is staged in a per-warp shared-memory buffer and read back as 16-byte broadcasts, so a node
  // costs ~120 shared-memory instructions instead of 96 shuffles + 160 loads.  The next node's inputs are loaded while the
  // current one is computed (its row pointers one node earlier), which hides the dependent global round trips.
  __shared__ float s_wn[2 * DN * DN];   // [in][out]
  __shared__ float s_bn[DN];
  __shared__ float s_w0[DN * EH];       // [i][o] over W0 columns 32..63
  __shared__ __align__(16) float s_vec[8][96];
  for (int idx = threadIdx.x; idx < 2 * DN * DN; idx += blockDim.x) {
    const int o = idx / (2 * DN), i = idx - o * 2 * DN;
    s_wn[i * DN + o] = node_w[idx];
  }
  for (int o = threadIdx.x; o < DN; o += blockDim.x) s_bn[o] = node_b[o];
  for (int idx = threadIdx.x; idx < EH * DN; idx += blockDim.x) {
    const int o = idx / DN, i = idx - o * DN;
    s_w0[i * EH + o] = w0[o * 160 + 32 + i];
  }
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  float wn[2 * DN];
#pragma unroll
  for (int i = 0; i < 2 * DN; ++i) wn[i] = s_wn[i * DN + lane];
  const int o2 = lane + 64 < EH ? lane + 64 : lane;                    // lanes >= 16 have no third output (result unused)
  const float bn = s_bn[lane];
  float* vec = s_vec[threadIdx.x >> 5];
  int ovf = 0;

  auto load_ptrs = [&](int64_t r, int32_t* p) { p[0] = in_ptr[r]; p[1] = in_ptr[r + 1]; p[2] = out_ptr[r]; p[3] = out_ptr[r + 1]; };
  // one direction's flow vector: the row sum written by the edge kernel, or its partials combined in fixed order
  auto load_flow = [&](int64_t r, int d, int64_t s0, int64_t s1) {
    const int64_t seg_base = d == 0 ? num_out : 0;
    const int64_t chunk_off = d == 0 ? chunks_out : 0;
    float v = 0.f;
    if (s1 > s0) {
      const int64_t ca = (s0 - seg_base) >> chunk_shift, cb = (s1 - 1 - seg_base) >> chunk_shift;
      if (ca == cb) {
        v = flow[r * 2 * DN + d * DN + lane];
      } else {
        const bool first_in_chunk = ((s0 - seg_base) & ((1 << chunk_shift) - 1)) == 0;
        // the partials of up to four granules are loaded together (independent loads), then added in granule order
        const int more = (int)(cb - ca);
        const float* pp = part + ((chunk_off + ca + 1) * 2) * DN + lane;
        const float v0 = part[((chunk_off + ca) * 2 + (first_in_chunk ? 0 : 1)) * DN + lane];
        const float v1 = pp[0];
        const float v2 = more >= 2 ? pp[2 * DN] : 0.f;
        const float v3 = more >= 3 ? pp[4 * DN] : 0.f;
        v = v0 + v1;
        if (more >= 2) v += v2;
        if (more >= 3) v += v3;
        for (int64_t t = ca + 4; t <= cb; ++t) v += part[((chunk_off + t) * 2) * DN + lane];
      }
    }
    return v;
  };
  auto load_inputs = [&](int64_t r, const int32_t* p, float* fl, float* pin) {
    fl[0] = load_flow(r, 0, p[0], p[1]);                                // flow_in
    fl[1] = load_flow(r, 1, p[2], p[3]);                                // flow_out
#pragma unroll
    for (int q = 0; q < 3; ++q) { const int o = lane + 32 * q; pin[q] = o < EH ? pinit[r * EH + o] : 0.f; }
  };

  int64_t r = warp;
  float fl[2] = {0.f, 0.f}, pin[3] = {0.f, 0.f, 0.f};
  int32_t pn[4] = {0, 0, 0, 0};
  if (r < num_nodes) {
    int32_t p[4];
    load_ptrs(r, p);
    load_inputs(r, p, fl, pin);
    if (r + nwarps < num_nodes) load_ptrs(r + nwarps, pn);
  }
  for (; r < num_nodes; r += nwarps) {
    const int64_t r1 = r + nwarps, r2 = r1 + nwarps;
    float nfl[2] = {0.f, 0.f}, npin[3] = {0.f, 0.f, 0.f};
    int32_t p2[4] = {0, 0, 0, 0};
    if (r1 < num_nodes) load_inputs(r1, pn, nfl, npin);
    if (r2 < num_nodes) load_ptrs(r2, p2);

    vec[lane] = fl[0];
    vec[DN + lane] = fl[1];
    __syncwarp();
    float acc = bn;
#pragma unroll
    for (int i = 0; i < 2 * DN / 4; ++i) {
      const float4 v = *reinterpret_cast<const float4*>(vec + 4 * i);
      acc = fmaf(v.x, wn[4 * i], acc);
      acc = fmaf(v.y, wn[4 * i + 1], acc);
      acc = fmaf(v.z, wn[4 * i + 2], acc);
      acc = fmaf(v.w, wn[4 * i + 3], acc);
    }
    const float xn = fmaxf(acc, 0.f);
    vec[2 * DN + lane] = xn;
    __syncwarp();
    if (x_out != nullptr) x_out[r * DN + lane] = xn;
    store_split_row(xl_next + r * 64, lane, xn, &ovf);
    float a[3] = {pin[0], pin[1], pin[2]};
#pragma unroll
    for (int i = 0; i < DN / 4; ++i) {
      const float4 v = *reinterpret_cast<const float4*>(vec + 2 * DN + 4 * i);
      const float* wr = s_w0 + 4 * i * EH;
      const float xs[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        a[0] = fmaf(xs[u], wr[u * EH + lane], a[0]);
        a[1] = fmaf(xs[u], wr[u * EH + lane + 32], a[1]);
        a[2] = fmaf(xs[u], wr[u * EH + o2], a[2]);
      }
    }
#pragma unroll
    for (int q = 0; q < 3; ++q) {
      const int o = lane + 32 * q;
      if (o < EH) prow[r * EH + o] = a[q];
    }
    __syncwarp();                                                       // vec is rewritten by the next node
    fl[0] = nfl[0]; fl[1] = nfl[1];
#pragma unroll
    for (int q = 0; q < 3; ++q) pin[q] = npin[q];
#pragma unroll
    for (int q = 0; q < 4; ++q) pn[q] = p2[q];
  }
  if (ovf) atomicOr(status, 1);
}

// e (fp32 [E,16], slot order) -> split rows [hi(16) | lo(16)] halfs
__global__ void split_edges_kernel(const float* __restrict__ e, int64_t num_edges, uint4* __restrict__ out,
                                   int32_t* __restrict__ status) {
  int ovf = 0;
  for (int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; s < num_edges; s += (int64_t)gridDim.x * blockDim.x) {
    const float4* p = reinterpret_cast<const float4*>(e + s * DE);
    uint32_t hi[8], lo[8];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const float4 v = __ldg(p + q);
      split2(v.x, v.y, hi[2 * q], lo[2 * q]);
      split2(v.z, v.w, hi[2 * q + 1], lo[2 * q + 1]);
      if (!(fmaxf(fmaxf(fabsf(v.x), fabsf(v.y)), fmaxf(fabsf(v.z), fabsf(v.w))) < 65000.f)) ovf = 1;
    }
    out[s * 4 + 0] = make_uint4(hi[0], hi[1], hi[2], hi[3]);
    out[s * 4 + 1] = make_uint4(hi[4], hi[5], hi[6], hi[7]);
    out[s * 4 + 2] = make_uint4(lo[0], lo[1], lo[2], lo[3]);
    out[s * 4 + 3] = make_uint4(lo[4], lo[5], lo[6], lo[7]);
  }
  if (ovf) atomicOr(status, 1);
}

// split rows -> fp32 (final edge state for callers that ask for it)
__global__ void unsplit_edges_kernel(const uint4* __restrict__ in, int64_t num_edges, float* __restrict__ e) {
  for (int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; s < num_edges; s += (int64_t)gridDim.x * blockDim.x) {
    const uint4 h0 = in[s * 4], h1 = in[s * 4 + 1], l0 = in[s * 4 + 2], l1 = in[s * 4 + 3];
    const uint32_t hw[8] = {h0.x, h0.y, h0.z, h0.w, h1.x, h1.y, h1.z, h1.w};
    const uint32_t lw[8] = {l0.x, l0.y, l0.z, l0.w, l1.x, l1.y, l1.z, l1.w};
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&hw[j]));
      const float2 b = __half22float2(*reinterpret_cast<const __half2*>(&lw[j]));
      e[s * DE + 2 * j] = a.x + b.x;
      e[s * DE + 2 * j + 1] = a.y + b.y;
    }
  }
}

// =================================================================== the edge kernel
struct TcArgs {
  const int32_t* slot_row; const int32_t* slot_col; const int32_t* slot_edge;
  int64_t num_edges, num_out;
  int32_t tiles_out, tiles_in;
  int64_t chunks_out;    // number of 32-slot chunks of the flow_out group
  const uint4* xi;       // [N][8]  x_init split rows (128 B)
  const uint4* xl;       // [N][8]  x_lat split rows
  const float* prow;     // [N][80] hoisted row term (incl. bias)
  const uint4* ei;       // [E][4]  e_init split rows (64 B)
  const uint4* es_in;    // [E][4]  e split rows
  uint4* es_out;         // may alias es_in
  float* flow; float* part; float* logits;
  const uint8_t* wimg_out; const uint8_t* wimg_in;
  int32_t* status;
  long long* trace;      // optional [16 stamps x 64 tiles] cycle trace of CTA 0 / group 0 (development)
};

__device__ __forceinline__ void ld_f32x16(float (&d)[16], const float* __restrict__ p) {
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const float4 v = *reinterpret_cast<const float4*>(p + 4 * q);
    d[4 * q] = v.x; d[4 * q + 1] = v.y; d[4 * q + 2] = v.z; d[4 * q + 3] = v.w;
  }
}
__device__ __forceinline__ void ld_f32x8(float (&d)[8], const float* __restrict__ p) {
#pragma unroll
  for (int q = 0; q < 2; ++q) {
    const float4 v = *reinterpret_cast<const float4*>(p + 4 * q);
    d[4 * q] = v.x; d[4 * q + 1] = v.y; d[4 * q + 2] = v.z; d[4 * q + 3] = v.w;
  }
}

// relu(acc + add) for 16 values, split to fp16 hi/lo words; accumulates the overflow detector
__device__ __forceinline__ void relu_split16(const uint32_t (&acc)[16], const float* add, uint32_t (&hi)[8],
                                             uint32_t (&lo)[8], uint32_t& ovf) {
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const float a = fmaxf(__uint_as_float(acc[2 * j]) + add[2 * j], 0.f);
    const float b = fmaxf(__uint_as_float(acc[2 * j + 1]) + add[2 * j + 1], 0.f);
    split2(a, b, hi[j], lo[j]);
    ovf |= hi[j] + 0x04000400u;        // fp16 inf (0x7C00) + 0x0400 sets the half's top bit
  }
}

__global__ void __launch_bounds__(NTHREADS, 1) mp_edge_tc_kernel(TcArgs a) {
  extern __shared__ __align__(128) uint8_t smem[];
  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);            // provably warp-uniform

  // CTAs [0, n_out_ctas) walk flow_out tiles, the rest walk flow_in tiles.
  const int total_tiles = a.tiles_out + a.tiles_in;
  int n_out_ctas = (int)(((int64_t)gridDim.x * a.tiles_out + total_tiles - 1) / total_tiles);
  if (a.tiles_out > 0 && n_out_ctas == 0) n_out_ctas = 1;
  if (a.tiles_in > 0 && n_out_ctas >= (int)gridDim.x) n_out_ctas = gridDim.x - 1;
  if (a.tiles_in == 0) n_out_ctas = gridDim.x;
  const bool dir_out = (int)blockIdx.x < n_out_ctas;
  const int cta_in_dir = dir_out ? blockIdx.x : blockIdx.x - n_out_ctas;
  const int ctas_in_dir = dir_out ? n_out_ctas : gridDim.x - n_out_ctas;
  const int tiles_dir = dir_out ? a.tiles_out : a.tiles_in;
  const int64_t seg_base = dir_out ? 0 : a.num_out;
  const int64_t seg_end = dir_out ? a.num_out : a.num_edges;

  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + SM_BAR);       // d_ready[group]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + SM_TMEM);
  {
    const uint4* src = reinterpret_cast<const uint4*>(dir_out ? a.wimg_out : a.wimg_in);
    uint4* dst = reinterpret_cast<uint4*>(smem);
    for (int i = tid; i < IMG_BYTES / 16; i += NTHREADS) dst[i] = __ldg(src + i);
  }
  if (tid == 0) {
    mbar_init(&bars[0], 1);
    mbar_init(&bars[1], 1);
    mbar_fence_init();
  }
  if (warp == 0) tmem_alloc<512>(tmem_slot);
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tbase = *tmem_slot;
  const float* s_f = reinterpret_cast<const float*>(smem + OFF_F32);

  // Two groups of 8 warps, one 128-edge tile in flight each; a group owns 256 TMEM columns.  Every
  // edge (TMEM lane) is served by TWO threads: half A (warps 0-3 of the group) and half B (warps 4-7)
  // split the columns of each epilogue, the operand loads and the bookkeeping.
  const int g = warp >> 3;
  const bool half_b = ((warp >> 2) & 1) != 0;
  const int wq = warp & 3;                                            // quarter of the tile = this warp's TMEM lanes
  const int gt = wq * 32 + lane;                                      // edge row inside the tile
  const uint32_t tcol = __shfl_sync(0xffffffffu, tbase, 0) + (uint32_t)g * 256u;   // uniform: MMA operand base
  const uint32_t tlane = tcol + ((uint32_t)(wq * 32) << 16);
  float* s_msg = reinterpret_cast<float*>(smem + SM_MSG) + (g * 4 + wq) * CHUNK * MSG_LD;
  const uint32_t stage_warp = smem_u32(smem + SM_STAGE + (g * TS + wq * CHUNK) * STAGE_ROW);
  const uint4* s_stage = reinterpret_cast<const uint4*>(smem + SM_STAGE + (g * TS + gt) * STAGE_ROW);
  uint64_t* d_ready = &bars[g];
  const int64_t chunk_off = dir_out ? 0 : a.chunks_out;
  const int dir_off = dir_out ? DN : 0;                               // cat(flow_in, flow_out), mpn.py:97
  uint32_t pd = 0;
  __half2 vmax = __floats2half2_rn(0.f, 0.f);                         // running max of every hi word (overflow check)

  // ---- MMA issue: run by the group's first warp (all lanes, uniform operands), one elected lane issues.
  const uint64_t dbase = smem_desc_kmajor(smem_u32(smem), 128, 256);   // + (byte offset >> 4) per slab
  auto step3 = [&](uint32_t d, uint32_t ah, uint32_t al, int off_h, int off_l, uint32_t idesc, bool first) {
    const uint64_t dh = dbase + (uint64_t)(off_h >> 4), dl = dbase + (uint64_t)(off_l >> 4);
    mma_ts(d, ah, dh, idesc, first ? 0u : 1u);
    mma_ts(d, ah, dl, idesc, 1u);
    mma_ts(d, al, dh, idesc, 1u);
  };
  auto issue_layer = [&](int layer) {
    tc_fence_after();
    if (elect_one()) {
      const uint32_t cb = tcol;
      if (layer == 1) {
#pragma unroll
        for (int ks = 0; ks < 4; ++ks)
          step3(cb + C_D1, cb + C_XCH + 8 * ks, cb + C_XCL + 8 * ks, OFF_L1H + ks * L1_SLAB, OFF_L1L + ks * L1_SLAB,
                idesc_f16(128, EH), ks == 0);
#pragma unroll
        for (int ks = 0; ks < 2; ++ks)
          step3(cb + C_D1, cb + C_EH + 8 * ks, cb + C_EL + 8 * ks, OFF_L1H + (4 + ks) * L1_SLAB,
                OFF_L1L + (4 + ks) * L1_SLAB, idesc_f16(128, EH), false);
      } else if (layer == 2) {
#pragma unroll
        for (int ks = 0; ks < L2_KS; ++ks)
          step3(cb + C_D2, cb + C_A2 + 16 * ks, cb + C_A2 + 16 * ks + 8, OFF_L2H + ks * L2_SLAB, OFF_L2L + ks * L2_SLAB,
                idesc_f16(128, DE), ks == 0);
      } else if (layer == 3) {
#pragma unroll
        for (int ks = 0; ks < 4; ++ks)
          step3(cb + C_D3, cb + C_XCH + 8 * ks, cb + C_XCL + 8 * ks, OFF_L3H + ks * L3_SLAB, OFF_L3L + ks * L3_SLAB,
                idesc_f16(128, FHP), ks == 0);
        step3(cb + C_D3, cb + C_A3, cb + C_A3 + 8, OFF_L3H + 4 * L3_SLAB, OFF_L3L + 4 * L3_SLAB, idesc_f16(128, FHP),
              false);
      } else {
#pragma unroll
        for (int ks = 0; ks < L4_KS; ++ks)
          step3(cb + C_D4, cb + C_A4 + 16 * ks, cb + C_A4 + 16 * ks + 8, OFF_L4H + ks * L4_SLAB, OFF_L4L + ks * L4_SLAB,
                idesc_f16(128, DN), ks == 0);
      }
      mma_commit(d_ready);
    }
    __syncwarp();
  };
  // operands written to TMEM by all 256 threads of the group -> visible to the MMAs of `layer`
  auto publish_and_issue = [&](int layer) {
    tc_wait_st();
    tc_fence_before();
    named_barrier(1 + g, 2 * TS);
    if (!half_b && wq == 0) issue_layer(layer);
  };
  // one accumulator chunk (16 fp32 columns) -> + add -> ReLU -> fp16 hi/lo, written back in place
  auto epilogue_chunk = [&](int col, const float* add) {
    uint32_t acc[16];
    tmem_ld16(tlane + col, acc);
    tc_wait_ld();
    uint32_t hi[8], lo[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      split2_relu(__uint_as_float(acc[2 * j]) + add[2 * j], __uint_as_float(acc[2 * j + 1]) + add[2 * j + 1], hi[j], lo[j]);
      vmax = __hmax2(vmax, *reinterpret_cast<const __half2*>(&hi[j]));
    }
    tmem_st8(tlane + col, hi);
    tmem_st8(tlane + col + 8, lo);
  };

  // Operand rows of the warp's 32 edges -> staging rows (x_init[c] 128 B at +0, x_lat[c] 128 B at +128,
  // e_init 64 B at +256, e 64 B at +320).  Lanes cooperate so that every request covers whole 128-B
  // lines.  Half A fetches (and later stores to TMEM) the node rows, half B the edge rows.
  auto prefetch_nodes = [&](int32_t c) {
    const int sub8 = lane >> 3, pc8 = lane & 7;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int row = i * 4 + sub8;
      const int32_t cr = __shfl_sync(0xffffffffu, c, row);
      cp_async16(stage_warp + row * STAGE_ROW + pc8 * 16, a.xi + (int64_t)cr * 8 + pc8);
      cp_async16(stage_warp + row * STAGE_ROW + 128 + pc8 * 16, a.xl + (int64_t)cr * 8 + pc8);
    }
  };
  auto prefetch_edges = [&](int64_t chunk_slot0, int64_t last_slot) {
    const int sub4 = lane >> 2, pc4 = lane & 3;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int row = i * 8 + sub4;
      int64_t sl = chunk_slot0 + row;
      sl = sl < last_slot ? sl : last_slot;                            // clamp: loads stay in range
      cp_async16(stage_warp + row * STAGE_ROW + 256 + pc4 * 16, a.ei + sl * 4 + pc4);
      cp_async16(stage_warp + row * STAGE_ROW + 320 + pc4 * 16, a.es_in + sl * 4 + pc4);
    }
  };

  // Per-tile indices are loaded two tiles ahead so that no load latency sits on the tile's chain.
  struct TileIdx { int64_t base; int cnt; int32_t r, x, nb; bool have; };   // x: col (half A) / slot_edge (half B)
  auto load_idx = [&](int p) {
    TileIdx t;
    t.have = 2 * p + g < tiles_dir;
    t.base = 0; t.cnt = 0; t.r = 0; t.x = 0; t.nb = -1;
    if (t.have) {
      t.base = seg_base + (int64_t)(2 * p + g) * TS;
      t.cnt = (int)(seg_end - t.base < TS ? seg_end - t.base : TS);
      const int64_t slot = gt < t.cnt ? t.base + gt : t.base + t.cnt - 1;   // clamp: loads stay in range
      t.r = a.slot_row[slot];
      if (!half_b) {
        t.x = a.slot_col[slot];
      } else {
        if (a.logits != nullptr) t.x = a.slot_edge[slot];
        // rows adjacent to this warp's chunk (lane 0: slot before, lane 31: slot after), for the row sums
        const int64_t cs = t.base + wq * CHUNK;
        const int cw = t.cnt - wq * CHUNK;
        // one predicated load (lane 0: the slot before the chunk, lane 31: the slot after it)
        const bool want = (lane == 0 && cw > 0 && cs > seg_base) || (lane == 31 && cw >= CHUNK && cs + CHUNK < seg_end);
        if (want) t.nb = a.slot_row[lane == 0 ? cs - 1 : cs + CHUNK];
      }
    }
    return t;
  };
  // Row sums of one tile's messages (in s_msg), this warp's 32 slots, in slot order (half B).
  auto row_sums = [&](const TileIdx& t) {
    const int64_t cs = t.base + wq * CHUNK;
    int cw = t.cnt - wq * CHUNK;
    cw = cw < 0 ? 0 : (cw > CHUNK ? CHUNK : cw);
    if (cw > 0) {                                                     // warp-uniform
      const int f = lane;                                             // lane = feature from here on
      const int64_t chunk_id = chunk_off + (cs - seg_base) / CHUNK;
      const int32_t r_prev = __shfl_sync(0xffffffffu, t.nb, 0);
      const int32_t r_next = __shfl_sync(0xffffffffu, t.nb, 31);
      const int32_t r_after = __shfl_down_sync(0xffffffffu, t.r, 1);
      const bool seg_end_here = lane < cw && (lane == cw - 1 || r_after != t.r);
      const unsigned ends = __ballot_sync(0xffffffffu, seg_end_here); // bit q: slot q closes a row segment
      float mv[CHUNK];
#pragma unroll
      for (int q = 0; q < CHUNK; ++q) mv[q] = s_msg[q * MSG_LD + f];
      float sum = 0.f;
      bool first_seg = true;
#pragma unroll
      for (int q = 0; q < CHUNK; ++q) {
        sum += mv[q];
        if ((ends >> q) & 1u) {                                       // warp-uniform
          const int32_t cur = __shfl_sync(0xffffffffu, t.r, q);
          const bool starts_before = first_seg && r_prev == cur;
          const bool continues = q == cw - 1 && r_next == cur;
          if (!starts_before && !continues) a.flow[(int64_t)cur * 2 * DN + dir_off + f] = sum;
          else a.part[(chunk_id * 2 + (first_seg ? 0 : 1)) * DN + f] = sum;
          sum = 0.f;
          first_seg = false;
        }
      }
    }
    __syncwarp();
  };

  TileIdx cur = load_idx(cta_in_dir);
  TileIdx nxt = load_idx(cta_in_dir + ctas_in_dir);
  TileIdx prev;
  prev.have = false; prev.base = 0; prev.cnt = 0; prev.r = 0; prev.x = 0; prev.nb = -1;
  if (cur.have) {
    if (!half_b) prefetch_nodes(cur.x);
    else prefetch_edges(cur.base + wq * CHUNK, cur.base + cur.cnt - 1);
  }
  int p = cta_in_dir;
  int trace_i = 0;
#define TC_STAMP(k) do { if (a.trace != nullptr && blockIdx.x == 0 && tid == 0 && trace_i < 64) a.trace[trace_i * 16 + (k)] = clock64(); } while (0)
  while (cur.have) {
    const bool valid = gt < cur.cnt;
    TC_STAMP(0);
    // hoisted row term of epilogue 1, this half's columns (A: 0..47, B: 48..79); consumed after layer 1
    float pr[48];
    {
      const float4* prp = reinterpret_cast<const float4*>(a.prow + (int64_t)cur.r * EH) + (half_b ? 12 : 0);
#pragma unroll
      for (int j = 0; j < 12; ++j) {
        if (half_b && j >= 8) break;
        const float4 v = __ldg(prp + j);
        pr[4 * j] = v.x; pr[4 * j + 1] = v.y; pr[4 * j + 2] = v.z; pr[4 * j + 3] = v.w;
      }
    }
    // ---- load phase: staged operands -> TMEM (A: node rows, B: edge rows), then layer 1
    cp_async_wait_all();
    __syncwarp();                                                     // rows were fetched by other lanes of this warp
    TC_STAMP(1);
    {
      auto st2 = [&](int col, int j) {
        const uint4 x = s_stage[j], y = s_stage[j + 1];
        const uint32_t w8[8] = {x.x, x.y, x.z, x.w, y.x, y.y, y.z, y.w};
        tmem_st8(tlane + col, w8);
      };
      if (!half_b) {
        st2(C_XCH + 0, 0);   st2(C_XCH + 8, 2);      // x_init hi
        st2(C_XCL + 0, 4);   st2(C_XCL + 8, 6);      // x_init lo
        st2(C_XCH + 16, 8);  st2(C_XCH + 24, 10);    // x_lat hi
        st2(C_XCL + 16, 12); st2(C_XCL + 24, 14);    // x_lat lo
      } else {
        st2(C_EH + 0, 16);   st2(C_EL + 0, 18);      // e_init hi / lo
        st2(C_EH + 8, 20);   st2(C_EL + 8, 22);      // e hi / lo
      }
    }
    publish_and_issue(1);
    TC_STAMP(2);
    // ---- while layer 1 runs: operands of the next tile, indices of the one after, row sums of the previous
    if (nxt.have) {
      if (!half_b) prefetch_nodes(nxt.x);
      else prefetch_edges(nxt.base + wq * CHUNK, nxt.base + nxt.cnt - 1);
    }
    p += ctas_in_dir;
    TileIdx nn = load_idx(p + ctas_in_dir);
    if (half_b && prev.have) row_sums(prev);
    TC_STAMP(3);

    // ---- epilogue 1: h = ReLU(D1 + prow[r]) -> layer-2 operand (A: chunks 0-2, B: chunks 3-4)
    mbar_wait(d_ready, pd); pd ^= 1;
    TC_STAMP(4);
    tc_fence_after();
    if (!half_b) {
#pragma unroll
      for (int ch = 0; ch < 3; ++ch) epilogue_chunk(C_D1 + 16 * ch, pr + 16 * ch);
    } else {
#pragma unroll
      for (int ch = 0; ch < 2; ++ch) epilogue_chunk(C_D1 + 48 + 16 * ch, pr + 16 * ch);
    }
    publish_and_issue(2);
    TC_STAMP(5);

    // ---- epilogue 2: e' = ReLU(D2 + b1). A: state + layer-3 operand; B: classifier
    mbar_wait(d_ready, pd); pd ^= 1;
    TC_STAMP(6);
    tc_fence_after();
    {
      uint32_t acc[16];
      tmem_ld16(tlane + C_D2, acc);
      float add[16];
      ld_f32x16(add, s_f + F_B1);
      tc_wait_ld();
      if (!half_b) {
        uint32_t hi[8], lo[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          split2_relu(__uint_as_float(acc[2 * j]) + add[2 * j], __uint_as_float(acc[2 * j + 1]) + add[2 * j + 1], hi[j], lo[j]);
          vmax = __hmax2(vmax, *reinterpret_cast<const __half2*>(&hi[j]));
        }
        tmem_st8(tlane + C_A3, hi);
        tmem_st8(tlane + C_A3 + 8, lo);
        publish_and_issue(3);
        TC_STAMP(7);
        if (valid) {
          uint4* dst = a.es_out + (cur.base + gt) * 4;
          dst[0] = make_uint4(hi[0], hi[1], hi[2], hi[3]);
          dst[1] = make_uint4(hi[4], hi[5], hi[6], hi[7]);
          dst[2] = make_uint4(lo[0], lo[1], lo[2], lo[3]);
          dst[3] = make_uint4(lo[4], lo[5], lo[6], lo[7]);
        }
      } else {
        // B only reads D2 (A writes the e' operand to its own columns); it still takes part in the group
        // barrier that precedes layer 3.
        publish_and_issue(3);
        if (valid && a.logits != nullptr) {                            // classifier 16 -> 8 -> 1 (fp32)
          float hc[CH], wv[CH];
          ld_f32x8(hc, s_f + F_CB0);
#pragma unroll
          for (int i = 0; i < DE; ++i) {
            const float ei = fmaxf(__uint_as_float(acc[i]) + add[i], 0.f);
            ld_f32x8(wv, s_f + F_CW0 + i * CH);
#pragma unroll
            for (int o = 0; o < CH; ++o) hc[o] = fmaf(ei, wv[o], hc[o]);
          }
          ld_f32x8(wv, s_f + F_CW1);
          float lg = s_f[F_CB1];
#pragma unroll
          for (int o = 0; o < CH; ++o) lg = fmaf(fmaxf(hc[o], 0.f), wv[o], lg);
          a.logits[cur.x] = lg;
        }
      }
    }
    // ---- epilogue 3: g = ReLU(D3 + fb0) -> layer-4 operand (A: chunks 0-1, B: chunks 2-3)
    TC_STAMP(8);
    mbar_wait(d_ready, pd); pd ^= 1;
    TC_STAMP(9);
    tc_fence_after();
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      const int ch = (half_b ? 2 : 0) + q;
      float add[16];
      ld_f32x16(add, s_f + F_FB0 + 16 * ch);
      epilogue_chunk(C_D3 + 16 * ch, add);
    }
    publish_and_issue(4);
    TC_STAMP(10);

    // ---- epilogue 4: m = ReLU(D4 + fb1) -> shared memory (A: features 0-15, B: 16-31); the row sums
    //      run under the next tile's layer 1
    mbar_wait(d_ready, pd); pd ^= 1;
    TC_STAMP(11);
    tc_fence_after();
    {
      const int ch = half_b ? 1 : 0;
      uint32_t acc[16];
      tmem_ld16(tlane + C_D4 + 16 * ch, acc);
      float add[16];
      ld_f32x16(add, s_f + F_FB1 + 16 * ch);
      tc_wait_ld();
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const float m = fmaxf(__uint_as_float(acc[j]) + add[j], 0.f);
        s_msg[lane * MSG_LD + 16 * ch + j] = valid ? m : 0.f;
      }
    }
    tc_fence_before();
    TC_STAMP(12);
    ++trace_i;
    prev = cur; cur = nxt; nxt = nn;
  }
  if (prev.have) {                                                    // row sums of the group's last tile
    named_barrier(1 + g, 2 * TS);
    if (half_b) row_sums(prev);
  }
  {
    // hi parts are truncated (rz), so a value beyond the fp16 range shows up as the largest finite
    // fp16 (0x7BFF = 65504) or inf: flag both (conservative).
    const uint32_t w = *reinterpret_cast<const uint32_t*>(&vmax);
    if ((w & 0x7FFFu) >= 0x7BFFu || ((w >> 16) & 0x7FFFu) >= 0x7BFFu) atomicOr(a.status, 1);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc<512>(tbase);
}

// =================================================================== variant 3: three tiles in flight
// Same arithmetic as mp_edge_tc_kernel, different resource plan:
//  * the gathered x[col] rows are fetched by TMA (tile::gather4, 4 rows per instruction, 8 lanes per warp issue)
//    straight into 128B-swizzled K-major shared-memory tiles and consumed by SS-mode MMAs (layers 1 and 3) - no
//    per-thread loads, no register / TMEM staging of the widest operand;
//  * the tile's own edge rows (e_init, e) go global -> registers -> TMEM by the half of the group that is idle
//    during epilogue 2;
//  * a tile needs 144 TMEM columns and 48 KB of shared memory, so THREE groups of 8 warps keep three tiles in
//    flight per SM (24 warps); every edge row is served by two threads (halves A/B split the columns);
//  * layer 1 of the next tile is issued as soon as layer 4 of the current one has retired, so epilogue 4 and the
//    row sums run under it; row sums are split over all 8 warps (16 features each, 16-slot partial granule);
//  * the classifier's first layer rides in the 8 padding columns of flow layer 0 (computed by the tensor core).
constexpr int NG3 = 3, GT3 = 2 * TS, NTHREADS3 = NG3 * GT3;
constexpr int CHUNK3 = 16, CHUNK3_SHIFT = 4;                                   // row-sum partial granule (slots)
constexpr int T3_D1 = 0, T3_D2 = 80, T3_A3 = 96, T3_D3 = 0, T3_D4 = 80, T3_EH = 112, T3_EL = 128, T3_COLS = 144;
constexpr int A_SLAB = TS * 32;                                                // one K=16 step of a 128-row A tile
// group region: x_init[col] tile (128 rows x [hi 64 B | lo 64 B], SW128), x_lat[col] tile, message buffers
constexpr int G3_XI = 0, G3_XL = 4 * A_SLAB, G3_MSG = 8 * A_SLAB, G3_BYTES = 12 * A_SLAB;
constexpr int SM3_GRP = (IMG_BYTES + 1023) / 1024 * 1024;                      // swizzled tiles need 1 KB alignment
constexpr int SM3_BAR = SM3_GRP + NG3 * G3_BYTES;                              // d_ready[3], xc_ready[3]
constexpr int SM3_TMEM = SM3_BAR + 8 * 8;
constexpr int SMEM3_BYTES = SM3_TMEM + 16;

__global__ void __launch_bounds__(NTHREADS3, 1) mp_edge_tc3_kernel(TcArgs a, const __grid_constant__ CUtensorMap tm_xi,
                                                                    const __grid_constant__ CUtensorMap tm_xl) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);

  const int total_tiles = a.tiles_out + a.tiles_in;
  int n_out_ctas = (int)(((int64_t)gridDim.x * a.tiles_out + total_tiles - 1) / total_tiles);
  if (a.tiles_out > 0 && n_out_ctas == 0) n_out_ctas = 1;
  if (a.tiles_in > 0 && n_out_ctas >= (int)gridDim.x) n_out_ctas = gridDim.x - 1;
  if (a.tiles_in == 0) n_out_ctas = gridDim.x;
  const bool dir_out = (int)blockIdx.x < n_out_ctas;
  const int cta_in_dir = dir_out ? blockIdx.x : blockIdx.x - n_out_ctas;
  const int ctas_in_dir = dir_out ? n_out_ctas : gridDim.x - n_out_ctas;
  const int tiles_dir = dir_out ? a.tiles_out : a.tiles_in;
  const int64_t seg_base = dir_out ? 0 : a.num_out;
  const int64_t seg_end = dir_out ? a.num_out : a.num_edges;

  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + SM3_BAR);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + SM3_TMEM);
  {
    const uint4* src = reinterpret_cast<const uint4*>(dir_out ? a.wimg_out : a.wimg_in);
    uint4* dst = reinterpret_cast<uint4*>(smem);
    for (int i = tid; i < IMG_BYTES / 16; i += NTHREADS3) dst[i] = __ldg(src + i);
  }
  if (tid == 0) {
    for (int i = 0; i < NG3; ++i) { mbar_init(&bars[i], 1); mbar_init(&bars[NG3 + i], 4); }
    mbar_fence_init();
  }
  if (warp == 0) tmem_alloc<512>(tmem_slot);
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tbase = *tmem_slot;
  const float* s_f = reinterpret_cast<const float*>(smem + OFF_F32);

  const int g = warp >> 3;
  const bool half_b = ((warp >> 2) & 1) != 0;
  const int hb = half_b ? 1 : 0;
  const int wq = warp & 3;
  const int gt = wq * 32 + lane;                                       // edge row inside the tile = TMEM lane
  const uint32_t tcol = __shfl_sync(0xffffffffu, tbase, 0) + (uint32_t)g * T3_COLS;
  const uint32_t tlane = tcol + ((uint32_t)(wq * 32) << 16);
  uint8_t* grp = smem + SM3_GRP + g * G3_BYTES;
  const uint32_t grp_addr = smem_u32(grp);
  // this warp's private message buffer: 32 rows x 16 features (its half of the columns), bank-conflict-free
  // both for the row-major writes (lane = row) and the feature-major reads of the row sums
  float* s_msg = reinterpret_cast<float*>(grp + G3_MSG) + (wq * 2 + hb) * 32 * 16;
  auto msg_at = [&](int row, int f) { return s_msg + ((row ^ ((row >> 4) & 1)) << 4) + ((f + (row >> 1)) & 15); };
  uint64_t* d_ready = &bars[g];
  uint64_t* xc_ready = &bars[NG3 + g];                                  // 4 gathering warps arrive + 32 KB of TMA bytes
  uint32_t px = 0;
  const int64_t chunk_off = dir_out ? 0 : a.chunks_out;
  const int dir_off = dir_out ? DN : 0;
  const int bar_grp = 1 + g;
  uint32_t pd = 0;
  __half2 vmax = __floats2half2_rn(0.f, 0.f);

  const uint64_t dzero = smem_desc_kmajor(0, 128, 256);
  const uint64_t dsw = smem_desc_sw128(0);
  const uint32_t img = smem_u32(smem);
  auto ss3 = [&](uint32_t d, uint32_t ah, uint32_t al, int off_h, int off_l, uint32_t idesc, bool first) {
    const uint64_t adh = dsw + (uint64_t)(ah >> 4), adl = dsw + (uint64_t)(al >> 4);
    const uint64_t bdh = dzero + (uint64_t)((img + off_h) >> 4), bdl = dzero + (uint64_t)((img + off_l) >> 4);
    mma_ss(d, adh, bdh, idesc, first ? 0u : 1u);
    mma_ss(d, adh, bdl, idesc, 1u);
    mma_ss(d, adl, bdh, idesc, 1u);
  };
  auto ts3 = [&](uint32_t d, uint32_t ah, uint32_t al, int off_h, int off_l, uint32_t idesc, bool first) {
    const uint64_t bdh = dzero + (uint64_t)((img + off_h) >> 4), bdl = dzero + (uint64_t)((img + off_l) >> 4);
    mma_ts(d, ah, bdh, idesc, first ? 0u : 1u);
    mma_ts(d, ah, bdl, idesc, 1u);
    mma_ts(d, al, bdh, idesc, 1u);
  };
  auto issue_layer = [&](int layer) {
    if (layer == 1) { mbar_wait(xc_ready, px); px ^= 1; }               // the tile's x[col] rows have landed (TMA)
    tc_fence_after();
    if (elect_one()) {
      const uint32_t cb = tcol;
      if (layer == 1) {
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {                               // K steps 0,1: x_init, 2,3: x_lat; hi at +0, lo at +64 B
          const uint32_t at = grp_addr + (ks >> 1) * (G3_XL - G3_XI) + (ks & 1) * 32;
          ss3(cb + T3_D1, at, at + 64, OFF_L1H + ks * L1_SLAB, OFF_L1L + ks * L1_SLAB, idesc_f16(128, EH), ks == 0);
        }
#pragma unroll
        for (int ks = 0; ks < 2; ++ks)
          ts3(cb + T3_D1, cb + T3_EH + 8 * ks, cb + T3_EL + 8 * ks, OFF_L1H + (4 + ks) * L1_SLAB,
              OFF_L1L + (4 + ks) * L1_SLAB, idesc_f16(128, EH), false);
      } else if (layer == 2) {
#pragma unroll
        for (int ks = 0; ks < L2_KS; ++ks)
          ts3(cb + T3_D2, cb + T3_D1 + 16 * ks, cb + T3_D1 + 16 * ks + 8, OFF_L2H + ks * L2_SLAB, OFF_L2L + ks * L2_SLAB,
              idesc_f16(128, DE), ks == 0);
      } else if (layer == 3) {
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
          const uint32_t at = grp_addr + (ks >> 1) * (G3_XL - G3_XI) + (ks & 1) * 32;
          ss3(cb + T3_D3, at, at + 64, OFF_L3H + ks * L3_SLAB, OFF_L3L + ks * L3_SLAB, idesc_f16(128, FHP), ks == 0);
        }
        ts3(cb + T3_D3, cb + T3_A3, cb + T3_A3 + 8, OFF_L3H + 4 * L3_SLAB, OFF_L3L + 4 * L3_SLAB, idesc_f16(128, FHP), false);
      } else {
#pragma unroll
        for (int ks = 0; ks < L4_KS; ++ks)
          ts3(cb + T3_D4, cb + T3_D3 + 16 * ks, cb + T3_D3 + 16 * ks + 8, OFF_L4H + ks * L4_SLAB, OFF_L4L + ks * L4_SLAB,
              idesc_f16(128, DN), ks == 0);
      }
      mma_commit(d_ready);
    }
    __syncwarp();
  };
  auto publish_and_issue = [&](int layer) {
    tc_wait_st();
    tc_fence_before();
    named_barrier(bar_grp, GT3);
    if (!half_b && wq == 0) issue_layer(layer);
  };
  auto epilogue_chunk = [&](int col, const float* add) {
    uint32_t acc[16];
    tmem_ld16(tlane + col, acc);
    tc_wait_ld();
    uint32_t hi[8], lo[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      split2_relu(__uint_as_float(acc[2 * j]) + add[2 * j], __uint_as_float(acc[2 * j + 1]) + add[2 * j + 1], hi[j], lo[j]);
      vmax = __hmax2(vmax, *reinterpret_cast<const __half2*>(&hi[j]));
    }
    tmem_st8(tlane + col, hi);
    tmem_st8(tlane + col + 8, lo);
  };

  // x[col] rows of this quarter's 32 edges -> the group's swizzled operand tiles, 4 rows per TMA gather; half A
  auto fetch_nodes = [&](int32_t c) {
    const int l4 = (lane & 7) * 4;
    const int32_t c0 = __shfl_sync(0xffffffffu, c, l4), c1 = __shfl_sync(0xffffffffu, c, l4 + 1);
    const int32_t c2 = __shfl_sync(0xffffffffu, c, l4 + 2), c3 = __shfl_sync(0xffffffffu, c, l4 + 3);
    if (lane == 0) mbar_arrive_expect_tx(xc_ready, 32 * 256);
    __syncwarp();
    if (lane < 8) {
      const uint32_t dst = grp_addr + G3_XI + (wq * 32 + l4) * 128;
      tma_gather4(dst, &tm_xi, xc_ready, c0, c1, c2, c3);
      tma_gather4(dst + (G3_XL - G3_XI), &tm_xl, xc_ready, c0, c1, c2, c3);
    }
  };
  // this thread's edge row [e_init | e] (split halves) -> TMEM operand columns; half B
  struct EdgeRow { uint4 v[8]; };
  auto load_edge_row = [&](int64_t base, int cnt) {
    EdgeRow r;
    const int64_t sl = base + (gt < cnt ? gt : cnt - 1);
    const uint4* p0 = a.ei + sl * 4;
    const uint4* p1 = a.es_in + sl * 4;
#pragma unroll
    for (int j = 0; j < 4; ++j) { r.v[j] = __ldg(p0 + j); r.v[4 + j] = p1[j]; }
    return r;
  };
  auto store_edge_row = [&](const EdgeRow& r) {
    const uint32_t h0[8] = {r.v[0].x, r.v[0].y, r.v[0].z, r.v[0].w, r.v[1].x, r.v[1].y, r.v[1].z, r.v[1].w};
    const uint32_t l0[8] = {r.v[2].x, r.v[2].y, r.v[2].z, r.v[2].w, r.v[3].x, r.v[3].y, r.v[3].z, r.v[3].w};
    const uint32_t h1[8] = {r.v[4].x, r.v[4].y, r.v[4].z, r.v[4].w, r.v[5].x, r.v[5].y, r.v[5].z, r.v[5].w};
    const uint32_t l1[8] = {r.v[6].x, r.v[6].y, r.v[6].z, r.v[6].w, r.v[7].x, r.v[7].y, r.v[7].z, r.v[7].w};
    tmem_st8(tlane + T3_EH, h0);
    tmem_st8(tlane + T3_EL, l0);
    tmem_st8(tlane + T3_EH + 8, h1);
    tmem_st8(tlane + T3_EL + 8, l1);
  };

  struct TileIdx { int64_t base; int cnt; int32_t r, x, nb; bool have; };   // x: col (half A) / slot_edge (half B)
  auto load_idx = [&](int p) {
    TileIdx t;
    t.have = NG3 * p + g < tiles_dir;
    t.base = 0; t.cnt = 0; t.r = 0; t.x = 0; t.nb = -1;
    if (t.have) {
      t.base = seg_base + (int64_t)(NG3 * p + g) * TS;
      t.cnt = (int)(seg_end - t.base < TS ? seg_end - t.base : TS);
      const int64_t slot = gt < t.cnt ? t.base + gt : t.base + t.cnt - 1;
      t.r = a.slot_row[slot];
      if (!half_b) t.x = a.slot_col[slot];
      else if (a.logits != nullptr) t.x = a.slot_edge[slot];
      const int64_t cs = t.base + wq * 32;
      const int cw = t.cnt - wq * 32;
      // one predicated load (lane 0: the slot before the quarter, lane 31: the slot after it)
      const bool want = (lane == 0 && cw > 0 && cs > seg_base) || (lane == 31 && cw >= 32 && cs + 32 < seg_end);
      if (want) t.nb = a.slot_row[lane == 0 ? cs - 1 : cs + 32];
    }
    return t;
  };
  // Row sums of this warp's 16 message features over its quarter's 32 slots, in slot order: lanes (sub, f) walk the
  // two 16-slot granules.  A row's segment that lies inside one granule is written to `flow`; pieces that touch a
  // granule edge go to the granule's two partial slots and are combined by the node kernel in fixed order.
  auto row_sums = [&](const TileIdx& t) {
    const int sub = lane >> 4, fl = lane & 15;
    int cwq = t.cnt - wq * 32;
    cwq = cwq < 0 ? 0 : (cwq > 32 ? 32 : cwq);
    int cw = cwq - 16 * sub;
    cw = cw < 0 ? 0 : (cw > 16 ? 16 : cw);
    const int32_t nb_prev = __shfl_sync(0xffffffffu, t.nb, 0);
    const int32_t nb_next = __shfl_sync(0xffffffffu, t.nb, 31);
    const int32_t r15 = __shfl_sync(0xffffffffu, t.r, 15);
    const int32_t r16 = __shfl_sync(0xffffffffu, t.r, 16);
    const int32_t r_prev = sub ? r15 : nb_prev;
    const int32_t r_next = sub ? nb_next : (cwq > 16 ? r16 : -1);
    const int32_t r_after = __shfl_down_sync(0xffffffffu, t.r, 1);
    const bool seg_end_here = fl < cw && (fl == cw - 1 || r_after != t.r);       // lane doubles as the row index here
    const unsigned ends = (__ballot_sync(0xffffffffu, seg_end_here) >> (16 * sub)) & 0xffffu;
    const int64_t chunk_id = chunk_off + ((t.base + wq * 32 - seg_base) >> CHUNK3_SHIFT) + sub;
    const int f = 16 * hb + fl;
    float sum = 0.f;
    bool first_seg = true;
#pragma unroll
    for (int q = 0; q < CHUNK3; ++q) {
      const int32_t rq = __shfl_sync(0xffffffffu, t.r, 16 * sub + q);
      sum += *msg_at(16 * sub + q, fl);
      if ((ends >> q) & 1u) {
        const bool starts_before = first_seg && r_prev == rq;
        const bool continues = q == cw - 1 && r_next == rq;
        if (!starts_before && !continues) a.flow[(int64_t)rq * 2 * DN + dir_off + f] = sum;
        else a.part[(chunk_id * 2 + (first_seg ? 0 : 1)) * DN + f] = sum;
        sum = 0.f;
        first_seg = false;
      }
    }
    __syncwarp();
  };
  auto load_prow16 = [&](float* add, int32_t r, int ch) {
    const float4* prp = reinterpret_cast<const float4*>(a.prow + (int64_t)r * EH + 16 * ch);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float4 v = __ldg(prp + j);
      add[4 * j] = v.x; add[4 * j + 1] = v.y; add[4 * j + 2] = v.z; add[4 * j + 3] = v.w;
    }
  };

  TileIdx cur = load_idx(cta_in_dir);
  TileIdx nxt = load_idx(cta_in_dir + ctas_in_dir);
  if (cur.have) {
    if (!half_b) {
      fetch_nodes(cur.x);
    } else {
      const EdgeRow er = load_edge_row(cur.base, cur.cnt);
      store_edge_row(er);
    }
    publish_and_issue(1);
  }
  int p = cta_in_dir;
  while (cur.have) {
    const bool valid = gt < cur.cnt;
    p += ctas_in_dir;
    TileIdx nn = load_idx(p + ctas_in_dir);
    if (nxt.have) {                                                   // next tile's hoisted row terms -> L1
      const char* pr = reinterpret_cast<const char*>(a.prow + (int64_t)nxt.r * EH) + (half_b ? 192 : 0);
      asm volatile("prefetch.global.L1 [%0];" ::"l"(pr));
      asm volatile("prefetch.global.L1 [%0];" ::"l"(pr + 127));
    }
    if (half_b && nn.have && (lane & 1) == 0) {                       // pull the edge rows two tiles ahead into L2
      const int64_t sl = nn.base + (gt < nn.cnt ? gt : nn.cnt - 1);
      asm volatile("prefetch.global.L2 [%0];" ::"l"(a.ei + sl * 4));
      asm volatile("prefetch.global.L2 [%0];" ::"l"(a.es_in + sl * 4));
    }

    // ---- epilogue 1: + hoisted x[row] term (A: columns 0..47, B: 48..79)
    {
      const int ch0 = half_b ? 3 : 0;
      float add[16];
      load_prow16(add, cur.r, ch0);
      mbar_wait(d_ready, pd); pd ^= 1;
      tc_fence_after();
      epilogue_chunk(T3_D1 + 16 * ch0, add);
      load_prow16(add, cur.r, ch0 + 1);
      epilogue_chunk(T3_D1 + 16 * (ch0 + 1), add);
      if (!half_b) {
        load_prow16(add, cur.r, 2);
        epilogue_chunk(T3_D1 + 32, add);
      }
    }
    publish_and_issue(2);

    // ---- epilogue 2 (half A): e' -> state + layer-3 operand; half B stages the next tile's edge rows in TMEM
    if (!half_b) {
      mbar_wait(d_ready, pd); pd ^= 1;
      tc_fence_after();
      uint32_t acc[16];
      tmem_ld16(tlane + T3_D2, acc);
      float add[16];
      ld_f32x16(add, s_f + F_B1);
      tc_wait_ld();
      uint32_t hi[8], lo[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        split2_relu(__uint_as_float(acc[2 * j]) + add[2 * j], __uint_as_float(acc[2 * j + 1]) + add[2 * j + 1], hi[j], lo[j]);
        vmax = __hmax2(vmax, *reinterpret_cast<const __half2*>(&hi[j]));
      }
      tmem_st8(tlane + T3_A3, hi);
      tmem_st8(tlane + T3_A3 + 8, lo);
      publish_and_issue(3);
      if (valid) {
        uint4* dst = a.es_out + (cur.base + gt) * 4;
        dst[0] = make_uint4(hi[0], hi[1], hi[2], hi[3]);
        dst[1] = make_uint4(hi[4], hi[5], hi[6], hi[7]);
        dst[2] = make_uint4(lo[0], lo[1], lo[2], lo[3]);
        dst[3] = make_uint4(lo[4], lo[5], lo[6], lo[7]);
      }
    } else {
      if (nxt.have) {                                                  // layer 1 (the only reader) has retired
        const EdgeRow er = load_edge_row(nxt.base, nxt.cnt);
        store_edge_row(er);
      }
      mbar_wait(d_ready, pd); pd ^= 1;
      tc_fence_after();
      publish_and_issue(3);
    }
    // ---- epilogue 3: g -> layer-4 operand (A: columns 0..31, B: 32..63; 56..63 carry the classifier's first layer)
    mbar_wait(d_ready, pd); pd ^= 1;
    tc_fence_after();
    if (!half_b && nxt.have) fetch_nodes(nxt.x);                       // layer 3 was the last reader of x[col]
    {
      const int ch0 = half_b ? 2 : 0;
      float add[16];
      ld_f32x16(add, s_f + F_FB0 + 16 * ch0);
      epilogue_chunk(T3_D3 + 16 * ch0, add);
      ld_f32x16(add, s_f + F_FB0 + 16 * (ch0 + 1));                    // B: entries 8..15 are zero (padding)
      if (!half_b) {
        epilogue_chunk(T3_D3 + 16, add);
      } else {
        uint32_t acc[16];
        tmem_ld16(tlane + T3_D3 + 48, acc);
        float cb0[CH], cw1[CH];
        ld_f32x8(cb0, s_f + F_CB0);
        ld_f32x8(cw1, s_f + F_CW1);
        tc_wait_ld();
        uint32_t hi[8], lo[8];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          split2_relu(__uint_as_float(acc[2 * j]) + add[2 * j], __uint_as_float(acc[2 * j + 1]) + add[2 * j + 1], hi[j], lo[j]);
          vmax = __hmax2(vmax, *reinterpret_cast<const __half2*>(&hi[j]));
        }
#pragma unroll
        for (int j = 4; j < 8; ++j) { hi[j] = 0u; lo[j] = 0u; }        // K padding of layer 4
        tmem_st8(tlane + T3_D3 + 48, hi);
        tmem_st8(tlane + T3_D3 + 56, lo);
        if (valid && a.logits != nullptr) {
          float lg = s_f[F_CB1];
#pragma unroll
          for (int o = 0; o < CH; ++o) lg = fmaf(fmaxf(__uint_as_float(acc[8 + o]) + cb0[o], 0.f), cw1[o], lg);
          a.logits[cur.x] = lg;
        }
      }
    }
    publish_and_issue(4);

    // ---- layer 4 retired -> its operand columns are free: start the next tile's layer 1, then finish this tile
    mbar_wait(d_ready, pd); pd ^= 1;
    tc_fence_after();
    if (nxt.have) publish_and_issue(1);
    {
      uint32_t acc[16];
      tmem_ld16(tlane + T3_D4 + 16 * hb, acc);
      float add[16];
      ld_f32x16(add, s_f + F_FB1 + 16 * hb);
      tc_wait_ld();
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const float m = fmaxf(__uint_as_float(acc[j]) + add[j], 0.f);
        *msg_at(lane, j) = valid ? m : 0.f;
      }
    }
    __syncwarp();
    row_sums(cur);
    cur = nxt; nxt = nn;
  }
  {
    const uint32_t w = *reinterpret_cast<const uint32_t*>(&vmax);
    if ((w & 0x7FFFu) >= 0x7BFFu || ((w >> 16) & 0x7FFFu) >= 0x7BFFu) atomicOr(a.status, 1);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc<512>(tbase);
}

// 2-D tensor map over split node rows [n][64 halfs] for the TMA row gather (box = one row, 128B swizzle)
static int make_row_map(CUtensorMap* out, const void* base, int64_t n) {
  typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                               const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                               CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  static EncodeFn encode = nullptr;
  if (encode == nullptr) {
    cudaDriverEntryPointQueryResult q;
    void* fn = nullptr;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) != cudaSuccess || fn == nullptr) return -1;
    encode = reinterpret_cast<EncodeFn>(fn);
  }
  const cuuint64_t dims[2] = {64, (cuuint64_t)n};
  const cuuint64_t strides[1] = {128};
  const cuuint32_t box[2] = {64, 1};
  const cuuint32_t estr[2] = {1, 1};
  return encode(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS ? 0 : -1;
}

int variant() {
  static int v = 0;
  if (v == 0) {
    const char* e = getenv("MPN_TC_VARIANT");
    v = (e != nullptr && e[0] == '2') ? 2 : 3;
  }
  return v;
}


long long* g_trace = nullptr;   // set by mpn_tc_set_trace (development only)

struct TcWorkspace {
  __half* xi; __half* xl[2];
  float* pinit; float* prow;
  uint4* ei; uint4* es;
  float* flow; float* part;
  uint8_t* wimg_out; uint8_t* wimg_in;
};

static int64_t carve(void* ws, int64_t n, int64_t e, TcWorkspace* out) {
  Carver cv(ws);
  const int64_t chunks = ceil_div(e, CHUNK3) + 8;          // sized for the finer granule of the two kernel variants
  TcWorkspace w;
  w.xi = cv.take<__half>(n * 64);
  w.xl[0] = cv.take<__half>(n * 64);
  w.xl[1] = cv.take<__half>(n * 64);
  w.pinit = cv.take<float>(n * EH);
  w.prow = cv.take<float>(n * EH);
  w.ei = cv.take<uint4>(e * 4);
  w.es = cv.take<uint4>(e * 4);
  w.flow = cv.take<float>(n * 2 * DN);
  w.part = cv.take<float>(chunks * 2 * DN);
  w.wimg_out = cv.take<uint8_t>(IMG_BYTES);
  w.wimg_in = cv.take<uint8_t>(IMG_BYTES);
  if (out) *out = w;
  return cv.off;
}

}  // namespace tc
}  // namespace mpn

using namespace mpn;

extern "C" {

/* development hook (not in the public header): device buffer of 64*16 int64 for a cycle trace */
void mpn_tc_set_trace(long long* d_buf) { tc::g_trace = d_buf; }

int64_t mpn_mp_tc_workspace(int64_t n, int64_t e) {
  return tc::carve(nullptr, n > 0 ? n : 1, e > 0 ? e : 1, nullptr) + 256;
}

int mpn_mp_forward_tc(const mpn_core_weights* w, const mpn_edge_layout* g, const float* x_init, const float* e_init,
                      int32_t num_steps, int32_t first_class_step, void* ws, float* logits, float* x_out,
                      float* e_out, int32_t* status, void* stream) {
  MPN_CHECK_ARG(w && g, "mp_forward_tc: null descriptor");
  MPN_CHECK_ARG(w->dn == 32 && w->de == 16 && w->edge_h == 80 && w->flow_h == 56 && w->cls_h == 8,
                "mp_forward_tc: built for widths dn=32 de=16 edge_h=80 flow_h=56 cls_h=8 (got %d %d %d %d %d)", w->dn,
                w->de, w->edge_h, w->flow_h, w->cls_h);
  MPN_CHECK_ARG(num_steps >= 1, "mp_forward_tc: num_steps must be >= 1 (use mpn_mp_forward for 0)");
  MPN_CHECK_ARG(ws && status, "mp_forward_tc: null workspace / status");
  const int64_t n = g->num_nodes, e = g->num_edges;
  cudaStream_t s = as_stream(stream);
  MPN_CUDA(cudaMemsetAsync(status, 0, 4, s));
  if (n == 0) return MPN_OK;
  tc::TcWorkspace m;
  tc::carve(ws, n, e > 0 ? e : 1, &m);

  static bool attr_set = false;
  if (!attr_set) {
    MPN_CUDA(cudaFuncSetAttribute(tc::mp_edge_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, tc::SMEM_BYTES));
    MPN_CUDA(cudaFuncSetAttribute(tc::mp_edge_tc3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, tc::SMEM3_BYTES));
    attr_set = true;
  }
  const int sms = sm_count();
  tc::pack_weights_kernel<<<16, 256, 0, s>>>(*w, m.wimg_out, m.wimg_in, tc::variant() == 3 ? 1 : 0); count_launch();
  const unsigned ngrid = (unsigned)std::min<int64_t>(ceil_div(n * 32, 256), (int64_t)sms * 4);
  const unsigned ngrid_node = (unsigned)std::min<int64_t>(ceil_div(n * 32, 256), (int64_t)sms * 2);   // node weights in registers: 2 CTAs/SM
  tc::prep_nodes_kernel<<<ngrid, 256, 0, s>>>(x_init, n, w->edge_w0, w->edge_b0, m.xi, m.xl[0], m.pinit, m.prow, status);
  count_launch();
  if (e > 0) {
    const unsigned egrid = (unsigned)std::min<int64_t>(ceil_div(e, 256), (int64_t)sms * 8);
    tc::split_edges_kernel<<<egrid, 256, 0, s>>>(e_init, e, m.ei, status); count_launch();
  }
  MPN_LAUNCH_CHECK();

  const int tiles_out = (int)ceil_div(g->num_out, tc::TS);
  const int tiles_in = (int)ceil_div(e - g->num_out, tc::TS);
  const int pairs = (tiles_out + 1) / 2 + (tiles_in + 1) / 2;
  int grid = sms;
  if (grid > pairs) grid = pairs;
  if (tiles_out > 0 && tiles_in > 0 && grid < 2) grid = 2;
  const int triples = (tiles_out + 2) / 3 + (tiles_in + 2) / 3;
  int grid3 = sms;
  if (grid3 > triples) grid3 = triples;
  if (tiles_out > 0 && tiles_in > 0 && grid3 < 2) grid3 = 2;

  CUtensorMap tm_xi, tm_xl[2];
  if (tc::variant() == 3 && e > 0) {
    if (tc::make_row_map(&tm_xi, m.xi, n) || tc::make_row_map(&tm_xl[0], m.xl[0], n) || tc::make_row_map(&tm_xl[1], m.xl[1], n)) {
      set_error("mpn_mp_forward_tc: cuTensorMapEncodeTiled failed");
      return MPN_ECUDA;
    }
  }
  const int chunk_shift = tc::variant() == 3 ? tc::CHUNK3_SHIFT : 5;
  const int64_t chunk = (int64_t)1 << chunk_shift;
  for (int step = 1; step <= num_steps; ++step) {
    const __half* xl_cur = m.xl[(step - 1) & 1];
    __half* xl_next = m.xl[step & 1];
    if (e > 0) {
      tc::TcArgs a;
      a.slot_row = g->slot_row; a.slot_col = g->slot_col; a.slot_edge = g->slot_edge;
      a.num_edges = e; a.num_out = g->num_out; a.tiles_out = tiles_out; a.tiles_in = tiles_in;
      a.chunks_out = ceil_div(g->num_out, chunk);
      a.xi = reinterpret_cast<const uint4*>(m.xi);
      a.xl = reinterpret_cast<const uint4*>(xl_cur);
      a.prow = m.prow;
      a.ei = m.ei;
      a.es_in = step == 1 ? m.ei : m.es;
      a.es_out = m.es;
      a.flow = m.flow; a.part = m.part;
      a.logits = (logits && step >= first_class_step) ? logits + (int64_t)(step - first_class_step) * e : nullptr;
      a.wimg_out = m.wimg_out; a.wimg_in = m.wimg_in;
      a.status = status;
      a.trace = (step == 2) ? tc::g_trace : nullptr;
      if (profiling()) profile_mark(0, true, s);
      if (tc::variant() == 3) tc::mp_edge_tc3_kernel<<<grid3, tc::NTHREADS3, tc::SMEM3_BYTES, s>>>(a, tm_xi, tm_xl[(step - 1) & 1]);
      else tc::mp_edge_tc_kernel<<<grid, tc::NTHREADS, tc::SMEM_BYTES, s>>>(a);
      count_launch();
      if (profiling()) profile_mark(0, false, s);
    }
    if (profiling()) profile_mark(1, true, s);
    tc::node_tc_kernel<<<ngrid_node, 256, 0, s>>>(g->out_ptr, g->in_ptr, n, g->num_out,
                                             (int32_t)ceil_div(g->num_out, chunk), chunk_shift, m.flow, m.part, w->node_w,
                                             w->node_b, w->edge_w0, m.pinit, xl_next, m.prow,
                                             step == num_steps ? x_out : nullptr, status);
    count_launch();
    if (profiling()) profile_mark(1, false, s);
    MPN_LAUNCH_CHECK();
  }
  if (e_out && e > 0) {
    const unsigned egrid = (unsigned)std::min<int64_t>(ceil_div(e, 256), (int64_t)sms * 8);
    tc::unsplit_edges_kernel<<<egrid, 256, 0, s>>>(m.es, e, e_out); count_launch();
    MPN_LAUNCH_CHECK();
  }
  return MPN_OK;
}

}  // extern "C"
