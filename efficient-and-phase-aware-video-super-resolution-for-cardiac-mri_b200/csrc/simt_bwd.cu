// Memory-bound kernels of the RefineNet BACKWARD path and of the optimiser (no tensor-core work).
//   lstm_bwd_pointwise  : adjoint of the ConvLSTM gate math (refine_net.py:258-265) -> pre-activation gate gradients
//   l1_multistage       : trainer loss  sum_k w_k * mean|out_k - target|  and its gradient
//                         (src/runner/trainers/acdc_vsr_refinenet_trainer.py:83-100 with nn.L1Loss)
//   head_last_bwd_data  : adjoint of the 64 -> 1 head conv (refine_net.py:203/205) wrt its input
//   head_last_bwd_weight: its weight / bias gradient
//   in_conv_prelu_bwd   : weight / bias / PReLU-slope gradient of _InBlock (refine_net.py:188-192)
//   posterm_bwd         : gradient of the positional-code input channels of _RefineBlock conv1 (:168-172)
//   cast / accumulate   : fp32 -> bf16 operand copies of accumulated gradients
//   adam                : torch.optim.Adam.step semantics on a flat parameter buffer
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>

#include "conv.h"
#include "ptx.cuh"
#include "simt.h"
#include "stencil.cuh"

namespace pvsr {

namespace {
__device__ __forceinline__ float bf16lo(uint32_t u) { return __uint_as_float(u << 16); }
__device__ __forceinline__ float bf16hi(uint32_t u) { return __uint_as_float(u & 0xFFFF0000u); }
__device__ __forceinline__ uint32_t pack2(float a, float b) {
  __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
  return v;
}
}  // namespace

// ------------------------------------------------------------------------------------------------
// lstm_bwd_pointwise.  One block per HALF 128-pixel tile of one cell; thread = (4 consecutive tile rows, 4 channels):
// warp = 4 row quads x 8 channel groups, so for a fixed channel the 4 row-quad lanes read 64 (fp32) / 32 (bf16)
// contiguous bytes of the tile-transposed tensors ([tile][ch][128]) = whole 32-byte sectors, and the 8 channel-group
// lanes of a pixel cover a whole 128-byte line of the NHWC dh / 64 bytes per gate of dgates.
// The first form (4 rows x 8 channels per thread, 128 registers, 2 blocks of 256 threads per SM = 25 % occupancy; ncu:
// 69 us per 6-cell launch, 22 % warps active, 33 % of DRAM peak, sm throughput 14 %) was latency-bound: halving the
// per-thread tile cuts the registers to 80 (3 blocks per SM) and issues all 32 loads of a thread up front.  tanh(c_t) uses the
// same MUFU form as the forward epilogue (conv3x3_tc.cu: tanh_from_scaled).
__device__ __forceinline__ float tanh_mufu(float x) {
  float e, r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(x * 2.885390081777927f));   // 2^(2 log2(e) x)
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(1.f + e));
  return fmaf(-2.f, r, 1.f);
}

__global__ void __launch_bounds__(256, 3) lstm_bwd_pointwise_kernel(const __grid_constant__ LstmBwdParams p) {
  // PDL: this kernel alternates with the tcgen05 data-gradient launches of the reverse wavefront
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  asm volatile("griddepcontrol.wait;" ::: "memory");
  const LstmBwdProb& pr = p.prob[blockIdx.y];
  const int tile = blockIdx.x >> 1;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int cg = (warp & 1) * 8 + (lane & 7);                       // channels [4*cg, 4*cg + 4)
  const int rq = (blockIdx.x & 1) * 16 + (warp >> 1) * 4 + (lane >> 3);   // tile rows [4*rq, 4*rq + 4)
  size_t pix[4];
  bool valid[4];
  if (p.wp > 0) {
    // padded raster: row r of tile t of an image is position 128 t + r = y * wp + x
    const int img = tile / p.tiles_per_img;
    const int pos0 = (tile - img * p.tiles_per_img) * 128 + rq * 4;
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const int y = (pos0 + r) / p.wp, x = (pos0 + r) - y * p.wp;
      valid[r] = y < p.H && x < p.W;
      pix[r] = (static_cast<size_t>(img) * p.H + y) * p.W + x;
    }
  } else {
    int t = tile;
    const int tx = t % p.tiles_x;
    t /= p.tiles_x;
    const int ty = t % p.tiles_y;
    const int img = t / p.tiles_y;
    const int TW = 1 << p.tw_log2;
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const int row = rq * 4 + r;
      const int y = ty * (128 >> p.tw_log2) + (row >> p.tw_log2), x = (tx << p.tw_log2) + (row & (TW - 1));
      valid[r] = y < p.H && x < p.W;
      pix[r] = (static_cast<size_t>(img) * p.H + y) * p.W + x;
    }
  }
  const __nv_bfloat16* gt = static_cast<const __nv_bfloat16*>(pr.gates) + static_cast<size_t>(tile) * 256 * 128 + rq * 4;
  const float* ct = pr.c + static_cast<size_t>(tile) * 64 * 128 + rq * 4;
  const float* cpt = pr.c_prev ? pr.c_prev + static_cast<size_t>(tile) * 64 * 128 + rq * 4 : nullptr;
  float* dct = pr.dc + static_cast<size_t>(tile) * 64 * 128 + rq * 4;

  // every load of the thread is issued before the first use: 4 (dh) + 4 x 7 independent requests in flight
  float4 dh4[4];
#pragma unroll
  for (int r = 0; r < 4; ++r)
    dh4[r] = valid[r] ? *reinterpret_cast<const float4*>(pr.dh + pix[r] * 64 + cg * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
  uint2 ui[4], uf[4], uo[4], ug[4];
  float4 cn4[4], cp4[4], dc4[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int ch = cg * 4 + j;
    ui[j] = *reinterpret_cast<const uint2*>(gt + ch * 128);
    uf[j] = *reinterpret_cast<const uint2*>(gt + (64 + ch) * 128);
    uo[j] = *reinterpret_cast<const uint2*>(gt + (128 + ch) * 128);
    ug[j] = *reinterpret_cast<const uint2*>(gt + (192 + ch) * 128);
    cn4[j] = *reinterpret_cast<const float4*>(ct + ch * 128);
    cp4[j] = cpt ? *reinterpret_cast<const float4*>(cpt + ch * 128) : make_float4(0.f, 0.f, 0.f, 0.f);
    dc4[j] = pr.dc_zero ? make_float4(0.f, 0.f, 0.f, 0.f) : *reinterpret_cast<const float4*>(dct + ch * 128);
  }
  const float dh[4][4] = {{dh4[0].x, dh4[0].y, dh4[0].z, dh4[0].w}, {dh4[1].x, dh4[1].y, dh4[1].z, dh4[1].w},
                          {dh4[2].x, dh4[2].y, dh4[2].z, dh4[2].w}, {dh4[3].x, dh4[3].y, dh4[3].z, dh4[3].w}};
  uint32_t oi[4][2], of[4][2], oo[4][2], og[4][2];   // [row][channel pair] packed bf16 (packed as soon as a pair is done)
#pragma unroll
  for (int jp = 0; jp < 2; ++jp) {
    float ai[4][2], af[4][2], ao[4][2], ag[4][2];
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const int j = jp * 2 + u;
      const float gi[4] = {bf16lo(ui[j].x), bf16hi(ui[j].x), bf16lo(ui[j].y), bf16hi(ui[j].y)};
      const float gf[4] = {bf16lo(uf[j].x), bf16hi(uf[j].x), bf16lo(uf[j].y), bf16hi(uf[j].y)};
      const float go[4] = {bf16lo(uo[j].x), bf16hi(uo[j].x), bf16lo(uo[j].y), bf16hi(uo[j].y)};
      const float gg[4] = {bf16lo(ug[j].x), bf16hi(ug[j].x), bf16lo(ug[j].y), bf16hi(ug[j].y)};
      const float cn[4] = {cn4[j].x, cn4[j].y, cn4[j].z, cn4[j].w};
      const float cp[4] = {cp4[j].x, cp4[j].y, cp4[j].z, cp4[j].w};
      const float dcin[4] = {dc4[j].x, dc4[j].y, dc4[j].z, dc4[j].w};
      float dco[4];
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        const float tc = tanh_mufu(cn[r]);
        const float dcv = fmaf(dh[r][j] * go[r], 1.f - tc * tc, dcin[r]);
        ao[r][u] = dh[r][j] * tc * go[r] * (1.f - go[r]);
        ai[r][u] = dcv * gg[r] * gi[r] * (1.f - gi[r]);
        af[r][u] = dcv * cp[r] * gf[r] * (1.f - gf[r]);
        ag[r][u] = dcv * gi[r] * (1.f - gg[r] * gg[r]);
        dco[r] = dcv * gf[r];
      }
      *reinterpret_cast<float4*>(dct + (cg * 4 + j) * 128) = make_float4(dco[0], dco[1], dco[2], dco[3]);
    }
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      oi[r][jp] = pack2(ai[r][0], ai[r][1]); of[r][jp] = pack2(af[r][0], af[r][1]);
      oo[r][jp] = pack2(ao[r][0], ao[r][1]); og[r][jp] = pack2(ag[r][0], ag[r][1]);
    }
  }
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    if (!valid[r]) continue;
    __nv_bfloat16* dst = static_cast<__nv_bfloat16*>(pr.dgates) + pix[r] * 256 + cg * 4;
    *reinterpret_cast<uint2*>(dst) = make_uint2(oi[r][0], oi[r][1]);
    *reinterpret_cast<uint2*>(dst + 64) = make_uint2(of[r][0], of[r][1]);
    *reinterpret_cast<uint2*>(dst + 128) = make_uint2(oo[r][0], oo[r][1]);
    *reinterpret_cast<uint2*>(dst + 192) = make_uint2(og[r][0], og[r][1]);
  }
}

// Second form (round 2): the first one reads the tile-transposed tensors with 8 different 128-byte lines per warp
// request (4 row quads x 8 channel groups), and ncu put it at 75 % L1 throughput / 47 % DRAM / 3.9 TB/s.  Here a lane
// IS a tile row (pixel), exactly like the tcgen05 epilogue that wrote those tensors: every load / store of gates, c,
// c_prev and dc is one fully used 128-byte (fp32) or 64-byte (bf16) line per warp request.  The two NHWC tensors are
// transposed through shared memory instead: dh (128 rows x 32 channels of this block, fp32, pitch 36 floats: the
// per-row float4 reads are conflict-free) is staged with 128-byte row reads; the bf16 gate gradients are collected in
// [row][4 gates x 32 channels] (pitch 132: the 8-byte per-lane writes are conflict-free) and leave as 64-byte pieces.
// Block = one tile x 32 channels (256 threads: 4 pixel groups x 2 channel groups of 16); same arithmetic per element
// as the first form (bit-identical results).
constexpr int kLb2DhPitch = 36;     // floats
constexpr int kLb2DgPitch = 132;    // bf16
constexpr int kLb2Smem = 128 * kLb2DhPitch * 4 + 128 * kLb2DgPitch * 2 + 128 * 8;
__device__ __forceinline__ float bf16_scalar(const unsigned short* p) { return __uint_as_float(static_cast<uint32_t>(*p) << 16); }

__global__ void __launch_bounds__(256, 4) lstm_bwd_pointwise_v2_kernel(const __grid_constant__ LstmBwdParams p) {
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  asm volatile("griddepcontrol.wait;" ::: "memory");
  extern __shared__ __align__(16) unsigned char lb2_smem[];
  float* s_dh = reinterpret_cast<float*>(lb2_smem);
  __nv_bfloat16* s_dg = reinterpret_cast<__nv_bfloat16*>(lb2_smem + 128 * kLb2DhPitch * 4);
  long long* s_pix = reinterpret_cast<long long*>(lb2_smem + 128 * kLb2DhPitch * 4 + 128 * kLb2DgPitch * 2);
  const LstmBwdProb& pr = p.prob[blockIdx.y];
  const int tile = blockIdx.x >> 1;
  const int ch_blk = (blockIdx.x & 1) * 32;                 // this block's 32 channels
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid < 128) {                                          // NHWC pixel index of every tile row (-1: padding)
    int y, x, img;
    if (p.wp > 0) {
      img = tile / p.tiles_per_img;
      const int pos = (tile - img * p.tiles_per_img) * 128 + tid;
      y = pos / p.wp;
      x = pos - y * p.wp;
    } else {
      int t = tile;
      const int tx = t % p.tiles_x;
      t /= p.tiles_x;
      const int ty = t % p.tiles_y;
      img = t / p.tiles_y;
      const int TW = 1 << p.tw_log2;
      y = ty * (128 >> p.tw_log2) + (tid >> p.tw_log2);
      x = (tx << p.tw_log2) + (tid & (TW - 1));
    }
    s_pix[tid] = (y < p.H && x < p.W) ? (static_cast<long long>(img) * p.H + y) * p.W + x : -1;
  }
  __syncthreads();
  // dh -> shared memory: a row's 32 channels = 128 contiguous bytes, 8 lanes x 16 bytes
#pragma unroll
  for (int pass = 0; pass < 4; ++pass) {
    const int row = pass * 32 + (tid >> 3), sub = tid & 7;
    const long long px = s_pix[row];
    const float4 v = px >= 0 ? *reinterpret_cast<const float4*>(pr.dh + px * 64 + ch_blk + sub * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
    *reinterpret_cast<float4*>(s_dh + row * kLb2DhPitch + sub * 4) = v;
  }
  const int row = (warp & 3) * 32 + lane;                   // this thread's tile row
  const int ch_thr = (warp >> 2) * 16;                      // its 16 channels inside the block's 32
  const unsigned short* gt = static_cast<const unsigned short*>(pr.gates) + static_cast<size_t>(tile) * 256 * 128 + row;
  const float* ct = pr.c + static_cast<size_t>(tile) * 64 * 128 + row;
  const float* cpt = pr.c_prev ? pr.c_prev + static_cast<size_t>(tile) * 64 * 128 + row : nullptr;
  float* dct = pr.dc + static_cast<size_t>(tile) * 64 * 128 + row;
  __syncthreads();
#pragma unroll 1
  for (int k = 0; k < 4; ++k) {
    const int lc = ch_thr + 4 * k;                          // channel inside the block
    const int ch = ch_blk + lc;
    float gi[4], gf[4], go[4], gg[4], cn[4], cp[4], dcin[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {                           // 28 independent requests in flight
      gi[j] = bf16_scalar(gt + (ch + j) * 128);
      gf[j] = bf16_scalar(gt + (64 + ch + j) * 128);
      go[j] = bf16_scalar(gt + (128 + ch + j) * 128);
      gg[j] = bf16_scalar(gt + (192 + ch + j) * 128);
      cn[j] = ct[(ch + j) * 128];
      cp[j] = cpt ? cpt[(ch + j) * 128] : 0.f;
      dcin[j] = pr.dc_zero ? 0.f : dct[(ch + j) * 128];
    }
    const float4 d4 = *reinterpret_cast<const float4*>(s_dh + row * kLb2DhPitch + lc);
    const float dh[4] = {d4.x, d4.y, d4.z, d4.w};
    float ai[4], af[4], ao[4], ag[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float tc = tanh_mufu(cn[j]);
      const float dcv = fmaf(dh[j] * go[j], 1.f - tc * tc, dcin[j]);
      ao[j] = dh[j] * tc * go[j] * (1.f - go[j]);
      ai[j] = dcv * gg[j] * gi[j] * (1.f - gi[j]);
      af[j] = dcv * cp[j] * gf[j] * (1.f - gf[j]);
      ag[j] = dcv * gi[j] * (1.f - gg[j] * gg[j]);
      dct[(ch + j) * 128] = dcv * gf[j];
    }
    __nv_bfloat16* dst = s_dg + row * kLb2DgPitch + lc;
    *reinterpret_cast<uint2*>(dst) = make_uint2(pack2(ai[0], ai[1]), pack2(ai[2], ai[3]));
    *reinterpret_cast<uint2*>(dst + 32) = make_uint2(pack2(af[0], af[1]), pack2(af[2], af[3]));
    *reinterpret_cast<uint2*>(dst + 64) = make_uint2(pack2(ao[0], ao[1]), pack2(ao[2], ao[3]));
    *reinterpret_cast<uint2*>(dst + 96) = make_uint2(pack2(ag[0], ag[1]), pack2(ag[2], ag[3]));
  }
  __syncthreads();
  // gate gradients -> NHWC [pixel][256]: one warp per row per pass, lane = (gate, 8-byte piece of its 64 bytes)
#pragma unroll 4
  for (int pass = 0; pass < 16; ++pass) {
    const int r = pass * 8 + warp;
    const long long px = s_pix[r];
    if (px < 0) continue;
    const uint2 v = *reinterpret_cast<const uint2*>(s_dg + r * kLb2DgPitch + lane * 4);
    __nv_bfloat16* dst = static_cast<__nv_bfloat16*>(pr.dgates) + px * 256 + (lane >> 3) * 64 + ch_blk + (lane & 7) * 4;
    *reinterpret_cast<uint2*>(dst) = v;
  }
}

// A/B switch (PVSR_LSTM_BWD_V1=1 selects the first form)
static int lstm_bwd_form() {
  static int form = -1;
  if (form < 0) {
    const char* e = getenv("PVSR_LSTM_BWD_V1");
    form = (e && e[0] == '1') ? 1 : 2;
  }
  return form;
}

int launch_lstm_bwd_pointwise(const LstmBwdParams& p, cudaStream_t s) {
  if (p.n_prob <= 0 || p.n_img <= 0) return 0;
  const int tiles = p.wp > 0 ? p.tiles_per_img : p.tiles_x * p.tiles_y;
  dim3 grid(static_cast<unsigned>(2 * p.n_img * tiles), static_cast<unsigned>(p.n_prob));   // two blocks per tile
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid;
  cfg.blockDim = dim3(256);
  cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = get_pdl() ? 1 : 0;
  if (lstm_bwd_form() == 2) {
    static bool attr_set = false;
    if (!attr_set) {
      cudaError_t e = cudaFuncSetAttribute(lstm_bwd_pointwise_v2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kLb2Smem);
      if (e != cudaSuccess) return static_cast<int>(e);
      attr_set = true;
    }
    cfg.dynamicSmemBytes = kLb2Smem;
    return static_cast<int>(cudaLaunchKernelEx(&cfg, lstm_bwd_pointwise_v2_kernel, p));
  }
  return static_cast<int>(cudaLaunchKernelEx(&cfg, lstm_bwd_pointwise_kernel, p));
}

// ------------------------------------------------------------------------------------------------
// l1_multistage: out [n_lists][n_per_list], target [n_per_list]; loss += sum_k w[k] * sum|out_k - target|,
// dout_k = w[k] * sign(out_k - target).  w[k] already holds discount / (T * N * H * W).
__global__ void __launch_bounds__(256) l1_multistage_kernel(const float4* __restrict__ out,
                                                            const float4* __restrict__ target,
                                                            const float* __restrict__ w, int n_lists, long long n4,
                                                            float* __restrict__ loss, float4* __restrict__ dout) {
  float acc = 0.f;
  const long long total = n4 * n_lists;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int k = static_cast<int>(i / n4);
    const long long e = i - k * n4;
    const float wk = w[k];
    const float4 o = out[i], t = __ldg(target + e);
    const float d0 = o.x - t.x, d1 = o.y - t.y, d2 = o.z - t.z, d3 = o.w - t.w;
    acc += wk * (fabsf(d0) + fabsf(d1) + fabsf(d2) + fabsf(d3));
    if (dout) {
      float4 g;
      g.x = d0 > 0.f ? wk : (d0 < 0.f ? -wk : 0.f);
      g.y = d1 > 0.f ? wk : (d1 < 0.f ? -wk : 0.f);
      g.z = d2 > 0.f ? wk : (d2 < 0.f ? -wk : 0.f);
      g.w = d3 > 0.f ? wk : (d3 < 0.f ? -wk : 0.f);
      dout[i] = g;
    }
  }
  __shared__ float part[8];
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x < 8) {
    float v = part[threadIdx.x];
#pragma unroll
    for (int d = 4; d > 0; d >>= 1) v += __shfl_xor_sync(0xffu, v, d);
    if (threadIdx.x == 0) atomicAdd(loss, v);
  }
}

// Element-wise form for list lengths that are not a multiple of 4 (x3 with odd LR sizes and odd T * N: the lists of the
// output tensor then start at addresses that are not 16-byte aligned).
__global__ void __launch_bounds__(256) l1_multistage_scalar_kernel(const float* __restrict__ out,
                                                                   const float* __restrict__ target,
                                                                   const float* __restrict__ w, int n_lists, long long n,
                                                                   float* __restrict__ loss, float* __restrict__ dout) {
  float acc = 0.f;
  const long long total = n * n_lists;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int k = static_cast<int>(i / n);
    const float wk = w[k];
    const float d = out[i] - __ldg(target + (i - k * n));
    acc += wk * fabsf(d);
    if (dout) dout[i] = d > 0.f ? wk : (d < 0.f ? -wk : 0.f);
  }
  __shared__ float part[8];
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x < 8) {
    float v = part[threadIdx.x];
#pragma unroll
    for (int d = 4; d > 0; d >>= 1) v += __shfl_xor_sync(0xffu, v, d);
    if (threadIdx.x == 0) atomicAdd(loss, v);
  }
}

int launch_l1_multistage(const float* out, const float* target, const float* w, int n_lists, long long n_per_list,
                         float* loss, float* dout, int num_sms, cudaStream_t s) {
  if (n_per_list <= 0 || n_lists == 0) return 0;
  const long long cap = 8LL * (num_sms > 0 ? num_sms : 148);
  if (n_per_list % 4 != 0) {
    long long blocks = (n_per_list * n_lists + 255) / 256;
    if (blocks > cap) blocks = cap;
    l1_multistage_scalar_kernel<<<static_cast<unsigned>(blocks), 256, 0, s>>>(out, target, w, n_lists, n_per_list, loss, dout);
    return static_cast<int>(cudaGetLastError());
  }
  const long long n4 = n_per_list / 4;
  long long blocks = (n4 * n_lists + 255) / 256;
  if (blocks > cap) blocks = cap;
  l1_multistage_kernel<<<static_cast<unsigned>(blocks), 256, 0, s>>>(
      reinterpret_cast<const float4*>(out), reinterpret_cast<const float4*>(target), w, n_lists, n4, loss,
      reinterpret_cast<float4*>(dout));
  return static_cast<int>(cudaGetLastError());
}

// ------------------------------------------------------------------------------------------------
// head_last_bwd_data: dIn[y, x, c] = sum_tap dOut[y - dy, x - dx] * w[c, tap]   (zero outside the image)
// 8 threads per pixel run (8 channels each; the 72 weights of a thread live in registers); see stencil.cuh.
__global__ void __launch_bounds__(kStencilThreads, 3) head_last_bwd_data_kernel(const float* __restrict__ dout,
                                                                    const float* __restrict__ w,
                                                                    __nv_bfloat16* __restrict__ din, unsigned n_rows,
                                                                    int H, int W) {
  const int cg = (threadIdx.x & 7) * 8;
  float wr[8][9], br[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    br[j] = 0.f;
#pragma unroll
    for (int t = 0; t < 9; ++t) wr[j][t] = __ldg(w + (cg + j) * 9 + t);   // parameter layout (1, 64, 3, 3)
  }
  stencil_1to64_runs<true, false>(dout, wr, br, 0.f, din, n_rows, H, W, cg);
}

int launch_head_last_bwd_data(const float* dout, const float* w, void* din_bf16, long long n_img, int H, int W,
                              cudaStream_t s) {
  const long long n_rows = n_img * H;
  if (n_rows == 0) return 0;
  if (n_rows * ((W + 7) / 8) >= (1LL << 31)) return static_cast<int>(cudaErrorInvalidValue);
  head_last_bwd_data_kernel<<<stencil_blocks(n_rows, W, 148 * 24), kStencilThreads, 0, s>>>(
      dout, w, static_cast<__nv_bfloat16*>(din_bf16), static_cast<unsigned>(n_rows), H, W);
  return static_cast<int>(cudaGetLastError());
}

// ------------------------------------------------------------------------------------------------
// head_last_bwd_weight: dW[c, tap] += sum_q in[q, c] * dOut[q - off(tap)],  db += sum dOut.
// A [16 taps x pixels] x [pixels x 64 channels] product on warp-level mma.sync (K = pixels).  Persistent blocks walk
// 8x32-pixel tiles: the `in` tile is staged with cp.async (144-byte pixel pitch), the dOut halo (10x34 fp32) with
// plain loads; warp w owns tile row w (two k16 steps).  The A operand D[tap][q] = dOut[q - off(tap)] is built in
// registers from the halo and split into bf16 hi + lo (the loss gradient keeps ~16 mantissa bits); the B operand
// in[q][c] comes from the tile via ldmatrix.trans.  The 16x64 fp32 accumulator (32 registers per thread) lives across
// all tiles of the block; one shared-memory and one global atomic pass at the end.
constexpr int kWTH = 8, kWTW = 32, kWPitch = 144, kWHaloW = 36;

__global__ void __launch_bounds__(256, 2) head_last_bwd_weight_kernel(const __nv_bfloat16* __restrict__ in,
                                                                      const float* __restrict__ dout,
                                                                      float* __restrict__ dw, float* __restrict__ db,
                                                                      int n_img, int H, int W, int tiles_x,
                                                                      int tiles_y) {
  __shared__ __align__(16) uint8_t tile[kWTH * kWTW * kWPitch];
  __shared__ float halo[(kWTH + 2) * kWHaloW];
  __shared__ float sacc[577];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint32_t tile_s = static_cast<uint32_t>(__cvta_generic_to_shared(tile));
  float acc[8][4];
#pragma unroll
  for (int nt = 0; nt < 8; ++nt)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[nt][j] = 0.f;
  float bsum = 0.f;
  const int total = n_img * tiles_y * tiles_x;
  for (int t = blockIdx.x; t < total; t += gridDim.x) {
    int tt = t;
    const int tx = tt % tiles_x;
    tt /= tiles_x;
    const int ty = tt % tiles_y;
    const int img = tt / tiles_y;
    const int y0 = ty * kWTH, x0 = tx * kWTW;
    const __nv_bfloat16* src = in + static_cast<size_t>(img) * H * W * 64;
    const float* dsrc = dout + static_cast<size_t>(img) * H * W;
#pragma unroll
    for (int it = 0; it < kWTH * kWTW * 8 / 256; ++it) {
      const int i = threadIdx.x + it * 256;
      const int hp = i >> 3, ck = i & 7;
      const int y = y0 + (hp >> 5), x = x0 + (hp & 31);
      const bool ok = y < H && x < W;
      const __nv_bfloat16* g = src + (static_cast<size_t>(ok ? y : 0) * W + (ok ? x : 0)) * 64 + ck * 8;
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(tile_s + hp * kWPitch + ck * 16), "l"(g),
                   "r"(ok ? 16 : 0)
                   : "memory");
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
    for (int i = threadIdx.x; i < (kWTH + 2) * (kWTW + 2); i += 256) {
      const int hy = i / (kWTW + 2), hx = i - hy * (kWTW + 2);
      const int y = y0 + hy - 1, x = x0 + hx - 1;
      const float v = (y >= 0 && y < H && x >= 0 && x < W) ? __ldg(dsrc + static_cast<size_t>(y) * W + x) : 0.f;
      halo[hy * kWHaloW + hx] = v;
      if (hy >= 1 && hy <= kWTH && hx >= 1 && hx <= kWTW) bsum += v;
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();

    const int ly = warp;
#pragma unroll
    for (int ks = 0; ks < 2; ++ks) {
      const int lx0 = ks * 16;
      // A fragments: rows = taps, columns = pixels of this k step
      uint32_t ahi[4], alo[4];
#pragma unroll
      for (int f = 0; f < 4; ++f) {
        const int tap = (lane >> 2) + (f & 1) * 8;
        const int k = (lane & 3) * 2 + (f >> 1) * 8;
        float v0 = 0.f, v1 = 0.f;
        if (tap < 9) {
          const int dy = tap / 3 - 1, dx = tap % 3 - 1;
          const float* hrow = halo + (ly - dy + 1) * kWHaloW + (lx0 + k - dx + 1);
          v0 = hrow[0];
          v1 = hrow[1];
        }
        const uint32_t hi = pack2(v0, v1);
        ahi[f] = hi;
        alo[f] = pack2(v0 - bf16lo(hi), v1 - bf16hi(hi));
      }
#pragma unroll
      for (int np = 0; np < 4; ++np) {
        uint32_t b0, b1, b2, b3;
        const uint32_t addr = tile_s + (ly * kWTW + lx0 + (lane & 7) + 8 * ((lane >> 3) & 1)) * kWPitch +
                              (16 * np + 8 * (lane >> 4)) * 2;
        asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];"
                     : "=r"(b0), "=r"(b1), "=r"(b2), "=r"(b3)
                     : "r"(addr));
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const uint32_t bb0 = h ? b2 : b0, bb1 = h ? b3 : b1;
          float* c = acc[2 * np + h];
          asm volatile(
              "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, "
              "{%0, %1, %2, %3};"
              : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
              : "r"(ahi[0]), "r"(ahi[1]), "r"(ahi[2]), "r"(ahi[3]), "r"(bb0), "r"(bb1));
          asm volatile(
              "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, "
              "{%0, %1, %2, %3};"
              : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
              : "r"(alo[0]), "r"(alo[1]), "r"(alo[2]), "r"(alo[3]), "r"(bb0), "r"(bb1));
        }
      }
    }
    __syncthreads();   // the tile and the halo are overwritten by the next iteration
  }
  for (int i = threadIdx.x; i < 577; i += 256) sacc[i] = 0.f;
  __syncthreads();
#pragma unroll
  for (int nt = 0; nt < 8; ++nt) {
    const int ch = nt * 8 + (lane & 3) * 2, tap = lane >> 2;
    atomicAdd(&sacc[ch * 9 + tap], acc[nt][0]);
    atomicAdd(&sacc[(ch + 1) * 9 + tap], acc[nt][1]);
    if (tap == 0) {
      atomicAdd(&sacc[ch * 9 + 8], acc[nt][2]);
      atomicAdd(&sacc[(ch + 1) * 9 + 8], acc[nt][3]);
    }
  }
  bsum = warp_sum(bsum);
  if (lane == 0) atomicAdd(&sacc[576], bsum);
  __syncthreads();
  for (int i = threadIdx.x; i < 576; i += 256) atomicAdd(dw + i, sacc[i]);
  if (threadIdx.x == 0) atomicAdd(db, sacc[576]);
}

// TMA form of head_last_bwd_weight: the `in` tile (64 ch x 32 x 8 px, 128B swizzle) arrives as one bulk tensor box
// through a 4-deep ring - three tiles (96 KB) in flight per SM where the cp.async form above loads, waits, multiplies,
// one tile at a time (same lesson as head_conv_last_tma_kernel).  The small fp32 dOut halo (34 x 10) of tile i + 1 is
// fetched into registers at the top of iteration i and parked in the other half of a double-buffered shared array at
// its end, so its latency hides behind the MMAs of tile i.  (A 3-D fp32 tensor map with a 36-element box faulted with
// "illegal instruction" at the cp.async.bulk.tensor - not pursued.)
constexpr int kWStages = 4;
constexpr int kWTileBytes = kWTH * kWTW * 128;                       // 32768
constexpr int kWHaloPitch = 1536;

__global__ void __launch_bounds__(256, 1) head_last_bwd_weight_tma_kernel(const __grid_constant__ CUtensorMap tm_in,
                                                                          const float* __restrict__ dout,
                                                                          float* __restrict__ dw, float* __restrict__ db,
                                                                          int H, int W, int tiles_x, int tiles_y,
                                                                          int total) {
  extern __shared__ uint8_t smw_raw[];
  uint8_t* sm = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smw_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* halos = sm + kWStages * kWTileBytes;
  uint64_t* full = reinterpret_cast<uint64_t*>(halos + kWStages * kWHaloPitch);
  float* sacc = reinterpret_cast<float*>(full + kWStages);           // [577]
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint32_t sm_s = ptx::smem_u32(sm);
  auto issue = [&](int t, int stage) {      // thread 0 only
    const int tx = t % tiles_x;
    const int q = t / tiles_x;
    const int ty = q % tiles_y, img = q / tiles_y;
    ptx::mbar_arrive_expect_tx(&full[stage], kWTileBytes);
    ptx::tma_load_4d(sm + stage * kWTileBytes, &tm_in, &full[stage], 0, tx * kWTW, ty * kWTH, img);
  };
  // this thread's (up to two) elements of the 34 x 10 dOut halo of tile t, zero outside the image
  auto halo_fetch = [&](long long t, float (&v)[2]) {
    const int tx = static_cast<int>(t % tiles_x);
    const long long q = t / tiles_x;
    const int ty = static_cast<int>(q % tiles_y);
    const float* dsrc = dout + (q / tiles_y) * H * W;
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      const int i = threadIdx.x + k * 256;
      const int hy = i / (kWTW + 2), hx = i - hy * (kWTW + 2);
      const int y = ty * kWTH + hy - 1, x = tx * kWTW + hx - 1;
      v[k] = (i < (kWTH + 2) * (kWTW + 2) && y >= 0 && y < H && x >= 0 && x < W)
                 ? __ldg(dsrc + static_cast<size_t>(y) * W + x) : 0.f;
    }
  };
  auto halo_store = [&](int buf, const float (&v)[2]) {
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      const int i = threadIdx.x + k * 256;
      if (i < (kWTH + 2) * (kWTW + 2)) {
        const int hy = i / (kWTW + 2), hx = i - hy * (kWTW + 2);
        reinterpret_cast<float*>(halos + buf * kWHaloPitch)[hy * kWHaloW + hx] = v[k];
      }
    }
  };
  if (threadIdx.x == 0) {
    ptx::prefetch_tmap(&tm_in);
    for (int i = 0; i < kWStages; ++i) ptx::mbar_init(&full[i], 1);
    ptx::fence_mbar_init();
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int i = 0; i < kWStages - 1; ++i) {
      const long long t = static_cast<long long>(blockIdx.x) + static_cast<long long>(i) * gridDim.x;
      if (t < total) issue(static_cast<int>(t), i);
    }
  }
  {
    float v[2];
    if (blockIdx.x < total) {
      halo_fetch(blockIdx.x, v);
      halo_store(0, v);
    }
  }
  __syncthreads();
  float acc[8][4];
#pragma unroll
  for (int nt = 0; nt < 8; ++nt)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[nt][j] = 0.f;
  float bsum = 0.f;
  int it = 0;
  for (long long tl = blockIdx.x; tl < total; tl += gridDim.x, ++it) {
    const int stage = it % kWStages;
    if (threadIdx.x == 0) {                 // slot (it + 3) % 4 held tile it - 1: released by the barrier below
      const long long tn = tl + static_cast<long long>(kWStages - 1) * gridDim.x;
      if (tn < total) issue(static_cast<int>(tn), (it + kWStages - 1) % kWStages);
    }
    float hnext[2];
    const bool has_next = tl + gridDim.x < total;
    if (has_next) halo_fetch(tl + gridDim.x, hnext);          // in flight during this tile's MMAs
    ptx::mbar_wait(&full[stage], (it / kWStages) & 1);
    const uint32_t tile_s = sm_s + stage * kWTileBytes;
    const float* halo = reinterpret_cast<const float*>(halos + (it & 1) * kWHaloPitch);
    bsum += halo[(1 + (threadIdx.x >> 5)) * kWHaloW + 1 + (threadIdx.x & 31)];      // one interior pixel per thread

    const int ly = warp;
#pragma unroll
    for (int ks = 0; ks < 2; ++ks) {
      const int lx0 = ks * 16;
      uint32_t ahi[4], alo[4];
#pragma unroll
      for (int f = 0; f < 4; ++f) {
        const int tap = (lane >> 2) + (f & 1) * 8;
        const int k = (lane & 3) * 2 + (f >> 1) * 8;
        float v0 = 0.f, v1 = 0.f;
        if (tap < 9) {
          const int dy = tap / 3 - 1, dx = tap % 3 - 1;
          const float* hrow = halo + (ly - dy + 1) * kWHaloW + (lx0 + k - dx + 1);
          v0 = hrow[0];
          v1 = hrow[1];
        }
        const uint32_t hi = pack2(v0, v1);
        ahi[f] = hi;
        alo[f] = pack2(v0 - bf16lo(hi), v1 - bf16hi(hi));
      }
      const int hp = ly * kWTW + lx0 + (lane & 7) + 8 * ((lane >> 3) & 1);
      const uint32_t arow = tile_s + hp * 128;
      const int sw = hp & 7;
#pragma unroll
      for (int np = 0; np < 4; ++np) {
        uint32_t b0, b1, b2, b3;
        asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];"
                     : "=r"(b0), "=r"(b1), "=r"(b2), "=r"(b3)
                     : "r"(arow + (((2 * np + (lane >> 4)) ^ sw) << 4)));
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const uint32_t bb0 = h ? b2 : b0, bb1 = h ? b3 : b1;
          float* c = acc[2 * np + h];
          asm volatile(
              "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, "
              "{%0, %1, %2, %3};"
              : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
              : "r"(ahi[0]), "r"(ahi[1]), "r"(ahi[2]), "r"(ahi[3]), "r"(bb0), "r"(bb1));
          asm volatile(
              "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, "
              "{%0, %1, %2, %3};"
              : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
              : "r"(alo[0]), "r"(alo[1]), "r"(alo[2]), "r"(alo[3]), "r"(bb0), "r"(bb1));
        }
      }
    }
    if (has_next) halo_store((it + 1) & 1, hnext);   // the other halo buffer: its readers finished an iteration ago
    __syncthreads();   // every read of this ring slot is done before thread 0 refills it next iteration
  }
  for (int i = threadIdx.x; i < 577; i += 256) sacc[i] = 0.f;
  __syncthreads();
#pragma unroll
  for (int nt = 0; nt < 8; ++nt) {
    const int ch = nt * 8 + (lane & 3) * 2, tap = lane >> 2;
    atomicAdd(&sacc[ch * 9 + tap], acc[nt][0]);
    atomicAdd(&sacc[(ch + 1) * 9 + tap], acc[nt][1]);
    if (tap == 0) {
      atomicAdd(&sacc[ch * 9 + 8], acc[nt][2]);
      atomicAdd(&sacc[(ch + 1) * 9 + 8], acc[nt][3]);
    }
  }
  bsum = warp_sum(bsum);
  if (lane == 0) atomicAdd(&sacc[576], bsum);
  __syncthreads();
  for (int i = threadIdx.x; i < 576; i += 256) atomicAdd(dw + i, sacc[i]);
  if (threadIdx.x == 0) atomicAdd(db, sacc[576]);
}

static int launch_head_last_bwd_weight_tma(const void* in_bf16, const float* dout, float* dw, float* db, long long n_img,
                                           int H, int W, int num_sms, cudaStream_t s) {
  const int tiles_x = (W + kWTW - 1) / kWTW, tiles_y = (H + kWTH - 1) / kWTH;
  const size_t smem = 1024 + kWStages * (kWTileBytes + kWHaloPitch) + 64 + 577 * sizeof(float) + 16;
  static bool attr = false;
  if (!attr) {
    cudaError_t e = cudaFuncSetAttribute(head_last_bwd_weight_tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         static_cast<int>(smem));
    if (e != cudaSuccess) return static_cast<int>(e);
    attr = true;
  }
  CUtensorMap tm_in;
  if (make_act_tmap(&tm_in, in_bf16, 64, W, H, n_img, kWTW, kWTH, 1) != 0) return static_cast<int>(cudaErrorInvalidValue);
  const long long total = n_img * tiles_x * tiles_y;
  if (total >= (1LL << 31)) return static_cast<int>(cudaErrorInvalidValue);
  const long long cap = num_sms > 0 ? num_sms : 148;
  head_last_bwd_weight_tma_kernel<<<static_cast<unsigned>(total < cap ? total : cap), 256, smem, s>>>(
      tm_in, dout, dw, db, H, W, tiles_x, tiles_y, static_cast<int>(total));
  return static_cast<int>(cudaGetLastError());
}

int launch_head_last_bwd_weight(const void* in_bf16, const float* dout, float* dw, float* db, long long n_img, int H,
                                int W, int num_sms, cudaStream_t s) {
  if (n_img * H * W == 0) return 0;
  if (get_head_tma())
    return launch_head_last_bwd_weight_tma(in_bf16, dout, dw, db, n_img, H, W, num_sms, s);
  const int tiles_x = (W + kWTW - 1) / kWTW, tiles_y = (H + kWTH - 1) / kWTH;
  long long blocks = n_img * tiles_x * tiles_y;
  const long long cap = 4LL * (num_sms > 0 ? num_sms : 148);
  if (blocks > cap) blocks = cap;
  head_last_bwd_weight_kernel<<<static_cast<unsigned>(blocks), 256, 0, s>>>(
      static_cast<const __nv_bfloat16*>(in_bf16), dout, dw, db, static_cast<int>(n_img), H, W, tiles_x, tiles_y);
  return static_cast<int>(cudaGetLastError());
}

// ------------------------------------------------------------------------------------------------
// in_conv_prelu_bwd: y = prelu(conv(x)),  g = dL/dy (fp32 NHWC 64 channels).  Recomputes the pre-activation.
//   dW[c, tap] += dpre * x[p + off(tap)],  db[c] += dpre,  dslope += g * min(pre, 0)    (torch prelu_backward rule)
__global__ void __launch_bounds__(256) in_conv_prelu_bwd_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                                const float* __restrict__ b,
                                                                const float* __restrict__ slope,
                                                                const float* __restrict__ g, float* __restrict__ dw,
                                                                float* __restrict__ db, float* __restrict__ dslope,
                                                                long long n_pix_total, int H, int W) {
  __shared__ float sw[9][64];
  __shared__ float sb[64];
  __shared__ float sacc[641];  // 576 dW + 64 db + 1 dslope
  for (int i = threadIdx.x; i < 576; i += blockDim.x) sw[i % 9][i / 9] = w[i];
  if (threadIdx.x < 64) sb[threadIdx.x] = b[threadIdx.x];
  for (int i = threadIdx.x; i < 641; i += blockDim.x) sacc[i] = 0.f;
  __syncthreads();
  const float a = slope[0];
  const int pl = threadIdx.x >> 3, cg = (threadIdx.x & 7) * 8;
  float acc[8][9], accb[8], acca = 0.f;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    accb[j] = 0.f;
#pragma unroll
    for (int t = 0; t < 9; ++t) acc[j][t] = 0.f;
  }
  for (long long q = static_cast<long long>(blockIdx.x) * 32 + pl; q < n_pix_total;
       q += static_cast<long long>(gridDim.x) * 32) {
    const int xw = static_cast<int>(q % W);
    const int yh = static_cast<int>((q / W) % H);
    const float* img = x + (q - static_cast<long long>(yh) * W - xw);
    float v[9];
#pragma unroll
    for (int t = 0; t < 9; ++t) {
      const int yy = yh + t / 3 - 1, xx = xw + t % 3 - 1;
      v[t] = (yy >= 0 && yy < H && xx >= 0 && xx < W) ? __ldg(img + static_cast<long long>(yy) * W + xx) : 0.f;
    }
    const float4 g0 = __ldg(reinterpret_cast<const float4*>(g + q * 64 + cg));
    const float4 g1 = __ldg(reinterpret_cast<const float4*>(g + q * 64 + cg + 4));
    const float gv[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float pre = sb[cg + j];
#pragma unroll
      for (int t = 0; t < 9; ++t) pre = fmaf(v[t], sw[t][cg + j], pre);
      const float dpre = pre > 0.f ? gv[j] : a * gv[j];
      acca += pre > 0.f ? 0.f : pre * gv[j];
      accb[j] += dpre;
#pragma unroll
      for (int t = 0; t < 9; ++t) acc[j][t] = fmaf(dpre, v[t], acc[j][t]);
    }
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) {
#pragma unroll
    for (int t = 0; t < 9; ++t) {
      float s = acc[j][t];
      s += __shfl_xor_sync(0xffffffffu, s, 8);
      s += __shfl_xor_sync(0xffffffffu, s, 16);
      if ((threadIdx.x & 31) < 8) atomicAdd(&sacc[(cg + j) * 9 + t], s);
    }
    float s = accb[j];
    s += __shfl_xor_sync(0xffffffffu, s, 8);
    s += __shfl_xor_sync(0xffffffffu, s, 16);
    if ((threadIdx.x & 31) < 8) atomicAdd(&sacc[576 + cg + j], s);
  }
  acca = warp_sum(acca);
  if ((threadIdx.x & 31) == 0) atomicAdd(&sacc[640], acca);
  __syncthreads();
  for (int i = threadIdx.x; i < 576; i += blockDim.x) atomicAdd(dw + i, sacc[i]);
  if (threadIdx.x < 64) atomicAdd(db + threadIdx.x, sacc[576 + threadIdx.x]);
  if (threadIdx.x == 0) atomicAdd(dslope, sacc[640]);
}

int launch_in_conv_prelu_bwd(const float* x, const float* w, const float* b, const float* slope, const float* g,
                             float* dw, float* db, float* dslope, long long n_img, int H, int W, int num_sms,
                             cudaStream_t s) {
  const long long n_pix = n_img * H * W;
  if (n_pix == 0) return 0;
  long long blocks = (n_pix + 31) / 32;
  const long long cap = 2LL * (num_sms > 0 ? num_sms : 148);
  if (blocks > cap) blocks = cap;
  in_conv_prelu_bwd_kernel<<<static_cast<unsigned>(blocks), 256, 0, s>>>(x, w, b, slope, g, dw, db, dslope, n_pix, H,
                                                                        W);
  return static_cast<int>(cudaGetLastError());
}

// ------------------------------------------------------------------------------------------------
// posterm_bwd: gradient of the positional-code input channels of conv1.  With S[img][cls][o] the sum of the conv1
// output gradient G over the pixels of border class cls,
//   dW1[o, (2F+1)*d + 2F, tap] += sum_{cls admitting tap} Q[d][cls][o],   Q[d][cls][o] = sum_img pos[b, f + d] * S.
// Step 1: one block per (image, chunk of rows), one thread per channel; per row the (left, middle, right) sums go to
// the row's classes, then the block adds pos-weighted sums into Q (only the few non-empty classes).  Step 2 folds Q
// over the classes that admit each tap.  img = f * B + b; the window of gradient frame f covers input frames
// frame0 + f + d.
constexpr int kPosChunks = 8;

__global__ void __launch_bounds__(192) posterm_bwd_sums_kernel(const __nv_bfloat16* __restrict__ g,
                                                               const float* __restrict__ pos, float* __restrict__ Q,
                                                               int B, int L, int frame0, int window, int H, int W,
                                                               int ch) {
  const int img = blockIdx.x, chunk = blockIdx.y;
  const int o = threadIdx.x;
  if (o >= ch) return;
  const int rows = (H + kPosChunks - 1) / kPosChunks;
  const int ya = chunk * rows, yb = min(H, ya + rows);
  float acc[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) acc[i] = 0.f;
  const __nv_bfloat16* base = g + static_cast<size_t>(img) * H * W * ch + o;
  for (int y = ya; y < yb; ++y) {
    const int rc = (y > 0 ? 1 : 0) | (y < H - 1 ? 2 : 0);
    const __nv_bfloat16* rowp = base + static_cast<size_t>(y) * W * ch;
    const float left = __bfloat162float(rowp[0]);
    const float right = W > 1 ? __bfloat162float(rowp[static_cast<size_t>(W - 1) * ch]) : 0.f;
    float mid = 0.f;
    for (int x = 1; x < W - 1; ++x) mid += __bfloat162float(rowp[static_cast<size_t>(x) * ch]);
    // classes: bit2 = x > 0, bit3 = x < W-1
    const int cl = rc | (W > 1 ? 8 : 0), cr = rc | 4, cm = rc | 12;
#pragma unroll
    for (int i = 0; i < 16; ++i)
      acc[i] += (i == cl ? left : 0.f) + (i == cm ? mid : 0.f) + ((i == cr && W > 1) ? right : 0.f);
  }
  const int f = img / B, b = img - f * B;
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    if (acc[i] == 0.f) continue;
    for (int d = 0; d < window; ++d)
      atomicAdd(Q + (static_cast<size_t>(d) * 16 + i) * ch + o, pos[static_cast<long long>(b) * L + frame0 + f + d] * acc[i]);
  }
}

__global__ void posterm_bwd_reduce_kernel(const float* __restrict__ Q, float* __restrict__ dw1, int window, int c_out,
                                          int c_in, int feat2, int ch) {
  const int gid = blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= c_out * window * 9) return;
  const int tap = gid % 9;
  const int d = (gid / 9) % window;
  const int o = gid / (9 * window);
  const int dy = tap / 3 - 1, dx = tap % 3 - 1;
  float s = 0.f;
#pragma unroll
  for (int cls = 0; cls < 16; ++cls) {
    const bool ok = (dy >= 0 || (cls & 1)) && (dy <= 0 || (cls & 2)) && (dx >= 0 || (cls & 4)) && (dx <= 0 || (cls & 8));
    if (ok) s += Q[(static_cast<size_t>(d) * 16 + cls) * ch + o];
  }
  atomicAdd(dw1 + (static_cast<long long>(o) * c_in + (feat2 + 1) * d + feat2) * 9 + tap, s);
}

int launch_posterm_bwd(const void* g_bf16, const float* pos, float* sums, float* dw1, int n_frames, int B, int L,
                       int frame0, int window, int H, int W, int c_out, int c_in, int feat2, int ch, cudaStream_t s) {
  if (n_frames * B == 0) return 0;
  if (ch > 192) return static_cast<int>(cudaErrorInvalidValue);
  // Q [window][16][ch] is the scratch buffer
  int e = static_cast<int>(cudaMemsetAsync(sums, 0, static_cast<size_t>(window) * 16 * ch * sizeof(float), s));
  if (e) return e;
  dim3 grid(n_frames * B, kPosChunks);
  posterm_bwd_sums_kernel<<<grid, 192, 0, s>>>(static_cast<const __nv_bfloat16*>(g_bf16), pos, sums, B, L, frame0,
                                               window, H, W, ch);
  e = static_cast<int>(cudaGetLastError());
  if (e) return e;
  const int total = c_out * window * 9;
  posterm_bwd_reduce_kernel<<<(total + 127) / 128, 128, 0, s>>>(sums, dw1, window, c_out, c_in, feat2, ch);
  return static_cast<int>(cudaGetLastError());
}

// ------------------------------------------------------------------------------------------------
// cast_f32_bf16: out = bf16(in) over n8 groups of 8 elements.
__global__ void cast_f32_bf16_kernel(const float4* __restrict__ in, uint4* __restrict__ out, long long n8) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n8) return;
  const float4 a = in[2 * i], b = in[2 * i + 1];
  out[i] = make_uint4(pack2(a.x, a.y), pack2(a.z, a.w), pack2(b.x, b.y), pack2(b.z, b.w));
}
int launch_cast_f32_bf16(const float* in, void* out, long long n, cudaStream_t s) {
  const long long n8 = n / 8;
  if (n8 == 0) return 0;
  cast_f32_bf16_kernel<<<static_cast<unsigned>((n8 + 255) / 256), 256, 0, s>>>(reinterpret_cast<const float4*>(in),
                                                                              static_cast<uint4*>(out), n8);
  return static_cast<int>(cudaGetLastError());
}

// ------------------------------------------------------------------------------------------------
// Adam (torch.optim.Adam.step, amsgrad=False, maximize=False) on flat fp32 buffers.
// state[0] = step count (as float), incremented by adam_tick before the update kernel reads it.
__global__ void adam_tick_kernel(float* state) { state[0] += 1.f; }

__global__ void __launch_bounds__(256) adam_kernel(float4* __restrict__ p, const float4* __restrict__ g,
                                                   float4* __restrict__ m, float4* __restrict__ v, long long n4,
                                                   float lr, float b1, float b2, float eps, float wd,
                                                   float grad_scale, const float* __restrict__ state) {
  const float step = state[0];
  const float bc1 = 1.f - powf(b1, step);
  const float bc2 = 1.f - powf(b2, step);
  const float step_size = lr / bc1;
  const float inv_sqrt_bc2 = rsqrtf(bc2);
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n4;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    float4 pp = p[i], gg = g[i], mm = m[i], vv = v[i];
    float* pf = reinterpret_cast<float*>(&pp);
    float* gf = reinterpret_cast<float*>(&gg);
    float* mf = reinterpret_cast<float*>(&mm);
    float* vf = reinterpret_cast<float*>(&vv);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      float gr = gf[k] * grad_scale;
      if (wd != 0.f) gr = fmaf(wd, pf[k], gr);
      mf[k] = fmaf(b1, mf[k], (1.f - b1) * gr);
      vf[k] = fmaf(b2, vf[k], (1.f - b2) * gr * gr);
      const float denom = sqrtf(vf[k]) * inv_sqrt_bc2 + eps;
      pf[k] -= step_size * (mf[k] / denom);
    }
    p[i] = pp; m[i] = mm; v[i] = vv;
  }
}

int launch_adam(float* p, const float* g, float* m, float* v, long long n, float lr, float b1, float b2, float eps,
                float wd, float grad_scale, float* state, int num_sms, cudaStream_t s) {
  if (n % 4 != 0) return static_cast<int>(cudaErrorInvalidValue);
  if (n == 0) return 0;
  adam_tick_kernel<<<1, 1, 0, s>>>(state);
  long long blocks = (n / 4 + 255) / 256;
  const long long cap = 8LL * (num_sms > 0 ? num_sms : 148);
  if (blocks > cap) blocks = cap;
  adam_kernel<<<static_cast<unsigned>(blocks), 256, 0, s>>>(reinterpret_cast<float4*>(p),
                                                           reinterpret_cast<const float4*>(g),
                                                           reinterpret_cast<float4*>(m), reinterpret_cast<float4*>(v),
                                                           n / 4, lr, b1, b2, eps, wd, grad_scale, state);
  return static_cast<int>(cudaGetLastError());
}

}  // namespace pvsr
