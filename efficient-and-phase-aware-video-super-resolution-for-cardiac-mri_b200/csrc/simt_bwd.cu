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

#include "simt.h"

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
// lstm_bwd_pointwise.  One block per 128-pixel tile of one cell; thread = (tile row, half of the channels).
// Tile-transposed tensors ([tile][ch][128]) are read coalesced along the row; dh / dgates are NHWC and each thread
// touches whole 32/64-byte runs of its own pixel.
__global__ void __launch_bounds__(256) lstm_bwd_pointwise_kernel(const __grid_constant__ LstmBwdParams p) {
  const LstmBwdProb& pr = p.prob[blockIdx.y];
  const int tile = blockIdx.x;
  const int row = threadIdx.x & 127;
  const int grp = threadIdx.x >> 7;
  int t = tile;
  const int tx = t % p.tiles_x;
  t /= p.tiles_x;
  const int ty = t % p.tiles_y;
  const int img = t / p.tiles_y;
  const int TW = 1 << p.tw_log2;
  const int y = ty * (128 >> p.tw_log2) + (row >> p.tw_log2), x = (tx << p.tw_log2) + (row & (TW - 1));
  const bool valid = y < p.H && x < p.W;
  const size_t pix = (static_cast<size_t>(img) * p.H + y) * p.W + x;
  const __nv_bfloat16* gt = static_cast<const __nv_bfloat16*>(pr.gates) + static_cast<size_t>(tile) * 256 * 128 + row;
  const float* ct = pr.c + static_cast<size_t>(tile) * 64 * 128 + row;
  const float* cpt = pr.c_prev ? pr.c_prev + static_cast<size_t>(tile) * 64 * 128 + row : nullptr;
  float* dct = pr.dc + static_cast<size_t>(tile) * 64 * 128 + row;
#pragma unroll 1
  for (int cc = 0; cc < 2; ++cc) {
    const int ch0 = grp * 32 + cc * 16;
    float dh[16];
    if (valid) {
      const float4* q = reinterpret_cast<const float4*>(pr.dh + pix * 64 + ch0);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float4 v = q[j];
        dh[4 * j] = v.x; dh[4 * j + 1] = v.y; dh[4 * j + 2] = v.z; dh[4 * j + 3] = v.w;
      }
    } else {
#pragma unroll
      for (int j = 0; j < 16; ++j) dh[j] = 0.f;
    }
    float ai[16], af[16], ao[16], ag[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      const int ch = ch0 + j;
      const float gi = __bfloat162float(gt[ch * 128]);
      const float gf = __bfloat162float(gt[(64 + ch) * 128]);
      const float go = __bfloat162float(gt[(128 + ch) * 128]);
      const float gg = __bfloat162float(gt[(192 + ch) * 128]);
      const float cn = ct[ch * 128];
      const float cp = cpt ? cpt[ch * 128] : 0.f;
      const float dcin = pr.dc_zero ? 0.f : dct[ch * 128];
      const float tc = tanhf(cn);
      const float dcv = fmaf(dh[j] * go, 1.f - tc * tc, dcin);
      ao[j] = dh[j] * tc * go * (1.f - go);
      ai[j] = dcv * gg * gi * (1.f - gi);
      af[j] = dcv * cp * gf * (1.f - gf);
      ag[j] = dcv * gi * (1.f - gg * gg);
      dct[ch * 128] = dcv * gf;
    }
    if (valid) {
      __nv_bfloat16* dst = static_cast<__nv_bfloat16*>(pr.dgates) + pix * 256 + ch0;
      uint4* q;
      q = reinterpret_cast<uint4*>(dst);
      q[0] = make_uint4(pack2(ai[0], ai[1]), pack2(ai[2], ai[3]), pack2(ai[4], ai[5]), pack2(ai[6], ai[7]));
      q[1] = make_uint4(pack2(ai[8], ai[9]), pack2(ai[10], ai[11]), pack2(ai[12], ai[13]), pack2(ai[14], ai[15]));
      q = reinterpret_cast<uint4*>(dst + 64);
      q[0] = make_uint4(pack2(af[0], af[1]), pack2(af[2], af[3]), pack2(af[4], af[5]), pack2(af[6], af[7]));
      q[1] = make_uint4(pack2(af[8], af[9]), pack2(af[10], af[11]), pack2(af[12], af[13]), pack2(af[14], af[15]));
      q = reinterpret_cast<uint4*>(dst + 128);
      q[0] = make_uint4(pack2(ao[0], ao[1]), pack2(ao[2], ao[3]), pack2(ao[4], ao[5]), pack2(ao[6], ao[7]));
      q[1] = make_uint4(pack2(ao[8], ao[9]), pack2(ao[10], ao[11]), pack2(ao[12], ao[13]), pack2(ao[14], ao[15]));
      q = reinterpret_cast<uint4*>(dst + 192);
      q[0] = make_uint4(pack2(ag[0], ag[1]), pack2(ag[2], ag[3]), pack2(ag[4], ag[5]), pack2(ag[6], ag[7]));
      q[1] = make_uint4(pack2(ag[8], ag[9]), pack2(ag[10], ag[11]), pack2(ag[12], ag[13]), pack2(ag[14], ag[15]));
    }
  }
}

int launch_lstm_bwd_pointwise(const LstmBwdParams& p, cudaStream_t s) {
  if (p.n_prob <= 0 || p.n_img <= 0) return 0;
  dim3 grid(static_cast<unsigned>(p.n_img * p.tiles_x * p.tiles_y), static_cast<unsigned>(p.n_prob));
  lstm_bwd_pointwise_kernel<<<grid, 256, 0, s>>>(p);
  return static_cast<int>(cudaGetLastError());
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

int launch_l1_multistage(const float* out, const float* target, const float* w, int n_lists, long long n_per_list,
                         float* loss, float* dout, int num_sms, cudaStream_t s) {
  if (n_per_list % 4 != 0) return static_cast<int>(cudaErrorInvalidValue);
  const long long n4 = n_per_list / 4;
  if (n4 == 0 || n_lists == 0) return 0;
  long long blocks = (n4 * n_lists + 255) / 256;
  const long long cap = 8LL * (num_sms > 0 ? num_sms : 148);
  if (blocks > cap) blocks = cap;
  l1_multistage_kernel<<<static_cast<unsigned>(blocks), 256, 0, s>>>(
      reinterpret_cast<const float4*>(out), reinterpret_cast<const float4*>(target), w, n_lists, n4, loss,
      reinterpret_cast<float4*>(dout));
  return static_cast<int>(cudaGetLastError());
}

// ------------------------------------------------------------------------------------------------
// head_last_bwd_data: dIn[y, x, c] = sum_tap dOut[y - dy, x - dx] * w[c, tap]   (zero outside the image)
// 8 threads per pixel (8 channels each); a warp writes 4 pixels x 128 B contiguous.
__global__ void __launch_bounds__(256) head_last_bwd_data_kernel(const float* __restrict__ dout,
                                                                 const float* __restrict__ w,
                                                                 __nv_bfloat16* __restrict__ din,
                                                                 long long n_pix_total, int H, int W) {
  __shared__ float sw[9][64];
  for (int i = threadIdx.x; i < 576; i += blockDim.x) sw[i % 9][i / 9] = w[i];  // parameter layout (1, 64, 3, 3)
  __syncthreads();
  const long long gid = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const long long pix = gid >> 3;
  const int cg = static_cast<int>(gid & 7) * 8;
  if (pix >= n_pix_total) return;
  const int xw = static_cast<int>(pix % W);
  const int yh = static_cast<int>((pix / W) % H);
  const float* img = dout + (pix - static_cast<long long>(yh) * W - xw);
  float acc[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) acc[j] = 0.f;
#pragma unroll
  for (int t = 0; t < 9; ++t) {
    const int yy = yh - (t / 3 - 1), xx = xw - (t % 3 - 1);
    const float v = (yy >= 0 && yy < H && xx >= 0 && xx < W) ? __ldg(img + static_cast<long long>(yy) * W + xx) : 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = fmaf(v, sw[t][cg + j], acc[j]);
  }
  *reinterpret_cast<uint4*>(din + pix * 64 + cg) =
      make_uint4(pack2(acc[0], acc[1]), pack2(acc[2], acc[3]), pack2(acc[4], acc[5]), pack2(acc[6], acc[7]));
}

int launch_head_last_bwd_data(const float* dout, const float* w, void* din_bf16, long long n_img, int H, int W,
                              cudaStream_t s) {
  const long long n_pix = n_img * H * W;
  if (n_pix == 0) return 0;
  const long long threads = n_pix * 8;
  head_last_bwd_data_kernel<<<static_cast<unsigned>((threads + 255) / 256), 256, 0, s>>>(
      dout, w, static_cast<__nv_bfloat16*>(din_bf16), n_pix, H, W);
  return static_cast<int>(cudaGetLastError());
}

// ------------------------------------------------------------------------------------------------
// head_last_bwd_weight: dW[c, tap] += sum_q in[q, c] * dOut[q - off(tap)],  db += sum dOut.
// Thread = (pixel lane, 8-channel group); 72 fp32 accumulators; block-level reduction in shared memory, then one
// atomic per (c, tap) per block.  Grid is persistent (a few blocks per SM).
__global__ void __launch_bounds__(256) head_last_bwd_weight_kernel(const __nv_bfloat16* __restrict__ in,
                                                                   const float* __restrict__ dout,
                                                                   float* __restrict__ dw, float* __restrict__ db,
                                                                   long long n_pix_total, int H, int W) {
  __shared__ float sacc[577];
  for (int i = threadIdx.x; i < 577; i += blockDim.x) sacc[i] = 0.f;
  __syncthreads();
  const int pl = threadIdx.x >> 3, cg = (threadIdx.x & 7) * 8;
  float acc[8][9];
#pragma unroll
  for (int j = 0; j < 8; ++j)
#pragma unroll
    for (int t = 0; t < 9; ++t) acc[j][t] = 0.f;
  float bsum = 0.f;
  for (long long q = static_cast<long long>(blockIdx.x) * 32 + pl; q < n_pix_total;
       q += static_cast<long long>(gridDim.x) * 32) {
    const int xw = static_cast<int>(q % W);
    const int yh = static_cast<int>((q / W) % H);
    const float* img = dout + (q - static_cast<long long>(yh) * W - xw);
    const uint4 u = __ldg(reinterpret_cast<const uint4*>(in + q * 64 + cg));
    const float f[8] = {bf16lo(u.x), bf16hi(u.x), bf16lo(u.y), bf16hi(u.y),
                        bf16lo(u.z), bf16hi(u.z), bf16lo(u.w), bf16hi(u.w)};
#pragma unroll
    for (int t = 0; t < 9; ++t) {
      const int yy = yh - (t / 3 - 1), xx = xw - (t % 3 - 1);
      const float v = (yy >= 0 && yy < H && xx >= 0 && xx < W) ? __ldg(img + static_cast<long long>(yy) * W + xx) : 0.f;
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j][t] = fmaf(f[j], v, acc[j][t]);
      if (t == 4 && cg == 0) bsum += v;
    }
  }
  // reduce over the 4 pixel lanes that share a warp (lane bits 3,4), then shared-memory atomics across warps
#pragma unroll
  for (int j = 0; j < 8; ++j)
#pragma unroll
    for (int t = 0; t < 9; ++t) {
      float v = acc[j][t];
      v += __shfl_xor_sync(0xffffffffu, v, 8);
      v += __shfl_xor_sync(0xffffffffu, v, 16);
      if ((threadIdx.x & 31) < 8) atomicAdd(&sacc[(cg + j) * 9 + t], v);
    }
  bsum += __shfl_xor_sync(0xffffffffu, bsum, 8);
  bsum += __shfl_xor_sync(0xffffffffu, bsum, 16);
  if ((threadIdx.x & 31) == 0) atomicAdd(&sacc[576], bsum);
  __syncthreads();
  for (int i = threadIdx.x; i < 576; i += blockDim.x) atomicAdd(dw + i, sacc[i]);
  if (threadIdx.x == 0) atomicAdd(db, sacc[576]);
}

int launch_head_last_bwd_weight(const void* in_bf16, const float* dout, float* dw, float* db, long long n_img, int H,
                                int W, int num_sms, cudaStream_t s) {
  const long long n_pix = n_img * H * W;
  if (n_pix == 0) return 0;
  long long blocks = (n_pix + 31) / 32;
  const long long cap = 4LL * (num_sms > 0 ? num_sms : 148);
  if (blocks > cap) blocks = cap;
  head_last_bwd_weight_kernel<<<static_cast<unsigned>(blocks), 256, 0, s>>>(
      static_cast<const __nv_bfloat16*>(in_bf16), dout, dw, db, n_pix, H, W);
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
// posterm_bwd, step 1: border-class sums of the conv1 output gradient.
//   S[img][cls][o] = sum over pixels of class cls of G[img, pixel, o]        (G bf16 NHWC with `ch` channels)
// One block per image, one thread per channel; per image row the (left, middle, right) sums are flushed to the
// shared accumulator of the row's classes.
__global__ void __launch_bounds__(192) posterm_bwd_sums_kernel(const __nv_bfloat16* __restrict__ g,
                                                               float* __restrict__ sums, int H, int W, int ch) {
  const int img = blockIdx.x;
  const int o = threadIdx.x;
  if (o >= ch) return;
  float acc[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) acc[i] = 0.f;
  const __nv_bfloat16* base = g + static_cast<size_t>(img) * H * W * ch + o;
  for (int y = 0; y < H; ++y) {
    const int rc = (y > 0 ? 1 : 0) | (y < H - 1 ? 2 : 0);
    const __nv_bfloat16* rowp = base + static_cast<size_t>(y) * W * ch;
    float left = __bfloat162float(rowp[0]);
    float right = W > 1 ? __bfloat162float(rowp[static_cast<size_t>(W - 1) * ch]) : 0.f;
    float mid = 0.f;
    for (int x = 1; x < W - 1; ++x) mid += __bfloat162float(rowp[static_cast<size_t>(x) * ch]);
    // classes: bit2 = x > 0, bit3 = x < W-1
    const int cl = rc | (W > 1 ? 8 : 0);
    const int cr = rc | 4;
    const int cm = rc | 12;
    // registers indexed by a runtime class: resolved with a short unrolled select to stay out of local memory
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      acc[i] += (i == cl ? left : 0.f) + (i == cm ? mid : 0.f) + ((i == cr && W > 1) ? right : 0.f);
    }
  }
#pragma unroll
  for (int i = 0; i < 16; ++i) sums[(static_cast<size_t>(img) * 16 + i) * ch + o] = acc[i];
}

// step 2: dW1[o, (2F+1)*d + 2F, tap] += sum_img pos[b(img), f(img) + d] * sum_{cls admitting tap} S[img][cls][o]
// img = f * B + b over the n_frames gradient frames; the window of gradient frame f covers input frames
// frame0 + f + d, d = 0..window-1.
__global__ void posterm_bwd_reduce_kernel(const float* __restrict__ sums, const float* __restrict__ pos,
                                          float* __restrict__ dw1, int n_frames, int B, int L, int frame0, int window,
                                          int c_out, int c_in, int feat2, int ch) {
  const int gid = blockIdx.x * blockDim.x + threadIdx.x;
  const int total = c_out * window * 9;
  if (gid >= total) return;
  const int tap = gid % 9;
  const int d = (gid / 9) % window;
  const int o = gid / (9 * window);
  const int dy = tap / 3 - 1, dx = tap % 3 - 1;
  float acc = 0.f;
  for (int f = 0; f < n_frames; ++f)
    for (int b = 0; b < B; ++b) {
      const float pc = pos[static_cast<long long>(b) * L + frame0 + f + d];
      const float* S = sums + (static_cast<size_t>(f) * B + b) * 16 * ch + o;
      float s = 0.f;
#pragma unroll
      for (int cls = 0; cls < 16; ++cls) {
        const bool ok = (dy >= 0 || (cls & 1)) && (dy <= 0 || (cls & 2)) && (dx >= 0 || (cls & 4)) &&
                        (dx <= 0 || (cls & 8));
        if (ok) s += S[static_cast<size_t>(cls) * ch];
      }
      acc = fmaf(pc, s, acc);
    }
  atomicAdd(dw1 + (static_cast<long long>(o) * c_in + (feat2 + 1) * d + feat2) * 9 + tap, acc);
}

int launch_posterm_bwd(const void* g_bf16, const float* pos, float* sums, float* dw1, int n_frames, int B, int L,
                       int frame0, int window, int H, int W, int c_out, int c_in, int feat2, int ch, cudaStream_t s) {
  if (n_frames * B == 0) return 0;
  if (ch > 192) return static_cast<int>(cudaErrorInvalidValue);
  posterm_bwd_sums_kernel<<<n_frames * B, 192, 0, s>>>(static_cast<const __nv_bfloat16*>(g_bf16), sums, H, W, ch);
  int e = static_cast<int>(cudaGetLastError());
  if (e) return e;
  const int total = c_out * window * 9;
  posterm_bwd_reduce_kernel<<<(total + 127) / 128, 128, 0, s>>>(sums, pos, dw1, n_frames, B, L, frame0, window, c_out,
                                                                c_in, feat2, ch);
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
