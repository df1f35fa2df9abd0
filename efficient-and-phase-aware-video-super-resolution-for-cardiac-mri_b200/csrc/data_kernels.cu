// Input-pipeline kernel: batches are cut out of cine volumes that stay RESIDENT in HBM.
//   cine_gather : what AcdcVSRRefineNetDataset.__getitem__ + the transform chain do per item on the host
//                 (reference src/data/datasets/acdc_vsr_refinenet_dataset.py:65-87: circular padding of the cardiac
//                 cycle, frame window; src/data/transforms.py:100-168 Normalize, :321-426 flips and RandomCropPatch),
//                 for a whole batch in one launch.  The random decisions are drawn on the host (same numpy stream as
//                 the host path) and arrive as one descriptor per sample; the kernel is a pure gather:
//                     out[f][n][y][x] = (float(vol_n[(t_first_n + f) mod T_n][ay*y + by][ax*x + bx]) - mean) / std
//                 (IEEE sub + div in fp32; fp64 volumes in fp64 then rounded once, which is what numpy's
//                 `(img - mean) / (std + 1e-10)` followed by ToTensor's `.float()` gives for either dtype)
//                 HBM-bound: 2-4 B read + 4 B written per output pixel, rows contiguous (reversed rows under a
//                 horizontal flip still cover whole 128 B lines per warp).
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/pvsr.h"
#include "simt.h"

namespace pvsr {

// Thread = 4 consecutive output pixels of one row (w % 4 == 0: one row/column division per 4 pixels, four independent
// loads in flight, one 16-byte store) or 1 pixel (any w).  First version (1 pixel per thread, grid-stride with a
// division per pixel): 1.96 TB/s on 32 whole cycles (ncu, profiles/r01/ncu_r01s5_loader_*.txt).
template <typename T, int V>
__global__ void __launch_bounds__(256) cine_gather_kernel(const T* __restrict__ vols,
                                                          const pvsr_cine_sample* __restrict__ samples, int n_samples,
                                                          int n_frames, int h, int w, double mean_d, double std_d,
                                                          float* __restrict__ out, const float* __restrict__ pos_codes,
                                                          float* __restrict__ pos_out) {
  // blockIdx.y = frame * n_samples + sample; blockIdx.x strides over the pixel groups of that image
  const int img = blockIdx.y;
  const int f = img / n_samples, n = img - f * n_samples;
  const pvsr_cine_sample s = samples[n];
  int t = (s.t_first + f) % s.T;
  if (t < 0) t += s.T;
  const T* __restrict__ src = vols + s.vol_off + static_cast<int64_t>(t) * s.Hs * s.Ws;
  float* __restrict__ dst = out + static_cast<int64_t>(img) * h * w;
  const int wg = w / V, n_grp = h * wg;
  for (int g = blockIdx.x * blockDim.x + threadIdx.x; g < n_grp; g += gridDim.x * blockDim.x) {
    const int y = g / wg, x = (g - y * wg) * V;
    const T* __restrict__ row = src + (s.ay * y + s.by) * s.Ws + s.bx;     // a frame is far below 2^31 elements
    float v[V];
    if constexpr (sizeof(T) == 8) {
#pragma unroll
      for (int k = 0; k < V; ++k)
        v[k] = static_cast<float>(__ddiv_rn(__dsub_rn(static_cast<double>(row[s.ax * (x + k)]), mean_d), std_d));
    } else {
      const float mean = static_cast<float>(mean_d), stdv = static_cast<float>(std_d);
#pragma unroll
      for (int k = 0; k < V; ++k) v[k] = static_cast<float>(row[s.ax * (x + k)]);
#pragma unroll
      for (int k = 0; k < V; ++k) v[k] = __fdiv_rn(__fsub_rn(v[k], mean), stdv);
    }
    if constexpr (V == 4) *reinterpret_cast<float4*>(dst + y * w + x) = make_float4(v[0], v[1], v[2], v[3]);
    else dst[y * w + x] = v[0];
  }
  if (pos_out != nullptr && blockIdx.x == 0 && threadIdx.x == 0)
    pos_out[static_cast<int64_t>(n) * n_frames + f] = s.pos_off >= 0 ? pos_codes[s.pos_off + t] : 0.f;
}

int launch_cine_gather(const void* vols, int dtype, const pvsr_cine_sample* samples, int n_samples, int n_frames, int h,
                       int w, double mean, double stdv, float* out, const float* pos_codes, float* pos_out,
                       cudaStream_t st) {
  if (n_samples <= 0 || n_frames <= 0 || h <= 0 || w <= 0) return 0;
  const long long imgs = static_cast<long long>(n_samples) * n_frames;
  if (imgs > 65535) return static_cast<int>(cudaErrorInvalidValue);
  const bool vec = (w % 4 == 0) && (reinterpret_cast<uintptr_t>(out) % 16 == 0);
  const int groups = vec ? h * (w / 4) : h * w;
  int bx = (groups + 256 * 2 - 1) / (256 * 2);         // ~2 groups per thread
  if (bx < 1) bx = 1;
  dim3 grid(bx, static_cast<unsigned>(imgs));
#define PVSR_GATHER(T) \
  do { \
    if (vec) cine_gather_kernel<T, 4><<<grid, 256, 0, st>>>(static_cast<const T*>(vols), samples, n_samples, n_frames, h, \
                                                            w, mean, stdv, out, pos_codes, pos_out); \
    else cine_gather_kernel<T, 1><<<grid, 256, 0, st>>>(static_cast<const T*>(vols), samples, n_samples, n_frames, h, w, \
                                                        mean, stdv, out, pos_codes, pos_out); \
  } while (0)
  switch (dtype) {
    case PVSR_DT_F32: PVSR_GATHER(float); break;
    case PVSR_DT_I16: PVSR_GATHER(int16_t); break;
    case PVSR_DT_U16: PVSR_GATHER(uint16_t); break;
    case PVSR_DT_U8: PVSR_GATHER(uint8_t); break;
    case PVSR_DT_F64: PVSR_GATHER(double); break;
    default: return static_cast<int>(cudaErrorInvalidValue);
  }
#undef PVSR_GATHER
  return static_cast<int>(cudaGetLastError());
}

// ---------------------------------------------------------------------------------------------------------------
// Single-channel ends of a net whose features are tensor-core operands (EDSR, edsr_net.py:29,33): the image (or the
// gradient of the output) becomes channel 0 of a zero-padded 64-channel bf16 K block; the 1-channel result of the tail
// conv is column 0 of its 16-column fp32 output.
__global__ void __launch_bounds__(256) pad_channel_bf16_kernel(const float* __restrict__ x, uint4* __restrict__ out,
                                                               long long n) {
  // 8 threads per pixel, one 16-byte store each: the pixel's 128-byte row is written by one quarter-warp
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const long long pix = i >> 3;
  if (pix >= n) return;
  uint4 v = make_uint4(0, 0, 0, 0);
  if ((i & 7) == 0) {
    __nv_bfloat16 h = __float2bfloat16(x[pix]);
    v.x = *reinterpret_cast<unsigned short*>(&h);
  }
  out[i] = v;
}
int launch_pad_channel_bf16(const float* x, void* out, long long n, cudaStream_t s) {
  if (n <= 0) return 0;
  const long long threads = n * 8;
  pad_channel_bf16_kernel<<<static_cast<unsigned>((threads + 255) / 256), 256, 0, s>>>(x, static_cast<uint4*>(out), n);
  return static_cast<int>(cudaGetLastError());
}

__global__ void __launch_bounds__(256) take_channel0_kernel(const float* __restrict__ in, int stride,
                                                            float* __restrict__ out, long long n) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < n) out[i] = in[i * stride];
}
int launch_take_channel0(const float* in, int stride, float* out, long long n, cudaStream_t s) {
  if (n <= 0) return 0;
  take_channel0_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, s>>>(in, stride, out, n);
  return static_cast<int>(cudaGetLastError());
}

// ---------------------------------------------------------------------------------------------------------------
// Table-driven pack / gather / scatter (include/pvsr.h: pvsr_table_job): blockIdx.y = job, blockIdx.x grid-strides over
// the job's elements.  One launch per training step replaces one pack + one bias gather + one transposed pack per
// layer (EDSR: 207 launches) resp. two scatters per layer.
__global__ void __launch_bounds__(256) table_kernel(const pvsr_table_job* __restrict__ jobs) {
  const pvsr_table_job j = jobs[blockIdx.y];
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  const int* __restrict__ idx = j.idx;
  const float* __restrict__ src = static_cast<const float*>(j.src);
  if (j.kind == PVSR_TJ_PACK) {
    // 2 consecutive elements per thread -> one 4-byte bf16x2 store (n is a multiple of 64)
    __nv_bfloat162* __restrict__ dst = static_cast<__nv_bfloat162*>(j.dst);
    for (long long e = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; 2 * e < j.n; e += stride) {
      const int2 i2 = reinterpret_cast<const int2*>(idx)[e];
      const float a = i2.x >= 0 ? __ldg(src + i2.x) : 0.f;
      const float b = i2.y >= 0 ? __ldg(src + i2.y) : 0.f;
      dst[e] = __floats2bfloat162_rn(a, b);
    }
  } else if (j.kind == PVSR_TJ_GATHER) {
    float* __restrict__ dst = static_cast<float*>(j.dst);
    for (long long e = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; e < j.n; e += stride) {
      const int i = idx[e];
      dst[e] = i >= 0 ? __ldg(src + i) : 0.f;
    }
  } else {
    float* __restrict__ dst = static_cast<float*>(j.dst);
    for (long long e = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; e < j.n; e += stride) {
      const int i = idx[e];
      if (i >= 0) atomicAdd(dst + i, j.scale * src[e]);
    }
  }
}
int launch_table(const pvsr_table_job* jobs, int n_jobs, long long max_n, cudaStream_t s) {
  if (n_jobs <= 0 || max_n <= 0) return 0;
  long long bx = (max_n + 256 * 8 - 1) / (256 * 8);      // ~8 elements per thread of the largest job
  if (bx > 1024) bx = 1024;
  table_kernel<<<dim3(static_cast<unsigned>(bx), static_cast<unsigned>(n_jobs)), 256, 0, s>>>(jobs);
  return static_cast<int>(cudaGetLastError());
}

// ---------------------------------------------------------------------------------------------------------------
// Per-frame scores of the runners in one pass over the SR / HR frames: the reference calls
// L1Loss, denormalize (src/utils.py:1-20), PSNR (metrics.py:20-36) and SSIM (metrics.py:86-113: five dense 11x11
// Gaussian convolutions) once per frame and reads every scalar back with .item().  Here two launches score N frames:
//   sums[n][0] = sum |sr - hr|                     (normalised frames, over the rectangle)
//   sums[n][1] = sum (D(sr) - D(hr))^2             D(x) = clamp(rint(x * std + mean), 0, 255)
//   sums[n][2] = sum of the SSIM map               (valid 11x11 window positions inside the rectangle)
// rects[n] = (h0, hn, w0, wn) restricts a frame to its cardiac bounding box (CardiacPSNR / CardiacSSIM, :116-165).
// Accumulation: fp32 inside a block, fp64 atomics across blocks.  HBM-bound: 8 B read per pixel and launch.
__device__ __forceinline__ float denorm255(float x, float mean, float stdv) {
  // two roundings like torch's `imgs * std + mean` (src/utils.py:19), then round-half-even like Tensor.round_()
  return fminf(fmaxf(rintf(__fadd_rn(__fmul_rn(x, stdv), mean)), 0.f), 255.f);
}
__device__ __forceinline__ float block_sum_256(float v, float* red) {
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  float t = threadIdx.x < 8 ? red[threadIdx.x] : 0.f;
  if (threadIdx.x < 32) {
#pragma unroll
    for (int d = 4; d > 0; d >>= 1) t += __shfl_xor_sync(0xffffffffu, t, d);
  }
  return t;   // valid in thread 0
}

__global__ void __launch_bounds__(256) frame_l1_mse_kernel(const float* __restrict__ sr, const float* __restrict__ hr,
                                                           const int* __restrict__ rects, int H, int W, float mean,
                                                           float stdv, double* __restrict__ sums) {
  __shared__ float red[8];
  const int n = blockIdx.y;
  const int h0 = rects ? rects[4 * n] : 0, hn = rects ? rects[4 * n + 1] : H;
  const int w0 = rects ? rects[4 * n + 2] : 0, wn = rects ? rects[4 * n + 3] : W;
  const int rw = wn - w0, area = (hn - h0) * rw;
  const float* a = sr + static_cast<long long>(n) * H * W;
  const float* b = hr + static_cast<long long>(n) * H * W;
  float l1 = 0.f, se = 0.f;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < area; i += gridDim.x * blockDim.x) {
    const int y = i / rw, x = i - y * rw;
    const int o = (h0 + y) * W + w0 + x;
    const float va = a[o], vb = b[o];
    l1 += fabsf(va - vb);
    const float dd = denorm255(va, mean, stdv) - denorm255(vb, mean, stdv);
    se = fmaf(dd, dd, se);
  }
  const float t0 = block_sum_256(l1, red);
  const float t1 = block_sum_256(se, red);
  if (threadIdx.x == 0) {
    atomicAdd(sums + 3 * n, static_cast<double>(t0));
    atomicAdd(sums + 3 * n + 1, static_cast<double>(t1));
  }
}

constexpr int kSsTH = 16, kSsTW = 32, kSsK = 11;
constexpr int kSsIH = kSsTH + kSsK - 1, kSsIW = kSsTW + kSsK - 1;      // 26 x 42 input tile

__global__ void __launch_bounds__(256) frame_ssim_kernel(const float* __restrict__ sr, const float* __restrict__ hr,
                                                         const int* __restrict__ rects, int H, int W, float mean,
                                                         float stdv, const float* __restrict__ window, float c1, float c2,
                                                         int tiles_x, double* __restrict__ sums) {
  __shared__ float ta[kSsIH][kSsIW + 1], tb[kSsIH][kSsIW + 1];
  __shared__ float hp[5][kSsIH][kSsTW + 1];
  __shared__ float wk[kSsK];
  __shared__ float red[8];
  const int n = blockIdx.y;
  const int h0 = rects ? rects[4 * n] : 0, hn = rects ? rects[4 * n + 1] : H;
  const int w0 = rects ? rects[4 * n + 2] : 0, wn = rects ? rects[4 * n + 3] : W;
  const int oh = hn - h0 - (kSsK - 1), ow = wn - w0 - (kSsK - 1);      // valid window positions
  const int ty = blockIdx.x / tiles_x, tx = blockIdx.x - ty * tiles_x;
  const int oy0 = ty * kSsTH, ox0 = tx * kSsTW;
  if (oy0 >= oh || ox0 >= ow) return;                                  // tile outside this frame's (smaller) rectangle
  if (threadIdx.x < kSsK) wk[threadIdx.x] = window[threadIdx.x];
  const float* a = sr + static_cast<long long>(n) * H * W;
  const float* b = hr + static_cast<long long>(n) * H * W;
  for (int i = threadIdx.x; i < kSsIH * kSsIW; i += 256) {
    const int ly = i / kSsIW, lx = i - ly * kSsIW;
    const int y = h0 + oy0 + ly, x = w0 + ox0 + lx;
    const bool ok = y < hn && x < wn;
    const int o = y * W + x;
    ta[ly][lx] = ok ? denorm255(a[o], mean, stdv) : 0.f;
    tb[ly][lx] = ok ? denorm255(b[o], mean, stdv) : 0.f;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < kSsIH * kSsTW; i += 256) {             // horizontal pass
    const int ly = i / kSsTW, lx = i - ly * kSsTW;
    float m1 = 0.f, m2 = 0.f, s11 = 0.f, s22 = 0.f, s12 = 0.f;
#pragma unroll
    for (int k = 0; k < kSsK; ++k) {
      const float w = wk[k], va = ta[ly][lx + k], vb = tb[ly][lx + k];
      m1 = fmaf(w, va, m1); m2 = fmaf(w, vb, m2);
      s11 = fmaf(w, va * va, s11); s22 = fmaf(w, vb * vb, s22); s12 = fmaf(w, va * vb, s12);
    }
    hp[0][ly][lx] = m1; hp[1][ly][lx] = m2; hp[2][ly][lx] = s11; hp[3][ly][lx] = s22; hp[4][ly][lx] = s12;
  }
  __syncthreads();
  float acc = 0.f;
  for (int i = threadIdx.x; i < kSsTH * kSsTW; i += 256) {             // vertical pass + SSIM map
    const int ly = i / kSsTW, lx = i - ly * kSsTW;
    if (oy0 + ly >= oh || ox0 + lx >= ow) continue;
    float q[5] = {0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int k = 0; k < kSsK; ++k) {
      const float w = wk[k];
#pragma unroll
      for (int c = 0; c < 5; ++c) q[c] = fmaf(w, hp[c][ly + k][lx], q[c]);
    }
    const float mu1 = q[0], mu2 = q[1];
    const float v1 = q[2] - mu1 * mu1, v2 = q[3] - mu2 * mu2, cov = q[4] - mu1 * mu2;
    acc += ((2.f * mu1 * mu2 + c1) * (2.f * cov + c2)) / ((mu1 * mu1 + mu2 * mu2 + c1) * (v1 + v2 + c2));
  }
  const float t = block_sum_256(acc, red);
  if (threadIdx.x == 0) atomicAdd(sums + 3 * n + 2, static_cast<double>(t));
}

int launch_frame_scores(const float* sr, const float* hr, const int* rects, long long n, int H, int W, float mean,
                        float stdv, const float* window, float value_range, double* sums, cudaStream_t s) {
  if (n <= 0) return 0;
  if (n > 65535) return static_cast<int>(cudaErrorInvalidValue);
  cudaError_t e = cudaMemsetAsync(sums, 0, static_cast<size_t>(n) * 3 * sizeof(double), s);
  if (e != cudaSuccess) return static_cast<int>(e);
  int bx = (H * W + 256 * 8 - 1) / (256 * 8);
  frame_l1_mse_kernel<<<dim3(bx < 1 ? 1 : bx, static_cast<unsigned>(n)), 256, 0, s>>>(sr, hr, rects, H, W, mean, stdv, sums);
  if (H >= kSsK && W >= kSsK) {
    const int tiles_x = (W - kSsK + 1 + kSsTW - 1) / kSsTW, tiles_y = (H - kSsK + 1 + kSsTH - 1) / kSsTH;
    const float c1 = (0.01f * value_range) * (0.01f * value_range), c2 = (0.03f * value_range) * (0.03f * value_range);
    frame_ssim_kernel<<<dim3(static_cast<unsigned>(tiles_x * tiles_y), static_cast<unsigned>(n)), 256, 0, s>>>(
        sr, hr, rects, H, W, mean, stdv, window, c1, c2, tiles_x, sums);
  }
  return static_cast<int>(cudaGetLastError());
}

// ---------------------------------------------------------------------------------------------------------------
// Bicubic up-sampling, the comparison baseline of the reference (src/model/nets/bicubic.py:15:
// nn.Upsample(scale_factor=s, mode='bicubic', align_corners=True)): source coordinate = dst * (in - 1) / (out - 1),
// Keys cubic convolution with A = -0.75, border indices clamped (torch upsample_bicubic2d semantics).
// HBM-bound: 4 B written per output pixel, the 4 x 4 input taps of neighbouring outputs hit L1 / L2.
__device__ __forceinline__ void cubic_coeffs(float t, float (&c)[4]) {
  const float A = -0.75f;
  const float x0 = t + 1.f, x3 = 2.f - t, u = 1.f - t;
  c[0] = ((A * x0 - 5.f * A) * x0 + 8.f * A) * x0 - 4.f * A;
  c[1] = ((A + 2.f) * t - (A + 3.f)) * t * t + 1.f;
  c[2] = ((A + 2.f) * u - (A + 3.f)) * u * u + 1.f;
  c[3] = ((A * x3 - 5.f * A) * x3 + 8.f * A) * x3 - 4.f * A;
}
__global__ void __launch_bounds__(256) bicubic_kernel(const float* __restrict__ in, float* __restrict__ out,
                                                      long long n_img, int h, int w, int H, int W, float sy,
                                                      float sx) {
  const long long total = n_img * H * W;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int X = static_cast<int>(i % W);
    const long long r = i / W;
    const int Y = static_cast<int>(r % H);
    const float* src = in + (r / H) * h * w;
    const float ry = sy * Y, rx = sx * X;
    int iy = static_cast<int>(floorf(ry)), ix = static_cast<int>(floorf(rx));
    iy = iy < h - 1 ? iy : h - 1;
    ix = ix < w - 1 ? ix : w - 1;
    const float ty = fminf(fmaxf(ry - iy, 0.f), 1.f), tx = fminf(fmaxf(rx - ix, 0.f), 1.f);
    float cy[4], cx[4];
    cubic_coeffs(ty, cy);
    cubic_coeffs(tx, cx);
    float acc = 0.f;
#pragma unroll
    for (int a = 0; a < 4; ++a) {
      const int yy = min(max(iy - 1 + a, 0), h - 1);
      float row = 0.f;
#pragma unroll
      for (int b = 0; b < 4; ++b) row += cx[b] * __ldg(src + static_cast<long long>(yy) * w + min(max(ix - 1 + b, 0), w - 1));
      acc += cy[a] * row;
    }
    out[i] = acc;
  }
}
int launch_bicubic(const float* in, float* out, long long n_img, int h, int w, int scale, cudaStream_t s) {
  const int H = h * scale, W = w * scale;
  const long long total = n_img * H * W;
  if (total <= 0) return 0;
  const float sy = H > 1 ? static_cast<float>(h - 1) / static_cast<float>(H - 1) : 0.f;
  const float sx = W > 1 ? static_cast<float>(w - 1) / static_cast<float>(W - 1) : 0.f;
  long long blocks = (total + 255) / 256;
  if (blocks > 148 * 32) blocks = 148 * 32;
  bicubic_kernel<<<static_cast<unsigned>(blocks), 256, 0, s>>>(in, out, n_img, h, w, H, W, sy, sx);
  return static_cast<int>(cudaGetLastError());
}

}  // namespace pvsr
