// Memory-bound kernels of the RefineNet path (no tensor-core work): vectorised, shared-memory staged.
//   in_conv_prelu   : _InBlock           (refine_net.py:188-192)  1 -> 64 channels, 3x3, PReLU
//   head_conv_last  : _OutBlock last conv(refine_net.py:203,205)  64 -> 1 channel at HR, fp32 out
//   posterm_build   : positional-code channel of _RefineBlock conv1 as a border-class additive table
//                     (refine_net.py:168-172: the code is a constant plane per frame; zero padding makes
//                      its 3x3 response depend only on which taps fall inside the image)
//   pack_weights    : fp32 parameter -> packed bf16 GEMM operand through a gather index
//   add_bf16        : feature updates x += h (refine_net.py:120-131) and head inputs x + h (:102,107)
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "simt.h"

namespace pvsr {

// ------------------------------------------------------------------------------------------------
// in_conv_prelu: x fp32 [n_img][H][W] -> out bf16 NHWC [n_img][H][W][64]
// 8 threads per pixel, 8 channels each; a warp writes 4 pixels x 128 B contiguous.
__global__ void __launch_bounds__(256) in_conv_prelu_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                            const float* __restrict__ b,
                                                            const float* __restrict__ slope,
                                                            __nv_bfloat16* __restrict__ out, long long n_pix_total,
                                                            int H, int W) {
  __shared__ float sw[9][64];
  __shared__ float sb[64];
  for (int i = threadIdx.x; i < 576; i += blockDim.x) {
    int o = i / 9, t = i % 9;  // parameter layout (64, 1, 3, 3)
    sw[t][o] = w[i];
  }
  if (threadIdx.x < 64) sb[threadIdx.x] = b[threadIdx.x];
  __syncthreads();
  const float a = slope[0];
  const long long gid = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const long long pix = gid >> 3;
  const int cg = static_cast<int>(gid & 7) * 8;
  if (pix >= n_pix_total) return;
  const int xw = static_cast<int>(pix % W);
  const int yh = static_cast<int>((pix / W) % H);
  const float* img = x + (pix - static_cast<long long>(yh) * W - xw);
  float v[9];
#pragma unroll
  for (int t = 0; t < 9; ++t) {
    const int yy = yh + t / 3 - 1, xx = xw + t % 3 - 1;
    v[t] = (yy >= 0 && yy < H && xx >= 0 && xx < W) ? __ldg(img + static_cast<long long>(yy) * W + xx) : 0.f;
  }
  float acc[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) acc[j] = sb[cg + j];
#pragma unroll
  for (int t = 0; t < 9; ++t)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = fmaf(v[t], sw[t][cg + j], acc[j]);
  uint32_t pk[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    float f0 = acc[2 * j], f1 = acc[2 * j + 1];
    f0 = f0 >= 0.f ? f0 : a * f0;
    f1 = f1 >= 0.f ? f1 : a * f1;
    __nv_bfloat162 h = __floats2bfloat162_rn(f0, f1);
    pk[j] = *reinterpret_cast<uint32_t*>(&h);
  }
  *reinterpret_cast<uint4*>(out + pix * 64 + cg) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
}

int launch_in_conv_prelu(const float* x, const float* w, const float* b, const float* slope, void* out,
                         long long n_img, int H, int W, cudaStream_t s) {
  const long long n_pix = n_img * H * W;
  if (n_pix == 0) return 0;
  const long long threads = n_pix * 8;
  const unsigned grid = static_cast<unsigned>((threads + 255) / 256);
  in_conv_prelu_kernel<<<grid, 256, 0, s>>>(x, w, b, slope, static_cast<__nv_bfloat16*>(out), n_pix, H, W);
  return static_cast<int>(cudaGetLastError());
}

// ------------------------------------------------------------------------------------------------
// head_conv_last: in bf16 NHWC [n_img][H][W][64] -> out fp32 [n_img][H][W]; out = b + sum_{tap,c} in*w
// Block = 8x32 output pixels; the (10 x 34) halo tile is staged in shared memory with a 144-byte pixel
// pitch (conflict-free 16-byte reads across consecutive pixels); one thread per output pixel.
constexpr int kLastTH = 8, kLastTW = 32, kLastPitch = 144;
constexpr int kLastHaloPix = (kLastTH + 2) * (kLastTW + 2);

__global__ void __launch_bounds__(256) head_conv_last_kernel(const __nv_bfloat16* __restrict__ in,
                                                             const float* __restrict__ w,
                                                             const float* __restrict__ b, float* __restrict__ out,
                                                             const float* __restrict__ target,
                                                             float* __restrict__ l1_partial, int H, int W,
                                                             int tiles_x, int tiles_y) {
  extern __shared__ __align__(16) uint8_t sm[];
  uint8_t* tile = sm;                                                   // kLastHaloPix * 144 B
  float* sw = reinterpret_cast<float*>(sm + kLastHaloPix * kLastPitch); // [9][64] fp32
  int t = blockIdx.x;
  const int tx = t % tiles_x;
  t /= tiles_x;
  const int ty = t % tiles_y;
  const int img = t / tiles_y;
  const int y0 = ty * kLastTH, x0 = tx * kLastTW;
  for (int i = threadIdx.x; i < 576; i += blockDim.x) {
    int c = i / 9, tap = i % 9;  // parameter layout (1, 64, 3, 3)
    sw[tap * 64 + c] = w[i];
  }
  const __nv_bfloat16* src = in + static_cast<size_t>(img) * H * W * 64;
  // 8 x 16-byte chunks per halo pixel
  for (int i = threadIdx.x; i < kLastHaloPix * 8; i += blockDim.x) {
    const int hp = i >> 3, ck = i & 7;
    const int hy = hp / (kLastTW + 2), hx = hp % (kLastTW + 2);
    const int y = y0 + hy - 1, x = x0 + hx - 1;
    uint4 v = make_uint4(0, 0, 0, 0);
    if (y >= 0 && y < H && x >= 0 && x < W)
      v = __ldg(reinterpret_cast<const uint4*>(src + (static_cast<size_t>(y) * W + x) * 64) + ck);
    *reinterpret_cast<uint4*>(tile + hp * kLastPitch + ck * 16) = v;
  }
  __syncthreads();
  const int ly = threadIdx.x / kLastTW, lx = threadIdx.x % kLastTW;
  const int y = y0 + ly, x = x0 + lx;
  float acc = b[0];
#pragma unroll
  for (int tap = 0; tap < 9; ++tap) {
    const uint8_t* prow = tile + ((ly + tap / 3) * (kLastTW + 2) + (lx + tap % 3)) * kLastPitch;
    const float* wt = sw + tap * 64;
#pragma unroll
    for (int ck = 0; ck < 8; ++ck) {
      const uint4 u = *reinterpret_cast<const uint4*>(prow + ck * 16);
      const __nv_bfloat162* h2 = reinterpret_cast<const __nv_bfloat162*>(&u);
      const float4 w0 = *reinterpret_cast<const float4*>(wt + ck * 8);
      const float4 w1 = *reinterpret_cast<const float4*>(wt + ck * 8 + 4);
      float2 f;
      f = __bfloat1622float2(h2[0]); acc = fmaf(f.x, w0.x, acc); acc = fmaf(f.y, w0.y, acc);
      f = __bfloat1622float2(h2[1]); acc = fmaf(f.x, w0.z, acc); acc = fmaf(f.y, w0.w, acc);
      f = __bfloat1622float2(h2[2]); acc = fmaf(f.x, w1.x, acc); acc = fmaf(f.y, w1.y, acc);
      f = __bfloat1622float2(h2[3]); acc = fmaf(f.x, w1.z, acc); acc = fmaf(f.y, w1.w, acc);
    }
  }
  float l1 = 0.f;
  if (y < H && x < W) {
    const size_t o = (static_cast<size_t>(img) * H + y) * W + x;
    out[o] = acc;
    if (target) l1 = fabsf(acc - target[o]);
  }
  if (l1_partial) {
    // |out - target| summed per image: warp shuffle, then one atomic per warp (nn.L1Loss numerator).
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) l1 += __shfl_xor_sync(0xffffffffu, l1, d);
    if ((threadIdx.x & 31) == 0) atomicAdd(l1_partial + img, l1);
  }
}

int launch_head_conv_last(const void* in, const float* w, const float* b, float* out, const float* target,
                          float* l1_partial, long long n_img, int H, int W, cudaStream_t s) {
  if (n_img == 0) return 0;
  const int tiles_x = (W + kLastTW - 1) / kLastTW, tiles_y = (H + kLastTH - 1) / kLastTH;
  const size_t smem = kLastHaloPix * kLastPitch + 576 * sizeof(float);
  static bool attr = false;
  if (!attr) {
    cudaError_t e = cudaFuncSetAttribute(head_conv_last_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         static_cast<int>(smem));
    if (e != cudaSuccess) return static_cast<int>(e);
    attr = true;
  }
  const long long blocks = n_img * tiles_x * tiles_y;
  head_conv_last_kernel<<<static_cast<unsigned>(blocks), 256, smem, s>>>(
      static_cast<const __nv_bfloat16*>(in), w, b, out, target, l1_partial, H, W, tiles_x, tiles_y);
  return static_cast<int>(cudaGetLastError());
}

// ------------------------------------------------------------------------------------------------
// posterm_build: table[img][cls][n] = bias1[o] + sum_j pos[b][f+j] * sum_{taps inside for cls} W1[o, 129*j+128, tap]
// img = f*B + b (frame-major), cls bit0: y>0, bit1: y<H-1, bit2: x>0, bit3: x<W-1; n = packed column (== o, zero
// beyond the real out channels).  W1 is the fp32 parameter (c_out, window*(2F+1), 3, 3).
__global__ void posterm_kernel(const float* __restrict__ w1, const float* __restrict__ b1,
                               const float* __restrict__ pos, float* __restrict__ table, int n_frames_out, int B,
                               int L, int window, int c_out, int c_in, int feat2, int n_total) {
  const long long gid = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const long long total = static_cast<long long>(n_frames_out) * B * 16 * n_total;
  if (gid >= total) return;
  const int n = static_cast<int>(gid % n_total);
  const int cls = static_cast<int>((gid / n_total) % 16);
  const long long img = gid / (static_cast<long long>(n_total) * 16);
  const int b = static_cast<int>(img % B);
  const int f = static_cast<int>(img / B);
  float acc = 0.f;
  if (n < c_out) {
    acc = b1[n];
    for (int j = 0; j < window; ++j) {
      const float p = pos[static_cast<long long>(b) * L + f + j];
      const float* wj = w1 + (static_cast<long long>(n) * c_in + (feat2 + 1) * j + feat2) * 9;
      float s = 0.f;
#pragma unroll
      for (int tap = 0; tap < 9; ++tap) {
        const int dy = tap / 3 - 1, dx = tap % 3 - 1;
        const bool ok = (dy >= 0 || (cls & 1)) && (dy <= 0 || (cls & 2)) && (dx >= 0 || (cls & 4)) &&
                        (dx <= 0 || (cls & 8));
        if (ok) s += wj[tap];
      }
      acc = fmaf(p, s, acc);
    }
  }
  table[gid] = acc;
}

int launch_posterm(const float* w1, const float* b1, const float* pos, float* table, int n_frames_out, int B, int L,
                   int window, int c_out, int c_in, int feat2, int n_total, cudaStream_t s) {
  const long long total = static_cast<long long>(n_frames_out) * B * 16 * n_total;
  if (total == 0) return 0;
  posterm_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, s>>>(w1, b1, pos, table, n_frames_out, B, L,
                                                                           window, c_out, c_in, feat2, n_total);
  return static_cast<int>(cudaGetLastError());
}

// ------------------------------------------------------------------------------------------------
// pack_weights: out[e] = bf16( idx[e] >= 0 ? w[idx[e]] : 0  +  idx2[e] >= 0 ? w[idx2[e]] : 0 )
__global__ void pack_weights_kernel(const float* __restrict__ w, const int* __restrict__ idx,
                                    const int* __restrict__ idx2, __nv_bfloat16* __restrict__ out, long long n) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float v = 0.f;
  const int a = idx[i];
  if (a >= 0) v = w[a];
  if (idx2) {
    const int c = idx2[i];
    if (c >= 0) v += w[c];
  }
  out[i] = __float2bfloat16(v);
}
int launch_pack_weights(const float* w, const int* idx, const int* idx2, void* out, long long n, cudaStream_t s) {
  if (n == 0) return 0;
  pack_weights_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, s>>>(w, idx, idx2,
                                                                            static_cast<__nv_bfloat16*>(out), n);
  return static_cast<int>(cudaGetLastError());
}

// gather fp32 (bias permutation / padding): out[i] = idx[i] >= 0 ? src[idx[i]] : 0
__global__ void gather_f32_kernel(const float* __restrict__ src, const int* __restrict__ idx, float* __restrict__ out,
                                  long long n) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int a = idx[i];
  out[i] = a >= 0 ? src[a] : 0.f;
}
int launch_gather_f32(const float* src, const int* idx, float* out, long long n, cudaStream_t s) {
  if (n == 0) return 0;
  gather_f32_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, s>>>(src, idx, out, n);
  return static_cast<int>(cudaGetLastError());
}

// ------------------------------------------------------------------------------------------------
// add_bf16: out = a + b over n8 groups of 8 bf16 (16-byte vectors), fp32 add, one rounding.
__global__ void add_bf16_kernel(const uint4* __restrict__ a, const uint4* __restrict__ b, uint4* __restrict__ out,
                                long long n8) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n8) return;
  const uint4 ua = a[i], ub = b[i];
  const __nv_bfloat162* ha = reinterpret_cast<const __nv_bfloat162*>(&ua);
  const __nv_bfloat162* hb = reinterpret_cast<const __nv_bfloat162*>(&ub);
  uint4 r;
  uint32_t* pr = reinterpret_cast<uint32_t*>(&r);
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const float2 fa = __bfloat1622float2(ha[j]), fb = __bfloat1622float2(hb[j]);
    __nv_bfloat162 h = __floats2bfloat162_rn(fa.x + fb.x, fa.y + fb.y);
    pr[j] = *reinterpret_cast<uint32_t*>(&h);
  }
  out[i] = r;
}
int launch_add_bf16(const void* a, const void* b, void* out, long long n_elems, cudaStream_t s) {
  const long long n8 = n_elems / 8;
  if (n8 == 0) return 0;
  add_bf16_kernel<<<static_cast<unsigned>((n8 + 255) / 256), 256, 0, s>>>(
      static_cast<const uint4*>(a), static_cast<const uint4*>(b), static_cast<uint4*>(out), n8);
  return static_cast<int>(cudaGetLastError());
}

}  // namespace pvsr
