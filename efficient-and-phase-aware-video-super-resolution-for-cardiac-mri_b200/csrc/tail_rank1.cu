// Rank-1 adjoint of the head's tail:  X --conv3x3 (64 -> 256, W2, b2)--> PixelShuffle(2) --conv3x3 (64 -> 1, w3)--> out
// (reference src/model/nets/refine_net.py:201-205: out_block.conv2 / pixelshuffle2 / conv3 for x4; the last conv +
// PixelShuffle(2) pair and the final conv of any power-of-two scale with at least two shuffle stages).
//
// The last conv has ONE output channel, so the gradient that flows back through it has rank <= 9 per pixel:
//     d a[p][c] = sum_t w3[c][t] * g[p - t]                       (g = dL/d out, one channel at HR; a = the 64-ch HR map)
// Substituting into the data / weight gradients of the 64 -> 256 conv collapses their K = 2304 / N = 256 contractions:
//     dX[z][ci]          = sum_o U_cls(z)[o][ci] * g[2z + o]                                  o in [-3, 4]^2  (64 offsets)
//     dW2[(c,q)][ci][t'] = sum_t w3[c][t] * R[q,t,t'][ci]
//     dw3[c][t]          = sum_q b2[(c,q)] Gq[q,t] + sum_{q,t',ci} W2[(c,q)][ci][t'] * R[q,t,t'][ci]
//     db2[(c,q)]         = sum_t w3[c][t] * Gq[q,t],      db3 = sum_p g[p]
// with   U_cls[o][ci]   = sum over (t', q, t) with -2t' + q - t = o and t' allowed by the border class of z of
//                         sum_c W2[(c,q)][ci][t'] w3[c][t]
//        R[q,t,t'][ci]  = sum over z with z - t' inside the image of g[2z + o(t',q,t)] * X[z][ci]
//                       = S[o][ci] - (row-edge sum) - (column-edge sum) + (corner sum)     (inclusion - exclusion)
//        S[o][ci]       = sum_z g[2z + o] * X[z][ci],     Gq[q,t] = sum_z g[2z + q - t]
// (zero padding of both convs = the "inside the image" conditions; tests/test_ops_gpu.py checks every border case
// against torch autograd).  So the whole backward of the tail needs: one 64-offset stencil of g per X pixel (a
// [pixels x 64] x [64 x 64] product on mma.sync, border pixels recomputed with their class tables) and ONE
// [64 offsets x pixels] x [pixels x 64] correlation per X pixel - 1/36 of the FLOPs of the dgrad + wgrad launches of the
// 64 -> 256 conv, and neither the 64-channel HR gradient (128 B per HR pixel written, then read twice) nor the HR
// activation is touched by the backward pass at all.  bench.py keeps counting the ALGORITHMIC FLOPs of the layer
// (SURVEY.md 8d); DESIGN.md states the executed ones.
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>

#include "simt.h"

namespace pvsr {

namespace {

constexpr int kO = 8;            // offsets per axis for r = 2: o in [-3, 4]
constexpr int kOmin = -3;
constexpr int kNO = kO * kO;     // 64
constexpr int kTH = 8, kTW = 16; // X-pixel tile of the two mma.sync kernels (128 pixels)
constexpr int kGhH = 2 * kTH + kO - 2;   // 22 halo rows of g per tile
constexpr int kGhW = 2 * kTW + kO - 2;   // 38 halo columns
constexpr int kGhPitch = 40;
constexpr int kXPitch = 144;     // bytes per X pixel row in shared memory (128 + 16: conflict-free ldmatrix)

__device__ __forceinline__ float bf_lo(uint32_t u) { return __uint_as_float(u << 16); }
__device__ __forceinline__ float bf_hi(uint32_t u) { return __uint_as_float(u & 0xFFFF0000u); }
__device__ __forceinline__ uint32_t pack_bf2(float a, float b) {
  __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}
__device__ __forceinline__ void mma16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, "
      "{%0, %1, %2, %3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// Border class of X pixel (zy, zx): bit 0 first row, bit 1 last row, bit 2 first column, bit 3 last column.
__device__ __forceinline__ int border_class(int zy, int zx, int H1, int W1) {
  return (zy == 0 ? 1 : 0) | (zy == H1 - 1 ? 2 : 0) | (zx == 0 ? 4 : 0) | (zx == W1 - 1 ? 8 : 0);
}

// g halo of one tile -> shared memory (zero outside the HR image): gh[r][c] = g[2*y0 + kOmin + r][2*x0 + kOmin + c]
// (all loads of a thread are issued before its first shared-memory store: the one-load-one-store loop spent 30 % of the
// kernel's samples on the store's long-scoreboard wait, profiles/r02/stalls_r02_tail_dx_kernel.txt)
__device__ __forceinline__ void load_g_halo(const float* __restrict__ gimg, int y0, int x0, int Hs, int Ws, float* gh) {
  constexpr int kIters = (kGhH * kGhW + 255) / 256;      // 256 threads per block
  float v[kIters];
#pragma unroll
  for (int k = 0; k < kIters; ++k) {
    const int i = threadIdx.x + k * 256;
    const int r = i / kGhW, c = i - r * kGhW;
    const int y = 2 * y0 + kOmin + r, x = 2 * x0 + kOmin + c;
    v[k] = (i < kGhH * kGhW && y >= 0 && y < Hs && x >= 0 && x < Ws) ? __ldg(gimg + static_cast<size_t>(y) * Ws + x) : 0.f;
  }
#pragma unroll
  for (int k = 0; k < kIters; ++k) {
    const int i = threadIdx.x + k * 256;
    if (i < kGhH * kGhW) gh[(i / kGhW) * kGhPitch + (i % kGhW)] = v[k];
  }
}

}  // namespace

// ------------------------------------------------------------------------------------------------
// Tables.  U[cls][o][ci] fp32 for the 16 border classes (class 0 = interior) and UT[ci][o] bf16 = class 0 as the
// K-major B operand of tail_dx_kernel.  One thread per (cls, o, ci).
__global__ void __launch_bounds__(256) tail_tables_kernel(const float* __restrict__ W2, const float* __restrict__ w3,
                                                          float* __restrict__ U, __nv_bfloat16* __restrict__ UT) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= 16 * kNO * 64) return;
  const int ci = i & 63, o = (i >> 6) & 63, cls = i >> 12;
  const int oy = o / kO + kOmin, ox = o % kO + kOmin;
  float sum = 0.f;
  for (int tpy = -1; tpy <= 1; ++tpy) {
    if ((tpy == 1 && (cls & 1)) || (tpy == -1 && (cls & 2))) continue;      // z - t' leaves the image
    const int sy = oy + 2 * tpy;                                            // = qy - ty
    for (int tpx = -1; tpx <= 1; ++tpx) {
      if ((tpx == 1 && (cls & 4)) || (tpx == -1 && (cls & 8))) continue;
      const int sx = ox + 2 * tpx;
      const int tp = (tpy + 1) * 3 + (tpx + 1);
      for (int ty = -1; ty <= 1; ++ty) {
        const int qy = sy + ty;
        if (qy < 0 || qy > 1) continue;
        for (int tx = -1; tx <= 1; ++tx) {
          const int qx = sx + tx;
          if (qx < 0 || qx > 1) continue;
          const int q = qy * 2 + qx, t = (ty + 1) * 3 + (tx + 1);
          float s = 0.f;
          for (int c = 0; c < 64; ++c) s = fmaf(__ldg(W2 + (static_cast<size_t>(c * 4 + q) * 64 + ci) * 9 + tp), __ldg(w3 + c * 9 + t), s);
          sum += s;
        }
      }
    }
  }
  U[i] = sum;
  if (cls == 0) UT[ci * kNO + o] = __float2bfloat16(sum);
}

// ------------------------------------------------------------------------------------------------
// dX, all pixels with the interior table: D[pixel][ci] = sum_o g[2 z + o] * U0[o][ci] as mma.sync m16n8k16 products
// (A = the g stencil of 16 consecutive pixels of a row built from the shared halo, split into bf16 hi + lo so that
// an arbitrary fp32 loss gradient keeps ~16 mantissa bits; B = UT held in registers for the life of the block).
// Warp w of a block owns row y0 + w of an 8 x 16 pixel tile.  The bf16 result tile is staged in shared memory and
// written with 16-byte stores.  Border pixels are overwritten afterwards by tail_dx_edge_kernel.
// SIGN = true: the loss gradient is known to be  scale * {-1, 0, +1}  (the fused L1 path: dout = w_k sign(out - target),
// all three lists of a stage share w_k): its sign is exact in bf16, so the lo half of the split and half of the MMAs go.
template <bool SIGN>
__global__ void __launch_bounds__(256, 2) tail_dx_kernel(const float* __restrict__ g, const __nv_bfloat16* __restrict__ UT,
                                                         __nv_bfloat16* __restrict__ dX, int n_img, int H1, int W1,
                                                         int tiles_x, int tiles_y, float scale) {
  __shared__ __align__(16) float gh[kGhH * kGhPitch];
  __shared__ __align__(16) uint8_t outt[kTH * kTW * kXPitch];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int Hs = 2 * H1, Ws = 2 * W1;
  // B fragments: b0 = UT[n = 8 nt + lane / 4][k = 16 ks + 2 (lane % 4) + {0, 1}], b1 = same with k + 8
  uint32_t bfrag[4][8][2];
#pragma unroll
  for (int ks = 0; ks < 4; ++ks)
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      const uint32_t* row = reinterpret_cast<const uint32_t*>(UT + (nt * 8 + (lane >> 2)) * kNO + ks * 16 + (lane & 3) * 2);
      bfrag[ks][nt][0] = __ldg(row);
      bfrag[ks][nt][1] = __ldg(row + 4);
    }
  const int total = n_img * tiles_y * tiles_x;
  for (int t = blockIdx.x; t < total; t += gridDim.x) {
    int tt = t;
    const int tx = tt % tiles_x;
    tt /= tiles_x;
    const int ty = tt % tiles_y;
    const int img = tt / tiles_y;
    const int y0 = ty * kTH, x0 = tx * kTW;
    load_g_halo(g + static_cast<size_t>(img) * Hs * Ws, y0, x0, Hs, Ws, gh);
    __syncthreads();
    float acc[8][4];
#pragma unroll
    for (int nt = 0; nt < 8; ++nt)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[nt][j] = 0.f;
    const int r0 = lane >> 2, c0 = (lane & 3) * 2;
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
      // k = 16 ks + col: col < 8 -> (oy index 2 ks, ox index col), col >= 8 -> (2 ks + 1, col - 8)
      uint32_t ahi[4], alo[4];
#pragma unroll
      for (int f = 0; f < 4; ++f) {
        const int row = r0 + (f & 1) * 8;          // pixel x0 + row of tile row `warp`
        const int oyi = 2 * ks + (f >> 1);
        float2 v = *reinterpret_cast<const float2*>(gh + (2 * warp + oyi) * kGhPitch + 2 * row + c0);
        if (SIGN) {
          v.x = v.x > 0.f ? 1.f : (v.x < 0.f ? -1.f : 0.f);
          v.y = v.y > 0.f ? 1.f : (v.y < 0.f ? -1.f : 0.f);
        }
        const uint32_t hi = pack_bf2(v.x, v.y);
        ahi[f] = hi;
        if (!SIGN) alo[f] = pack_bf2(v.x - bf_lo(hi), v.y - bf_hi(hi));
      }
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) {
        mma16816(acc[nt], ahi, bfrag[ks][nt][0], bfrag[ks][nt][1]);
        if (!SIGN) mma16816(acc[nt], alo, bfrag[ks][nt][0], bfrag[ks][nt][1]);
      }
    }
    if (SIGN) {
#pragma unroll
      for (int nt = 0; nt < 8; ++nt)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[nt][j] *= scale;
    }
    // stage: pixel (warp, r0 / r0 + 8), channels 8 nt + c0, c0 + 1
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      *reinterpret_cast<uint32_t*>(outt + (warp * kTW + r0) * kXPitch + (nt * 8 + c0) * 2) = pack_bf2(acc[nt][0], acc[nt][1]);
      *reinterpret_cast<uint32_t*>(outt + (warp * kTW + r0 + 8) * kXPitch + (nt * 8 + c0) * 2) = pack_bf2(acc[nt][2], acc[nt][3]);
    }
    __syncthreads();
    __nv_bfloat16* dimg = dX + static_cast<size_t>(img) * H1 * W1 * 64;
#pragma unroll
    for (int it = 0; it < kTH * kTW * 8 / 256; ++it) {
      const int i = threadIdx.x + it * 256;
      const int px = i >> 3, ck = i & 7;
      const int y = y0 + px / kTW, x = x0 + px % kTW;
      if (y < H1 && x < W1)
        *reinterpret_cast<uint4*>(dimg + (static_cast<size_t>(y) * W1 + x) * 64 + ck * 8) =
            *reinterpret_cast<const uint4*>(outt + px * kXPitch + ck * 16);
    }
    __syncthreads();     // gh / outt are overwritten by the next tile
  }
}

// Border pixels (first / last row or column) again, exactly, with the table of their class (fp32 SIMT: they are
// 2 (H1 + W1) - 4 of H1 * W1 pixels).  Thread = (border pixel, channel).
__device__ __forceinline__ bool border_pixel(int e, int H1, int W1, int* zy, int* zx) {
  // enumeration: top row, bottom row (H1 > 1), left column without corners, right column without corners (W1 > 1)
  if (e < W1) { *zy = 0; *zx = e; return true; }
  e -= W1;
  if (H1 > 1) {
    if (e < W1) { *zy = H1 - 1; *zx = e; return true; }
    e -= W1;
  }
  const int inner = H1 - 2 > 0 ? H1 - 2 : 0;
  if (e < inner) { *zy = 1 + e; *zx = 0; return true; }
  e -= inner;
  if (W1 > 1 && e < inner) { *zy = 1 + e; *zx = W1 - 1; return true; }
  return false;
}
__host__ __device__ inline int border_count(int H1, int W1) {
  const int inner = H1 - 2 > 0 ? H1 - 2 : 0;
  return W1 + (H1 > 1 ? W1 : 0) + inner + (W1 > 1 ? inner : 0);
}

__global__ void __launch_bounds__(256) tail_dx_edge_kernel(const float* __restrict__ g, const float* __restrict__ U,
                                                           __nv_bfloat16* __restrict__ dX, long long n_img, int H1,
                                                           int W1) {
  const int per_img = border_count(H1, W1);
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const int ci = static_cast<int>(i & 63);
  const long long pe = i >> 6;
  if (pe >= n_img * per_img) return;
  const long long img = pe / per_img;
  int zy, zx;
  if (!border_pixel(static_cast<int>(pe - img * per_img), H1, W1, &zy, &zx)) return;
  const int Hs = 2 * H1, Ws = 2 * W1;
  const float* gimg = g + static_cast<size_t>(img) * Hs * Ws;
  const float* Uc = U + static_cast<size_t>(border_class(zy, zx, H1, W1)) * kNO * 64 + ci;
  float sum = 0.f;
#pragma unroll 1
  for (int oyi = 0; oyi < kO; ++oyi) {
    const int y = 2 * zy + kOmin + oyi;
    if (y < 0 || y >= Hs) continue;
#pragma unroll
    for (int oxi = 0; oxi < kO; ++oxi) {
      const int x = 2 * zx + kOmin + oxi;
      if (x < 0 || x >= Ws) continue;
      sum = fmaf(__ldg(gimg + static_cast<size_t>(y) * Ws + x), __ldg(Uc + (oyi * kO + oxi) * 64), sum);
    }
  }
  dX[(static_cast<size_t>(img) * H1 * W1 + static_cast<size_t>(zy) * W1 + zx) * 64 + ci] = __float2bfloat16(sum);
}

// Fast form for images of at least 4 x 4 pixels: the four border runs WITHOUT their corners (top / bottom row, left /
// right column) have one border class each, so a thread (channel ci, 8 pixels of a 16-pixel run segment) reads each
// table entry U[cls][o][ci] once for its 8 pixels; the g stencils of the segment are staged in shared memory.  The first
// form above re-read the 16 KB class table per pixel (L1-bound: 0.1 ms per launch, as long as tail_dx_kernel itself);
// it still serves the 4 corners (grid = corners only) and small images.
constexpr int kRunSeg = 16;
__global__ void __launch_bounds__(128) tail_dx_runs_kernel(const float* __restrict__ g, const float* __restrict__ U,
                                                           __nv_bfloat16* __restrict__ dX, int H1, int W1, int segs_row,
                                                           int segs_col) {
  __shared__ float gp[kRunSeg][kNO];
  const int img = blockIdx.y;
  int b = blockIdx.x;
  int run, seg;                       // run 0 top row, 1 bottom row, 2 left column, 3 right column
  if (b < 2 * segs_row) { run = b / segs_row; seg = b - run * segs_row; }
  else { b -= 2 * segs_row; run = 2 + b / segs_col; seg = b - (run - 2) * segs_col; }
  const int len = run < 2 ? W1 - 2 : H1 - 2;
  const int p0 = 1 + seg * kRunSeg;                      // first pixel of the segment along the run (corners excluded)
  const int n = len - seg * kRunSeg < kRunSeg ? len - seg * kRunSeg : kRunSeg;
  const int Hs = 2 * H1, Ws = 2 * W1;
  const float* gimg = g + static_cast<size_t>(img) * Hs * Ws;
  auto pix = [&](int i, int* zy, int* zx) {
    if (run == 0) { *zy = 0; *zx = p0 + i; }
    else if (run == 1) { *zy = H1 - 1; *zx = p0 + i; }
    else if (run == 2) { *zy = p0 + i; *zx = 0; }
    else { *zy = p0 + i; *zx = W1 - 1; }
  };
  for (int i = threadIdx.x; i < kRunSeg * kNO; i += 128) {
    const int pi = i >> 6, o = i & 63;
    float v = 0.f;
    if (pi < n) {
      int zy, zx;
      pix(pi, &zy, &zx);
      const int y = 2 * zy + kOmin + (o >> 3), x = 2 * zx + kOmin + (o & 7);
      if (y >= 0 && y < Hs && x >= 0 && x < Ws) v = __ldg(gimg + static_cast<size_t>(y) * Ws + x);
    }
    gp[pi][o] = v;
  }
  __syncthreads();
  const int ci = threadIdx.x & 63, pg = threadIdx.x >> 6;     // pixels [8 pg, 8 pg + 8) of the segment
  int zy0, zx0;
  pix(0, &zy0, &zx0);
  const float* Uc = U + static_cast<size_t>(border_class(zy0, zx0, H1, W1)) * kNO * 64 + ci;
  float acc[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) acc[k] = 0.f;
#pragma unroll 4
  for (int o = 0; o < kNO; ++o) {
    const float u = __ldg(Uc + o * 64);
#pragma unroll
    for (int k = 0; k < 8; ++k) acc[k] = fmaf(gp[pg * 8 + k][o], u, acc[k]);
  }
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const int pi = pg * 8 + k;
    if (pi < n) {
      int zy, zx;
      pix(pi, &zy, &zx);
      dX[(static_cast<size_t>(img) * H1 * W1 + static_cast<size_t>(zy) * W1 + zx) * 64 + ci] = __float2bfloat16(acc[k]);
    }
  }
}

// The four corners of every image (first form's arithmetic).  Thread = (image, corner, channel).
__global__ void __launch_bounds__(256) tail_dx_corners_kernel(const float* __restrict__ g, const float* __restrict__ U,
                                                              __nv_bfloat16* __restrict__ dX, long long n_img, int H1, int W1) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n_img * 4 * 64) return;
  const int ci = static_cast<int>(i & 63), cor = static_cast<int>((i >> 6) & 3);
  const long long img = i >> 8;
  const int zy = cor < 2 ? 0 : H1 - 1, zx = (cor & 1) ? W1 - 1 : 0;
  const int Hs = 2 * H1, Ws = 2 * W1;
  const float* gimg = g + static_cast<size_t>(img) * Hs * Ws;
  const float* Uc = U + static_cast<size_t>(border_class(zy, zx, H1, W1)) * kNO * 64 + ci;
  float sum = 0.f;
  for (int o = 0; o < kNO; ++o) {
    const int y = 2 * zy + kOmin + (o >> 3), x = 2 * zx + kOmin + (o & 7);
    if (y >= 0 && y < Hs && x >= 0 && x < Ws) sum = fmaf(__ldg(gimg + static_cast<size_t>(y) * Ws + x), __ldg(Uc + o * 64), sum);
  }
  dX[(static_cast<size_t>(img) * H1 * W1 + static_cast<size_t>(zy) * W1 + zx) * 64 + ci] = __float2bfloat16(sum);
}

// ------------------------------------------------------------------------------------------------
// S[o][ci] += sum_z g[2 z + o] * X[z][ci]  and  Gs[o] += sum_z g[2 z + o]  over all pixels of all images: an
// [64 offsets x pixels] x [pixels x 64 channels] product on mma.sync (K = pixels).  Persistent blocks walk 8 x 16 pixel
// tiles; the X tile is staged with cp.async (144-byte pixel pitch), the B operand comes from it via ldmatrix.trans, the A
// operand (g stencil, bf16 hi + lo) is built from the shared halo.  Warp w accumulates offsets [16 (w % 4), +16) x
// channels [32 (w / 4), +32) across all tiles of the block; one atomic pass at the end.
template <bool SIGN>
__global__ void __launch_bounds__(256, 2) tail_corr_kernel(const float* __restrict__ g, const __nv_bfloat16* __restrict__ X,
                                                           float* __restrict__ S, float* __restrict__ Gs, int n_img,
                                                           int H1, int W1, int tiles_x, int tiles_y, float scale) {
  __shared__ float gh[kGhH * kGhPitch];
  __shared__ __align__(16) uint8_t xt[kTH * kTW * kXPitch];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int mt = warp & 3, nh = warp >> 2;
  const int Hs = 2 * H1, Ws = 2 * W1;
  const uint32_t xt_s = static_cast<uint32_t>(__cvta_generic_to_shared(xt));
  float acc[4][4];
#pragma unroll
  for (int nt = 0; nt < 4; ++nt)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[nt][j] = 0.f;
  float gsum = 0.f;                  // threads 0..63: offset o = threadIdx.x
  const int total = n_img * tiles_y * tiles_x;
  for (int t = blockIdx.x; t < total; t += gridDim.x) {
    int tt = t;
    const int tx = tt % tiles_x;
    tt /= tiles_x;
    const int ty = tt % tiles_y;
    const int img = tt / tiles_y;
    const int y0 = ty * kTH, x0 = tx * kTW;
    const __nv_bfloat16* src = X + static_cast<size_t>(img) * H1 * W1 * 64;
#pragma unroll
    for (int it = 0; it < kTH * kTW * 8 / 256; ++it) {
      const int i = threadIdx.x + it * 256;
      const int px = i >> 3, ck = i & 7;
      const int y = y0 + px / kTW, x = x0 + px % kTW;
      const bool ok = y < H1 && x < W1;
      const __nv_bfloat16* p = src + (static_cast<size_t>(ok ? y : 0) * W1 + (ok ? x : 0)) * 64 + ck * 8;
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(xt_s + px * kXPitch + ck * 16), "l"(p),
                   "r"(ok ? 16 : 0)
                   : "memory");
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
    load_g_halo(g + static_cast<size_t>(img) * Hs * Ws, y0, x0, Hs, Ws, gh);
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();
    if (threadIdx.x < kNO) {
      const int oyi = threadIdx.x / kO, oxi = threadIdx.x % kO;
      const int ny = H1 - y0 < kTH ? H1 - y0 : kTH, nx = W1 - x0 < kTW ? W1 - x0 : kTW;
      float s = 0.f;
      for (int y = 0; y < ny; ++y)
        for (int x = 0; x < nx; ++x) s += gh[(2 * y + oyi) * kGhPitch + 2 * x + oxi];
      gsum += s;
    }
    const int r0 = lane >> 2, c0 = (lane & 3) * 2;
#pragma unroll
    for (int ks = 0; ks < kTH; ++ks) {           // k step = the 16 pixels of tile row ks
      // A (row-major, rows = offsets 16 mt + {r0, r0 + 8} = (oy index 2 mt / 2 mt + 1, ox index r0), columns = pixels)
      uint32_t ahi[4], alo[4];
#pragma unroll
      for (int f = 0; f < 4; ++f) {
        const int oyi = 2 * mt + (f & 1);
        const int zx = c0 + (f >> 1) * 8;
        const float* row = gh + (2 * ks + oyi) * kGhPitch + 2 * zx + r0;
        float v0 = row[0], v1 = row[2];
        if (SIGN) {
          v0 = v0 > 0.f ? 1.f : (v0 < 0.f ? -1.f : 0.f);
          v1 = v1 > 0.f ? 1.f : (v1 < 0.f ? -1.f : 0.f);
        }
        const uint32_t hi = pack_bf2(v0, v1);
        ahi[f] = hi;
        if (!SIGN) alo[f] = pack_bf2(v0 - bf_lo(hi), v1 - bf_hi(hi));
      }
#pragma unroll
      for (int np = 0; np < 2; ++np) {
        uint32_t b0, b1, b2, b3;
        const uint32_t addr = xt_s + (ks * kTW + (lane & 7) + 8 * ((lane >> 3) & 1)) * kXPitch +
                              (nh * 32 + 16 * np + 8 * (lane >> 4)) * 2;
        asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];"
                     : "=r"(b0), "=r"(b1), "=r"(b2), "=r"(b3)
                     : "r"(addr));
        mma16816(acc[2 * np], ahi, b0, b1);
        if (!SIGN) mma16816(acc[2 * np], alo, b0, b1);
        mma16816(acc[2 * np + 1], ahi, b2, b3);
        if (!SIGN) mma16816(acc[2 * np + 1], alo, b2, b3);
      }
    }
    __syncthreads();     // the tile and the halo are overwritten by the next iteration
  }
  if (SIGN) {
#pragma unroll
    for (int nt = 0; nt < 4; ++nt)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[nt][j] *= scale;
  }
#pragma unroll
  for (int nt = 0; nt < 4; ++nt) {
    const int o = mt * 16 + (lane >> 2), ci = nh * 32 + nt * 8 + (lane & 3) * 2;
    atomicAdd(S + o * 64 + ci, acc[nt][0]);
    atomicAdd(S + o * 64 + ci + 1, acc[nt][1]);
    atomicAdd(S + (o + 8) * 64 + ci, acc[nt][2]);
    atomicAdd(S + (o + 8) * 64 + ci + 1, acc[nt][3]);
  }
  if (threadIdx.x < kNO) atomicAdd(Gs + threadIdx.x, gsum);
}

// Edge sums E[set][o][ci]: the same correlation restricted to one edge set of pixels.  set 0 first row, 1 last row,
// 2 first column, 3 last column, 4..7 the corners (0,0), (0,W1-1), (H1-1,0), (H1-1,W1-1).  blockIdx.y = set,
// blockIdx.x walks chunks of `chunk` pixels of that set over all images; per sub-chunk of 32 pixels the g stencils and
// the X rows are staged in shared memory, then thread (channel, group of 16 offsets) accumulates from there
// (the first form looped over global loads per pixel: 0.2 ms per launch, more than the main correlation kernel).
constexpr int kEdgeSub = 32;
__global__ void __launch_bounds__(256) tail_corr_edge_kernel(const float* __restrict__ g, const __nv_bfloat16* __restrict__ X,
                                                             float* __restrict__ E, long long n_img, int H1, int W1,
                                                             int chunk) {
  __shared__ float gp[kEdgeSub][kNO];
  __shared__ float xs[kEdgeSub][64];
  const int set = blockIdx.y;
  const int per_img = set < 2 ? W1 : (set < 4 ? H1 : 1);
  const long long n_pix = n_img * per_img;
  const long long p0 = static_cast<long long>(blockIdx.x) * chunk;
  if (p0 >= n_pix) return;
  const long long p1 = p0 + chunk < n_pix ? p0 + chunk : n_pix;
  const int ci = threadIdx.x & 63, og = threadIdx.x >> 6;      // offsets [16 og, 16 og + 16)
  const int Hs = 2 * H1, Ws = 2 * W1;
  float acc[16];
#pragma unroll
  for (int k = 0; k < 16; ++k) acc[k] = 0.f;
  for (long long q0 = p0; q0 < p1; q0 += kEdgeSub) {
    const int n = static_cast<int>(p1 - q0 < kEdgeSub ? p1 - q0 : kEdgeSub);
    for (int i = threadIdx.x; i < kEdgeSub * kNO; i += 256) {
      const int pi = i >> 6, c = i & 63;
      float gv = 0.f, xv = 0.f;
      if (pi < n) {
        const long long p = q0 + pi;
        const long long img = p / per_img;
        const int e = static_cast<int>(p - img * per_img);
        int zy, zx;
        switch (set) {
          case 0: zy = 0; zx = e; break;
          case 1: zy = H1 - 1; zx = e; break;
          case 2: zy = e; zx = 0; break;
          case 3: zy = e; zx = W1 - 1; break;
          case 4: zy = 0; zx = 0; break;
          case 5: zy = 0; zx = W1 - 1; break;
          case 6: zy = H1 - 1; zx = 0; break;
          default: zy = H1 - 1; zx = W1 - 1; break;
        }
        const int yy = 2 * zy + kOmin + (c >> 3), xx = 2 * zx + kOmin + (c & 7);
        if (yy >= 0 && yy < Hs && xx >= 0 && xx < Ws) gv = __ldg(g + static_cast<size_t>(img) * Hs * Ws + static_cast<size_t>(yy) * Ws + xx);
        xv = __bfloat162float(X[(static_cast<size_t>(img) * H1 * W1 + static_cast<size_t>(zy) * W1 + zx) * 64 + c]);
      }
      gp[pi][c] = gv;
      xs[pi][c] = xv;
    }
    __syncthreads();
    for (int pi = 0; pi < n; ++pi) {
      const float x = xs[pi][ci];
#pragma unroll
      for (int k = 0; k < 16; ++k) acc[k] = fmaf(gp[pi][og * 16 + k], x, acc[k]);
    }
    __syncthreads();
  }
  float* dst = E + (static_cast<size_t>(set) * kNO + og * 16) * 64 + ci;
#pragma unroll
  for (int k = 0; k < 16; ++k) atomicAdd(dst + k * 64, acc[k]);
}

// ------------------------------------------------------------------------------------------------
// Parameter gradients from S, E, Gs (see the file header).  Three thread ranges:
//   [0, 256*64*9)                 dW2[(c,q)][ci][t'] += sum_t w3[c][t] R[q,t,t'][ci]
//   next 64*9*4*9                 dw3[c][t] += sum_ci W2[(c,q)][ci][t'] R[q,t,t'][ci]        (atomic over q, t')
//   next 256 + 576 + 1            db2, the bias part of dw3, db3
__device__ __forceinline__ float tail_R(const float* __restrict__ S, const float* __restrict__ E, int q, int t, int tp,
                                        int ci) {
  const int tpy = tp / 3 - 1, tpx = tp % 3 - 1, ty = t / 3 - 1, tx = t % 3 - 1, qy = q >> 1, qx = q & 1;
  const int o = (-2 * tpy + qy - ty - kOmin) * kO + (-2 * tpx + qx - tx - kOmin);
  float v = S[o * 64 + ci];
  const int rset = tpy == 1 ? 0 : 1, cset = tpx == 1 ? 2 : 3;
  if (tpy != 0) v -= E[(rset * kNO + o) * 64 + ci];
  if (tpx != 0) v -= E[(cset * kNO + o) * 64 + ci];
  if (tpy != 0 && tpx != 0) {
    const int cor = 4 + (tpy == 1 ? 0 : 2) + (tpx == 1 ? 0 : 1);
    v += E[(cor * kNO + o) * 64 + ci];
  }
  return v;
}
__device__ __forceinline__ float tail_Gq(const float* __restrict__ Gs, int q, int t) {
  const int ty = t / 3 - 1, tx = t % 3 - 1, qy = q >> 1, qx = q & 1;
  return Gs[(qy - ty - kOmin) * kO + (qx - tx - kOmin)];
}

__global__ void __launch_bounds__(256) tail_finish_kernel(const float* __restrict__ S, const float* __restrict__ E,
                                                          const float* __restrict__ Gs, const float* __restrict__ W2,
                                                          const float* __restrict__ b2, const float* __restrict__ w3,
                                                          float* __restrict__ dW2, float* __restrict__ db2,
                                                          float* __restrict__ dw3, float* __restrict__ db3) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  constexpr int n1 = 256 * 64 * 9, n2 = 64 * 9 * 4 * 9, n3 = 256 + 576 + 1;
  if (i < n1) {
    if (!dW2) return;
    const int tp = i % 9, ci = (i / 9) & 63, co = i / (9 * 64);
    const int c = co >> 2, q = co & 3;
    float s = 0.f;
#pragma unroll
    for (int t = 0; t < 9; ++t) s = fmaf(w3[c * 9 + t], tail_R(S, E, q, t, tp, ci), s);
    dW2[i] += s;
    return;
  }
  i -= n1;
  if (i < n2) {
    if (!dw3) return;
    const int tp = i % 9, q = (i / 9) & 3, t = (i / 36) % 9, c = i / 324;
    const float* w = W2 + static_cast<size_t>(c * 4 + q) * 64 * 9 + tp;
    float s = 0.f;
    for (int ci = 0; ci < 64; ++ci) s = fmaf(w[ci * 9], tail_R(S, E, q, t, tp, ci), s);
    atomicAdd(dw3 + c * 9 + t, s);
    return;
  }
  i -= n2;
  if (i >= n3) return;
  if (i < 256) {
    if (!db2) return;
    const int c = i >> 2, q = i & 3;
    float s = 0.f;
#pragma unroll
    for (int t = 0; t < 9; ++t) s = fmaf(w3[c * 9 + t], tail_Gq(Gs, q, t), s);
    db2[i] += s;
  } else if (i < 256 + 576) {
    if (!dw3) return;
    const int k = i - 256, c = k / 9, t = k % 9;
    float s = 0.f;
#pragma unroll
    for (int q = 0; q < 4; ++q) s = fmaf(b2[c * 4 + q], tail_Gq(Gs, q, t), s);
    atomicAdd(dw3 + k, s);
  } else if (db3) {
    float s = 0.f;
#pragma unroll
    for (int q = 0; q < 4; ++q) s += tail_Gq(Gs, q, 4);     // the four HR parity classes partition the image
    db3[0] += s;
  }
}

// ================================================================================================
// FORWARD of the same tail as ONE composite convolution.  Both convs are linear with nothing between them but the
// pixel shuffle, and the second has a single output channel, so for every X pixel z and sub-pixel q in {0,1}^2
//     out[2z + q] = b3 + biasT_cls(z)[q] + sum_{d in [-2,2]^2} sum_ci Kf_cls(z)[q][d][ci] * X[z + d][ci]
//     Kf_cls[q][d][ci] = sum over taps t of the last conv with 2z + q + t inside the HR image, d1 = floor((q + t) / 2),
//                        q' = (q + t) mod 2, t' = d - d1 in [-1,1]^2  of  sum_c w3[c][t] W2[(c,q')][ci][t']
// (zero padding of the 64-channel HR map = the "inside" condition = the border class of z; zero padding of X is the
// zero fill of the halo tile).  A 64 -> 4 channel 5x5 convolution: 12.8 kFLOP per X pixel instead of 295 k + 4.6 k,
// and the 64-channel HR map (128 B per HR pixel written by the shuffle conv, read by the last conv) never exists.
// N = 4 is far too narrow for tcgen05 (N >= 16 at M = 128); warp-level mma.sync m16n8k16 with the upper half of the
// n8 tile zero does it at ~25.6 kFLOP executed per pixel.  Interior table on the tensor cores for every pixel, border
// pixels recomputed exactly afterwards with their class tables (SIMT, ~3 % of the pixels).
namespace {
constexpr int kFT = 16;                         // tile = 16 x 16 X pixels
constexpr int kFH = kFT + 4;                    // halo tile 20 x 20
constexpr int kFK = 25 * 64;                    // K of the composite conv
constexpr int kFBPitch = kFK + 8;               // bf16 elements per table row in shared memory (conflict-free)
constexpr int kFSmem = kFH * kFH * kXPitch + 4 * kFBPitch * 2;
}  // namespace

// Kf fp32 [16 cls][4 q][25 d][64 ci], KB bf16 [4 q][25*64] (class 0), biasT fp32 [16][4] (b3 included).
__global__ void __launch_bounds__(256) tail_fwd_tables_kernel(const float* __restrict__ W2, const float* __restrict__ b2,
                                                              const float* __restrict__ w3, const float* __restrict__ b3,
                                                              float* __restrict__ Kf, __nv_bfloat16* __restrict__ KB,
                                                              float* __restrict__ biasT) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < 16 * 4) {
    const int cls = i >> 2, q = i & 3, qy = q >> 1, qx = q & 1;
    float s = b3[0];
    for (int ty = -1; ty <= 1; ++ty) {
      const int d1y = (qy + ty + 2) / 2 - 1, qpy = (qy + ty + 2) & 1;
      if ((d1y == -1 && (cls & 1)) || (d1y == 1 && (cls & 2))) continue;
      for (int tx = -1; tx <= 1; ++tx) {
        const int d1x = (qx + tx + 2) / 2 - 1, qpx = (qx + tx + 2) & 1;
        if ((d1x == -1 && (cls & 4)) || (d1x == 1 && (cls & 8))) continue;
        const int t = (ty + 1) * 3 + (tx + 1), qp = qpy * 2 + qpx;
        for (int c = 0; c < 64; ++c) s = fmaf(w3[c * 9 + t], b2[c * 4 + qp], s);
      }
    }
    biasT[i] = s;
  }
  if (i >= 16 * 4 * 25 * 64) return;
  const int ci = i & 63, d = (i >> 6) % 25, q = (i / (64 * 25)) & 3, cls = i / (64 * 25 * 4);
  const int dy = d / 5 - 2, dx = d % 5 - 2, qy = q >> 1, qx = q & 1;
  float sum = 0.f;
  for (int ty = -1; ty <= 1; ++ty) {
    const int d1y = (qy + ty + 2) / 2 - 1, qpy = (qy + ty + 2) & 1;
    if ((d1y == -1 && (cls & 1)) || (d1y == 1 && (cls & 2))) continue;
    const int tpy = dy - d1y;
    if (tpy < -1 || tpy > 1) continue;
    for (int tx = -1; tx <= 1; ++tx) {
      const int d1x = (qx + tx + 2) / 2 - 1, qpx = (qx + tx + 2) & 1;
      if ((d1x == -1 && (cls & 4)) || (d1x == 1 && (cls & 8))) continue;
      const int tpx = dx - d1x;
      if (tpx < -1 || tpx > 1) continue;
      const int t = (ty + 1) * 3 + (tx + 1), qp = qpy * 2 + qpx, tp = (tpy + 1) * 3 + (tpx + 1);
      float s = 0.f;
      for (int c = 0; c < 64; ++c) s = fmaf(__ldg(w3 + c * 9 + t), __ldg(W2 + (static_cast<size_t>(c * 4 + qp) * 64 + ci) * 9 + tp), s);
      sum += s;
    }
  }
  Kf[i] = sum;
  if (cls == 0) KB[q * kFK + d * 64 + ci] = __float2bfloat16(sum);
}

// out fp32 [n_img][2 H1][2 W1] for every X pixel with the interior table.  Block = 16 x 16 pixel tile, warp w = tile
// rows 2w, 2w + 1 (two m16 tiles); per (tap d, 16-channel step) one B fragment from the shared table serves both.
__global__ void __launch_bounds__(256, 2) tail_fwd_kernel(const __nv_bfloat16* __restrict__ X, const __nv_bfloat16* __restrict__ KB,
                                                          const float* __restrict__ biasT, float* __restrict__ out,
                                                          int n_img, int H1, int W1, int tiles_x, int tiles_y) {
  extern __shared__ __align__(16) uint8_t fsm[];
  uint8_t* xt = fsm;                                                 // [20 * 20][144 B]
  __nv_bfloat16* bs = reinterpret_cast<__nv_bfloat16*>(fsm + kFH * kFH * kXPitch);   // [4][kFBPitch]
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint32_t xt_s = static_cast<uint32_t>(__cvta_generic_to_shared(xt));
  for (int i = threadIdx.x; i < 4 * kFK / 8; i += 256) {             // 16-byte chunks of the table
    const int q = i / (kFK / 8), k8 = i - q * (kFK / 8);
    *reinterpret_cast<uint4*>(bs + q * kFBPitch + k8 * 8) = __ldg(reinterpret_cast<const uint4*>(KB + q * kFK + k8 * 8));
  }
  const float bias0 = biasT[(lane & 3) < 2 ? (lane & 3) * 2 : 0], bias1 = biasT[(lane & 3) < 2 ? (lane & 3) * 2 + 1 : 0];
  const int Hs = 2 * H1, Ws = 2 * W1;
  const int total = n_img * tiles_y * tiles_x;
  for (int t = blockIdx.x; t < total; t += gridDim.x) {
    int tt = t;
    const int tx = tt % tiles_x;
    tt /= tiles_x;
    const int ty = tt % tiles_y;
    const int img = tt / tiles_y;
    const int y0 = ty * kFT, x0 = tx * kFT;
    const __nv_bfloat16* src = X + static_cast<size_t>(img) * H1 * W1 * 64;
    for (int i = threadIdx.x; i < kFH * kFH * 8; i += 256) {
      const int hp = i >> 3, ck = i & 7;
      const int y = y0 - 2 + hp / kFH, x = x0 - 2 + hp % kFH;
      const bool ok = y >= 0 && y < H1 && x >= 0 && x < W1;
      const __nv_bfloat16* p = src + (static_cast<size_t>(ok ? y : 0) * W1 + (ok ? x : 0)) * 64 + ck * 8;
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(xt_s + hp * kXPitch + ck * 16), "l"(p),
                   "r"(ok ? 16 : 0)
                   : "memory");
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();

    float acc[2][4];
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[mt][j] = 0.f;
    const uint32_t* brow = reinterpret_cast<const uint32_t*>(bs + (lane >> 2 & 3) * kFBPitch) + (lane & 3);
    const bool bval = lane < 16;                                     // n8 columns 4..7 are zero
    const uint32_t a_lane = ((lane & 7) + 8 * ((lane >> 3) & 1)) * kXPitch + 16 * (lane >> 4);
#pragma unroll 1
    for (int d = 0; d < 25; ++d) {
      const int dy = d / 5, dx = d - dy * 5;
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) {
        const int kk = d * 64 + ks * 16;
        const uint32_t b0 = bval ? brow[kk / 2] : 0u, b1 = bval ? brow[kk / 2 + 4] : 0u;
#pragma unroll
        for (int mt = 0; mt < 2; ++mt) {
          uint32_t a[4];
          const uint32_t addr = xt_s + ((2 * warp + mt + dy) * kFH + dx) * kXPitch + a_lane + ks * 32;
          asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];"
                       : "=r"(a[0]), "=r"(a[1]), "=r"(a[2]), "=r"(a[3])
                       : "r"(addr));
          mma16816(acc[mt], a, b0, b1);
        }
      }
    }
    // columns 2 (lane % 4) + {0, 1} = sub-pixel (qy = lane % 4, qx = 0 / 1) for lane % 4 < 2
    if ((lane & 3) < 2) {
      const int qy = lane & 3;
      float* oimg = out + static_cast<size_t>(img) * Hs * Ws;
#pragma unroll
      for (int mt = 0; mt < 2; ++mt) {
        const int zy = y0 + 2 * warp + mt;
        if (zy >= H1) continue;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int zx = x0 + (lane >> 2) + 8 * h;
          if (zx < W1)
            *reinterpret_cast<float2*>(oimg + static_cast<size_t>(2 * zy + qy) * Ws + 2 * zx) =
                make_float2(acc[mt][2 * h] + bias0, acc[mt][2 * h + 1] + bias1);
        }
      }
    }
    __syncthreads();     // the tile is overwritten by the next iteration
  }
}

// Border pixels again with the table of their class: warp = one border pixel, lane = two channels.
__global__ void __launch_bounds__(256) tail_fwd_edge_kernel(const __nv_bfloat16* __restrict__ X, const float* __restrict__ Kf,
                                                            const float* __restrict__ biasT, float* __restrict__ out,
                                                            long long n_img, int H1, int W1) {
  const int per_img = border_count(H1, W1);
  const long long pe = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (pe >= n_img * per_img) return;
  const long long img = pe / per_img;
  int zy, zx;
  if (!border_pixel(static_cast<int>(pe - img * per_img), H1, W1, &zy, &zx)) return;
  const int cls = border_class(zy, zx, H1, W1);
  const __nv_bfloat16* ximg = X + static_cast<size_t>(img) * H1 * W1 * 64;
  const float* K = Kf + static_cast<size_t>(cls) * 4 * kFK + 2 * lane;
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll 1
  for (int d = 0; d < 25; ++d) {
    const int y = zy + d / 5 - 2, x = zx + d % 5 - 2;
    if (y < 0 || y >= H1 || x < 0 || x >= W1) continue;
    const __nv_bfloat162 xv = *reinterpret_cast<const __nv_bfloat162*>(ximg + (static_cast<size_t>(y) * W1 + x) * 64 + 2 * lane);
    const float2 xf = __bfloat1622float2(xv);
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const float2 kv = __ldg(reinterpret_cast<const float2*>(K + q * kFK + d * 64));
      acc[q] = fmaf(xf.x, kv.x, fmaf(xf.y, kv.y, acc[q]));
    }
  }
#pragma unroll
  for (int q = 0; q < 4; ++q)
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) acc[q] += __shfl_xor_sync(0xffffffffu, acc[q], s);
  if (lane < 4) {
    const int qy = lane >> 1, qx = lane & 1;
    const float v = lane == 0 ? acc[0] : (lane == 1 ? acc[1] : (lane == 2 ? acc[2] : acc[3]));
    out[(static_cast<size_t>(img) * 2 * H1 + 2 * zy + qy) * (2 * W1) + 2 * zx + qx] = v + biasT[cls * 4 + lane];
  }
}

// ================================================================================================
// FORWARD, 36-channel form (tcgen05).  The same composite as above, factored the other way round:
//     B[z'][(q',t)] = bias36[(q',t)] + sum_{t',ci} Wc[t'][(q',t)][ci] X[z'+t'][ci]       (a 3x3 conv 64 -> 36 channels)
//     Wc[t'][(q',t)][ci] = sum_c w3[c][t] W2[(c,q')][ci][t'],   bias36[(q',t)] = sum_c w3[c][t] b2[(c,q')]
//     out[p] = b3 + sum over taps t with p + t = 2 z' + q' inside the HR image of B[z'][(q',t)]
// B[z'][(q',t)] is what HR position 2z'+q' of the (never materialised) 64-channel map contributes to output pixel
// 2z'+q'-t.  The zero padding of that map is simply "positions outside the image contribute nothing", so there are NO
// border classes, and the conv is an ordinary N = 48 launch of conv3x3_halo_kernel (weights resident, 41.5 kFLOP per
// pixel on the tensor cores instead of the 25.6 k of the mma.sync form - but at tcgen05 rate, without re-reading the tile
// from shared memory per tap).  The 9-tap gather that follows reads 144 B and writes 16 B per input pixel.
__global__ void __launch_bounds__(256) tail36_weights_kernel(const float* __restrict__ W2, const float* __restrict__ b2,
                                                             const float* __restrict__ w3, __nv_bfloat16* __restrict__ Wc,
                                                             float* __restrict__ bias48) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < 48) {
    float s = 0.f;
    if (i < 36) {
      const int qp = i / 9, t = i % 9;
      for (int c = 0; c < 64; ++c) s = fmaf(w3[c * 9 + t], b2[c * 4 + qp], s);
    }
    bias48[i] = s;
  }
  if (i >= 9 * 48 * 64) return;
  const int ci = i & 63, n = (i >> 6) % 48, tp = i / (64 * 48);
  float s = 0.f;
  if (n < 36) {
    const int qp = n / 9, t = n % 9;
    for (int c = 0; c < 64; ++c) s = fmaf(__ldg(w3 + c * 9 + t), __ldg(W2 + (static_cast<size_t>(c * 4 + qp) * 64 + ci) * 9 + tp), s);
  }
  Wc[i] = __float2bfloat16(s);          // packed operand layout: [K block = tap t'][N = 48][64 channels]
}

// out[img][py][px] = b3 + sum_t B[z'][(q',t)],  (2 z' + q') = (py, px) + t inside the HR image.
// B is bf16 [n_img][H1][W1][48] (the fp32 accumulator rounded once, like the 64-channel HR map it replaces was).  A block
// stages the records of an 8 x 32 input-pixel tile + a one-pixel ring in shared memory with 16-byte loads (the first
// form gathered nine scattered 4-byte values per thread straight from global memory: 99 % L1 throughput, 0.66 ms per
// inference step under ncu), then every thread sums the nine values of four output pixels from there.
constexpr int kGTH = 8;                                  // input rows per tile
constexpr int kGRH = kGTH + 2;                           // staged record rows
constexpr int kGRec = 48;                                // values per record
__device__ __forceinline__ float rec_value(const __nv_bfloat16& v) { return __bfloat162float(v); }
__device__ __forceinline__ float rec_value(const float& v) { return v; }
// T = bf16 (inference plans: 96 B per record) or float (training plans: the loss gradient is sign(out - target), so the
// records keep the fp32 accumulator - 8 x 16 input pixels per tile to stay inside 48 KB of static shared memory).
template <typename T, int kGTW>
__global__ void __launch_bounds__(8 * kGTW) tail36_gather_kernel(const T* __restrict__ B, const float* __restrict__ b3,
                                                                 float* __restrict__ out, int H1, int W1, int tiles_x,
                                                                 int tiles_y) {
  constexpr int kGRW = kGTW + 2;
  constexpr int kVec = 16 / static_cast<int>(sizeof(T));                // values per 16-byte load
  constexpr int kChunks = kGRec / kVec;                                 // 16-byte loads per record
  constexpr int kThreads = 8 * kGTW;
  __shared__ __align__(16) T rec[kGRH * kGRW * kGRec];                  // 32 640 B (bf16, 32 wide) / 34 560 B (fp32, 16 wide)
  const int tx = blockIdx.x % tiles_x, ty = (blockIdx.x / tiles_x) % tiles_y;
  const long long img = blockIdx.x / (tiles_x * tiles_y);
  const int y0 = ty * kGTH, x0 = tx * kGTW;
  const T* Bimg = B + static_cast<size_t>(img) * H1 * W1 * kGRec;
  for (int i = threadIdx.x; i < kGRH * kGRW * kChunks; i += kThreads) {
    const int r = i / kChunks, ck = i - r * kChunks;
    const int zy = y0 - 1 + r / kGRW, zx = x0 - 1 + r % kGRW;
    uint4 v = make_uint4(0, 0, 0, 0);
    if (zy >= 0 && zy < H1 && zx >= 0 && zx < W1)
      v = __ldg(reinterpret_cast<const uint4*>(Bimg + (static_cast<size_t>(zy) * W1 + zx) * kGRec) + ck);
    reinterpret_cast<uint4*>(rec + r * kGRec)[ck] = v;
  }
  __syncthreads();
  const int Hs = 2 * H1, Ws = 2 * W1;
  const float bias = b3[0];
  float* oimg = out + static_cast<size_t>(img) * Hs * Ws;
  // 16 x 2 kGTW output pixels per tile, thread = (row ly, 4 consecutive columns)
  constexpr int kThreadsPerRow = kGTW / 2;
  const int ly = threadIdx.x / kThreadsPerRow, lx0 = (threadIdx.x % kThreadsPerRow) * 4;
  const int py = 2 * y0 + ly;
  if (py >= Hs) return;
  float o4[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int lx = lx0 + k;
    float sacc = bias;
#pragma unroll
    for (int tyy = -1; tyy <= 1; ++tyy) {
      const int y = ly + tyy;                       // tile-local HR row of the contributing position, -1 .. 16
      const int gy = py + tyy;
      if (gy < 0 || gy >= Hs) continue;
      const int ry = (y + 2) >> 1;                  // staged record row: input row y0 - 1 + ry
#pragma unroll
      for (int txx = -1; txx <= 1; ++txx) {
        const int x = lx + txx;
        const int gx = 2 * x0 + x;
        if (gx < 0 || gx >= Ws) continue;
        const int rx = (x + 2) >> 1;
        const int n = ((gy & 1) * 2 + (gx & 1)) * 9 + (tyy + 1) * 3 + (txx + 1);
        sacc += rec_value(rec[(ry * kGRW + rx) * kGRec + n]);
      }
    }
    o4[k] = sacc;
  }
  const int px = 2 * x0 + lx0;
  if (px + 3 < Ws && (Ws & 3) == 0) {
    *reinterpret_cast<float4*>(oimg + static_cast<size_t>(py) * Ws + px) = make_float4(o4[0], o4[1], o4[2], o4[3]);
  } else {
#pragma unroll
    for (int k = 0; k < 4; ++k)
      if (px + k < Ws) oimg[static_cast<size_t>(py) * Ws + px + k] = o4[k];
  }
}

// ------------------------------------------------------------------------------------------------ launchers
// Optional dynamic shared-memory padding of the two mma.sync kernels (PVSR_TAIL_EXCLUSIVE=1): 23 KB + 16 KB + the 193 KB
// of a wgrad_tc_kernel CTA do not fit one SM, so the kernels then never share an SM with a tcgen05 CTA.  Built while hunting
// a hang of the two-branch graph (it was not the cause: plan.cpp, launch of tail_corr); off by default.
constexpr int kTailExclusiveSmem = 16 * 1024;
static int tail_pad_bytes() {
  static int pad = -1;
  if (pad < 0) {
    const char* e = getenv("PVSR_TAIL_EXCLUSIVE");
    pad = (e && e[0] == '1') ? kTailExclusiveSmem : 0;
  }
  return pad;
}
// scratch layout (floats): U [16][64][64] | S [64][64] | E [8][64][64] | Gs [64] | UT bf16 [64][64]
constexpr size_t kTailU = 0, kTailS = kTailU + 16 * kNO * 64, kTailE = kTailS + kNO * 64, kTailG = kTailE + 8 * kNO * 64,
                 kTailUT = kTailG + kNO, kTailFloats = kTailUT + kNO * 64 / 2;

size_t tail_scratch_bytes() { return kTailFloats * sizeof(float); }

int launch_tail_tables(const float* W2, const float* w3, void* scratch, cudaStream_t s) {
  float* f = static_cast<float*>(scratch);
  tail_tables_kernel<<<16 * kNO * 64 / 256, 256, 0, s>>>(W2, w3, f + kTailU, reinterpret_cast<__nv_bfloat16*>(f + kTailUT));
  return static_cast<int>(cudaGetLastError());
}

int launch_tail_zero_sums(void* scratch, cudaStream_t s) {
  float* f = static_cast<float*>(scratch);
  return static_cast<int>(cudaMemsetAsync(f + kTailS, 0, (kTailUT - kTailS) * sizeof(float), s));
}

int launch_tail_dx(const float* g, const void* scratch, void* dx_bf16, long long n_img, int H1, int W1, int num_sms,
                   cudaStream_t s, float sign_scale) {
  if (n_img <= 0) return 0;
  const float* f = static_cast<const float*>(scratch);
  const int tiles_x = (W1 + kTW - 1) / kTW, tiles_y = (H1 + kTH - 1) / kTH;
  const long long total = n_img * tiles_x * tiles_y;
  if (total >= (1LL << 31)) return static_cast<int>(cudaErrorInvalidValue);
  const long long cap = 4LL * (num_sms > 0 ? num_sms : 148);
  const unsigned grid = static_cast<unsigned>(total < cap ? total : cap);
  if (sign_scale != 0.f)
    tail_dx_kernel<true><<<grid, 256, tail_pad_bytes(), s>>>(g, reinterpret_cast<const __nv_bfloat16*>(f + kTailUT),
                                                             static_cast<__nv_bfloat16*>(dx_bf16), static_cast<int>(n_img),
                                                             H1, W1, tiles_x, tiles_y, sign_scale);
  else
    tail_dx_kernel<false><<<grid, 256, tail_pad_bytes(), s>>>(g, reinterpret_cast<const __nv_bfloat16*>(f + kTailUT),
                                                              static_cast<__nv_bfloat16*>(dx_bf16), static_cast<int>(n_img),
                                                              H1, W1, tiles_x, tiles_y, 1.f);
  int e = static_cast<int>(cudaGetLastError());
  if (e) return e;
  if (H1 >= 4 && W1 >= 4 && n_img <= 65535) {
    const int segs_row = (W1 - 2 + kRunSeg - 1) / kRunSeg, segs_col = (H1 - 2 + kRunSeg - 1) / kRunSeg;
    dim3 grid(static_cast<unsigned>(2 * segs_row + 2 * segs_col), static_cast<unsigned>(n_img));
    tail_dx_runs_kernel<<<grid, 128, 0, s>>>(g, f + kTailU, static_cast<__nv_bfloat16*>(dx_bf16), H1, W1, segs_row, segs_col);
    e = static_cast<int>(cudaGetLastError());
    if (e) return e;
    tail_dx_corners_kernel<<<static_cast<unsigned>((n_img * 256 + 255) / 256), 256, 0, s>>>(
        g, f + kTailU, static_cast<__nv_bfloat16*>(dx_bf16), n_img, H1, W1);
    return static_cast<int>(cudaGetLastError());
  }
  const long long threads = n_img * border_count(H1, W1) * 64;
  tail_dx_edge_kernel<<<static_cast<unsigned>((threads + 255) / 256), 256, 0, s>>>(g, f + kTailU,
                                                                                  static_cast<__nv_bfloat16*>(dx_bf16),
                                                                                  n_img, H1, W1);
  return static_cast<int>(cudaGetLastError());
}

int launch_tail_corr(const float* g, const void* x_bf16, void* scratch, long long n_img, int H1, int W1, int num_sms,
                     cudaStream_t s, float sign_scale) {
  if (n_img <= 0) return 0;
  float* f = static_cast<float*>(scratch);
  const int tiles_x = (W1 + kTW - 1) / kTW, tiles_y = (H1 + kTH - 1) / kTH;
  const long long total = n_img * tiles_x * tiles_y;
  if (total >= (1LL << 31)) return static_cast<int>(cudaErrorInvalidValue);
  const long long cap = 2LL * (num_sms > 0 ? num_sms : 148);
  const unsigned nblk = static_cast<unsigned>(total < cap ? total : cap);
  if (sign_scale != 0.f)
    tail_corr_kernel<true><<<nblk, 256, tail_pad_bytes(), s>>>(g, static_cast<const __nv_bfloat16*>(x_bf16), f + kTailS,
                                                               f + kTailG, static_cast<int>(n_img), H1, W1, tiles_x, tiles_y,
                                                               sign_scale);
  else
    tail_corr_kernel<false><<<nblk, 256, tail_pad_bytes(), s>>>(g, static_cast<const __nv_bfloat16*>(x_bf16), f + kTailS,
                                                                f + kTailG, static_cast<int>(n_img), H1, W1, tiles_x, tiles_y,
                                                                1.f);
  int e = static_cast<int>(cudaGetLastError());
  if (e) return e;
  const int chunk = 256;
  const long long longest = n_img * (H1 > W1 ? H1 : W1);
  dim3 grid(static_cast<unsigned>((longest + chunk - 1) / chunk), 8);
  tail_corr_edge_kernel<<<grid, 256, 0, s>>>(g, static_cast<const __nv_bfloat16*>(x_bf16), f + kTailE, n_img, H1, W1, chunk);
  return static_cast<int>(cudaGetLastError());
}

int launch_tail_finish(const void* scratch, const float* W2, const float* b2, const float* w3, float* dW2, float* db2,
                       float* dw3, float* db3, cudaStream_t s) {
  const float* f = static_cast<const float*>(scratch);
  constexpr int n = 256 * 64 * 9 + 64 * 9 * 4 * 9 + 256 + 576 + 1;
  tail_finish_kernel<<<(n + 255) / 256, 256, 0, s>>>(f + kTailS, f + kTailE, f + kTailG, W2, b2, w3, dW2, db2, dw3, db3);
  return static_cast<int>(cudaGetLastError());
}

// ---- forward: tables live in the packed-parameter buffer (they change only when the weights do)
// layout (bytes): Kf fp32 [16][4][1600] | biasT fp32 [16][4] | KB bf16 [4][1600]
constexpr size_t kTailFwdKf = 0, kTailFwdBias = kTailFwdKf + 16 * 4 * kFK * 4, kTailFwdKB = kTailFwdBias + 16 * 4 * 4,
                 kTailFwdBytes = kTailFwdKB + 4 * kFK * 2;

size_t tail_fwd_table_bytes() { return kTailFwdBytes; }

int launch_tail_fwd_tables(const float* W2, const float* b2, const float* w3, const float* b3, void* tables,
                           cudaStream_t s) {
  uint8_t* t = static_cast<uint8_t*>(tables);
  tail_fwd_tables_kernel<<<(16 * 4 * kFK + 255) / 256, 256, 0, s>>>(
      W2, b2, w3, b3, reinterpret_cast<float*>(t + kTailFwdKf), reinterpret_cast<__nv_bfloat16*>(t + kTailFwdKB),
      reinterpret_cast<float*>(t + kTailFwdBias));
  return static_cast<int>(cudaGetLastError());
}

int launch_tail_fwd(const void* x_bf16, const void* tables, float* out, long long n_img, int H1, int W1, int num_sms,
                    cudaStream_t s) {
  if (n_img <= 0) return 0;
  static bool attr = false;
  if (!attr) {
    cudaError_t e = cudaFuncSetAttribute(tail_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kFSmem);
    if (e != cudaSuccess) return static_cast<int>(e);
    attr = true;
  }
  const uint8_t* t = static_cast<const uint8_t*>(tables);
  const int tiles_x = (W1 + kFT - 1) / kFT, tiles_y = (H1 + kFT - 1) / kFT;
  const long long total = n_img * tiles_x * tiles_y;
  if (total >= (1LL << 31)) return static_cast<int>(cudaErrorInvalidValue);
  const long long cap = 2LL * (num_sms > 0 ? num_sms : 148);
  tail_fwd_kernel<<<static_cast<unsigned>(total < cap ? total : cap), 256, kFSmem, s>>>(
      static_cast<const __nv_bfloat16*>(x_bf16), reinterpret_cast<const __nv_bfloat16*>(t + kTailFwdKB),
      reinterpret_cast<const float*>(t + kTailFwdBias), out, static_cast<int>(n_img), H1, W1, tiles_x, tiles_y);
  int e = static_cast<int>(cudaGetLastError());
  if (e) return e;
  const long long warps = n_img * border_count(H1, W1);
  tail_fwd_edge_kernel<<<static_cast<unsigned>((warps * 32 + 255) / 256), 256, 0, s>>>(
      static_cast<const __nv_bfloat16*>(x_bf16), reinterpret_cast<const float*>(t + kTailFwdKf),
      reinterpret_cast<const float*>(t + kTailFwdBias), out, n_img, H1, W1);
  return static_cast<int>(cudaGetLastError());
}

// 36-channel forward: packed operand bf16 [9][48][64] | bias fp32 [48] (inside the plan's packed-parameter buffer)
size_t tail36_weight_bytes() { return 9 * 48 * 64 * 2; }
int launch_tail36_weights(const float* W2, const float* b2, const float* w3, void* wc_bf16, float* bias48, cudaStream_t s) {
  tail36_weights_kernel<<<(9 * 48 * 64 + 255) / 256, 256, 0, s>>>(W2, b2, w3, static_cast<__nv_bfloat16*>(wc_bf16), bias48);
  return static_cast<int>(cudaGetLastError());
}
template <typename T, int kGTW>
static int launch_tail36_gather_t(const void* B, const float* b3, float* out, long long n_img, int H1, int W1, cudaStream_t s) {
  const int tiles_x = (W1 + kGTW - 1) / kGTW, tiles_y = (H1 + kGTH - 1) / kGTH;
  const long long blocks = n_img * tiles_x * tiles_y;
  if (blocks >= (1LL << 31)) return static_cast<int>(cudaErrorInvalidValue);
  tail36_gather_kernel<T, kGTW><<<static_cast<unsigned>(blocks), 8 * kGTW, 0, s>>>(static_cast<const T*>(B), b3, out, H1, W1,
                                                                                   tiles_x, tiles_y);
  return static_cast<int>(cudaGetLastError());
}
int launch_tail36_gather(const void* B, int fp32_records, const float* b3, float* out, long long n_img, int H1, int W1,
                         cudaStream_t s) {
  if (n_img <= 0) return 0;
  return fp32_records ? launch_tail36_gather_t<float, 16>(B, b3, out, n_img, H1, W1, s)
                      : launch_tail36_gather_t<__nv_bfloat16, 32>(B, b3, out, n_img, H1, W1, s);
}

}  // namespace pvsr
