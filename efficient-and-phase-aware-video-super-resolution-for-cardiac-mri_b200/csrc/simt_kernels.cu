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

#include <stdlib.h>

#include "conv.h"
#include "ptx.cuh"
#include "simt.h"
#include "stencil.cuh"

namespace pvsr {

// ------------------------------------------------------------------------------------------------
// in_conv_prelu: x fp32 [n_img][H][W] -> out bf16 NHWC [n_img][H][W][64]
// 8 threads per pixel run, 8 channels each (their 72 weights + 8 biases live in registers); see stencil.cuh.
__global__ void __launch_bounds__(kStencilThreads, 3) in_conv_prelu_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                               const float* __restrict__ b,
                                                               const float* __restrict__ slope,
                                                               __nv_bfloat16* __restrict__ out, unsigned n_rows,
                                                               int H, int W) {
  const int cg = static_cast<int>(threadIdx.x & 7) * 8;
  float wr[8][9], br[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    br[j] = __ldg(b + cg + j);
#pragma unroll
    for (int t = 0; t < 9; ++t) wr[j][t] = __ldg(w + (cg + j) * 9 + t);   // parameter layout (64, 1, 3, 3)
  }
  stencil_1to64_runs<false, true>(x, wr, br, slope[0], out, n_rows, H, W, cg);
}

int launch_in_conv_prelu(const float* x, const float* w, const float* b, const float* slope, void* out,
                         long long n_img, int H, int W, cudaStream_t s) {
  const long long n_rows = n_img * H;
  if (n_rows == 0) return 0;
  if (n_rows * ((W + 7) / 8) >= (1LL << 31)) return static_cast<int>(cudaErrorInvalidValue);
  in_conv_prelu_kernel<<<stencil_blocks(n_rows, W, 148 * 24), kStencilThreads, 0, s>>>(
      x, w, b, slope, static_cast<__nv_bfloat16*>(out), static_cast<unsigned>(n_rows), H, W);
  return static_cast<int>(cudaGetLastError());
}

// ------------------------------------------------------------------------------------------------
// head_conv_last: in bf16 NHWC [n_img][H][W][64] -> out fp32 [n_img][H][W]; out = b + sum_{tap,c} in*w
//
// HBM-bound (128 B in, 4 B out per pixel).  Every INPUT pixel q contributes to its 9 neighbours through
//   P[q][tap] = sum_c in[q][c] * w[c][tap]            (a [pixels x 64] x [64 x 9] product),
// and out[p] = b + sum_tap P[p + off(tap)][tap].  A block stages a 16x32 halo tile (512 pixels, 144-byte pixel
// pitch: conflict-free ldmatrix rows) once, its 8 warps compute P for 64 pixels each with warp-level
// mma.sync.m16n8k16 (bf16 operands, fp32 accumulate; the 64x9 weight matrix lives in 16 registers per thread as B
// fragments), P goes to shared memory as [tap][pixel] and the 14x30 interior outputs gather 9 floats each.
// Versus one thread per output pixel re-reading 9x128 B of inputs and all 576 weights from shared memory this cuts
// shared-memory traffic ~7x and removes the FFMA bottleneck.
constexpr int kLastHaloH = 16, kLastHaloW = 32, kLastPitch = 144;
constexpr int kLastTH = kLastHaloH - 2, kLastTW = kLastHaloW - 2;
constexpr int kLastHaloPix = kLastHaloH * kLastHaloW;   // 512
constexpr int kLastPStride = 516;                       // [tap][516]: conflict-free fragment stores and row gathers

__device__ __forceinline__ uint32_t f2_to_bf16x2(float lo, float hi) {
  __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&h);
}

// Persistent, double-buffered: a block walks tiles grid-stride and issues the cp.async loads of tile i+1 into the other
// halo buffer before computing tile i, so every SM always has one 64 KB tile in flight.  Measured stand-alone
// (profiles/membound_bench.py): 4.05 TB/s = 0.64 of the copy bandwidth of the same box - exactly what the former
// single-stage form (one tile per block, two blocks per SM) reached, i.e. load/compute overlap is NOT what limits it;
// ~19 MB are in flight at that rate, so the bound is on the memory side of the access pattern (open).
__device__ __forceinline__ void head_last_issue_tile(const __nv_bfloat16* __restrict__ in, uint32_t tile_s, int t,
                                                     int H, int W, int tiles_x, int tiles_y) {
  const int tx = t % tiles_x;
  t /= tiles_x;
  const int ty = t % tiles_y;
  const int img = t / tiles_y;
  const int y0 = ty * kLastTH - 1, x0 = tx * kLastTW - 1;
  const __nv_bfloat16* src = in + static_cast<size_t>(img) * H * W * 64;
  // 16 x 16 B per thread, all in flight; src-size 0 zero-fills pixels outside the image = the conv padding
#pragma unroll
  for (int it = 0; it < kLastHaloPix * 8 / 256; ++it) {
    const int i = threadIdx.x + it * 256;
    const int hp = i >> 3, ck = i & 7;
    const int y = y0 + (hp >> 5), x = x0 + (hp & 31);
    const bool ok = y >= 0 && y < H && x >= 0 && x < W;
    const __nv_bfloat16* g = src + (static_cast<size_t>(ok ? y : 0) * W + (ok ? x : 0)) * 64 + ck * 8;
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(tile_s + hp * kLastPitch + ck * 16), "l"(g),
                 "r"(ok ? 16 : 0)
                 : "memory");
  }
  asm volatile("cp.async.commit_group;" ::: "memory");
}

__global__ void __launch_bounds__(256, 1) head_conv_last_kernel(const __nv_bfloat16* __restrict__ in,
                                                                const float* __restrict__ w,
                                                                const float* __restrict__ b, float* __restrict__ out,
                                                                const float* __restrict__ target,
                                                                float* __restrict__ l1_partial, int H, int W,
                                                                int tiles_x, int tiles_y, int n_tiles) {
  extern __shared__ __align__(16) uint8_t sm[];
  constexpr int kTileBytes = kLastHaloPix * kLastPitch;                      // 512 * 144 B
  float* P = reinterpret_cast<float*>(sm + 2 * kTileBytes);                  // [9][516] fp32
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint32_t sm_s = static_cast<uint32_t>(__cvta_generic_to_shared(sm));

  // B fragments of the [64 x 16] weight matrix (taps 9..15 are zero): parameter layout (1, 64, 3, 3) -> w[c*9 + tap]
  uint32_t bw[4][2][2];
  {
    const int n = lane >> 2, k0 = (lane & 3) * 2;
#pragma unroll
    for (int ks = 0; ks < 4; ++ks)
#pragma unroll
      for (int nt = 0; nt < 2; ++nt) {
        const int tap = nt * 8 + n;
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {
          const int c = ks * 16 + k0 + hh * 8;
          const float lo = tap < 9 ? __ldg(w + c * 9 + tap) : 0.f;
          const float hi = tap < 9 ? __ldg(w + (c + 1) * 9 + tap) : 0.f;
          bw[ks][nt][hh] = f2_to_bf16x2(lo, hi);
        }
      }
  }
  const float bias = b[0];

  int t = blockIdx.x;
  if (t < n_tiles) head_last_issue_tile(in, sm_s, t, H, W, tiles_x, tiles_y);
  for (int it = 0; t < n_tiles; t += gridDim.x, ++it) {
    const int buf = it & 1;
    const int tn = t + gridDim.x;
    if (tn < n_tiles) {
      head_last_issue_tile(in, sm_s + (buf ^ 1) * kTileBytes, tn, H, W, tiles_x, tiles_y);
      asm volatile("cp.async.wait_group 1;" ::: "memory");
    } else {
      asm volatile("cp.async.wait_group 0;" ::: "memory");
    }
    __syncthreads();                     // tile `t` landed for every thread; P of the previous tile fully consumed
    const uint32_t tile_s = sm_s + buf * kTileBytes;
    int q = t;
    const int tx = q % tiles_x;
    q /= tiles_x;
    const int ty = q % tiles_y;
    const int img = q / tiles_y;
    const int y0 = ty * kLastTH - 1, x0 = tx * kLastTW - 1;                  // image coordinates of halo pixel (0, 0)

    // P for this warp's 64 halo pixels: 4 m16 tiles x 4 k16 steps x 2 n8 tiles
#pragma unroll
    for (int mt = 0; mt < 4; ++mt) {
      const int px0 = warp * 64 + mt * 16;
      float acc[2][4];
#pragma unroll
      for (int nt = 0; nt < 2; ++nt)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[nt][j] = 0.f;
      const uint32_t arow = tile_s + (px0 + (lane & 7) + 8 * ((lane >> 3) & 1)) * kLastPitch + 16 * (lane >> 4);
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) {
        uint32_t a0, a1, a2, a3;
        asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];"
                     : "=r"(a0), "=r"(a1), "=r"(a2), "=r"(a3)
                     : "r"(arow + ks * 32));
#pragma unroll
        for (int nt = 0; nt < 2; ++nt)
          asm volatile(
              "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, "
              "{%0, %1, %2, %3};"
              : "+f"(acc[nt][0]), "+f"(acc[nt][1]), "+f"(acc[nt][2]), "+f"(acc[nt][3])
              : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(bw[ks][nt][0]), "r"(bw[ks][nt][1]));
      }
      const int r = px0 + (lane >> 2), c = (lane & 3) * 2;
      P[c * kLastPStride + r] = acc[0][0];
      P[(c + 1) * kLastPStride + r] = acc[0][1];
      P[c * kLastPStride + r + 8] = acc[0][2];
      P[(c + 1) * kLastPStride + r + 8] = acc[0][3];
      if ((lane & 3) == 0) {
        P[8 * kLastPStride + r] = acc[1][0];
        P[8 * kLastPStride + r + 8] = acc[1][2];
      }
    }
    __syncthreads();                     // P complete; halo buffer `buf` free for the loads issued next iteration

    float l1 = 0.f;
    for (int i = threadIdx.x; i < kLastTH * kLastTW; i += 256) {
      const int ly = i / kLastTW, lx = i - ly * kLastTW;
      const int y = y0 + 1 + ly, x = x0 + 1 + lx;
      float acc = bias;
#pragma unroll
      for (int tap = 0; tap < 9; ++tap) acc += P[tap * kLastPStride + (ly + tap / 3) * kLastHaloW + lx + tap % 3];
      if (y < H && x < W) {
        const size_t o = (static_cast<size_t>(img) * H + y) * W + x;
        out[o] = acc;
        if (target) l1 += fabsf(acc - target[o]);
      }
    }
    if (l1_partial) {
      // |out - target| summed per image: warp shuffle, then one atomic per warp (nn.L1Loss numerator).
#pragma unroll
      for (int d = 16; d > 0; d >>= 1) l1 += __shfl_xor_sync(0xffffffffu, l1, d);
      if (lane == 0) atomicAdd(l1_partial + img, l1);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// head_conv_last, TMA form.  Same arithmetic; the halo tile arrives as ONE cp.async.bulk.tensor box (64 ch x 32 x 16
// pixels, 128B swizzle, zero fill outside the image = the conv padding) into a 3-deep ring, so that two 64 KB tiles
// are in flight per SM while a third is being multiplied (the cp.async form above holds one: 148 x 64 KB in flight
// cap it at ~4 TB/s by Little's law; 144-byte pixel pitch + 2 stages already used 165 KB of shared memory).
// ldmatrix rows address the swizzled tile directly: 16-byte chunk ck of halo pixel hp lives at hp*128 + ((ck ^ (hp&7))*16).
constexpr int kLastStages = 3;
constexpr int kLastTileBytes = kLastHaloPix * 128;          // 65536

__global__ void __launch_bounds__(256, 1) head_conv_last_tma_kernel(const __grid_constant__ CUtensorMap tmap,
                                                                    const float* __restrict__ w,
                                                                    const float* __restrict__ b, float* __restrict__ out,
                                                                    const float* __restrict__ target,
                                                                    float* __restrict__ l1_partial, int H, int W,
                                                                    int tiles_x, int tiles_y, int n_tiles) {
  extern __shared__ uint8_t sm_raw[];
  uint8_t* sm = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(sm_raw) + 1023) & ~uintptr_t(1023));
  float* P = reinterpret_cast<float*>(sm + kLastStages * kLastTileBytes);      // [9][516] fp32
  uint64_t* full = reinterpret_cast<uint64_t*>(sm + kLastStages * kLastTileBytes + 9 * kLastPStride * 4);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint32_t sm_s = ptx::smem_u32(sm);

  auto issue = [&](int t, int stage) {      // thread 0 only
    const int tx = t % tiles_x;
    const int q = t / tiles_x;
    const int ty = q % tiles_y, img = q / tiles_y;
    ptx::mbar_arrive_expect_tx(&full[stage], kLastTileBytes);
    ptx::tma_load_4d(sm + stage * kLastTileBytes, &tmap, &full[stage], 0, tx * kLastTW - 1, ty * kLastTH - 1, img);
  };
  if (threadIdx.x == 0) {
    ptx::prefetch_tmap(&tmap);
    for (int i = 0; i < kLastStages; ++i) ptx::mbar_init(&full[i], 1);
    ptx::fence_mbar_init();
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int i = 0; i < kLastStages - 1; ++i) {
      const long long t = static_cast<long long>(blockIdx.x) + static_cast<long long>(i) * gridDim.x;
      if (t < n_tiles) issue(static_cast<int>(t), i);
    }
  }

  // B fragments of the [64 x 16] weight matrix (taps 9..15 are zero): parameter layout (1, 64, 3, 3) -> w[c*9 + tap]
  uint32_t bw[4][2][2];
  {
    const int n = lane >> 2, k0 = (lane & 3) * 2;
#pragma unroll
    for (int ks = 0; ks < 4; ++ks)
#pragma unroll
      for (int nt = 0; nt < 2; ++nt) {
        const int tap = nt * 8 + n;
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {
          const int c = ks * 16 + k0 + hh * 8;
          const float lo = tap < 9 ? __ldg(w + c * 9 + tap) : 0.f;
          const float hi = tap < 9 ? __ldg(w + (c + 1) * 9 + tap) : 0.f;
          bw[ks][nt][hh] = f2_to_bf16x2(lo, hi);
        }
      }
  }
  const float bias = b[0];

  int it = 0;
  for (long long tl = blockIdx.x; tl < n_tiles; tl += gridDim.x, ++it) {
    const int t = static_cast<int>(tl);
    const int stage = it % kLastStages;
    // ring slot (it + 2) % 3 held tile it - 1, whose last reader passed the barrier that closed iteration it - 1
    if (threadIdx.x == 0) {
      const long long tn = tl + static_cast<long long>(kLastStages - 1) * gridDim.x;
      if (tn < n_tiles) issue(static_cast<int>(tn), (it + kLastStages - 1) % kLastStages);
    }
    ptx::mbar_wait(&full[stage], (it / kLastStages) & 1);
    const uint32_t tile_s = sm_s + stage * kLastTileBytes;
    int q = t;
    const int tx = q % tiles_x;
    q /= tiles_x;
    const int ty = q % tiles_y;
    const int img = q / tiles_y;
    const int y0 = ty * kLastTH - 1, x0 = tx * kLastTW - 1;                  // image coordinates of halo pixel (0, 0)

    // P for this warp's 64 halo pixels: 4 m16 tiles x 4 k16 steps x 2 n8 tiles
#pragma unroll
    for (int mt = 0; mt < 4; ++mt) {
      const int px0 = warp * 64 + mt * 16;
      float acc[2][4];
#pragma unroll
      for (int nt = 0; nt < 2; ++nt)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[nt][j] = 0.f;
      const int hp = px0 + (lane & 7) + 8 * ((lane >> 3) & 1);
      const uint32_t arow = tile_s + hp * 128;
      const int sw = hp & 7, ck0 = lane >> 4;
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) {
        uint32_t a0, a1, a2, a3;
        asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];"
                     : "=r"(a0), "=r"(a1), "=r"(a2), "=r"(a3)
                     : "r"(arow + (((ks * 2 + ck0) ^ sw) << 4)));
#pragma unroll
        for (int nt = 0; nt < 2; ++nt)
          asm volatile(
              "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, "
              "{%0, %1, %2, %3};"
              : "+f"(acc[nt][0]), "+f"(acc[nt][1]), "+f"(acc[nt][2]), "+f"(acc[nt][3])
              : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(bw[ks][nt][0]), "r"(bw[ks][nt][1]));
      }
      const int r = px0 + (lane >> 2), c = (lane & 3) * 2;
      P[c * kLastPStride + r] = acc[0][0];
      P[(c + 1) * kLastPStride + r] = acc[0][1];
      P[c * kLastPStride + r + 8] = acc[0][2];
      P[(c + 1) * kLastPStride + r + 8] = acc[0][3];
      if ((lane & 3) == 0) {
        P[8 * kLastPStride + r] = acc[1][0];
        P[8 * kLastPStride + r + 8] = acc[1][2];
      }
    }
    __syncthreads();                     // P complete; every ldmatrix read of this ring slot is done

    float l1 = 0.f;
    for (int i = threadIdx.x; i < kLastTH * kLastTW; i += 256) {
      const int ly = i / kLastTW, lx = i - ly * kLastTW;
      const int y = y0 + 1 + ly, x = x0 + 1 + lx;
      float acc = bias;
#pragma unroll
      for (int tap = 0; tap < 9; ++tap) acc += P[tap * kLastPStride + (ly + tap / 3) * kLastHaloW + lx + tap % 3];
      if (y < H && x < W) {
        const size_t o = (static_cast<size_t>(img) * H + y) * W + x;
        out[o] = acc;
        if (target) l1 += fabsf(acc - target[o]);
      }
    }
    if (l1_partial) {
#pragma unroll
      for (int d = 16; d > 0; d >>= 1) l1 += __shfl_xor_sync(0xffffffffu, l1, d);
      if (lane == 0) atomicAdd(l1_partial + img, l1);
    }
    __syncthreads();                     // P consumed before the next tile's products overwrite it
  }
}

static int g_head_tma = 1;   // 1: TMA ring form of head_conv_last, 0: cp.async double-buffer form (A/B: PVSR_HEAD_TMA)
void set_head_tma(int enable) { g_head_tma = enable ? 1 : 0; }
int get_head_tma() { return g_head_tma; }

static int launch_head_conv_last_tma(const void* in, const float* w, const float* b, float* out, const float* target,
                                     float* l1_partial, long long n_img, int H, int W, cudaStream_t s) {
  const int tiles_x = (W + kLastTW - 1) / kLastTW, tiles_y = (H + kLastTH - 1) / kLastTH;
  const size_t smem = 1024 + kLastStages * kLastTileBytes + 9 * kLastPStride * sizeof(float) + 64;
  static bool attr = false;
  static int num_sms = 0;
  if (!attr) {
    cudaError_t e = cudaFuncSetAttribute(head_conv_last_tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         static_cast<int>(smem));
    if (e != cudaSuccess) return static_cast<int>(e);
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev);
    attr = true;
  }
  CUtensorMap tmap;
  if (make_act_tmap(&tmap, in, 64, W, H, n_img, kLastHaloW, kLastHaloH, 1) != 0) return static_cast<int>(cudaErrorInvalidValue);
  const long long blocks = n_img * tiles_x * tiles_y;
  if (blocks >= (1LL << 31)) return static_cast<int>(cudaErrorInvalidValue);
  const unsigned grid = static_cast<unsigned>(blocks < num_sms ? blocks : num_sms);
  head_conv_last_tma_kernel<<<grid, 256, smem, s>>>(tmap, w, b, out, target, l1_partial, H, W, tiles_x, tiles_y,
                                                   static_cast<int>(blocks));
  return static_cast<int>(cudaGetLastError());
}

int launch_head_conv_last(const void* in, const float* w, const float* b, float* out, const float* target,
                          float* l1_partial, long long n_img, int H, int W, cudaStream_t s) {
  if (n_img == 0) return 0;
  if (g_head_tma) return launch_head_conv_last_tma(in, w, b, out, target, l1_partial, n_img, H, W, s);
  const int tiles_x = (W + kLastTW - 1) / kLastTW, tiles_y = (H + kLastTH - 1) / kLastTH;
  const size_t smem = 2 * kLastHaloPix * kLastPitch + 9 * kLastPStride * sizeof(float);
  static bool attr = false;
  static int num_sms = 0;
  if (!attr) {
    cudaError_t e = cudaFuncSetAttribute(head_conv_last_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         static_cast<int>(smem));
    if (e != cudaSuccess) return static_cast<int>(e);
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev);
    attr = true;
  }
  const long long blocks = n_img * tiles_x * tiles_y;
  if (blocks >= (1LL << 31)) return static_cast<int>(cudaErrorInvalidValue);
  const unsigned grid = static_cast<unsigned>(blocks < num_sms ? blocks : num_sms);
  head_conv_last_kernel<<<grid, 256, smem, s>>>(static_cast<const __nv_bfloat16*>(in), w, b, out, target, l1_partial,
                                               H, W, tiles_x, tiles_y, static_cast<int>(blocks));
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
