// 3x3 single-channel <-> 64-channel stencils of the path's two HBM-bound ends, shared by in_conv_prelu (1 -> 64 conv +
// PReLU, refine_net.py:188-192) and head_last_bwd_data (adjoint of the 64 -> 1 head conv, refine_net.py:203,205).
//
// Thread = 8 channels x a run of 8 consecutive pixels of one image row: the 3 x 10 input window is loaded once (30
// independent loads in flight, 3.75 per pixel instead of 9), row / column validity is evaluated once per run, and the
// inner loop is 72 FFMA + 4 packs + one 16-byte store per pixel.  8 threads cover the 64 channels of a pixel, so every
// store instruction of a warp writes four full 128-byte lines.  The previous one-pixel-per-iteration form spent ~3/4 of
// its issue slots on index and predicate arithmetic (ncu: 301 instructions per pixel-thread, long-scoreboard stalls on
// the 9 dependent loads).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace pvsr {

// FLIP = false: out[y, x, c] = act(b[c] + sum_t in[y + ky - 1, x + kx - 1] * w[c][t])
// FLIP = true : out[y, x, c] =        sum_t in[y - ky + 1, x - kx + 1] * w[c][t]        (data gradient)
template <bool FLIP, bool PRELU>
__device__ __forceinline__ void stencil_1to64_runs(const float* __restrict__ in, const float (&wr)[8][9],
                                                   const float (&br)[8], float slope, __nv_bfloat16* __restrict__ out,
                                                   unsigned n_rows, int H, int W, int cg) {
  const unsigned runs_per_row = static_cast<unsigned>(W + 7) >> 3;
  const unsigned n_runs = n_rows * runs_per_row;
  const unsigned rpb = blockDim.x >> 3;                       // runs per block
  const unsigned stride = gridDim.x * rpb;
  for (unsigned run = blockIdx.x * rpb + (threadIdx.x >> 3); run < n_runs; run += stride) {
    const unsigned row = run / runs_per_row;                 // img * H + y
    const int x0 = static_cast<int>(run - row * runs_per_row) * 8;
    const int y = static_cast<int>(row % static_cast<unsigned>(H));
    const float* r1 = in + static_cast<size_t>(row) * W;
    const bool up = y > 0, dn = y < H - 1;
    float v[3][10];
#pragma unroll
    for (int j = 0; j < 10; ++j) {
      const int xx = x0 - 1 + j;
      const bool okx = xx >= 0 && xx < W;
      v[0][j] = (okx && up) ? __ldg(r1 - W + xx) : 0.f;
      v[1][j] = okx ? __ldg(r1 + xx) : 0.f;
      v[2][j] = (okx && dn) ? __ldg(r1 + W + xx) : 0.f;
    }
    __nv_bfloat16* o = out + (static_cast<size_t>(row) * W + x0) * 64 + cg;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      uint32_t pk[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float f0 = br[2 * j], f1 = br[2 * j + 1];
#pragma unroll
        for (int t = 0; t < 9; ++t) {
          const int ky = t / 3, kx = t % 3;
          const float val = FLIP ? v[2 - ky][i + 2 - kx] : v[ky][i + kx];
          f0 = fmaf(val, wr[2 * j][t], f0);
          f1 = fmaf(val, wr[2 * j + 1][t], f1);
        }
        if (PRELU) {
          f0 = f0 >= 0.f ? f0 : slope * f0;
          f1 = f1 >= 0.f ? f1 : slope * f1;
        }
        __nv_bfloat162 h = __floats2bfloat162_rn(f0, f1);
        pk[j] = *reinterpret_cast<uint32_t*>(&h);
      }
      if (x0 + i < W) *reinterpret_cast<uint4*>(o + i * 64) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
    }
  }
}

constexpr int kStencilThreads = 128;   // 16 runs per block; 3 blocks per SM at <= 168 registers
inline unsigned stencil_blocks(long long n_rows, int W, int num_ctas_cap) {
  const long long runs = n_rows * ((W + 7) / 8);
  const int rpb = kStencilThreads / 8;
  long long blocks = (runs + rpb - 1) / rpb;
  if (blocks > num_ctas_cap) blocks = num_ctas_cap;
  return static_cast<unsigned>(blocks);
}

}  // namespace pvsr
