// Weight-gradient launch of the 3x3 convs (and their bias gradients) on the sm_100a tensor cores.
//
//   dWp[(src,tap,cb), n, c] = sum_{img, pixel} dY[img, pixel, n] * X_src[img, pixel + tap, 64*cb + c]
//
// i.e. a GEMM whose reduction dimension is the pixel index.  Both operands are read with the SAME TMA boxes as the
// forward pass (64 channels x 128 pixels, 128B swizzle) and fed to tcgen05.mma as MN-major operands
// (M = input channels of up to two K blocks, N = up to 128 output channels, K = 16 pixels per instruction).
// The result has the layout of the forward packed operand (fp32), so the packing index doubles as scatter index.
// This is the wgrad of torch's conv backward for refine_net.py:149,151,199-205,235.
#pragma once
#include "conv.h"

namespace pvsr {

constexpr int kWgUnits = 4;    // K blocks (source, tap, channel block) handled by one CTA
constexpr int kWgChunks = 2;   // 64-column chunks of dY handled by one CTA

struct WgUnit {
  int kind;       // 0: TMA view of an activation tensor, 1: constant ones (bias gradient)
  SrcView view;
  int dx, dy;     // tap offset
  int out_kb;     // K-block index of the result inside the packed layout (ones unit: unused)
};

struct WgJob {
  int n_units;
  int n_chunks;
  WgUnit unit[kWgUnits];
  SrcView dy[kWgChunks];    // 64-channel views of the output gradient (pixel-unshuffled views for PS convs)
  int col0[kWgChunks];      // first packed column of each chunk
  int n_total;              // packed columns per K block
  long long dw_off;         // element offset of this layer's packed gradient in the fp32 gradient buffer
  long long db_off;         // element offset of this layer's packed bias gradient (ones unit)
};

struct WgParams {
  int H, W;
  int tw_log2, tiles_x, tiles_y;
  int n_img;                // images reduced over
  int n_jobs, n_splits;     // grid = n_jobs * n_splits
  const WgJob* jobs;        // device memory
  float* grad;              // fp32 packed-gradient buffer (accumulated with red.add)
};

int launch_wgrad(const ConvMaps& maps, const WgParams& p, cudaStream_t stream);
int launch_scatter_add(float* param_grad, const int* idx, const int* idx2, const float* packed, long long n,
                       cudaStream_t stream);

}  // namespace pvsr
