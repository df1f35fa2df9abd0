// Weight-gradient launch of the 3x3 convs (and their bias gradients) on the sm_100a tensor cores.
//
//   dWp[(src,tap,cb), n, c] = sum_{img, pixel} dY[img, pixel, n] * X_src[img, pixel + tap, 64*cb + c]
//
// i.e. a GEMM whose reduction dimension is the pixel index.  Both operands are read with the SAME TMA boxes as the
// forward pass (64 channels x 128 pixels, 128B swizzle) and fed to tcgen05.mma as MN-major operands.  A job is worked
// by a CTA PAIR (cta_group::2): M = 256 input channels (each CTA stages up to 4 units of 64 channels = two M = 128
// halves per MMA pair), N = up to 256 output channels of which each CTA stages one half, K = 16 pixels per
// instruction.  Per CTA and 128-pixel tile that is 6 boxes (96 KB) for 2 x 8 MMAs of 128 x 256 x 16 - half the
// shared-memory traffic per FLOP of the former single-CTA M = 128 / N = 128 form, which was bound by operand delivery.
// The result has the layout of the forward packed operand (fp32), so the packing index doubles as scatter index.
// This is the wgrad of torch's conv backward for refine_net.py:149,151,199-205,235.
#pragma once
#include "conv.h"

namespace pvsr {

constexpr int kWgUnits = 4;    // K blocks (source, tap, channel block) staged by one CTA of the pair
constexpr int kWgSlots = 2;    // 64-column dY boxes staged by one CTA of the pair

struct WgUnit {
  int kind;       // 0: TMA view of an activation tensor, 1: constant ones (bias gradient)
  SrcView view;
  int dx, dy;     // tap offset
  int out_kb;     // K-block index of the result inside the packed layout (ones unit: unused)
};

struct WgJob {
  int n_units[2];               // units of CTA 0 / CTA 1 (the pair issues ceil(max / 2) MMA pairs per tile)
  WgUnit unit[2][kWgUnits];
  int n_slots[2];               // dY boxes CTA 0 / 1 loads (0..2)
  int n_half;                   // D columns contributed by each CTA: N = 2 * n_half (multiple of 8, <= 128)
  SrcView dy[2][kWgSlots];      // 64-channel views of the output gradient (pixel-unshuffled views for PS convs)
  int col0[2][kWgSlots];        // packed column of the first channel of each box
  int ncols[2][kWgSlots];       // real columns of each box (<= 64; the rest is never written back)
  int n_total;                  // packed columns per K block
  long long dw_off;             // element offset of this layer's packed gradient in the fp32 gradient buffer
  long long db_off;             // element offset of this layer's packed bias gradient (ones unit)
};

struct WgParams {
  int H, W;
  int tw_log2, tiles_x, tiles_y;
  int n_img;                // images reduced over
  int n_jobs;
  int n_heavy;              // jobs [0, n_heavy) issue two MMA pairs per tile and get n_splits tile subsets each,
  int n_splits;             // jobs [n_heavy, n_jobs) issue one pair and get n_splits_light (about half as many):
  int n_splits_light;       // grid = n_heavy * n_splits + (n_jobs - n_heavy) * n_splits_light clusters of 2 CTAs
  const WgJob* jobs;        // device memory
  float* grad;              // fp32 packed-gradient buffer (accumulated with red.add)
};

int launch_wgrad(const ConvMaps& maps, const WgParams& p, cudaStream_t stream);
// MMA pairs per tile of a job (its cost class): 1 or 2.
inline int wg_job_pairs(const WgJob& j) { return ((j.n_units[0] > j.n_units[1] ? j.n_units[0] : j.n_units[1]) + 1) >> 1; }
// Orders jobs heavy-first (stable) and returns the number of heavy ones.
int sort_wgrad_jobs(WgJob* jobs, int n);
// Picks the split counts (heavy / light) minimising waves x tiles-per-split on num_sms / 2 cluster slots.
void choose_wgrad_splits(int n_heavy, int n_light, long long total_tiles, int num_sms, int* s_heavy, int* s_light);
int launch_scatter_add(float* param_grad, const int* idx, const int* idx2, const float* packed, long long n,
                       cudaStream_t stream, float scale = 1.f);

}  // namespace pvsr
