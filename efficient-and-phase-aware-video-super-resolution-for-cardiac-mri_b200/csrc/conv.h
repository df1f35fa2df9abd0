// Host/device shared definitions of the implicit-GEMM 3x3 convolution launch.
//
// One launch computes, for every output image `img` of every problem `z`:
//   D[pixel, n] = sum_{src, tap, c} A_src[pixel + tap, c] * Wp[(src, tap, cblock), n, c]
// with A read by TMA from NHWC bf16 activation tensors (zero fill outside the image = padding 1)
// and Wp the packed bf16 weight matrix (api.cpp: pvsr_pack_index_host).  This is the GEMM view of
// torch.nn.Conv2d(k=3, padding=1) at reference src/model/nets/refine_net.py:149,151,199-205,235, and - with
// transposed/flipped packed weights - of its data gradient.
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace pvsr {

constexpr int kMaxSrc = 10;   // refine conv1: 5 frames x {forward h, backward h}
constexpr int kMaxProb = 6;   // one ConvLSTM wavefront: up to 3 layers x 2 directions share one launch
constexpr int kMaxMaps = 4;   // distinct activation tensors (TMA descriptors) one launch may read
constexpr int kTileM = 128;   // output pixels per tile (UMMA M)
constexpr int kBlockK = 64;   // bf16 channels per K block (one 128-byte swizzle row)
constexpr int kBiasStage = 2304;   // fp32 biases of one launch staged in shared memory (>= kMaxProb * 256; 9 * 256)

enum EpiKind : int {
  EPI_STORE = 0,  // bias (+ border-class term) (+ residual) -> NHWC bf16 and/or fp32
  EPI_PS = 1,     // bias -> pixel-shuffled NHWC bf16 (columns grouped per sub-pixel)
  EPI_LSTM = 2,   // ConvLSTM gates + state update (refine_net.py:258-265)
  EPI_GRAD = 3,   // fp32 accumulation (+=) into one or two NHWC 64-channel gradient tensors
};

// One 64-channel-block source of the A operand.  TMA coordinates of output pixel (x, y) of image `img`, tap
// (dx, dy), channel block cb:  (ch0 + 64*cb, mul*(x+dx) + off_x, mul*(y+dy) + off_y, img_base + img).
// mul > 1 reads a pixel-UNshuffled view of a higher-resolution tensor (element stride `mul` in the map).
struct SrcView {
  int map;       // index into ConvMaps::act
  int img_base;
  int ch0;
  int mul;
  int off_x, off_y;
};

struct ConvProblem {
  int n_src;                  // A sources of this problem (each 64*kb_per_src channels)
  SrcView src[kMaxSrc];
  int w_row_base;             // first row of this problem's weights in the packed weight matrix
  const float* bias;          // [n_total] fp32 in packed column order (nullptr = none)
  // EPI_STORE / EPI_PS
  __nv_bfloat16* out_bf16;    // NHWC, out_ch channels per pixel (nullptr = skip)
  float* out_f32;             // NHWC fp32 (nullptr = skip)
  const __nv_bfloat16* res;   // optional residual, same shape as out_bf16
  const float* posterm;       // optional [n_img][16][n_total] border-class additive term
  const __nv_bfloat16* mask;  // optional, shape of out_bf16: the value is zeroed where mask <= 0 (ReLU adjoint, EDSR)
  // EPI_GRAD: fp32 [n_img][H][W][64]; split: cols [0,64) -> grad0, [64,128) -> grad1; else cols [0,64) -> both
  float* grad0;
  float* grad1;
  // EPI_LSTM
  const float* c_in;          // cell state, tile-transposed [tile][64][128] fp32 (nullptr = zeros)
  float* c_out;               // same layout (may alias c_in)
  __nv_bfloat16* h_out;       // NHWC [n_img][H][W][64]
  __nv_bfloat16* gates_out;   // optional, tile-transposed [tile][256][128] bf16 post-activation i,f,o,g
};

struct ConvParams {
  int H, W;          // image height/width (input == output resolution)
  int tw_log2;       // tile is TH x TW pixels, TW = 1 << tw_log2, TH = 128 / TW
  int tiles_x, tiles_y;
  int n_img;         // output images per problem
  int n_prob;        // 1..kMaxProb
  int taps;          // 9 (3x3) or 1 (centre tap only: 1x1 conv)
  int kb_per_src;    // 64-channel K blocks per source per tap
  int k16_last;      // number of K=16 MMAs issued for the last K block of a source (1..4)
  int n_tiles_n;     // N tiles of BN columns
  int n_total;       // packed weight rows per K block (= n_tiles_n * BN)
  int n_store;       // columns per N tile that are real outputs (<= BN, multiple of 16)
  int out_ch;        // channels per pixel of the output tensor
  int ps_r;          // pixel-shuffle factor for EPI_PS
  int grad_split;    // EPI_GRAD column routing (see ConvProblem)
  int relu;          // EPI_STORE: max(x, 0) after the bias, before mask / residual (edsr_net.py:50 relu1)
  float out_scale;   // EPI_STORE: (acc + bias) * out_scale before mask / residual; 0 = 1 (edsr_net.py:56 res_scale)
  const float* prelu; // EPI_STORE / EPI_PS: device scalar a of nn.PReLU(num_parameters=1): x > 0 ? x : a x after the bias
                      // (drf_net.py:55-57,65,82-105), or nullptr
  int ps_ch;         // EPI_PS: channels per shuffled pixel (0 = 64): column q * ps_ch + c -> sub-pixel q, channel c
  int halo;          // != 0: padded-raster slab kernel (below); tiles_x = 1, tiles_y = pr_tiles
  int pr_wp;         // padded row pitch Wp >= W + 1: output position p = y * Wp + x, tile t covers [128 t, 128 t + 128)
  int pr_rows;       // rows of the activation box = rows a tile can span + 2 (the maps carry (64, Wp, pr_rows, 1) boxes)
  int w_resident;    // slab kernel: keep the whole weight operand in shared memory when it fits (single problem, one N tile)
  ConvProblem prob[kMaxProb];
};

struct ConvMaps {
  CUtensorMap act[kMaxMaps];  // 4D (C, W, H, images), box (64, TW*mul, TH*mul, 1), element strides (1,mul,mul,1)
  CUtensorMap w;              // 2D (64, rows), box (64, BN)
};

// CTA pairs (tcgen05 cta_group::2) on/off; the weight tensor map box must be (64, bn / cta_pair_factor()).
void set_cta_pair(int enable);
int get_cta_pair();
inline int cta_pair_factor() { return get_cta_pair() ? 2 : 1; }

// Programmatic dependent launch (griddepcontrol) on the tcgen05 launches and the LSTM gate-adjoint kernel: 0 = off.
void set_pdl(int enable);
int get_pdl();
void set_pdl_explicit();      // the caller chose (pvsr_set_pdl): plans no longer pick PDL for small batches themselves
int pdl_is_explicit();
constexpr long long kPdlAutoPixels = 8192;   // B * h * w up to which an inference plan captures its graph with PDL

// Resident weight operand of narrow slab launches (conv3x3_tc.cu: halo_resident_blocks).  1 = on (default).
void set_w_resident(int enable);
int get_w_resident();

// One table_kernel launch for all operand packs of a plan / all gradient scatters of its backward pass. 1 = on.
void set_pack_table(int enable);
int get_pack_table();

// Rank-1 adjoint of the head's tail (tail_rank1.cu) in training plans.  1 = on (default).
void set_tail_rank1(int enable);
int get_tail_rank1();

// Composite forward of the head's tail (tail_rank1.cu: one 64 -> 4 channel 5x5 conv instead of conv + shuffle + conv).
void set_tail_fwd(int enable);
int get_tail_fwd();

// Two-branch schedules of training plans (plan.cpp: Ctx::side): weight-gradient launches and the HBM-bound head
// kernels on a second stream / graph branch.  1 = on (default), 0 = single chain (A/B measurements).
void set_two_branch(int enable);
int get_two_branch();

// Slab ("halo") variant of the 3x3 launches: one activation slab per source instead of nine shifted boxes. 0 = off.
void set_halo_mode(int mode);
int get_halo_mode();

// Padded-raster geometry of a slab launch.  Output pixels are numbered p = y * Wp + x with a row pitch Wp > W, so
// that every row ends in at least one column the TMA box zero-fills (x >= W): the left / right neighbour of a border
// pixel is then a zero of the previous / same row and all nine taps of a source are row-shifted views of ONE slab of
// `rows` image rows (the rows a 128-position tile touches, plus one above and one below).  Wp = 64 for W = 63 is the
// special case "tile = whole rows"; any W works (Wp = 33 for 32-pixel training patches: 9 tiles instead of 8 per
// image, but 4.9x less activation traffic than nine 16 KB boxes).
struct PrGeom {
  int wp, rows, tiles;   // pitch, box rows, tiles per image
};
constexpr int kPrMaxSlabBytes = 56 * 1024;
inline int pr_rows_for(int H, int W, int wp) {
  (void)W;
  const int tiles = (H * wp + kTileM - 1) / kTileM;
  int rmax = 1;
  for (int t = 0; t < tiles; ++t) {
    const int o = (t * kTileM) % wp;
    const int r = (o + kTileM - 1) / wp + 1;
    rmax = r > rmax ? r : rmax;
  }
  return rmax + 2;
}
// Picks Wp minimising (tiles per image, slab positions); false if no pitch satisfies the TMA box / smem limits.
inline bool choose_pr(int H, int W, int max_mul, PrGeom* g) {
  long long best_tiles = -1, best_pos = -1;
  int cand[10], nc = 0;
  for (int i = 1; i <= 8; ++i) cand[nc++] = W + i;
  int p2 = 8;
  while (p2 < W + 1) p2 <<= 1;
  cand[nc++] = p2;
  for (int i = 0; i < nc; ++i) {
    const int wp = cand[i];
    const int rows = pr_rows_for(H, W, wp);
    if (wp * max_mul > 256 || rows * max_mul > 256) continue;
    const long long pos = static_cast<long long>(rows) * wp;
    if (pos * 128 > kPrMaxSlabBytes) continue;
    const long long tiles = (static_cast<long long>(H) * wp + kTileM - 1) / kTileM;
    if (best_tiles < 0 || tiles < best_tiles || (tiles == best_tiles && pos < best_pos)) {
      best_tiles = tiles; best_pos = pos;
      g->wp = wp; g->rows = rows; g->tiles = static_cast<int>(tiles);
    }
  }
  return best_tiles > 0;
}
// Launches the tcgen05 kernel.  Returns a cudaError_t as int.
int launch_conv3x3(int bn, int epi, const ConvMaps& maps, const ConvParams& p, int num_sms, cudaStream_t stream);

// Host helpers (tensormap.cpp)
int make_act_tmap(CUtensorMap* out, const void* base, int channels, int W, int H, long long images, int tw, int th,
                  int mul = 1);   // box (64, tw * mul, th * mul, 1); tw / th need not be powers of two
int make_weight_tmap(CUtensorMap* out, const void* base, long long rows, int bn);

inline void choose_tile(int H, int W, int* tw_log2_out, int max_tw = 128) {
  // Pick TW in {8,...,max_tw} (TH = 128/TW) minimising padded work; ties -> wider tiles (longer TMA rows).
  long long best = -1;
  int best_l = 3;
  for (int l = 7; l >= 3; --l) {
    int tw = 1 << l, th = 128 >> l;
    if (tw > max_tw) continue;
    long long tiles = (long long)((W + tw - 1) / tw) * ((H + th - 1) / th);
    if (best < 0 || tiles < best) { best = tiles; best_l = l; }
  }
  *tw_log2_out = best_l;
}

// Tile -> pixel map of the tile-transposed ConvLSTM state / gate tensors ([tile][channel][128 rows]): the padded raster
// (row r of tile t of an image = position 128 t + r = y * wp + x) whenever the slab kernel runs the cells, else
// (wp = 0) the box kernel's TH x TW rectangles of choose_tile.  One definition for the forward epilogue, the gate
// adjoint kernel, the workspace sizing and the test helpers (pvsr_lstm_tile_geometry).
inline void lstm_tile_geometry(int H, int W, int* wp, int* tiles_per_img) {
  PrGeom g{0, 0, 0};
  if (get_halo_mode() != 0 && choose_pr(H, W, 1, &g)) {
    *wp = g.wp;
    *tiles_per_img = g.tiles;
    return;
  }
  int l;
  choose_tile(H, W, &l);
  const int tw = 1 << l, th = kTileM >> l;
  *wp = 0;
  *tiles_per_img = ((W + tw - 1) / tw) * ((H + th - 1) / th);
}

}  // namespace pvsr
