/*
 * pvsr.h - C ABI of the B200-native RefineNet hot path (phase-aware cardiac cine-MRI video super-resolution).
 *
 * The reference (cmlab-mira/Efficient-and-Phase-aware-Video-Super-resolution-for-Cardiac-MRI) has NO FFI on this
 * path: its boundary is the Python nn.Module `RefineNet` (src/model/nets/refine_net.py:10-135) whose arithmetic is
 * done by torch.nn layers.  This header is the boundary a maintainer would bind instead of those layers; every entry
 * point names the reference call site it replaces.  Host side: Python (ctypes), see INTEGRATION.md.
 *
 * Conventions
 *   - plain C, no torch types; all pointers are DEVICE pointers unless the name ends in `_host`;
 *   - bf16 tensors are passed as `void*` (16-bit storage), activations are NHWC, images are stacked frame-major:
 *     image index = frame * batch + sample;
 *   - `stream` is a cudaStream_t passed as void*; work is enqueued, never synchronised;
 *   - return value: 0 = OK, >0 = cudaError_t, <0 = pvsr error (see pvsr_last_error()); no C++ exception crosses;
 *   - buffers are borrowed: the caller keeps them alive until the stream work has completed.
 */
#ifndef PVSR_H_
#define PVSR_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PVSR_VERSION 100
#define PVSR_MAX_SRC 10
#define PVSR_MAX_LAYERS 8
#define PVSR_MAX_HEAD_CONVS 4

enum { PVSR_EPI_STORE = 0, PVSR_EPI_PS = 1, PVSR_EPI_LSTM = 2, PVSR_EPI_GRAD = 3 };

/* ---- library ------------------------------------------------------------------------------------------------ */
int pvsr_version(void);
const char* pvsr_last_error(void);
/* 0 when the current CUDA device can run the kernels (compute capability 10.x); negative otherwise. */
int pvsr_device_check(void);
/* tcgen05 conv launches as CTA pairs (cta_group::2: two SMs share one weight tile, M = 256 pixels per MMA).
 * Default on; 0 selects the single-CTA kernel (A/B measurements, debugging).  Process-wide. */
int pvsr_set_cta_pair(int enable);
int pvsr_get_cta_pair(void);
/* Slab ("halo") variant of the 3x3 launches on the padded raster: output pixels are numbered p = y * Wp + x with a pitch
 * Wp > W, a tile is 128 consecutive positions, each source slab (the rows a tile touches + one above and below) is loaded
 * once and the nine taps are row-shifted shared-memory views of it; works for any image width.  0 = off (nine shifted TMA
 * boxes per source), 1 = on (default).  Also selects the tile -> pixel map of the ConvLSTM state tensors
 * (pvsr_lstm_tile_geometry): set it BEFORE the first forward of a model.  Process-wide. */
int pvsr_set_halo_mode(int mode);
int pvsr_get_halo_mode(void);
/* Programmatic dependent launch (griddepcontrol) between consecutive launches of a schedule: the next kernel's prologue
 * (barrier init, TMEM allocation) overlaps the tail of the previous one.  Default off (no gain under CUDA-graph replay at
 * 8+ sequences per step on B200: profiles/r01, r02); inference plans of at most 8192 LR pixels per frame batch (one or two
 * ACDC sequences: 143 dependent launches of <= 162 tiles, +3.5 %) capture their graph with it on their own unless this
 * function was called (the caller's choice then holds for every plan).  Set BEFORE the first run of a plan (captured
 * CUDA graphs keep the setting they were captured with).  Process-wide. */
int pvsr_set_pdl(int enable);
int pvsr_get_pdl(void);
/* Resident weight operand: slab launches with one problem and one N tile whose packed weights fit next to two activation
 * slabs (refine conv2: 27 K blocks x 4 KB per CTA; its data gradient) load them ONCE per persistent CTA instead of once
 * per tile - those launches are bound by L2 -> shared-memory delivery.  1 = on (default); env PVSR_W_RESIDENT. */
int pvsr_set_w_resident(int enable);
int pvsr_get_w_resident(void);
/* pvsr_plan_pack as ONE table launch (pvsr_run_table) instead of one launch per operand, and the gradient scatter at the
 * end of pvsr_plan_backward likewise.  1 = on (default); env PVSR_PACK_TABLE. */
int pvsr_set_pack_table(int enable);
int pvsr_get_pack_table(void);
/* Two-branch schedules of training plans: the tcgen05 weight-gradient launches and the HBM-bound 64 <-> 1 channel head
 * kernels run on a second stream (a second branch of the captured CUDA graph) underneath the dependent chain of data
 * gradients.  1 = on (default), 0 = one chain (A/B switch; env PVSR_TWO_BRANCH).  Set BEFORE the first run of a plan. */
int pvsr_set_two_branch(int enable);
int pvsr_get_two_branch(void);
/* head_conv_last form: 1 = 3-stage TMA ring (default), 0 = cp.async double buffer (A/B switch; env PVSR_HEAD_TMA). */
int pvsr_set_head_tma(int enable);
int pvsr_get_head_tma(void);

/* Debug aid (env PVSR_TRACE_LAUNCH=1): every launch of an EAGER plan run is followed by an event on its stream;
 * pvsr_debug_dump_trace() prints (stderr) the first launch of each branch that has not completed - callable from a
 * watchdog thread while the device hangs.  pvsr_debug_clear_trace() forgets the recorded launches. */
int pvsr_debug_dump_trace(void);
void pvsr_debug_clear_trace(void);

/* ---- host-side packing logic (pure CPU; usable without a GPU) ------------------------------------------------ */
/* Tile choice of the implicit GEMM: tile = (128 >> tw_log2) x (1 << tw_log2) output pixels. */
int pvsr_choose_tile(int H, int W, int* tw_log2_out);

/* Gather index of a packed conv weight operand.  The packed operand is bf16 [n_src*taps*kb_per_src][n_total][64]
 * (K block major; K block = (source, tap, 64-channel block)); element e takes parameter element idx[e] of the fp32
 * (c_out, c_in, kh, kw) Conv2d weight (refine_net.py:149,151,154,199-205,235), or 0 when idx[e] < 0.
 *   src_ch_off[s] : first input channel of source s in the Conv2d weight
 *   src_ch        : real channels per source (<= 64*kb_per_src; the rest is zero padding)
 *   taps          : 9 (3x3) or 1 (1x1 conv, or centre tap)
 *   ps_r          : 0 = column n is output channel n; r>0 = pixel-shuffle order: column q*64+c is channel c*r*r+q
 *   transpose_flip: 1 = data-gradient operand (swap c_out/c_in roles, spatially flipped taps)                    */
typedef struct pvsr_pack_spec {
  int c_out, c_in, kh, kw;
  int n_src;
  int src_ch_off[PVSR_MAX_SRC];
  int src_ch;
  int kb_per_src;
  int taps;
  int n_total;
  int ps_r;
  int transpose_flip;
  int k_ps_r;   /* transpose_flip only: K channel ic of source s is parameter output channel ic*k_ps_r^2 + s
                   (sources = sub-pixels of a pixel-unshuffled gradient); 0 = src_ch_off[s] + ic */
  int src_col_off[PVSR_MAX_SRC]; /* transpose_flip only: column n of source s is parameter input channel
                   n + src_col_off[s] (window gather of the refine conv1 data gradient) */
  int ps_ch;    /* ps_r > 0: channels per shuffled pixel (0 = 64): column q*ps_ch + c is channel c*r*r + q */
} pvsr_pack_spec;
int64_t pvsr_pack_index_count(const pvsr_pack_spec* spec);
int pvsr_pack_index_host(const pvsr_pack_spec* spec, int32_t* idx_host);
/* Column -> output-channel index (or -1) for biases in packed order. */
int pvsr_pack_bias_index_host(const pvsr_pack_spec* spec, int32_t* idx_host);

/* ---- per-op entry points ------------------------------------------------------------------------------------- */
/* out[e] = bf16(w[idx[e]] (+ w[idx2[e]]))  -- parameter -> tensor-core operand (idx2 may be NULL). */
int pvsr_pack_weights(const float* w, const int32_t* idx, const int32_t* idx2, void* out_bf16, int64_t n,
                      void* stream);
int pvsr_gather_f32(const float* src, const int32_t* idx, float* out, int64_t n, void* stream);

/* _InBlock (refine_net.py:188-192): conv3x3 1->64 + bias + PReLU.  x fp32 [n_img][H][W] -> bf16 [n_img][H][W][64]. */
int pvsr_in_conv_prelu_fwd(const float* x, const float* w, const float* b, const float* slope, void* out_bf16,
                           int64_t n_img, int H, int W, void* stream);

/* Generic tcgen05 implicit-GEMM conv3x3 (padding 1).  One descriptor = one problem.
 * Sources are 64-channel-block views of up to PVSR_MAX_VIEWS bf16 NHWC tensors.  A view with mul > 1 reads the
 * pixel-UNshuffled image (every mul-th pixel starting at (off_x, off_y)) of a tensor that is mul x larger than the
 * output - the adjoint of nn.PixelShuffle (refine_net.py:200,204), used by the data/weight gradients of the head. */
#define PVSR_MAX_VIEWS 4
typedef struct pvsr_act_view {
  const void* ptr;         /* bf16 [images][H][W][channels] */
  int channels;
  int W, H;
  int64_t images;
  int mul;                 /* 1, or the pixel-shuffle factor for unshuffled views */
} pvsr_act_view;

typedef struct pvsr_conv_desc {
  int epi;                 /* PVSR_EPI_* */
  int bn;                  /* N tile: 64, 128, 144, 192 or 256 (LSTM: 256) */
  int H, W;                /* output image size */
  int64_t n_img;           /* output images */
  int n_views;
  pvsr_act_view views[PVSR_MAX_VIEWS];
  int n_src;
  int src_view[PVSR_MAX_SRC];
  int src_img_base[PVSR_MAX_SRC];
  int src_ch0[PVSR_MAX_SRC];
  int src_off_x[PVSR_MAX_SRC];
  int src_off_y[PVSR_MAX_SRC];
  int kb_per_src;          /* 64-channel K blocks per source */
  int k16_last;            /* K=16 slices used in the last K block of a source (4 = all) */
  int taps;                /* 9 or 1 */
  const void* w_packed;    /* bf16 [rows][64] */
  int64_t w_rows;
  int w_row_base;
  int n_tiles_n;
  const float* bias;       /* packed order, n_tiles_n*bn entries, or NULL */
  /* PVSR_EPI_STORE / PVSR_EPI_PS */
  void* out_bf16;
  float* out_f32;
  const void* res;         /* bf16 residual with the shape of out_bf16, or NULL */
  const float* posterm;    /* [n_img][16][n_tiles_n*bn] or NULL */
  int out_ch;
  int n_store;
  int ps_r;
  /* PVSR_EPI_GRAD: fp32 [n_img][H][W][64] accumulated in place.  grad_split: columns [0,64) -> grad0 and
   * [64,128) -> grad1; otherwise columns [0,64) are added to both (NULL = skip). */
  float* grad0;
  float* grad1;
  int grad_split;
  /* PVSR_EPI_LSTM (ConvLSTMCell.forward, refine_net.py:247-267) */
  const float* c_in;       /* tile-transposed fp32 state or NULL (= zeros) */
  float* c_out;
  void* h_out;             /* bf16 [n_img][H][W][64] */
  void* gates_out;         /* optional bf16 [tiles][256][128] */
  /* PVSR_EPI_STORE extras (EDSR residual blocks, edsr_net.py:44-58) */
  int relu;                /* 1: max(x, 0) after the bias (conv1 + relu1) */
  const void* mask;        /* bf16, shape of out_bf16: result zeroed where mask <= 0 (adjoint of relu1), or NULL */
  float out_scale;         /* (acc + bias) * out_scale before relu / mask / residual; 0 = 1 (res_scale, :56) */
  /* PVSR_EPI_STORE / PVSR_EPI_PS: nn.PReLU(num_parameters=1) fused behind the bias (DRFNet, drf_net.py:55-57,65-105):
   * device pointer to the slope a (x > 0 ? x : a x, before mask / residual), or NULL */
  const float* prelu;
} pvsr_conv_desc;
int pvsr_conv3x3_fwd(const pvsr_conv_desc* d, void* stream);
/* Weight (+ bias) gradient of a conv described like pvsr_conv_desc: X sources (source, tap, channel block) against
 * the output gradient dY given as `n_dy` 64-column chunks (views; pixel-unshuffled views for conv+PixelShuffle).
 * Result: fp32, ACCUMULATED, in the layout of the forward packed operand [n_src*taps*kb_per_src][n_total][64]
 * (scatter to the parameter with the packing index: pvsr_scatter_add) and [n_total] for the bias. */
#define PVSR_MAX_DY 36
typedef struct pvsr_wgrad_desc {
  int H, W;
  int64_t n_img;
  int n_views;
  pvsr_act_view views[PVSR_MAX_VIEWS];
  int n_src;
  int src_view[PVSR_MAX_SRC];
  int src_img_base[PVSR_MAX_SRC];
  int src_ch0[PVSR_MAX_SRC];
  int src_off_x[PVSR_MAX_SRC];
  int src_off_y[PVSR_MAX_SRC];
  int kb_per_src;
  int taps;
  int n_dy;
  int dy_view[PVSR_MAX_DY];
  int dy_img_base[PVSR_MAX_DY];
  int dy_ch0[PVSR_MAX_DY];
  int dy_off_x[PVSR_MAX_DY];
  int dy_off_y[PVSR_MAX_DY];
  int n_total;
  int with_bias;
  float* dw_packed;
  float* db_packed;
  int n_splits;            /* 0 = automatic */
  void* job_scratch;       /* device scratch of pvsr_wgrad_scratch_bytes() bytes */
} pvsr_wgrad_desc;
int64_t pvsr_wgrad_scratch_bytes(void);
int pvsr_conv3x3_wgrad(const pvsr_wgrad_desc* d, void* stream);
/* Two-phase form for CUDA-graph capture: upload = 1 builds the job list and copies it to d->job_scratch
 * (synchronises the stream; call once, outside capture); upload = 0 launches against the jobs already resident in
 * d->job_scratch (no host<->device traffic, capturable).  The descriptor must be identical in both calls. */
int pvsr_conv3x3_wgrad_staged(const pvsr_wgrad_desc* d, int upload, void* stream);
/* Weight gradients of n_desc convs in ONE launch.  All descriptors share descs[0]'s views, image size, image count
 * and job_scratch; descs[k].dw_packed / db_packed must lie in the same fp32 allocation as descs[0].dw_packed (results
 * are addressed relative to it).  Same two-phase protocol as pvsr_conv3x3_wgrad_staged; at most 1024 jobs. */
int pvsr_conv3x3_wgrad_multi(const pvsr_wgrad_desc* descs, int n_desc, int upload, void* stream);
/* Table-driven parameter traffic: one launch runs n_jobs independent element-wise jobs (device-resident table):
 *   PVSR_TJ_PACK   : dst bf16[e] = bf16(src f32[idx[e]])  (0 when idx[e] < 0)   - pvsr_pack_weights for many layers
 *   PVSR_TJ_GATHER : dst f32[e]  = src f32[idx[e]]        (0 when idx[e] < 0)   - pvsr_gather_f32
 *   PVSR_TJ_SCATTER: dst f32[idx[e]] += scale * src f32[e] for idx[e] >= 0      - pvsr_scatter_add_scaled
 * max_n = the largest n of the table (sizes the grid). */
#define PVSR_TJ_PACK 0
#define PVSR_TJ_GATHER 1
#define PVSR_TJ_SCATTER 2
typedef struct pvsr_table_job {
  const void* src;
  const int32_t* idx;
  void* dst;
  int64_t n;
  float scale;
  int kind;
} pvsr_table_job;
int pvsr_run_table(const pvsr_table_job* jobs_dev, int n_jobs, int64_t max_n, void* stream);
/* Per-frame scores of n SR / HR frame pairs (fp32 [n][H][W], normalised) in two launches - what the runners compute
 * per frame with L1Loss + denormalize (src/utils.py:1-20) + PSNR (src/model/metrics.py:20-36) + SSIM (:86-113) and, with
 * rects, CardiacPSNR / CardiacSSIM (:116-165).  sums fp64 [n][3] (zeroed here):
 *   [0] sum |sr - hr|,  [1] sum (D(sr) - D(hr))^2 with D(x) = clamp(rint(x*std + mean), 0, 255),
 *   [2] sum of the SSIM map over the valid positions of the 11-tap window `window11` (device, normalised), constants
 *       c1 = (0.01 value_range)^2, c2 = (0.03 value_range)^2.
 * rects int32 [n][4] = (h0, hn, w0, wn) per frame, or NULL for the whole frame. */
int pvsr_frame_scores(const float* sr, const float* hr, const int32_t* rects, int64_t n, int H, int W, float mean,
                      float std, const float* window11, float value_range, double* sums, void* stream);
/* Bicubic baseline (src/model/nets/bicubic.py:15: nn.Upsample(scale_factor, 'bicubic', align_corners=True)):
 * in fp32 [n_img][h][w] -> out fp32 [n_img][h*scale][w*scale]. */
int pvsr_bicubic_upsample(const float* in, float* out, int64_t n_img, int h, int w, int scale, void* stream);
/* x fp32 [n] -> out bf16 [n][64] with channel 0 = x and channels 1..63 = 0: single-channel images / gradients as a
 * 64-channel K block of the tensor-core conv (EDSR head conv edsr_net.py:29 and the adjoint of its tail conv :33). */
int pvsr_pad_channel_bf16(const float* x, void* out_bf16, int64_t n, void* stream);
/* in fp32 [n][stride] -> out fp32 [n] = in[:, 0]  (channel 0 of a 16-column conv output). */
int pvsr_take_channel0_f32(const float* in, int stride, float* out, int64_t n, void* stream);
/* param_grad[idx[e]] += packed[e] for idx[e] >= 0 (idx2 optional second target). */
int pvsr_scatter_add(float* param_grad, const int32_t* idx, const int32_t* idx2, const float* packed, int64_t n,
                     void* stream);
/* param_grad[idx[e]] += scale * packed[e]  (gradient of a conv whose output is multiplied by res_scale). */
int pvsr_scatter_add_scaled(float* param_grad, const int32_t* idx, const float* packed, int64_t n, float scale,
                            void* stream);

/* ---- input pipeline: batches cut out of HBM-resident cine volumes --------------------------------
 * Replaces, per batch, AcdcVSRRefineNetDataset.__getitem__ (src/data/datasets/acdc_vsr_refinenet_dataset.py:49-89:
 * frame window of the circularly padded cycle, positional code slice) and the transform chain applied to it
 * (src/data/transforms.py:100-168 Normalize, :321-426 RandomHorizontalFlip / RandomVerticalFlip / RandomCropPatch).
 * `volumes`: every sequence stored as [T][Hs][Ws] (one channel) in one device buffer of element type `vol_dtype`.
 * One descriptor per sample of the batch (device memory); out fp32 [n_frames][n_samples][h][w],
 *     out[f][n][y][x] = (float(vol_n[(t_first + f) mod T][ay*y + by][ax*x + bx]) - mean) / std      (IEEE fp32)
 * pos_out (optional) fp32 [n_samples][n_frames] = pos_codes[pos_off + (t_first + f) mod T]. */
#define PVSR_DT_F32 0
#define PVSR_DT_I16 1
#define PVSR_DT_U16 2
#define PVSR_DT_U8 3
#define PVSR_DT_F64 4   /* normalised in fp64, rounded to fp32 once (numpy semantics of the reference's Normalize) */
typedef struct pvsr_cine_sample {
  int64_t vol_off;         /* element offset of frame 0 of the sequence inside `volumes` */
  int64_t pos_off;         /* element offset of the sequence's code [T] inside `pos_codes`, or -1 */
  int32_t T;               /* phases of the cardiac cycle */
  int32_t t_first;         /* cycle index of output frame 0 (may be negative; wraps modulo T) */
  int32_t Hs, Ws;          /* frame size of the stored volume */
  int32_t ay, by;          /* source row = ay * y + by (ay = -1 under a vertical flip) */
  int32_t ax, bx;          /* source column = ax * x + bx */
} pvsr_cine_sample;
int pvsr_cine_gather(const void* volumes, int vol_dtype, const pvsr_cine_sample* samples, int n_samples, int n_frames,
                     int h, int w, double mean, double std, float* out, const float* pos_codes, float* pos_out,
                     void* stream);

/* Number of fp32 elements of a tile-transposed ConvLSTM cell-state buffer for n_img images of H x W. */
int64_t pvsr_lstm_state_elems(int64_t n_img, int H, int W);
/* Tile -> pixel map of the tile-transposed ConvLSTM tensors ([tile][channel][128 rows]) under the current kernel
 * selection: *wp > 0: row r of tile t of an image is the padded-raster position 128 t + r = y * wp + x (positions with
 * x >= W or y >= H are unused); *wp == 0: the TH x TW rectangles of pvsr_choose_tile.  *tiles_per_img tiles per image. */
int pvsr_lstm_tile_geometry(int H, int W, int* wp, int* tiles_per_img);

/* _RefineBlock positional-code term (refine_net.py:168-172) as a border-class table, bias included. */
int pvsr_refine_posterm(const float* w1, const float* b1, const float* pos, float* table, int n_frames_out, int B,
                        int L, int window, int c_out, int c_in, int feat2, int n_total, void* stream);

/* _OutBlock last conv (refine_net.py:203/205): bf16 [n_img][H][W][64] -> fp32 [n_img][H][W].
 * If l1_partial != NULL also accumulates sum|out - target| per image (nn.L1Loss numerator). */
int pvsr_head_conv_last_fwd(const void* in_bf16, const float* w, const float* b, float* out, const float* target,
                            float* l1_partial, int64_t n_img, int H, int W, void* stream);

int pvsr_add_bf16(const void* a, const void* b, void* out, int64_t n_elems, void* stream);

/* Backward of the head's tail  X -conv3x3 64->256 (w2, b2)-> PixelShuffle(2) -conv3x3 64->1 (w3)-> out  in one call
 * (refine_net.py:201-205 under autograd), exploiting that the last conv has ONE output channel (csrc/tail_rank1.cu):
 *   dout fp32 [n_img][2 H1][2 W1], x bf16 NHWC [n_img][H1][W1][64], w2 fp32 (256,64,3,3), b2 (256), w3 (1,64,3,3)
 *   -> dx bf16 NHWC [n_img][H1][W1][64] (written), dw2 / db2 / dw3 / db3 (parameter layout, ACCUMULATED; NULL = skip).
 * scratch: pvsr_head_tail_scratch_bytes() bytes of device memory.  sign_scale != 0: the caller promises
 * dout = sign_scale * {-1, 0, +1} element by element (see pvsr_plan_set_sign_gradient); 0 = arbitrary gradient. */
int64_t pvsr_head_tail_scratch_bytes(void);
int pvsr_head_tail_bwd(const float* dout, const void* x_bf16, const float* w2, const float* b2, const float* w3,
                       void* dx_bf16, float* dw2, float* db2, float* dw3, float* db3, void* scratch, int64_t n_img,
                       int H1, int W1, float sign_scale, void* stream);
/* Training plans use that form for the last conv + PixelShuffle(2) + final conv of x4 / x8 heads instead of the
 * tcgen05 dgrad / wgrad launches of the 64 -> 256 conv and the two adjoint kernels of the 64 -> 1 conv.
 * 1 = on (default), 0 = the conv-by-conv backward (A/B switch; env PVSR_TAIL_RANK1). */
int pvsr_set_tail_rank1(int enable);
int pvsr_get_tail_rank1(void);
/* Forward of the same tail as ONE composite 64 -> 4 channel 5x5 convolution (both convs are linear, the second has one
 * output channel): out fp32 [n_img][2 H1][2 W1] from x bf16 NHWC [n_img][H1][W1][64]; the 64-channel HR map is never
 * materialised.  tables: pvsr_head_tail_fwd_table_bytes() of device memory, refreshed by pvsr_head_tail_fwd_tables
 * whenever w2 / b2 / w3 / b3 change (pvsr_plan_pack does it for plans). */
int64_t pvsr_head_tail_fwd_table_bytes(void);
int pvsr_head_tail_fwd_tables(const float* w2, const float* b2, const float* w3, const float* b3, void* tables,
                              void* stream);
int pvsr_head_tail_fwd(const void* x_bf16, const void* tables, float* out, int64_t n_img, int H1, int W1, void* stream);
/* Plans, x4 / x8 heads (training plans only together with pvsr_set_tail_rank1(1)): 2 = the composite as a 64 -> 36
 * channel tcgen05 conv (B[z'][(q',t)] = what HR position 2z'+q' contributes to output pixel 2z'+q'-t) + a 9-tap gather
 * (default), 1 = the 5x5 composite on mma.sync above, 0 = conv + shuffle + conv (A/B switch; env PVSR_TAIL_FWD). */
int pvsr_set_tail_fwd(int enable);
int pvsr_get_tail_fwd(void);

/* ---- backward / optimiser ops ---------------------------------------------------------------------------------- */
/* Adjoint of the ConvLSTM gate math (refine_net.py:258-265) for one cell step over n_img images of H x W:
 * dh fp32 NHWC [n_img][H][W][64]; gates bf16 / c, c_prev, dc fp32 tile-transposed (c_prev NULL = zeros; dc is
 * in/out, dc_zero != 0 treats the incoming value as zero); dgates bf16 NHWC [n_img][H][W][256] (pre-activation). */
int pvsr_lstm_cell_bwd_pointwise(const float* dh, const void* gates, const float* c, const float* c_prev, float* dc,
                                 int dc_zero, void* dgates, int64_t n_img, int H, int W, void* stream);
/* Trainer loss (acdc_vsr_refinenet_trainer.py:83-100 with nn.L1Loss): out fp32 [n_lists][n_per_list], target fp32
 * [n_per_list], w fp32 [n_lists] (device) = per-list weight discount/(T*N*H*W).  *loss += sum_k w_k*sum|out_k - t|;
 * dout (may be NULL) = w_k * sign(out_k - target).  Any n_per_list (16-byte vector form when it is a multiple of 4). */
int pvsr_l1_multistage(const float* out, const float* target, const float* w, int n_lists, int64_t n_per_list,
                       float* loss, float* dout, void* stream);
/* _OutBlock last conv backward: dout fp32 [n_img][H][W] -> din bf16 [n_img][H][W][64]; dw (1,64,3,3), db (1) +=. */
int pvsr_head_conv_last_bwd_data(const float* dout, const float* w, void* din_bf16, int64_t n_img, int H, int W,
                                 void* stream);
int pvsr_head_conv_last_bwd_weight(const void* in_bf16, const float* dout, float* dw, float* db, int64_t n_img, int H,
                                   int W, void* stream);
/* _InBlock backward: x fp32 [n_img][H][W], g = dL/dy fp32 NHWC [n_img][H][W][64]; dw (64,1,3,3), db (64), dslope += */
int pvsr_in_conv_prelu_bwd(const float* x, const float* w, const float* b, const float* slope, const float* g,
                           float* dw, float* db, float* dslope, int64_t n_img, int H, int W, void* stream);
/* Gradient of the positional-code channels of _RefineBlock conv1: g = dL/d(conv1 out) bf16 NHWC with `ch` channels
 * for n_frames*B images (frame-major); the window of gradient frame f covers input frames frame0+f .. +window-1.
 * sums: scratch fp32 [window][16][ch]; dw1 (c_out, c_in, 3, 3) += on the pos channels only. */
int pvsr_refine_posterm_bwd(const void* g_bf16, const float* pos, float* sums, float* dw1, int n_frames, int B, int L,
                            int frame0, int window, int H, int W, int c_out, int c_in, int feat2, int ch,
                            void* stream);
int pvsr_cast_f32_bf16(const float* in, void* out_bf16, int64_t n, void* stream);
/* nn.PReLU(num_parameters=1) of the DRFNet row (drf_net.py:55-57,65,82-105) on bf16 streams, n a multiple of 8.
 * Inference fuses it into the conv epilogue (pvsr_conv_desc.prelu); training stores the pre-activation z and runs
 *   fwd: y = z > 0 ? z : a z        bwd: dz = g (z > 0 ? 1 : a)  (dz may alias g),  dslope[0] += sum g min(z, 0). */
int pvsr_prelu_fwd_bf16(const void* z_bf16, const float* slope, void* y_bf16, int64_t n, void* stream);
int pvsr_prelu_bwd_bf16(const void* g_bf16, const void* z_bf16, const float* slope, void* dz_bf16, float* dslope,
                        int64_t n, void* stream);
/* torch.optim.Adam.step (amsgrad off) on flat fp32 buffers of n elements (n % 4 == 0); g is scaled by grad_scale
 * first (1/world_size for averaged data-parallel gradients); state[0] (device float) is the step counter. */
int pvsr_adam_step(float* p, const float* g, float* m, float* v, int64_t n, float lr, float beta1, float beta2,
                   float eps, float weight_decay, float grad_scale, float* state, void* stream);

/* ---- whole-network plan: RefineNet.forward (refine_net.py:61-135) --------------------------------------------- */
typedef struct pvsr_net_config {
  int batch;              /* N: cine sequences processed together */
  int n_frames;           /* L = T + 2U input frames */
  int n_updated;          /* U = num_updated_frames */
  int h, w;               /* LR frame size */
  int scale;              /* upscale_factor: 2, 3, 4 or 8 */
  int n_stages;           /* num_stages */
  int window;             /* refine_window_size (odd) */
  int n_layers;           /* len(num_features); every entry must be 64 */
  int pos_enc;            /* positional_encoding */
  int memory;             /* ConvLSTMCell memory flag */
  int all_heads;          /* 1: all 3*n_stages output lists (reference behaviour); 0: only the last list */
  int save_for_backward;  /* 1: keep per-step states for pvsr_plan_backward */
} pvsr_net_config;

typedef struct pvsr_net_params { /* fp32 device pointers, reference state_dict layout */
  const float* in_w; const float* in_b; const float* in_slope;
  const float* lstm_w[2][PVSR_MAX_LAYERS]; const float* lstm_b[2][PVSR_MAX_LAYERS]; /* [0]=forward,[1]=backward */
  const float* ref_w1; const float* ref_b1; const float* ref_w2; const float* ref_b2;
  const float* head_w[PVSR_MAX_HEAD_CONVS]; const float* head_b[PVSR_MAX_HEAD_CONVS];
} pvsr_net_params;

typedef struct pvsr_plan pvsr_plan;
int pvsr_plan_create(const pvsr_net_config* cfg, pvsr_plan** out);      /* host only */
void pvsr_plan_destroy(pvsr_plan* p);
int64_t pvsr_plan_workspace_bytes(const pvsr_plan* p);
int64_t pvsr_plan_packed_bytes(const pvsr_plan* p);
int64_t pvsr_plan_output_elems(const pvsr_plan* p);    /* fp32 elements: [lists][T][N][H*s][W*s] */
int pvsr_plan_num_lists(const pvsr_plan* p);
int64_t pvsr_plan_num_launches(const pvsr_plan* p);    /* kernels launched by one pvsr_plan_forward */
double pvsr_plan_flops(const pvsr_plan* p);            /* conv FLOPs executed by one forward */
/* fp32 parameters -> packed bf16 operands (+ biases in packed order); call after every parameter update. */
int pvsr_plan_pack(pvsr_plan* p, const pvsr_net_params* params, void* packed, void* stream);
/* lr: fp32 [L][N][h][w]; pos: fp32 [N][L]; out: fp32 [lists][T][N][H*s][W*s].
 * use_graph != 0 replays a CUDA graph captured on first use for this (workspace, packed, lr, pos, out) tuple. */
int pvsr_plan_forward(pvsr_plan* p, const pvsr_net_params* params, const void* packed, const float* lr,
                      const float* pos, float* out, void* workspace, int use_graph, void* stream);

/* ---- training: backward of the plan (cfg.save_for_backward = 1, cfg.all_heads = 1) ------------------------------
 * Gradients are ACCUMULATED (+=) into fp32 buffers with the parameter layouts; NULL entries are skipped. */
typedef struct pvsr_net_grads {
  float* in_w; float* in_b; float* in_slope;
  float* lstm_w[2][PVSR_MAX_LAYERS]; float* lstm_b[2][PVSR_MAX_LAYERS];
  float* ref_w1; float* ref_b1; float* ref_w2; float* ref_b2;
  float* head_w[PVSR_MAX_HEAD_CONVS]; float* head_b[PVSR_MAX_HEAD_CONVS];
} pvsr_net_grads;
/* dout: fp32 [lists][T][N][H*s][W*s] = dL/d(out) of the preceding pvsr_plan_forward on the same workspace (which
 * must not have been overwritten since); lr / pos: the inputs of that forward.  Runs loss.backward()'s work:
 * head / refine / ConvLSTM (truncated BPTT over the T gradient frames, refine_net.py:74-93,179-183) / in-block
 * data and weight gradients (reference: autograd through refine_net.py:61-135). */
int pvsr_plan_backward(pvsr_plan* p, const pvsr_net_params* params, const void* packed, const float* lr,
                       const float* pos, const float* dout, const pvsr_net_grads* grads, void* workspace,
                       int use_graph, void* stream);
/* Promise about the NEXT pvsr_plan_backward calls of this plan: dL/d(out) of list k is scales[k] * {-1, 0, +1} element
 * by element (what pvsr_l1_multistage writes: the L1 gradient w_k sign(out - target)).  The rank-1 tail adjoint then uses
 * the exact sign as its bf16 operand and drops the hi/lo split of an arbitrary fp32 gradient (half the MMAs).  scales =
 * NULL (or n_lists = 0) withdraws the promise (arbitrary gradients: the autograd path).  Host pointer, copied. */
int pvsr_plan_set_sign_gradient(pvsr_plan* plan, const float* scales, int n_lists);
int64_t pvsr_plan_num_launches_bwd(const pvsr_plan* p);
double pvsr_plan_flops_bwd(const pvsr_plan* p);
/* Backward launch classes: 0 head last-conv adjoints, 1 head dgrad, 2 head wgrad, 3 refine dgrad, 4 refine wgrad,
 * 5 ConvLSTM pointwise adjoint, 6 ConvLSTM dgrad, 7 ConvLSTM wgrad, 8 misc (memsets, casts, in-block, scatter). */
#define PVSR_NUM_CLASSES_BWD 9
int pvsr_plan_class_stats_bwd(const pvsr_plan* p, int64_t* launches, double* flops);
int pvsr_plan_profile_bwd(pvsr_plan* p, const pvsr_net_params* params, const void* packed, const float* lr,
                          const float* pos, const float* dout, const pvsr_net_grads* grads, void* workspace,
                          double* ms_by_class, void* stream);

/* Launch classes for accounting: 0 in_conv, 1 ConvLSTM cells, 2 refine conv1, 3 refine conv2, 4 head conv+shuffle,
 * 5 head last conv, 6 misc (adds, posterm).  Arrays must hold PVSR_NUM_CLASSES entries. */
#define PVSR_NUM_CLASSES 7
int pvsr_plan_class_stats(const pvsr_plan* p, int64_t* launches, double* flops);
/* One eager forward with CUDA events around every launch (synchronises `stream`); summed ms per launch class. */
int pvsr_plan_profile(pvsr_plan* p, const pvsr_net_params* params, const void* packed, const float* lr,
                      const float* pos, float* out, void* workspace, double* ms_by_class, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* PVSR_H_ */
