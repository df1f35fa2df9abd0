// Launchers of the memory-bound kernels (simt_kernels.cu). All return a cudaError_t as int.
#pragma once
#include <cuda_runtime.h>

struct pvsr_cine_sample;
struct pvsr_table_job;

namespace pvsr {

int launch_in_conv_prelu(const float* x, const float* w, const float* b, const float* slope, void* out_bf16_nhwc,
                         long long n_img, int H, int W, cudaStream_t s);
int launch_head_conv_last(const void* in_bf16_nhwc, const float* w, const float* b, float* out, const float* target,
                          float* l1_partial, long long n_img, int H, int W, cudaStream_t s);
void set_head_tma(int enable);   // A/B switch of the head_conv_last forms (TMA ring / cp.async)
int get_head_tma();
int launch_posterm(const float* w1, const float* b1, const float* pos, float* table, int n_frames_out, int B, int L,
                   int window, int c_out, int c_in, int feat2, int n_total, cudaStream_t s);
int launch_pack_weights(const float* w, const int* idx, const int* idx2, void* out_bf16, long long n, cudaStream_t s);
int launch_gather_f32(const float* src, const int* idx, float* out, long long n, cudaStream_t s);
int launch_cine_gather(const void* vols, int dtype, const struct pvsr_cine_sample* samples, int n_samples, int n_frames,
                       int h, int w, double mean, double stdv, float* out, const float* pos_codes, float* pos_out,
                       cudaStream_t s);
int launch_pad_channel_bf16(const float* x, void* out_bf16, long long n, cudaStream_t s);
int launch_take_channel0(const float* in, int stride, float* out, long long n, cudaStream_t s);
int launch_table(const struct pvsr_table_job* jobs, int n_jobs, long long max_n, cudaStream_t s);
int launch_frame_scores(const float* sr, const float* hr, const int* rects, long long n, int H, int W, float mean,
                        float stdv, const float* window, float value_range, double* sums, cudaStream_t s);
int launch_bicubic(const float* in, float* out, long long n_img, int h, int w, int scale, cudaStream_t s);
int launch_add_bf16(const void* a, const void* b, void* out, long long n_elems, cudaStream_t s);


// ---- backward / optimiser kernels (simt_bwd.cu)
struct LstmBwdProb {
  const float* dh;          // fp32 NHWC [n_img][H][W][64]: total gradient wrt h_t
  const void* gates;        // bf16 tile-transposed [tile][256][128]: post-activation i, f, o, g of the forward pass
  const float* c;           // fp32 tile-transposed [tile][64][128]: c_t
  const float* c_prev;      // c_{t-1} (nullptr = zeros)
  float* dc;                // in: gradient wrt c_t from the later step; out: gradient wrt c_{t-1}
  int dc_zero;              // 1: the incoming dc is zero (first backward step of the cell)
  void* dgates;             // bf16 NHWC [n_img][H][W][256]: pre-activation gate gradients
};
struct LstmBwdParams {
  int H, W, tw_log2, tiles_x, tiles_y, n_img, n_prob;
  int wp, tiles_per_img;    // tile -> pixel map of the tile-transposed tensors (conv.h lstm_tile_geometry; wp = 0: box tiles)
  LstmBwdProb prob[6];
};
int launch_lstm_bwd_pointwise(const LstmBwdParams& p, cudaStream_t s);
int launch_l1_multistage(const float* out, const float* target, const float* w, int n_lists, long long n_per_list,
                         float* loss, float* dout, int num_sms, cudaStream_t s);
int launch_head_last_bwd_data(const float* dout, const float* w, void* din_bf16, long long n_img, int H, int W,
                              cudaStream_t s);
int launch_head_last_bwd_weight(const void* in_bf16, const float* dout, float* dw, float* db, long long n_img, int H,
                                int W, int num_sms, cudaStream_t s);
int launch_in_conv_prelu_bwd(const float* x, const float* w, const float* b, const float* slope, const float* g,
                             float* dw, float* db, float* dslope, long long n_img, int H, int W, int num_sms,
                             cudaStream_t s);
int launch_posterm_bwd(const void* g_bf16, const float* pos, float* sums, float* dw1, int n_frames, int B, int L,
                       int frame0, int window, int H, int W, int c_out, int c_in, int feat2, int ch, cudaStream_t s);
// Rank-1 adjoint of conv3x3 (64 -> 256) + PixelShuffle(2) + conv3x3 (64 -> 1) (tail_rank1.cu)
size_t tail_scratch_bytes();
int launch_tail_tables(const float* W2, const float* w3, void* scratch, cudaStream_t s);
int launch_tail_zero_sums(void* scratch, cudaStream_t s);
// sign_scale != 0: g is known to be sign_scale * {-1, 0, +1} (fused L1 gradient) - no bf16 lo half, half the MMAs
int launch_tail_dx(const float* g, const void* scratch, void* dx_bf16, long long n_img, int H1, int W1, int num_sms,
                   cudaStream_t s, float sign_scale = 0.f);
int launch_tail_corr(const float* g, const void* x_bf16, void* scratch, long long n_img, int H1, int W1, int num_sms,
                     cudaStream_t s, float sign_scale = 0.f);
int launch_tail_finish(const void* scratch, const float* W2, const float* b2, const float* w3, float* dW2, float* db2,
                       float* dw3, float* db3, cudaStream_t s);
// ... and its forward as one composite 64 -> 4 channel 5x5 convolution (tables: tail_fwd_table_bytes() of device memory)
size_t tail_fwd_table_bytes();
int launch_tail_fwd_tables(const float* W2, const float* b2, const float* w3, const float* b3, void* tables,
                           cudaStream_t s);
int launch_tail_fwd(const void* x_bf16, const void* tables, float* out, long long n_img, int H1, int W1, int num_sms,
                    cudaStream_t s);
// ... and as a 64 -> 36 channel tcgen05 conv + 9-tap gather (packed operand [9][48][64] bf16, bias fp32 [48])
size_t tail36_weight_bytes();
int launch_tail36_weights(const float* W2, const float* b2, const float* w3, void* wc_bf16, float* bias48, cudaStream_t s);
int launch_tail36_gather(const void* B, int fp32_records, const float* b3, float* out, long long n_img, int H1, int W1,
                         cudaStream_t s);
int launch_cast_f32_bf16(const float* in, void* out, long long n, cudaStream_t s);
// drf_kernels.cu: nn.PReLU(num_parameters=1) on bf16 streams (n a multiple of 8)
int launch_prelu_fwd(const void* z, const float* slope, void* y, long long n, int num_sms, cudaStream_t s);
int launch_prelu_bwd(const void* g, const void* z, const float* slope, void* dz, float* dslope, long long n, int num_sms,
                     cudaStream_t s);
int launch_adam(float* p, const float* g, float* m, float* v, long long n, float lr, float b1, float b2, float eps,
                float wd, float grad_scale, float* state, int num_sms, cudaStream_t s);

}  // namespace pvsr
