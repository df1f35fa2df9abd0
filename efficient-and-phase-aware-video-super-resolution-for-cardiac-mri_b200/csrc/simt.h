// Launchers of the memory-bound kernels (simt_kernels.cu). All return a cudaError_t as int.
#pragma once
#include <cuda_runtime.h>

namespace pvsr {

int launch_in_conv_prelu(const float* x, const float* w, const float* b, const float* slope, void* out_bf16_nhwc,
                         long long n_img, int H, int W, cudaStream_t s);
int launch_head_conv_last(const void* in_bf16_nhwc, const float* w, const float* b, float* out, const float* target,
                          float* l1_partial, long long n_img, int H, int W, cudaStream_t s);
int launch_posterm(const float* w1, const float* b1, const float* pos, float* table, int n_frames_out, int B, int L,
                   int window, int c_out, int c_in, int feat2, int n_total, cudaStream_t s);
int launch_pack_weights(const float* w, const int* idx, const int* idx2, void* out_bf16, long long n, cudaStream_t s);
int launch_gather_f32(const float* src, const int* idx, float* out, long long n, cudaStream_t s);
int launch_add_bf16(const void* a, const void* b, void* out, long long n_elems, cudaStream_t s);

}  // namespace pvsr
