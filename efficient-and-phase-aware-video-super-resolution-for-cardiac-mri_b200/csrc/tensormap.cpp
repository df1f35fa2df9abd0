// TMA descriptor construction (host). cuTensorMapEncodeTiled is fetched through the runtime so the
// library does not link against libcuda directly (it must load on the GPU-less build box).
#include <cuda.h>
#include <cuda_runtime.h>
#include <cudaTypedefs.h>

#include <mutex>

#include "conv.h"

namespace pvsr {

static PFN_cuTensorMapEncodeTiled_v12000 get_encode_fn() {
  static PFN_cuTensorMapEncodeTiled_v12000 fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(p);
  });
  return fn;
}

// NHWC bf16 activation tensor viewed as (C, W, H, images); box = (64 channels, tw, th, 1 image).
// mul > 1: pixel-unshuffled view - every mul-th pixel in x and y (element strides), box spans tw*mul x th*mul.
int make_act_tmap(CUtensorMap* out, const void* base, int channels, int W, int H, long long images, int tw, int th,
                  int mul) {
  auto fn = get_encode_fn();
  if (!fn) return -100;
  cuuint64_t dims[4] = {(cuuint64_t)channels, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)images};
  cuuint64_t strides[3] = {(cuuint64_t)channels * 2, (cuuint64_t)W * channels * 2, (cuuint64_t)H * W * channels * 2};
  if (tw * mul > 256 || th * mul > 256) return -101;
  cuuint32_t box[4] = {64, (cuuint32_t)(tw * mul), (cuuint32_t)(th * mul), 1};
  cuuint32_t estr[4] = {1, (cuuint32_t)mul, (cuuint32_t)mul, 1};
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : -(1000 + (int)r);
}

// Packed weights: row-major [rows][64] bf16; box = (64, bn rows) - or half of them per CTA of a pair.
int make_weight_tmap(CUtensorMap* out, const void* base, long long rows, int bn) {
  auto fn = get_encode_fn();
  if (!fn) return -100;
  bn /= cta_pair_factor();
  cuuint64_t dims[2] = {64, (cuuint64_t)rows};
  cuuint64_t strides[1] = {128};
  cuuint32_t box[2] = {64, (cuuint32_t)bn};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : -(1000 + (int)r);
}

}  // namespace pvsr
