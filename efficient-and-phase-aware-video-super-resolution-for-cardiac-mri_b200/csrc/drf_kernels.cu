// Pointwise kernels of the DRFNet widening row (reference src/model/nets/drf_net.py): every conv of
// that net is followed by nn.PReLU(num_parameters=1, init=0.2) (:55-57, :65, :82-105).  Inference fuses the activation
// into the conv epilogue (ConvParams::prelu); training plans store the pre-activation z once (bf16) and run
//   prelu_fwd : y = z > 0 ? z : a z                                  (the next conv's TMA operand)
//   prelu_bwd : dz = g (z > 0 ? 1 : a)  and  da += sum g min(z, 0)   (g = dL/dy accumulated over all consumers of y)
// Both are HBM-bound streams of 16-byte vectors (8 bf16): 4 B / element forward, 6 B / element backward.
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "simt.h"

namespace pvsr {

__device__ __forceinline__ void unpack8(const uint4& u, float (&f)[8]) {
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const float2 v = __bfloat1622float2(h[j]);
    f[2 * j] = v.x;
    f[2 * j + 1] = v.y;
  }
}
__device__ __forceinline__ uint4 pack8(const float (&f)[8]) {
  uint4 r;
  uint32_t* p = reinterpret_cast<uint32_t*>(&r);
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    __nv_bfloat162 h = __floats2bfloat162_rn(f[2 * j], f[2 * j + 1]);
    p[j] = *reinterpret_cast<uint32_t*>(&h);
  }
  return r;
}

__global__ void __launch_bounds__(256) prelu_fwd_kernel(const uint4* __restrict__ z, const float* __restrict__ slope,
                                                        uint4* __restrict__ y, long long n8) {
  const float a = __ldg(slope);
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n8;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    float f[8];
    unpack8(z[i], f);
#pragma unroll
    for (int j = 0; j < 8; ++j) f[j] = f[j] > 0.f ? f[j] : a * f[j];
    y[i] = pack8(f);
  }
}

// dz may alias g (in place).  da: one fp32 atomicAdd per block (the order of the block sums is not fixed; the result
// is a sum of ~1e3 partials of an fp32 reduction - well below the bf16 noise of g).
__global__ void __launch_bounds__(256) prelu_bwd_kernel(const uint4* g, const uint4* __restrict__ z,
                                                        const float* __restrict__ slope, uint4* dz,
                                                        float* __restrict__ dslope, long long n8) {
  const float a = __ldg(slope);
  float acc = 0.f;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n8;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    float gf[8], zf[8];
    unpack8(g[i], gf);
    unpack8(z[i], zf);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const bool pos = zf[j] > 0.f;
      acc += pos ? 0.f : gf[j] * zf[j];
      gf[j] = pos ? gf[j] : a * gf[j];
    }
    dz[i] = pack8(gf);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  __shared__ float part[8];
  if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) s += part[w];
    if (dslope) atomicAdd(dslope, s);
  }
}

static unsigned stream_grid(long long n8, int num_sms) {
  const long long want = (n8 + 255) / 256;
  const long long cap = static_cast<long long>(num_sms > 0 ? num_sms : 148) * 8;      // 8 resident blocks per SM
  return static_cast<unsigned>(want < cap ? (want > 0 ? want : 1) : cap);
}

int launch_prelu_fwd(const void* z, const float* slope, void* y, long long n, int num_sms, cudaStream_t s) {
  const long long n8 = n / 8;
  if (n8 == 0) return 0;
  prelu_fwd_kernel<<<stream_grid(n8, num_sms), 256, 0, s>>>(static_cast<const uint4*>(z), slope, static_cast<uint4*>(y), n8);
  return static_cast<int>(cudaGetLastError());
}
int launch_prelu_bwd(const void* g, const void* z, const float* slope, void* dz, float* dslope, long long n, int num_sms,
                     cudaStream_t s) {
  const long long n8 = n / 8;
  if (n8 == 0) return 0;
  prelu_bwd_kernel<<<stream_grid(n8, num_sms), 256, 0, s>>>(static_cast<const uint4*>(g), static_cast<const uint4*>(z), slope,
                                                            static_cast<uint4*>(dz), dslope, n8);
  return static_cast<int>(cudaGetLastError());
}

}  // namespace pvsr
