// Weight gradients of the 3x3 convs as pixel-reduction GEMMs on tcgen05 (see wgrad.h).
//
// CTA = (job, split).  A job is up to 4 K blocks of X (two MMA pairs of M = 2 x 64 input channels) against up to
// 128 columns of dY; a split is an interleaved subset of the 128-pixel tiles.  Per tile the producer warp issues one
// TMA box per X unit (shifted by its tap, zero fill = conv padding) and per dY chunk; the MMA warp issues 8
// K=16-pixel MN-major MMAs per pair, accumulating in TMEM over ALL tiles of the split; the 4 epilogue warps then add
// the [128 x N] fp32 blocks into the packed gradient buffer with red.global.add.f32 (coalesced over input channels).
// A "ones" unit (constant 1.0 tile) yields the bias gradient in the same pass.
#include "ptx.cuh"
#include "wgrad.h"

namespace pvsr {

using namespace ptx;

constexpr int kWgThreads = 192;
constexpr int kWgSlot = kTileM * kBlockK * 2;                      // 16 KB: one 64-channel x 128-pixel box
constexpr int kWgStageBytes = (kWgUnits + kWgChunks) * kWgSlot;    // 96 KB
constexpr int kWgStages = 2;
constexpr int kWgSmem = kWgStages * kWgStageBytes + 1024 + 256;
constexpr int kWgTmemCols = 256;

// MN-major operand, 128B swizzle: 64 elements contiguous along M/N, pixel rows 128 B apart, 8-row groups 1024 B
// apart (SBO), next 64-element M/N atom one slot (16 KB) further (LBO).
__device__ __forceinline__ uint64_t make_desc_mn_sw128(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFFu) >> 4);
  d |= static_cast<uint64_t>(kWgSlot >> 4) << 16;   // LBO
  d |= static_cast<uint64_t>(1024 >> 4) << 32;      // SBO
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(2) << 61;
  return d;
}

__global__ void __launch_bounds__(kWgThreads, 1)
wgrad_tc_kernel(const __grid_constant__ ConvMaps maps, const __grid_constant__ WgParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kWgStages * kWgStageBytes);
  uint64_t* full = bars;
  uint64_t* empty = bars + kWgStages;
  uint64_t* tfull = bars + 2 * kWgStages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tfull + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int job_id = blockIdx.x / p.n_splits, split = blockIdx.x % p.n_splits;
  const WgJob& job = p.jobs[job_id];
  const int tiles_per_img = p.tiles_x * p.tiles_y;
  const int total_tiles = p.n_img * tiles_per_img;
  const int my_tiles = total_tiles > split ? (total_tiles - split + p.n_splits - 1) / p.n_splits : 0;
  const int n_units = job.n_units, n_chunks = job.n_chunks;
  const int n_pairs = (n_units + 1) >> 1;

  if (warp == 0 && elect_one()) {
    prefetch_tmap(&maps.act[0]);
    for (int i = 0; i < kWgStages; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], 1);
    }
    mbar_init(tfull, 1);
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc<kWgTmemCols>(tmem_slot);
  // constant-ones tiles (bias gradient units): written once, never touched by TMA
  for (int u = 0; u < n_units; ++u) {
    if (job.unit[u].kind != 1) continue;
    for (int s = 0; s < kWgStages; ++s) {
      uint4* dst = reinterpret_cast<uint4*>(smem + s * kWgStageBytes + u * kWgSlot);
      for (int i = threadIdx.x; i < kWgSlot / 16; i += kWgThreads)
        dst[i] = make_uint4(0x3F803F80u, 0x3F803F80u, 0x3F803F80u, 0x3F803F80u);
    }
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_launch_dependents();
  pdl_wait();

  if (warp == 0) {
    if (elect_one()) {
      int n_tma = n_chunks;
      for (int u = 0; u < n_units; ++u) n_tma += job.unit[u].kind == 0 ? 1 : 0;
      int stage = 0;
      uint32_t phase = 0;
      for (int i = 0; i < my_tiles; ++i) {
        int t = split + i * p.n_splits;
        const int img = t / tiles_per_img;
        t -= img * tiles_per_img;
        const int ty = t / p.tiles_x, tx = t - ty * p.tiles_x;
        const int x0 = tx << p.tw_log2, y0 = ty * (kTileM >> p.tw_log2);
        uint8_t* st = smem + stage * kWgStageBytes;
        mbar_wait(&empty[stage], phase ^ 1);
        mbar_arrive_expect_tx(&full[stage], n_tma * kWgSlot);
        for (int u = 0; u < n_units; ++u) {
          const WgUnit& un = job.unit[u];
          if (un.kind != 0) continue;
          const SrcView& v = un.view;
          tma_load_4d(st + u * kWgSlot, &maps.act[v.map], &full[stage], v.ch0, v.mul * (x0 + un.dx) + v.off_x,
                      v.mul * (y0 + un.dy) + v.off_y, v.img_base + img);
        }
        for (int c = 0; c < n_chunks; ++c) {
          const SrcView& v = job.dy[c];
          tma_load_4d(st + (kWgUnits + c) * kWgSlot, &maps.act[v.map], &full[stage], v.ch0, v.mul * x0 + v.off_x,
                      v.mul * y0 + v.off_y, v.img_base + img);
        }
        if (++stage == kWgStages) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    // M = 128 (two 64-channel units), N = 64 * n_chunks, both operands MN-major (bits 15/16)
    const uint32_t idesc = make_idesc_bf16(64 * n_chunks) | (1u << 15) | (1u << 16);
    int stage = 0;
    uint32_t phase = 0;
    for (int i = 0; i < my_tiles; ++i) {
      mbar_wait(&full[stage], phase);
      tc_fence_after();
      if (elect_one()) {
        const uint32_t st = smem_u32(smem + stage * kWgStageBytes);
        const uint64_t bdesc = make_desc_mn_sw128(st + kWgUnits * kWgSlot);
        for (int pr = 0; pr < n_pairs; ++pr) {
          const uint64_t adesc = make_desc_mn_sw128(st + 2 * pr * kWgSlot);
          for (int k = 0; k < kTileM / 16; ++k)   // 16 pixel rows = 2048 bytes per K step
            mma_bf16_ss(tmem_base + pr * 128, adesc + 128 * k, bdesc + 128 * k, idesc, (i | k) != 0);
        }
        mma_commit(&empty[stage]);
        if (i == my_tiles - 1) mma_commit(tfull);
      }
      __syncwarp();
      if (++stage == kWgStages) { stage = 0; phase ^= 1; }
    }
  } else if (my_tiles > 0) {
    const int quad = warp & 3;
    const int m = quad * 32 + lane;
    mbar_wait(tfull, 0);
    tc_fence_after();
    for (int pr = 0; pr < n_pairs; ++pr) {
      const int u = 2 * pr + (m >> 6);     // warp-uniform
      if (u >= n_units) continue;
      const WgUnit& un = job.unit[u];
      const int cin = m & 63;
      for (int c = 0; c < n_chunks; ++c) {
        float* dst = un.kind == 0
                         ? p.grad + job.dw_off + (static_cast<long long>(un.out_kb) * job.n_total + job.col0[c]) * 64 + cin
                         : p.grad + job.db_off + job.col0[c];
#pragma unroll 1
        for (int g = 0; g < 4; ++g) {
          uint32_t v[16];
          tmem_ld16(tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + pr * 128 + c * 64 + g * 16, v);
          tmem_ld_wait();
          if (un.kind == 0) {
#pragma unroll
            for (int j = 0; j < 16; ++j) atomicAdd(dst + (g * 16 + j) * 64, __uint_as_float(v[j]));
          } else if (cin == 0) {
#pragma unroll
            for (int j = 0; j < 16; ++j) atomicAdd(dst + g * 16 + j, __uint_as_float(v[j]));
          }
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<kWgTmemCols>(tmem_base);
}

int launch_wgrad(const ConvMaps& maps, const WgParams& p, cudaStream_t stream) {
  static bool attr = false;
  if (!attr) {
    cudaError_t e = cudaFuncSetAttribute(wgrad_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kWgSmem);
    if (e != cudaSuccess) return static_cast<int>(e);
    attr = true;
  }
  if (p.n_jobs <= 0 || p.n_splits <= 0) return 0;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(p.n_jobs * p.n_splits);
  cfg.blockDim = dim3(kWgThreads);
  cfg.dynamicSmemBytes = kWgSmem;
  cfg.stream = stream;
  cudaLaunchAttribute la[1];
  la[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  la[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = la;
  cfg.numAttrs = get_pdl() ? 1 : 0;
  return static_cast<int>(cudaLaunchKernelEx(&cfg, wgrad_tc_kernel, maps, p));
}

// param_grad[idx[e]] += packed[e]  (idx < 0: padding).  The packing index is injective on real entries.
__global__ void scatter_add_kernel(float* __restrict__ param_grad, const int* __restrict__ idx,
                                   const int* __restrict__ idx2, const float* __restrict__ packed, long long n) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float v = packed[i];
  const int a = idx[i];
  if (a >= 0) atomicAdd(param_grad + a, v);
  if (idx2) {
    const int b = idx2[i];
    if (b >= 0) atomicAdd(param_grad + b, v);
  }
}

int launch_scatter_add(float* param_grad, const int* idx, const int* idx2, const float* packed, long long n,
                       cudaStream_t stream) {
  if (n == 0) return 0;
  scatter_add_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, stream>>>(param_grad, idx, idx2, packed, n);
  return static_cast<int>(cudaGetLastError());
}

}  // namespace pvsr
