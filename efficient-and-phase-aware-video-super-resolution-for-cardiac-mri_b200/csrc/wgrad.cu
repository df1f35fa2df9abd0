// Weight gradients of the 3x3 convs as pixel-reduction GEMMs on tcgen05 (see wgrad.h).
//
// Cluster of 2 CTAs = (job, split).  A split is an interleaved subset of the 128-pixel tiles.  Per tile each CTA's
// producer warp issues one TMA box per X unit (shifted by its tap, zero fill = conv padding) and per dY box it owns; all
// boxes of both CTAs are counted on the leader's `full` barrier.  One thread of the leader issues, per MMA pair, 8
// K=16-pixel MN-major cta_group::2 MMAs (M = 256: two units of each CTA; N = 2 * n_half: each CTA's dY boxes),
// accumulating in the TMEM of both CTAs over ALL tiles of the split; the 4 epilogue warps of each CTA then add their
// [128 x N] fp32 blocks into the packed gradient buffer with red.global.add.f32 (coalesced over input channels).
// A "ones" unit (constant 1.0 tile) yields the bias gradient in the same pass.
#include "ptx.cuh"
#include "wgrad.h"

namespace pvsr {

using namespace ptx;

constexpr int kWgThreads = 192;
constexpr int kWgSlot = kTileM * kBlockK * 2;                      // 16 KB: one 64-channel x 128-pixel box
constexpr int kWgStageBytes = (kWgUnits + kWgSlots) * kWgSlot;     // 96 KB
constexpr int kWgStages = 2;
constexpr int kWgSmem = kWgStages * kWgStageBytes + 1024 + 256;
constexpr int kWgTmemCols = 512;                                   // 2 MMA pairs x up to 256 columns

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
  uint64_t* full = bars;                    // [stages]  TMA (both CTAs) -> MMA (leader)
  uint64_t* empty = bars + kWgStages;       // [stages]  MMA -> TMA (multicast to both CTAs)
  uint64_t* tfull = bars + 2 * kWgStages;   // MMA -> epilogue (multicast)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tfull + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int rank = static_cast<int>(cluster_ctarank());
  const int cid = blockIdx.x >> 1;
  // heavy jobs (two MMA pairs per tile) first with n_splits tile subsets each, then the light ones with n_splits_light
  const int heavy_clusters = p.n_heavy * p.n_splits;
  const bool heavy = cid < heavy_clusters;
  const int nsp = heavy ? p.n_splits : p.n_splits_light;
  const int job_id = heavy ? cid / nsp : p.n_heavy + (cid - heavy_clusters) / nsp;
  const int split = heavy ? cid % nsp : (cid - heavy_clusters) % nsp;
  const WgJob& job = p.jobs[job_id];
  const int tiles_per_img = p.tiles_x * p.tiles_y;
  const int total_tiles = p.n_img * tiles_per_img;
  const int my_tiles = total_tiles > split ? (total_tiles - split + nsp - 1) / nsp : 0;
  const int n_units = job.n_units[rank], n_slots = job.n_slots[rank];
  const int max_units = job.n_units[0] > job.n_units[1] ? job.n_units[0] : job.n_units[1];
  const int n_pairs = (max_units + 1) >> 1;
  const int n_cols = 2 * job.n_half;

  if (warp == 0 && elect_one()) {
    prefetch_tmap(&maps.act[0]);
    for (int i = 0; i < kWgStages; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], 1);
    }
    mbar_init(tfull, 1);
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc_cg2<kWgTmemCols>(tmem_slot);
  // constant-ones tiles (bias gradient units): written once, never touched by TMA
  for (int u = 0; u < n_units; ++u) {
    if (job.unit[rank][u].kind != 1) continue;
    for (int s = 0; s < kWgStages; ++s) {
      uint4* dst = reinterpret_cast<uint4*>(smem + s * kWgStageBytes + u * kWgSlot);
      for (int i = threadIdx.x; i < kWgSlot / 16; i += kWgThreads)
        dst[i] = make_uint4(0x3F803F80u, 0x3F803F80u, 0x3F803F80u, 0x3F803F80u);
    }
  }
  fence_proxy_async();
  tc_fence_before();
  cluster_sync();   // the peer's barriers must be initialised before anything signals them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_launch_dependents();
  pdl_wait();

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer (each CTA loads its own boxes)
    if (elect_one()) {
      int n_tma[2];
      for (int r = 0; r < 2; ++r) {
        n_tma[r] = job.n_slots[r];
        for (int u = 0; u < job.n_units[r]; ++u) n_tma[r] += job.unit[r][u].kind == 0 ? 1 : 0;
      }
      const uint32_t tx_bytes = static_cast<uint32_t>(n_tma[0] + n_tma[1]) * kWgSlot;
      int stage = 0;
      uint32_t phase = 0;
      for (int i = 0; i < my_tiles; ++i) {
        int t = split + i * nsp;
        const int img = t / tiles_per_img;
        t -= img * tiles_per_img;
        const int ty = t / p.tiles_x, tx = t - ty * p.tiles_x;
        const int x0 = tx << p.tw_log2, y0 = ty * (kTileM >> p.tw_log2);
        uint8_t* st = smem + stage * kWgStageBytes;
        mbar_wait(&empty[stage], phase ^ 1);
        if (rank == 0) mbar_arrive_expect_tx(&full[stage], tx_bytes);
        const uint32_t bar = mapa(smem_u32(&full[stage]), 0);
        for (int u = 0; u < n_units; ++u) {
          const WgUnit& un = job.unit[rank][u];
          if (un.kind != 0) continue;
          const SrcView& v = un.view;
          tma_load_4d_cg2(st + u * kWgSlot, &maps.act[v.map], bar, v.ch0, v.mul * (x0 + un.dx) + v.off_x,
                          v.mul * (y0 + un.dy) + v.off_y, v.img_base + img);
        }
        for (int c = 0; c < n_slots; ++c) {
          const SrcView& v = job.dy[rank][c];
          tma_load_4d_cg2(st + (kWgUnits + c) * kWgSlot, &maps.act[v.map], bar, v.ch0, v.mul * x0 + v.off_x,
                          v.mul * y0 + v.off_y, v.img_base + img);
        }
        if (++stage == kWgStages) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer: one thread of the leader CTA
    if (rank == 0 && elect_one()) {
      // M = 256 (two 64-channel units of each CTA), N = n_cols (n_half from each CTA), both operands MN-major
      const uint32_t idesc = make_idesc_bf16(n_cols, 256) | (1u << 15) | (1u << 16);
      const uint64_t desc0 = make_desc_mn_sw128(smem_u32(smem));
      int stage = 0;
      uint32_t phase = 0;
      for (int i = 0; i < my_tiles; ++i) {
        mbar_wait(&full[stage], phase);
        tc_fence_after();
        const uint64_t st = desc0 + static_cast<uint32_t>((stage * kWgStageBytes) >> 4);
        const uint64_t bdesc = st + ((kWgUnits * kWgSlot) >> 4);
        for (int pr = 0; pr < n_pairs; ++pr) {
          const uint64_t adesc = st + ((2 * pr * kWgSlot) >> 4);
          const uint32_t d = tmem_base + pr * 256;
          mma_bf16_ss_cg2(d, adesc, bdesc, idesc, i != 0);
#pragma unroll
          for (int k = 1; k < kTileM / 16; ++k)   // 16 pixel rows = 2048 bytes per K step
            mma_bf16_ss_cg2(d, adesc + 128 * k, bdesc + 128 * k, idesc, 1);
        }
        mma_commit_cg2(&empty[stage]);
        if (i == my_tiles - 1) mma_commit_cg2(tfull);
        if (++stage == kWgStages) { stage = 0; phase ^= 1; }
      }
    }
    __syncwarp();
  } else if (my_tiles > 0) {
    // ------------------------------------------------------------------ epilogue: this CTA's 128 x n_cols blocks
    const int quad = warp & 3;
    const int m = quad * 32 + lane;
    mbar_wait(tfull, 0);
    tc_fence_after();
    for (int pr = 0; pr < n_pairs; ++pr) {
      const int u = 2 * pr + (m >> 6);     // warp-uniform
      if (u >= n_units) continue;
      const WgUnit& un = job.unit[rank][u];
      const int cin = m & 63;
      for (int hh = 0; hh < 2; ++hh)
        for (int c = 0; c < kWgSlots; ++c) {
          const int ncol = c < job.n_slots[hh] ? job.ncols[hh][c] : 0;   // real columns of this box (multiple of 8)
          const int dcol = hh * job.n_half + c * 64;                      // first D column of the box
          float* dst = un.kind == 0 ? p.grad + job.dw_off +
                                          (static_cast<long long>(un.out_kb) * job.n_total + job.col0[hh][c]) * 64 + cin
                                    : p.grad + job.db_off + job.col0[hh][c];
#pragma unroll 1
          for (int g = 0; g * 16 < ncol; ++g) {
            uint32_t v[16];
            tmem_ld16(tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + pr * 256 + dcol + g * 16, v);
            tmem_ld_wait();
            const int lim = ncol - g * 16;   // 8 or >= 16
            if (un.kind == 0) {
#pragma unroll
              for (int j = 0; j < 16; ++j)
                if (j < lim) atomicAdd(dst + (g * 16 + j) * 64, __uint_as_float(v[j]));
            } else if (cin == 0) {
#pragma unroll
              for (int j = 0; j < 16; ++j)
                if (j < lim) atomicAdd(dst + g * 16 + j, __uint_as_float(v[j]));
            }
          }
        }
    }
  }

  tc_fence_before();
  cluster_sync();   // neither CTA may leave (or free TMEM) while its peer still reads its operands / barriers
  if (warp == 1) tmem_dealloc_cg2<kWgTmemCols>(tmem_base);
}

int launch_wgrad(const ConvMaps& maps, const WgParams& p, cudaStream_t stream) {
  static bool attr = false;
  if (!attr) {
    cudaError_t e = cudaFuncSetAttribute(wgrad_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kWgSmem);
    if (e != cudaSuccess) return static_cast<int>(e);
    attr = true;
  }
  const int clusters = p.n_heavy * p.n_splits + (p.n_jobs - p.n_heavy) * p.n_splits_light;
  if (p.n_jobs <= 0 || clusters <= 0) return 0;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(2 * clusters);
  cfg.blockDim = dim3(kWgThreads);
  cfg.dynamicSmemBytes = kWgSmem;
  cfg.stream = stream;
  cudaLaunchAttribute la[2];
  la[0].id = cudaLaunchAttributeClusterDimension;
  la[0].val.clusterDim.x = 2;
  la[0].val.clusterDim.y = 1;
  la[0].val.clusterDim.z = 1;
  la[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  la[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = la;
  cfg.numAttrs = get_pdl() ? 2 : 1;
  return static_cast<int>(cudaLaunchKernelEx(&cfg, wgrad_tc_kernel, maps, p));
}

int sort_wgrad_jobs(WgJob* jobs, int n) {
  int n_heavy = 0;
  for (int i = 0; i < n; ++i)
    if (wg_job_pairs(jobs[i]) >= 2) {
      const WgJob j = jobs[i];
      for (int k = i; k > n_heavy; --k) jobs[k] = jobs[k - 1];
      jobs[n_heavy++] = j;
    }
  return n_heavy;
}

void choose_wgrad_splits(int n_heavy, int n_light, long long total_tiles, int num_sms, int* s_heavy, int* s_light) {
  // One CTA per SM is resident (193 KB of shared memory) and a split is worked by a CTA pair: the launch runs in
  // ceil(clusters / (num_sms / 2)) waves of (total_tiles / s) tiles each - pick the split count minimising waves / s
  // (ties: fewer splits = fewer epilogues).  The former "about two waves" rule produced grids such as 300 CTAs on 148
  // SMs: a third wave for 2 % of the work.  Measured: giving the light (one MMA pair per tile) jobs half as many
  // splits is SLOWER (LSTM layer 1.08 -> 1.34 ms): with two 96 KB stages per CTA a cluster is bound by load latency,
  // not by its MMA count, so light and heavy clusters take nearly the same time per tile and what counts is how many
  // clusters are resident.  Both classes therefore get the same split count.
  const long long slots = num_sms / 2 > 0 ? num_sms / 2 : 1;
  const long long n_jobs = n_heavy + n_light;
  double best = 1e30;
  long long best_s = 1;
  const long long max_s = total_tiles < 32 ? total_tiles : 32;
  // Launches with more jobs than cluster slots (EDSR: the weight gradients of all 65 body convs in one launch, 325
  // jobs) run several waves whatever the split count; there every extra split only adds a full accumulator epilogue
  // (red.add of 2 x 128 x 256 fp32 per CTA pair, ~8 tiles worth of time - measured: 32 splits of 4 tiles each ran the
  // launch at 411 TFLOP/s), so the epilogue is charged per wave.  Launches that fit one wave keep the former rule.
  const long long epi = n_jobs > slots ? 8 : 0;
  for (long long s = 1; s <= max_s; ++s) {
    const long long waves = (n_jobs * s + slots - 1) / slots;
    const double t = static_cast<double>(waves) * static_cast<double>((total_tiles + s - 1) / s + epi);
    if (t < best * 0.999) { best = t; best_s = s; }
  }
  *s_heavy = *s_light = static_cast<int>(best_s);
}

// param_grad[idx[e]] += packed[e]  (idx < 0: padding).  The packing index is injective on real entries.
__global__ void scatter_add_kernel(float* __restrict__ param_grad, const int* __restrict__ idx,
                                   const int* __restrict__ idx2, const float* __restrict__ packed, long long n,
                                   float scale) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float v = packed[i] * scale;
  const int a = idx[i];
  if (a >= 0) atomicAdd(param_grad + a, v);
  if (idx2) {
    const int b = idx2[i];
    if (b >= 0) atomicAdd(param_grad + b, v);
  }
}

int launch_scatter_add(float* param_grad, const int* idx, const int* idx2, const float* packed, long long n,
                       cudaStream_t stream, float scale) {
  if (n == 0) return 0;
  scatter_add_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, stream>>>(param_grad, idx, idx2, packed, n,
                                                                                 scale);
  return static_cast<int>(cudaGetLastError());
}

}  // namespace pvsr
