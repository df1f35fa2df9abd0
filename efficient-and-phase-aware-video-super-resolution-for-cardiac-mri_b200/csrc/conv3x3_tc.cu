// 3x3 convolution (padding 1) as an implicit GEMM on the sm_100a tensor cores.
//
//   warp 0      : TMA producer  - per K block one 4D box (64 ch x TW x TH px, shifted by the tap; TMA zero-fills
//                 outside the image, which *is* the conv padding) + one 2D box of packed weights (BN x 64).
//   warp 1      : MMA issuer    - tcgen05.mma (M=128, N=BN, K=16) x 4 per K block, fp32 accumulator in TMEM,
//                 two accumulator stages so the epilogue of tile i overlaps the MMAs of tile i+1.
//   warps 2..9  : epilogue      - tcgen05.ld -> bias / ConvLSTM gates / residual / pixel-shuffle -> global.
// Persistent: grid = min(#tiles, #SMs); tiles are walked round-robin.
//
// Reference semantics reproduced (file:line in /root/reference/src/model/nets/refine_net.py):
//   ConvLSTMCell.forward 247-267 (EPI_LSTM), _RefineBlock.body 147-155 (EPI_STORE), _OutBlock 194-205 (EPI_PS).
#include "conv.h"
#include "ptx.cuh"

namespace pvsr {

using namespace ptx;

constexpr int kNumEpiWarps = 8;
constexpr int kNumThreads = 64 + kNumEpiWarps * 32;  // 320
constexpr int kABytes = kTileM * kBlockK * 2;        // 16384
constexpr int kAccStride = 256;                      // TMEM columns between the two accumulator stages
constexpr int kTmemCols = 512;

// CG = CTAs cooperating on one MMA: 1, or 2 = a CTA pair (tcgen05 cta_group::2): M = 256 pixels (128 per CTA), the
// weight tile is split over the pair (each CTA stages BN/2 rows), halving its L2->SMEM traffic and stage footprint.
template <int BN, int CG>
struct Cfg {
  static constexpr int kBBytes = (BN / CG) * kBlockK * 2;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kStages = (200 * 1024) / kStageBytes > 8 ? 8 : (200 * 1024) / kStageBytes;
  static constexpr int kSmemBytes = kStages * kStageBytes + 1024 /*align*/ + 256 /*barriers*/ +
                                    kBiasStage * 4 /*biases of the launch (LSTM: pre-scaled)*/;
  static_assert(kBBytes % 1024 == 0, "B stage must keep 1024B alignment for the 128B swizzle");
  static_assert(BN % 16 == 0 && BN <= 256, "UMMA N constraint for M=128");
};

// MUFU-only transcendental building blocks (ex2.approx / rcp.approx: ~1 ulp, no slow paths, no branches).
__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float rcp_approx(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
constexpr float kLog2e = 1.4426950408889634f;
// sigmoid(x) with t = -log2(e) * x precomputed:  1 / (1 + 2^t)
__device__ __forceinline__ float sigmoid_from_scaled(float t) { return rcp_approx(1.f + ex2_approx(t)); }
// tanh(x) with t = 2 * log2(e) * x precomputed:  1 - 2 / (1 + 2^t)   (saturates correctly at +-inf)
__device__ __forceinline__ float tanh_from_scaled(float t) { return fmaf(-2.f, rcp_approx(1.f + ex2_approx(t)), 1.f); }

__device__ __forceinline__ void red_add_v4(float* addr, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
__device__ __forceinline__ uint32_t pack_bf16x2(float a, float b) {
  __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}
__device__ __forceinline__ void unpack_bf16x8(const uint4& u, float* f) {
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float2 t = __bfloat1622float2(h[i]);
    f[2 * i] = t.x;
    f[2 * i + 1] = t.y;
  }
}

template <int CG>
__device__ __forceinline__ void mma_ss(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  if constexpr (CG == 2) mma_bf16_ss_cg2(tmem_d, adesc, bdesc, idesc, acc);
  else mma_bf16_ss(tmem_d, adesc, bdesc, idesc, acc);
}
template <int CG>
__device__ __forceinline__ void mma_commit_t(uint64_t* bar) {
  if constexpr (CG == 2) mma_commit_cg2(bar);
  else mma_commit(bar);
}

struct TileCoord {
  int z, img, y0, x0, nt, tile_lin, ty;
  bool valid;   // false: padding tile of an odd pair (loaded and multiplied like its neighbour, never stored)
};
// Work item `it` (N tile fastest, then the group of CG M tiles, then the problem) -> the M tile of CTA `rank` of the
// group.  Box kernel: a group = CG consecutive tiles of the (image, ty, tx) raster.  Slab kernel (p.halo): a group =
// the SAME tile of CG consecutive images, so that both CTAs of a pair share the tile's row phase inside the slab (one
// MMA descriptor offset serves both).
__device__ __forceinline__ TileCoord decode_item(const ConvParams& p, int it, int groups, int tiles_m, int cg, int rank) {
  TileCoord c;
  c.nt = it % p.n_tiles_n;
  it /= p.n_tiles_n;
  const int grp = it % groups;
  c.z = it / groups;
  if (p.halo) {
    c.ty = grp % p.tiles_y;
    int img = (grp / p.tiles_y) * cg + rank;
    c.valid = img < p.n_img;
    if (!c.valid) img = p.n_img - 1;
    c.img = img;
    c.x0 = 0;
    c.y0 = 0;
    c.tile_lin = img * p.tiles_y + c.ty;
    return c;
  }
  int m = grp * cg + rank;
  c.valid = m < tiles_m;
  if (!c.valid) m = tiles_m - 1;
  const int tx = m % p.tiles_x;
  m /= p.tiles_x;
  const int ty = m % p.tiles_y;
  c.ty = ty;
  c.img = m / p.tiles_y;
  c.x0 = tx << p.tw_log2;
  c.y0 = ty * (kTileM >> p.tw_log2);
  c.tile_lin = (c.img * p.tiles_y + ty) * p.tiles_x + tx;
  return c;
}

// Epilogue of one accumulator tile for one epilogue thread (TMEM lane = tile row `row`, pixel (x, y)); `half`
// selects which half of the column chunks this warp handles.
template <int BN, int EPI>
__device__ __forceinline__ void epilogue_tile(const ConvParams& p, const ConvProblem& pr, const TileCoord& tc,
                                              uint32_t taddr, int row, int y, int x, bool valid, int half,
                                              const float* bias_s, const uint4 (&res_pre)[4], bool has_pre,
                                              const float (&c_pre)[16]) {
  if constexpr (EPI == EPI_LSTM) {
    // Columns: [i | f | o | g] x 64 channels (refine_net.py:258). This warp: channels [32*half, 32*half+32).
    const float* cin = pr.c_in ? pr.c_in + (static_cast<size_t>(tc.tile_lin) * 64) * kTileM + row : nullptr;
    float* cout = pr.c_out + (static_cast<size_t>(tc.tile_lin) * 64) * kTileM + row;
    __nv_bfloat16* gout =
        pr.gates_out ? pr.gates_out + (static_cast<size_t>(tc.tile_lin) * 256) * kTileM + row : nullptr;
    __nv_bfloat16* hrow = pr.h_out + ((static_cast<size_t>(tc.img) * p.H + y) * p.W + x) * 64;
    const float4* bs4 = reinterpret_cast<const float4*>(bias_s + tc.z * 256);
    // c_{t-1} of the first 16 channels was fetched before the accumulator wait (c_pre, epilogue_prefetch); the second
    // 16 are requested while the first chunk is processed - the state loads are off the per-tile critical path.
    float cnext[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) cnext[j] = c_pre[j];
#pragma unroll 1
    for (int cc = 0; cc < 2; ++cc) {
      const int ch0 = half * 32 + cc * 16;
      uint32_t vi[16], vf[16], vo[16], vg[16];
      tmem_ld16(taddr + ch0, vi);
      tmem_ld16(taddr + 64 + ch0, vf);
      tmem_ld16(taddr + 128 + ch0, vo);
      tmem_ld16(taddr + 192 + ch0, vg);
      float cprev[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) cprev[j] = cnext[j];
      if (cc == 0 && cin) {
#pragma unroll
        for (int j = 0; j < 16; ++j) cnext[j] = cin[(ch0 + 16 + j) * kTileM];
      }
      tmem_ld_wait();
      uint32_t hp[8];
#pragma unroll
      for (int j4 = 0; j4 < 4; ++j4) {
        const float4 bi = bs4[(ch0 >> 2) + j4], bf = bs4[16 + (ch0 >> 2) + j4];
        const float4 bo = bs4[32 + (ch0 >> 2) + j4], bg = bs4[48 + (ch0 >> 2) + j4];
        const float bia[4] = {bi.x, bi.y, bi.z, bi.w}, bfa[4] = {bf.x, bf.y, bf.z, bf.w};
        const float boa[4] = {bo.x, bo.y, bo.z, bo.w}, bga[4] = {bg.x, bg.y, bg.z, bg.w};
        float hn[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int j = j4 * 4 + u;
          const float gi = sigmoid_from_scaled(fmaf(__uint_as_float(vi[j]), -kLog2e, bia[u]));
          const float gf = sigmoid_from_scaled(fmaf(__uint_as_float(vf[j]), -kLog2e, bfa[u]));
          const float go = sigmoid_from_scaled(fmaf(__uint_as_float(vo[j]), -kLog2e, boa[u]));
          const float gg = tanh_from_scaled(fmaf(__uint_as_float(vg[j]), 2.f * kLog2e, bga[u]));
          const float cn = fmaf(gf, cprev[j], gi * gg);
          hn[u] = go * tanh_from_scaled(cn * (2.f * kLog2e));
          if (tc.valid) cout[(ch0 + j) * kTileM] = cn;   // padding tiles of an odd pair never store
          if (gout && tc.valid) {
            gout[(ch0 + j) * kTileM] = __float2bfloat16(gi);
            gout[(64 + ch0 + j) * kTileM] = __float2bfloat16(gf);
            gout[(128 + ch0 + j) * kTileM] = __float2bfloat16(go);
            gout[(192 + ch0 + j) * kTileM] = __float2bfloat16(gg);
          }
        }
        hp[j4 * 2] = pack_bf16x2(hn[0], hn[1]);
        hp[j4 * 2 + 1] = pack_bf16x2(hn[2], hn[3]);
      }
      if (valid) {
        uint4* dst = reinterpret_cast<uint4*>(hrow + ch0);
        dst[0] = make_uint4(hp[0], hp[1], hp[2], hp[3]);
        dst[1] = make_uint4(hp[4], hp[5], hp[6], hp[7]);
      }
    }
  } else if constexpr (EPI == EPI_GRAD) {
    // fp32 accumulation into NHWC 64-channel gradient tensors (data gradients of the convs).
    const int nchunks = p.n_store >> 4;
    const size_t pixoff = ((static_cast<size_t>(tc.img) * p.H + y) * p.W + x) * 64;
#pragma unroll 1
    for (int ck = half; ck < nchunks; ck += 2) {
      uint32_t v[16];
      tmem_ld16(taddr + ck * 16, v);
      tmem_ld_wait();
      if (valid) {
        float* d0 = p.grad_split ? ((ck < 4) ? pr.grad0 : pr.grad1) : pr.grad0;
        float* d1 = p.grad_split ? nullptr : pr.grad1;
        const int c0 = (ck & 3) * 16;
        // Fire-and-forget vector reductions: several tiles of one launch (cells of a reverse wavefront) may add
        // into the same gradient pixels, so a plain read-modify-write would race.
        if (d0) {
          float* q = d0 + pixoff + c0;
#pragma unroll
          for (int j = 0; j < 4; ++j)
            red_add_v4(q + 4 * j, __uint_as_float(v[4 * j]), __uint_as_float(v[4 * j + 1]),
                       __uint_as_float(v[4 * j + 2]), __uint_as_float(v[4 * j + 3]));
        }
        if (d1) {
          float* q = d1 + pixoff + c0;
#pragma unroll
          for (int j = 0; j < 4; ++j)
            red_add_v4(q + 4 * j, __uint_as_float(v[4 * j]), __uint_as_float(v[4 * j + 1]),
                       __uint_as_float(v[4 * j + 2]), __uint_as_float(v[4 * j + 3]));
        }
      }
    }
  } else {
    // EPI_STORE / EPI_PS: 16-column chunks, alternating between the two warp halves.
    const int nchunks = p.n_store >> 4;
    const int cls = (y > 0 ? 1 : 0) | (y < p.H - 1 ? 2 : 0) | (x > 0 ? 4 : 0) | (x < p.W - 1 ? 8 : 0);
    const float* pterm =
        pr.posterm ? pr.posterm + (static_cast<size_t>(tc.img) * 16 + cls) * p.n_total + tc.nt * BN : nullptr;
    // bias_s: this launch's biases staged in shared memory by the prologue ([problem][n_total]); a dependent global
    // load per 16-column chunk was the top stall of the short-K head launches (ncu: 29 % of samples)
    const float* bias = pr.bias ? bias_s + tc.z * p.n_total + tc.nt * BN : nullptr;
    auto chunk = [&](int ck, bool pre, const uint4& rpre0, const uint4& rpre1) {
      uint32_t v[16];
      tmem_ld16(taddr + ck * 16, v);
      tmem_ld_wait();
      float f[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) f[j] = __uint_as_float(v[j]);
      if (bias) {
        const float4* b4 = reinterpret_cast<const float4*>(bias + ck * 16);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float4 b = b4[j];
          f[4 * j] += b.x; f[4 * j + 1] += b.y; f[4 * j + 2] += b.z; f[4 * j + 3] += b.w;
        }
      }
      if (valid) {
        if (pterm) {
          const float4* t4 = reinterpret_cast<const float4*>(pterm + ck * 16);
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float4 b = __ldg(t4 + j);
            f[4 * j] += b.x; f[4 * j + 1] += b.y; f[4 * j + 2] += b.z; f[4 * j + 3] += b.w;
          }
        }
        size_t off;
        if constexpr (EPI == EPI_PS) {
          // column = q*64 + c with q = i*r + j  ->  HR pixel (y*r+i, x*r+j), channel c (PixelShuffle, :200,204)
          const int col = tc.nt * BN + ck * 16;
          const int pc = p.ps_ch ? p.ps_ch : 64;
          const int q = col / pc, c0 = col - q * pc;
          const int r = p.ps_r;
          const int qi = q / r, qj = q - qi * r;
          off = ((static_cast<size_t>(tc.img) * p.H * r + (y * r + qi)) * (p.W * r) + (x * r + qj)) * pc + c0;
        } else {
          off = ((static_cast<size_t>(tc.img) * p.H + y) * p.W + x) * p.out_ch + tc.nt * BN + ck * 16;
        }
        if (p.out_scale != 0.f) {
#pragma unroll
          for (int j = 0; j < 16; ++j) f[j] *= p.out_scale;
        }
        if (p.relu) {
#pragma unroll
          for (int j = 0; j < 16; ++j) f[j] = fmaxf(f[j], 0.f);
        }
        if (p.prelu) {
          const float a = __ldg(p.prelu);
#pragma unroll
          for (int j = 0; j < 16; ++j) f[j] = f[j] > 0.f ? f[j] : a * f[j];
        }
        if (pr.mask) {
          float mf[16];
          const uint4* mp = reinterpret_cast<const uint4*>(pr.mask + off);
          unpack_bf16x8(mp[0], mf);
          unpack_bf16x8(mp[1], mf + 8);
#pragma unroll
          for (int j = 0; j < 16; ++j) f[j] = mf[j] > 0.f ? f[j] : 0.f;
        }
        if (pr.res) {
          float rf[16];
          if (pre) {
            unpack_bf16x8(rpre0, rf);
            unpack_bf16x8(rpre1, rf + 8);
          } else {
            const uint4* rp = reinterpret_cast<const uint4*>(pr.res + off);
            unpack_bf16x8(rp[0], rf);
            unpack_bf16x8(rp[1], rf + 8);
          }
#pragma unroll
          for (int j = 0; j < 16; ++j) f[j] += rf[j];
        }
        if (pr.out_bf16) {
          uint4* dst = reinterpret_cast<uint4*>(pr.out_bf16 + off);
          dst[0] = make_uint4(pack_bf16x2(f[0], f[1]), pack_bf16x2(f[2], f[3]), pack_bf16x2(f[4], f[5]),
                              pack_bf16x2(f[6], f[7]));
          dst[1] = make_uint4(pack_bf16x2(f[8], f[9]), pack_bf16x2(f[10], f[11]), pack_bf16x2(f[12], f[13]),
                              pack_bf16x2(f[14], f[15]));
        }
        if (pr.out_f32) {
          float4* dst = reinterpret_cast<float4*>(pr.out_f32 + off);
#pragma unroll
          for (int j = 0; j < 4; ++j) dst[j] = make_float4(f[4 * j], f[4 * j + 1], f[4 * j + 2], f[4 * j + 3]);
        }
      }
    };
    if constexpr (EPI == EPI_STORE && BN == 64) {
      // Narrow tiles finish their MMAs in ~2.6k cycles: the residual was fetched BEFORE the accumulator wait
      // (res_pre, see epilogue_prefetch) so that its global-load latency is not on the per-tile critical path.
      if (half < nchunks) chunk(half, has_pre, res_pre[0], res_pre[1]);
      if (half + 2 < nchunks) chunk(half + 2, has_pre, res_pre[2], res_pre[3]);
    } else {
      const uint4 z = make_uint4(0, 0, 0, 0);
#pragma unroll 1
      for (int ck = half; ck < nchunks; ck += 2) chunk(ck, false, z, z);
    }
  }
}

// Biases of every problem of the launch -> shared memory [problem][n_total], by the 256 epilogue threads only (the
// TMA and MMA warps start right after the barrier setup; the first accumulator is ~5 us away).  All (<= 6) global loads
// of a thread are issued before the first store: the former one-load-per-iteration loop in the CTA prologue cost 5
// dependent global-latency round trips per launch.  EPI_LSTM: gate biases pre-multiplied so that each gate costs one
// FFMA + ex2 + add + rcp: i, f, o: -log2(e) * b (sigmoid); g: 2 * log2(e) * b (tanh).
template <int EPI>
__device__ __forceinline__ void stage_bias(const ConvParams& p, float* bias_s) {
  if constexpr (EPI == EPI_GRAD) return;
  const int e = static_cast<int>(threadIdx.x) - 64;          // 0..255
  const int n_tot = EPI == EPI_LSTM ? 256 : p.n_total;
  const int count = p.n_prob * n_tot;                        // <= kBiasStage (launch_conv3x3)
  constexpr int kIters = kBiasStage / 256;
  float v[kIters];
#pragma unroll
  for (int k = 0; k < kIters; ++k) {
    const int i = e + k * 256;
    v[k] = 0.f;
    if (i < count) {
      const int z = i / n_tot, n = i - z * n_tot;
      const float* b = p.prob[z].bias;
      if (b) v[k] = EPI == EPI_LSTM ? b[n] * (n < 192 ? -kLog2e : 2.f * kLog2e) : b[n];
    }
  }
#pragma unroll
  for (int k = 0; k < kIters; ++k) {
    const int i = e + k * 256;
    if (i < count) bias_s[i] = v[k];
  }
  asm volatile("bar.sync 1, 256;" ::: "memory");             // epilogue warps only
}

// EPI_STORE, BN = 64: this thread's residual values (2 chunks x 16 bf16) loaded ahead of the accumulator wait.
template <int BN, int EPI>
__device__ __forceinline__ bool epilogue_prefetch(const ConvParams& p, const ConvProblem& pr, const TileCoord& tc, int y,
                                                  int x, bool valid, int half, uint4 (&res_pre)[4], float (&c_pre)[16],
                                                  int row) {
  if constexpr (EPI == EPI_LSTM) {
    // first 16 channels of this thread's half of c_{t-1} (tile-transposed: [tile][channel][128 rows])
    if (pr.c_in) {
      const float* cin = pr.c_in + (static_cast<size_t>(tc.tile_lin) * 64 + half * 32) * kTileM + row;
#pragma unroll
      for (int j = 0; j < 16; ++j) c_pre[j] = cin[j * kTileM];
    } else {
#pragma unroll
      for (int j = 0; j < 16; ++j) c_pre[j] = 0.f;
    }
    return false;
  } else if constexpr (EPI == EPI_STORE && BN == 64) {
    if (pr.res == nullptr || !valid) return false;
    const int nchunks = p.n_store >> 4;
    const size_t off = ((static_cast<size_t>(tc.img) * p.H + y) * p.W + x) * p.out_ch + tc.nt * BN;
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int ck = half + 2 * i;
      if (ck < nchunks) {
        const uint4* rp = reinterpret_cast<const uint4*>(pr.res + off + ck * 16);
        res_pre[2 * i] = rp[0];
        res_pre[2 * i + 1] = rp[1];
      }
    }
    return true;
  } else {
    return false;
  }
}

template <int BN, int EPI, int CG>
__global__ void __launch_bounds__(kNumThreads, 1)
conv3x3_tc_kernel(const __grid_constant__ ConvMaps maps, const __grid_constant__ ConvParams p) {
  using C = Cfg<BN, CG>;
  constexpr int S = C::kStages;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + S * kABytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + S * C::kStageBytes);
  uint64_t* full = bars;            // [S]  TMA -> MMA
  uint64_t* empty = bars + S;       // [S]  MMA -> TMA
  uint64_t* tfull = bars + 2 * S;   // [2]  MMA -> epilogue
  uint64_t* tempty = tfull + 2;     // [2]  epilogue -> MMA
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);
  float* bias_s = reinterpret_cast<float*>(smem + S * C::kStageBytes + 256);  // [n_prob][256], LSTM only

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int rank = CG == 2 ? static_cast<int>(cluster_ctarank()) : 0;
  const int tiles_m = p.n_img * p.tiles_y * p.tiles_x;
  const int groups = (tiles_m + CG - 1) / CG;
  const int total_items = p.n_prob * groups * p.n_tiles_n;
  const int item0 = blockIdx.x / CG, item_step = gridDim.x / CG;

  if (warp == 0 && elect_one()) {
    prefetch_tmap(&maps.act[0]);
    prefetch_tmap(&maps.w);
    for (int i = 0; i < S; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull[i], 1);
      mbar_init(&tempty[i], CG * kNumEpiWarps * 32);   // pair: the epilogue warps of both CTAs release the leader
    }
    fence_mbar_init();
  }
  if (warp == 1) {
    if constexpr (CG == 2) tmem_alloc_cg2<kTmemCols>(tmem_slot);
    else tmem_alloc<kTmemCols>(tmem_slot);
  }
  // PDL: everything above touches no global memory; the next launch may start its own prologue as our CTAs retire.
  pdl_launch_dependents();
  pdl_wait();
  tc_fence_before();
  if constexpr (CG == 2) cluster_sync();   // the peer's barriers must be initialised before anything signals them
  else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      for (int t = item0; t < total_items; t += item_step) {
        const TileCoord tc = decode_item(p, t, groups, tiles_m, CG, rank);
        const ConvProblem& pr = p.prob[tc.z];
        int wrow = pr.w_row_base + tc.nt * BN + rank * (BN / CG);
        for (int s = 0; s < pr.n_src; ++s) {
          const SrcView& sv = pr.src[s];
          const int img = sv.img_base + tc.img;
          const CUtensorMap* tm = &maps.act[sv.map];
          for (int ti = 0; ti < p.taps; ++ti) {
            const int tap = (p.taps == 9) ? ti : 4;
            const int cx = sv.mul * (tc.x0 + tap % 3 - 1) + sv.off_x;
            const int cy = sv.mul * (tc.y0 + tap / 3 - 1) + sv.off_y;
            for (int cb = 0; cb < p.kb_per_src; ++cb) {
              mbar_wait(&empty[stage], phase ^ 1);
              if constexpr (CG == 2) {
                // both CTAs' boxes are counted on the leader's barrier (its MMA warp consumes both halves)
                if (rank == 0) mbar_arrive_expect_tx(&full[stage], 2 * C::kStageBytes);
                const uint32_t bar = mapa(smem_u32(&full[stage]), 0);
                tma_load_4d_cg2(smem_a + stage * kABytes, tm, bar, sv.ch0 + cb * kBlockK, cx, cy, img);
                tma_load_2d_cg2(smem_b + stage * C::kBBytes, &maps.w, bar, 0, wrow);
              } else {
                mbar_arrive_expect_tx(&full[stage], C::kStageBytes);
                tma_load_4d(smem_a + stage * kABytes, tm, &full[stage], sv.ch0 + cb * kBlockK, cx, cy, img);
                tma_load_2d(smem_b + stage * C::kBBytes, &maps.w, &full[stage], 0, wrow);
              }
              wrow += p.n_total;
              if (++stage == S) { stage = 0; phase ^= 1; }
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer: ONE thread of the leader CTA runs
    // the whole loop (no per-K-block elect / warp sync: for N <= 144 the issue loop, not the tensor pipe, was the limit)
    if (rank == 0 && elect_one()) {
      constexpr uint32_t idesc = make_idesc_bf16(BN, 128 * CG);
      const uint64_t adesc0 = make_desc_k_sw128(smem_u32(smem_a));
      const uint64_t bdesc0 = make_desc_k_sw128(smem_u32(smem_b));
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int t = item0; t < total_items; t += item_step, ++it) {
        const int as = it & 1;
        const uint32_t aphase = (it >> 1) & 1;
        mbar_wait(&tempty[as], aphase ^ 1);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + as * kAccStride;
        const int kb_per_tile = p.prob[decode_item(p, t, groups, tiles_m, CG, 0).z].n_src * p.taps * p.kb_per_src;
        int cb = 0;
        for (int kb = 0; kb < kb_per_tile; ++kb) {
          mbar_wait(&full[stage], phase);
          tc_fence_after();
          // +32 bytes per K=16 slice inside the 128-byte swizzle row (address field is in 16-byte units)
          const uint64_t adesc = adesc0 + static_cast<uint32_t>(stage * (kABytes >> 4));
          const uint64_t bdesc = bdesc0 + static_cast<uint32_t>(stage * (C::kBBytes >> 4));
          const int nk16 = (cb == p.kb_per_src - 1) ? p.k16_last : 4;
          mma_ss<CG>(tmem_d, adesc, bdesc, idesc, kb != 0);
          if (nk16 == 4) {
            mma_ss<CG>(tmem_d, adesc + 2, bdesc + 2, idesc, 1);
            mma_ss<CG>(tmem_d, adesc + 4, bdesc + 4, idesc, 1);
            mma_ss<CG>(tmem_d, adesc + 6, bdesc + 6, idesc, 1);
          } else {
            for (int k = 1; k < nk16; ++k) mma_ss<CG>(tmem_d, adesc + 2 * k, bdesc + 2 * k, idesc, 1);
          }
          mma_commit_t<CG>(&empty[stage]);                       // frees the stage (in both CTAs of a pair)
          if (kb == kb_per_tile - 1) mma_commit_t<CG>(&tfull[as]);
          if (++cb == p.kb_per_src) cb = 0;
          if (++stage == S) { stage = 0; phase ^= 1; }
        }
      }
    }
    __syncwarp();
  } else {
    // ------------------------------------------------------------------ epilogue
    const int ew = warp - 2;
    const int quad = warp & 3;       // TMEM lane quadrant this warp may access
    const int half = ew >> 2;        // which half of the column chunks this warp handles
    const int row = quad * 32 + lane;
    const int TW = 1 << p.tw_log2;
    const int ly = row >> p.tw_log2, lx = row & (TW - 1);
    stage_bias<EPI>(p, bias_s);
    int it = 0;
    for (int t = item0; t < total_items; t += item_step, ++it) {
      const int as = it & 1;
      const uint32_t aphase = (it >> 1) & 1;
      const TileCoord tc = decode_item(p, t, groups, tiles_m, CG, rank);
      const ConvProblem& pr = p.prob[tc.z];
      const int y = tc.y0 + ly, x = tc.x0 + lx;
      const bool valid = tc.valid && (y < p.H) && (x < p.W);
      uint4 res_pre[4];
      float c_pre[16];
      const bool has_pre = epilogue_prefetch<BN, EPI>(p, pr, tc, y, x, valid, half, res_pre, c_pre, row);
      mbar_wait(&tfull[as], aphase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + as * kAccStride + (static_cast<uint32_t>(quad * 32) << 16);

      epilogue_tile<BN, EPI>(p, pr, tc, taddr, row, y, x, valid, half, bias_s, res_pre, has_pre, c_pre);
      // All TMEM reads of this accumulator stage are complete (wait::ld above): hand it back to the MMA warp.
      tc_fence_before();
      if constexpr (CG == 2) mbar_arrive_cluster_relaxed(mapa(smem_u32(&tempty[as]), 0));
      else mbar_arrive(&tempty[as]);
    }
  }

  tc_fence_before();
  if constexpr (CG == 2) {
    cluster_sync();   // neither CTA may leave (or free TMEM) while its peer still reads its operands / barriers
    if (warp == 1) tmem_dealloc_cg2<kTmemCols>(tmem_base);
  } else {
    __syncthreads();
    if (warp == 1) tmem_dealloc<kTmemCols>(tmem_base);
  }
}

// =====================================================================================================================
// Slab ("halo") variant on the padded raster (conv.h: PrGeom).  Output positions p = y * Wp + x with Wp > W: a tile is
// 128 consecutive positions, and the nine taps of a source are nine row-shifted views of ONE slab holding the image
// rows the tile touches plus one above and one below (TMA box (64 ch, Wp, rows): the columns x >= W and the rows
// outside the image arrive as zeros and serve as left / right / top / bottom padding of every row).  The slab is
// loaded once per (source, channel block) - ~4.5x less activation traffic than nine shifted boxes - and the MMA A
// descriptor simply starts o + (dy+1)*Wp + dx rows into it (o = the tile's offset inside its first row; 128 B per
// row; the 128B swizzle is a function of the absolute shared-memory address, so shifted starts stay consistent with
// what TMA wrote - measured: the descriptor's base-offset field must stay 0, setting it to the row phase gives wrong
// products).  One zeroed 1 KB guard before and after every slab catches the -1 / +1 row of the corner taps.  Weights
// stream through their own ring.  Positions with x >= W or y >= H are computed and never stored.
constexpr int kHaloMaxNA = 3;       // activation slabs in flight: 3 when the weight stages are small, else 2
constexpr int kHaloGuard = 1024;
constexpr int kHaloMaxNB = 8;       // weight stages in flight
constexpr int kHaloBudget = 200 * 1024;

// Weight stage = G consecutive taps (one row of the 3x3 stencil when G = 3): narrow-N launches spend only 32-64 tensor
// cycles per MMA, so one barrier round trip per tap would bound them; one wait + one commit per G taps does not.
__host__ __device__ constexpr int halo_group(int b_bytes) { return b_bytes <= 8192 ? 3 : 1; }
__host__ __device__ inline int halo_slab_bytes(int positions) { return (positions * 128 + 1023) & ~1023; }
__host__ __device__ inline int halo_num_b(int slab_bytes, int b_bytes, int na) {
  const int n = (kHaloBudget - na * (slab_bytes + kHaloGuard) - kHaloGuard) / b_bytes;
  return n > kHaloMaxNB ? kHaloMaxNB : n;
}
// three slabs if that still leaves >= 6 (ungrouped) weight stages
__host__ __device__ inline int halo_num_a(int slab_bytes, int b_bytes) {
  return halo_num_b(slab_bytes, b_bytes, 3) >= 6 ? 3 : 2;
}

// Resident weights: a launch with ONE problem and one N tile whose whole packed weight operand (this CTA's share) fits
// next to two activation slabs keeps it in shared memory for the life of the persistent CTA instead of re-streaming it
// from L2 for every tile.  The narrow-N launches are bound by L2 -> SMEM delivery, not by the tensor pipe: refine conv2
// (N = 64, 27 K blocks) moved 110 KB of weights + 96 KB of activations per 2.6 k-cycle tile = 5.2 KB / cycle chip-wide
// against the ~6.3 KB / cycle the L2 slices deliver (B300_MICROARCH.md: LTS throughput cap); resident: 96 KB.
__host__ __device__ inline int halo_resident_blocks(const ConvParams& p, int b_bytes, int slab_bytes) {
  if (p.w_resident == 0 || p.n_prob != 1 || p.n_tiles_n != 1) return 0;
  const int kb = p.prob[0].n_src * 9 * p.kb_per_src;
  const long long need = static_cast<long long>(kb) * b_bytes + 2LL * (slab_bytes + kHaloGuard) + kHaloGuard;
  return need <= kHaloBudget ? kb : 0;
}

template <int BN, int EPI, int CG>
__global__ void __launch_bounds__(kNumThreads, 1)
conv3x3_halo_kernel(const __grid_constant__ ConvMaps maps, const __grid_constant__ ConvParams p) {
  using C = Cfg<BN, CG>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int Wp = p.pr_wp;                          // padded-raster pitch (conv.h: PrGeom)
  const int box_bytes = p.pr_rows * Wp * 128;      // what one TMA box delivers
  const int slab_bytes = halo_slab_bytes(p.pr_rows * Wp);
  constexpr int G = halo_group(C::kBBytes);        // taps per weight stage
  constexpr int kBStage = G * C::kBBytes;
  const int RES = halo_resident_blocks(p, C::kBBytes, slab_bytes);   // > 0: K blocks held resident (all of them)
  const int NA = RES ? 2 : halo_num_a(slab_bytes, C::kBBytes);
  const int NB = RES ? 1 : halo_num_b(slab_bytes, kBStage, NA);
  // [guard][slab 0][guard] .. [slab NA-1][guard][B stage 0 .. NB-1][barriers][LSTM biases]
  uint8_t* slab0 = smem + kHaloGuard;
  const int slab_pitch = slab_bytes + kHaloGuard;
  uint8_t* smem_b = smem + kHaloGuard + NA * slab_pitch;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kHaloBudget);
  uint64_t* afull = bars;                    // [3]
  uint64_t* aempty = bars + 3;               // [3]
  uint64_t* bfull = bars + 6;                // [8]
  uint64_t* bempty = bars + 14;              // [8]
  uint64_t* tfull = bars + 22;               // [2]
  uint64_t* tempty = bars + 24;              // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 26);
  float* bias_s = reinterpret_cast<float*>(smem + kHaloBudget + 256);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int rank = CG == 2 ? static_cast<int>(cluster_ctarank()) : 0;
  const int tiles_m = p.n_img * p.tiles_y * p.tiles_x;
  const int groups = ((p.n_img + CG - 1) / CG) * p.tiles_y;   // pairs are formed across images (decode_item)
  const int total_items = p.n_prob * groups * p.n_tiles_n;
  const int item0 = blockIdx.x / CG, item_step = gridDim.x / CG;

  if (warp == 0 && elect_one()) {
    prefetch_tmap(&maps.act[0]);
    prefetch_tmap(&maps.w);
    for (int i = 0; i < kHaloMaxNA; ++i) { mbar_init(&afull[i], 1); mbar_init(&aempty[i], 1); }
    for (int i = 0; i < kHaloMaxNB; ++i) { mbar_init(&bfull[i], 1); mbar_init(&bempty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&tfull[i], 1); mbar_init(&tempty[i], CG * kNumEpiWarps * 32); }
    fence_mbar_init();
  }
  if (warp == 1) {
    if constexpr (CG == 2) tmem_alloc_cg2<kTmemCols>(tmem_slot);
    else tmem_alloc<kTmemCols>(tmem_slot);
  }
  // zero guards (read by the MMA through the async proxy, never written by TMA): 1 KB in front of the first slab, and
  // behind every slab the bytes from the end of the TMA box to the next slab (round-up slack + 1 KB)
  for (int i = threadIdx.x; i < kHaloGuard / 16; i += kNumThreads)
    reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  {
    const int tail16 = (slab_pitch - box_bytes) / 16;
    for (int i = threadIdx.x; i < NA * tail16; i += kNumThreads) {
      const int g = i / tail16, o = i - g * tail16;
      reinterpret_cast<uint4*>(slab0 + g * slab_pitch + box_bytes)[o] = make_uint4(0, 0, 0, 0);
    }
  }
  fence_proxy_async();
  pdl_launch_dependents();
  pdl_wait();
  tc_fence_before();
  if constexpr (CG == 2) cluster_sync();
  else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (elect_one()) {
      int sa = 0, sb = 0;
      uint32_t pa = 0, pb = 0;
      if (RES && item0 < total_items) {
        // the whole weight operand once: bfull[0] collects every box (both CTAs' halves on the leader's barrier)
        const int wcol = p.prob[0].w_row_base + rank * (BN / CG);
        if (CG == 1 || rank == 0) mbar_arrive_expect_tx(&bfull[0], CG * RES * C::kBBytes);
        for (int kb = 0; kb < RES; ++kb) {
          uint8_t* bdst = smem_b + kb * C::kBBytes;
          if constexpr (CG == 2) tma_load_2d_cg2(bdst, &maps.w, mapa(smem_u32(&bfull[0]), 0), 0, wcol + kb * p.n_total);
          else tma_load_2d(bdst, &maps.w, &bfull[0], 0, wcol + kb * p.n_total);
        }
      }
      for (int t = item0; t < total_items; t += item_step) {
        const TileCoord tc = decode_item(p, t, groups, tiles_m, CG, rank);
        const ConvProblem& pr = p.prob[tc.z];
        const int wcol = pr.w_row_base + tc.nt * BN + rank * (BN / CG);
        const int row0 = (tc.ty * kTileM) / Wp;    // first image row the tile touches
        for (int s = 0; s < pr.n_src; ++s) {
          const SrcView& sv = pr.src[s];
          const CUtensorMap* tm = &maps.act[sv.map];
          for (int cb = 0; cb < p.kb_per_src; ++cb) {
            mbar_wait(&aempty[sa], pa ^ 1);
            uint8_t* dst = slab0 + sa * slab_pitch;
            const int cx = sv.off_x, cy = sv.mul * (row0 - 1) + sv.off_y;   // image rows row0 - 1 .. row0 + pr_rows - 2
            if constexpr (CG == 2) {
              if (rank == 0) mbar_arrive_expect_tx(&afull[sa], 2 * box_bytes);
              tma_load_4d_cg2(dst, tm, mapa(smem_u32(&afull[sa]), 0), sv.ch0 + cb * kBlockK, cx, cy,
                              sv.img_base + tc.img);
            } else {
              mbar_arrive_expect_tx(&afull[sa], box_bytes);
              tma_load_4d(dst, tm, &afull[sa], sv.ch0 + cb * kBlockK, cx, cy, sv.img_base + tc.img);
            }
            if (++sa == NA) { sa = 0; pa ^= 1; }
            for (int tg = 0; tg < (RES ? 0 : 9 / G); ++tg) {
              mbar_wait(&bempty[sb], pb ^ 1);
              if constexpr (CG == 2) {
                if (rank == 0) mbar_arrive_expect_tx(&bfull[sb], 2 * kBStage);
              } else {
                mbar_arrive_expect_tx(&bfull[sb], kBStage);
              }
#pragma unroll
              for (int g = 0; g < G; ++g) {
                const int wrow = wcol + ((s * 9 + tg * G + g) * p.kb_per_src + cb) * p.n_total;
                uint8_t* bdst = smem_b + sb * kBStage + g * C::kBBytes;
                if constexpr (CG == 2) tma_load_2d_cg2(bdst, &maps.w, mapa(smem_u32(&bfull[sb]), 0), 0, wrow);
                else tma_load_2d(bdst, &maps.w, &bfull[sb], 0, wrow);
              }
              if (++sb == NB) { sb = 0; pb ^= 1; }
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer: one thread of the leader CTA
    if (rank == 0 && elect_one()) {
      constexpr uint32_t idesc = make_idesc_bf16(BN, 128 * CG);
      const uint32_t slab_base = smem_u32(slab0);
      const uint64_t bdesc0 = make_desc_k_sw128(smem_u32(smem_b));
      int sa = 0, sb = 0;
      uint32_t pa = 0, pb = 0;
      int it = 0;
      if (RES && item0 < total_items) {
        mbar_wait(&bfull[0], 0);          // resident weights have landed (in both CTAs of a pair)
        tc_fence_after();
      }
      for (int t = item0; t < total_items; t += item_step, ++it) {
        const int as = it & 1;
        const uint32_t aphase = (it >> 1) & 1;
        mbar_wait(&tempty[as], aphase ^ 1);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + as * kAccStride;
        const TileCoord tc0 = decode_item(p, t, groups, tiles_m, CG, 0);   // both CTAs of a pair: same row phase
        const int n_src = p.prob[tc0.z].n_src;
        // Both CTAs of a pair work on the same tile index of two images: one descriptor offset serves both.
        const int o0 = tc0.ty * kTileM - ((tc0.ty * kTileM) / Wp) * Wp;
        uint32_t acc = 0;   // the first MMA of the tile overwrites the accumulator
        for (int s = 0; s < n_src; ++s)
          for (int cb = 0; cb < p.kb_per_src; ++cb) {
            mbar_wait(&afull[sa], pa);
            // tap (dx, dy) of output position m reads slab row o0 + m + (dy + 1) * Wp + dx (row -1: the zero guard)
            const uint32_t slab = slab_base + sa * slab_pitch + (o0 - 1) * 128;
            const int nk16 = (cb == p.kb_per_src - 1) ? p.k16_last : 4;
            const bool last_a = (s == n_src - 1) && (cb == p.kb_per_src - 1);
            if (RES) {
              tc_fence_after();
              const uint64_t bsrc = bdesc0 + static_cast<uint32_t>(((s * 9 * p.kb_per_src + cb) * C::kBBytes) >> 4);
              const uint32_t bstep = static_cast<uint32_t>((p.kb_per_src * C::kBBytes) >> 4);   // next tap, same channel block
#pragma unroll
              for (int tap = 0; tap < 9; ++tap) {
                const int dyp = tap / 3, dxp = tap - 3 * dyp;
                const uint64_t adesc = make_desc_k_sw128(slab + (dyp * Wp + dxp) * 128);
                const uint64_t bdesc = bsrc + tap * bstep;
                mma_ss<CG>(tmem_d, adesc, bdesc, idesc, acc);
                acc = 1;
                if (nk16 == 4) {
                  mma_ss<CG>(tmem_d, adesc + 2, bdesc + 2, idesc, 1);
                  mma_ss<CG>(tmem_d, adesc + 4, bdesc + 4, idesc, 1);
                  mma_ss<CG>(tmem_d, adesc + 6, bdesc + 6, idesc, 1);
                } else {
                  for (int k = 1; k < nk16; ++k) mma_ss<CG>(tmem_d, adesc + 2 * k, bdesc + 2 * k, idesc, 1);
                }
              }
              mma_commit_t<CG>(&aempty[sa]);
              if (last_a) mma_commit_t<CG>(&tfull[as]);
              if (++sa == NA) { sa = 0; pa ^= 1; }
              continue;
            }
            for (int tg = 0; tg < 9 / G; ++tg) {
              mbar_wait(&bfull[sb], pb);
              tc_fence_after();
#pragma unroll
              for (int g = 0; g < G; ++g) {
                const int tap = tg * G + g;
                const int dyp = G == 3 ? tg : tap / 3, dxp = G == 3 ? g : tap - 3 * dyp;
                const uint64_t adesc = make_desc_k_sw128(slab + (dyp * Wp + dxp) * 128);
                const uint64_t bdesc = bdesc0 + static_cast<uint32_t>((sb * kBStage + g * C::kBBytes) >> 4);
                mma_ss<CG>(tmem_d, adesc, bdesc, idesc, acc);
                acc = 1;
                if (nk16 == 4) {
                  mma_ss<CG>(tmem_d, adesc + 2, bdesc + 2, idesc, 1);
                  mma_ss<CG>(tmem_d, adesc + 4, bdesc + 4, idesc, 1);
                  mma_ss<CG>(tmem_d, adesc + 6, bdesc + 6, idesc, 1);
                } else {
                  for (int k = 1; k < nk16; ++k) mma_ss<CG>(tmem_d, adesc + 2 * k, bdesc + 2 * k, idesc, 1);
                }
              }
              mma_commit_t<CG>(&bempty[sb]);
              if (++sb == NB) { sb = 0; pb ^= 1; }
            }
            mma_commit_t<CG>(&aempty[sa]);
            if (last_a) mma_commit_t<CG>(&tfull[as]);
            if (++sa == NA) { sa = 0; pa ^= 1; }
          }
      }
    }
    __syncwarp();
  } else {
    // ------------------------------------------------------------------ epilogue (identical to the box kernel)
    const int ew = warp - 2;
    const int quad = warp & 3;
    const int half = ew >> 2;
    const int row = quad * 32 + lane;
    stage_bias<EPI>(p, bias_s);
    int it = 0;
    for (int t = item0; t < total_items; t += item_step, ++it) {
      const int as = it & 1;
      const uint32_t aphase = (it >> 1) & 1;
      const TileCoord tc = decode_item(p, t, groups, tiles_m, CG, rank);
      const ConvProblem& pr = p.prob[tc.z];
      const int pos = tc.ty * kTileM + row;        // padded-raster position of this thread's output pixel
      const int y = pos / Wp, x = pos - y * Wp;
      const bool valid = tc.valid && (y < p.H) && (x < p.W);
      uint4 res_pre[4];
      float c_pre[16];
      const bool has_pre = epilogue_prefetch<BN, EPI>(p, pr, tc, y, x, valid, half, res_pre, c_pre, row);
      mbar_wait(&tfull[as], aphase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + as * kAccStride + (static_cast<uint32_t>(quad * 32) << 16);
      epilogue_tile<BN, EPI>(p, pr, tc, taddr, row, y, x, valid, half, bias_s, res_pre, has_pre, c_pre);
      tc_fence_before();
      if constexpr (CG == 2) mbar_arrive_cluster_relaxed(mapa(smem_u32(&tempty[as]), 0));
      else mbar_arrive(&tempty[as]);
    }
  }

  tc_fence_before();
  if constexpr (CG == 2) {
    cluster_sync();
    if (warp == 1) tmem_dealloc_cg2<kTmemCols>(tmem_base);
  } else {
    __syncthreads();
    if (warp == 1) tmem_dealloc<kTmemCols>(tmem_base);
  }
}

static int g_cta_pair = 1;   // 1: use CTA pairs (cta_group::2) whenever the grid allows it

template <int BN, int EPI, int CG, bool HALO>
static int launch_t(const ConvMaps& maps, const ConvParams& p, int num_sms, cudaStream_t stream) {
  using C = Cfg<BN, CG>;
  auto kern = HALO ? conv3x3_halo_kernel<BN, EPI, CG> : conv3x3_tc_kernel<BN, EPI, CG>;
  constexpr int kSmem = HALO ? kHaloBudget + 1024 + 256 + kBiasStage * 4 : C::kSmemBytes;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem);
    if (e != cudaSuccess) return static_cast<int>(e);
    attr_set = true;
  }
  const long long tiles_m = 1LL * p.n_img * p.tiles_y * p.tiles_x;
  const long long groups = HALO ? 1LL * ((p.n_img + CG - 1) / CG) * p.tiles_y : (tiles_m + CG - 1) / CG;
  const long long items = 1LL * p.n_prob * groups * p.n_tiles_n;
  if (items <= 0) return 0;
  const long long max_groups = num_sms / CG;
  const int grid = static_cast<int>((items < max_groups ? items : max_groups) * CG);
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(kNumThreads);
  cfg.dynamicSmemBytes = kSmem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  int na = 0;
  if (CG == 2) {
    attr[na].id = cudaLaunchAttributeClusterDimension;
    attr[na].val.clusterDim.x = CG;
    attr[na].val.clusterDim.y = 1;
    attr[na].val.clusterDim.z = 1;
    ++na;
  }
  if (get_pdl()) {
    attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[na].val.programmaticStreamSerializationAllowed = 1;
    ++na;
  }
  cfg.attrs = attr;
  cfg.numAttrs = na;
  return static_cast<int>(cudaLaunchKernelEx(&cfg, kern, maps, p));
}

template <int BN, int EPI>
static int launch_cg(const ConvMaps& maps, const ConvParams& p, int num_sms, cudaStream_t stream) {
  if (p.halo) {
    if (g_cta_pair && num_sms >= 2) return launch_t<BN, EPI, 2, true>(maps, p, num_sms, stream);
    return launch_t<BN, EPI, 1, true>(maps, p, num_sms, stream);
  }
  if (g_cta_pair && num_sms >= 2) return launch_t<BN, EPI, 2, false>(maps, p, num_sms, stream);
  return launch_t<BN, EPI, 1, false>(maps, p, num_sms, stream);
}

static int g_halo = 1;   // 0: nine shifted TMA boxes per source, 1: one slab per source + shifted descriptors
void set_halo_mode(int mode) { g_halo = mode; }
int get_halo_mode() { return g_halo; }

// Programmatic dependent launch between consecutive kernels.  Measured (profiles/r02/README.md): +3.5 % at one sequence
// per step (143 dependent launches of <= 162 tiles: the prologue of launch n + 1 hides behind the tail of launch n),
// -1.7 % at 8 sequences, nothing at 32 - so it is off, and inference plans of <= kPdlAutoPixels pixels per frame batch
// switch it on for their own graph capture unless the caller took the decision (pvsr_set_pdl / PVSR_PDL).
static int g_pdl = 0;
static int g_pdl_explicit = 0;
void set_pdl(int enable) { g_pdl = enable ? 1 : 0; }
int get_pdl() { return g_pdl; }
void set_pdl_explicit() { g_pdl_explicit = 1; }
int pdl_is_explicit() { return g_pdl_explicit; }

static int g_w_resident = 1;   // resident weight operand for launches where it fits (halo_resident_blocks)
void set_w_resident(int enable) { g_w_resident = enable ? 1 : 0; }
int get_w_resident() { return g_w_resident; }

static int g_pack_table = 1;
void set_pack_table(int enable) { g_pack_table = enable ? 1 : 0; }
int get_pack_table() { return g_pack_table; }

static int g_tail_rank1 = 1;
void set_tail_rank1(int enable) { g_tail_rank1 = enable ? 1 : 0; }
int get_tail_rank1() { return g_tail_rank1; }

static int g_tail_fwd = 2;   // 0: conv + shuffle + conv, 1: composite 5x5 conv on mma.sync, 2: 36-channel tcgen05 conv + gather
void set_tail_fwd(int mode) { g_tail_fwd = mode < 0 ? 0 : (mode > 2 ? 2 : mode); }
int get_tail_fwd() { return g_tail_fwd; }

static int g_two_branch = 0;   // off by default: 3 % gain, and one placement of the tail correlation hung the graph (plan.cpp)
void set_two_branch(int enable) { g_two_branch = enable ? 1 : 0; }
int get_two_branch() { return g_two_branch; }

void set_cta_pair(int enable) { g_cta_pair = enable ? 1 : 0; }
int get_cta_pair() { return g_cta_pair; }

int launch_conv3x3(int bn, int epi, const ConvMaps& maps, const ConvParams& p, int num_sms, cudaStream_t stream) {
  if ((epi == EPI_STORE || epi == EPI_PS) && p.n_prob * p.n_total > kBiasStage)   // bias staging area
    return static_cast<int>(cudaErrorInvalidValue);
  if (epi == EPI_LSTM && bn == 256) return launch_cg<256, EPI_LSTM>(maps, p, num_sms, stream);
  if (epi == EPI_STORE) {
    switch (bn) {
      case 48: return launch_cg<48, EPI_STORE>(maps, p, num_sms, stream);
      case 64: return launch_cg<64, EPI_STORE>(maps, p, num_sms, stream);
      case 144: return launch_cg<144, EPI_STORE>(maps, p, num_sms, stream);
      case 256: return launch_cg<256, EPI_STORE>(maps, p, num_sms, stream);
    }
  }
  if (epi == EPI_PS) {
    switch (bn) {
      case 192: return launch_cg<192, EPI_PS>(maps, p, num_sms, stream);
      case 256: return launch_cg<256, EPI_PS>(maps, p, num_sms, stream);
    }
  }
  if (epi == EPI_GRAD) {
    switch (bn) {
      case 64: return launch_cg<64, EPI_GRAD>(maps, p, num_sms, stream);
      case 128: return launch_cg<128, EPI_GRAD>(maps, p, num_sms, stream);
    }
  }
  return static_cast<int>(cudaErrorInvalidValue);
}

}  // namespace pvsr
