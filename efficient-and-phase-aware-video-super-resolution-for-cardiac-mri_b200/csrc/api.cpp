// C ABI (include/pvsr.h): library info, host-side packing logic and the per-op entry points.
#include <cuda.h>
#include <cuda_runtime.h>

#include <cstdarg>
#include <cstdio>
#include <algorithm>
#include <cstring>
#include <vector>

#include "../../include/pvsr.h"
#include "conv.h"
#include "internal.h"
#include "simt.h"

namespace pvsr {

static thread_local char g_err[512] = "";

int set_error(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}
int check_cuda(int e, const char* what) {
  if (e == 0) return 0;
  return set_error(e, "%s: %s", what, cudaGetErrorString(static_cast<cudaError_t>(e)));
}

int device_num_sms() {
  static int sms = 0;
  if (sms == 0) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  }
  return sms;
}

int fill_conv_params(const pvsr_conv_desc* d, ConvParams* p) {
  memset(p, 0, sizeof(*p));
  if (d->n_src < 1 || d->n_src > kMaxSrc) return set_error(-2, "n_src %d out of range", d->n_src);
  if (d->n_views < 1 || d->n_views > kMaxMaps) return set_error(-2, "n_views %d out of range", d->n_views);
  if (d->taps != 9 && d->taps != 1) return set_error(-2, "taps must be 9 or 1");
  if (d->k16_last < 1 || d->k16_last > 4) return set_error(-2, "k16_last must be 1..4");
  p->H = d->H;
  p->W = d->W;
  int max_mul = 1;
  for (int v = 0; v < d->n_views; ++v) max_mul = d->views[v].mul > max_mul ? d->views[v].mul : max_mul;
  choose_tile(d->H, d->W, &p->tw_log2, max_mul == 1 ? 128 : (max_mul == 2 ? 128 : 64));
  const int tw = 1 << p->tw_log2, th = kTileM >> p->tw_log2;
  p->tiles_x = (d->W + tw - 1) / tw;
  p->tiles_y = (d->H + th - 1) / th;
  p->n_img = static_cast<int>(d->n_img);
  p->n_prob = 1;
  p->taps = d->taps;
  p->kb_per_src = d->kb_per_src;
  p->k16_last = d->k16_last;
  p->n_tiles_n = d->n_tiles_n;
  p->n_total = d->n_tiles_n * d->bn;
  p->n_store = d->n_store;
  p->out_ch = d->out_ch;
  p->ps_r = d->ps_r;
  p->grad_split = d->grad_split;
  p->relu = d->relu;
  p->out_scale = d->out_scale;
  p->prelu = d->prelu;
  if (d->ps_r > 0) p->ps_ch = p->n_total / (d->ps_r * d->ps_r);
  ConvProblem& pr = p->prob[0];
  pr.n_src = d->n_src;
  for (int i = 0; i < d->n_src; ++i) {
    const int v = d->src_view[i];
    if (v < 0 || v >= d->n_views) return set_error(-2, "src_view out of range");
    pr.src[i] = SrcView{v, d->src_img_base[i], d->src_ch0[i], d->views[v].mul, d->src_off_x[i], d->src_off_y[i]};
  }
  pr.w_row_base = d->w_row_base;
  pr.bias = d->bias;
  pr.out_bf16 = static_cast<__nv_bfloat16*>(d->out_bf16);
  pr.out_f32 = d->out_f32;
  pr.res = static_cast<const __nv_bfloat16*>(d->res);
  pr.posterm = d->posterm;
  pr.mask = static_cast<const __nv_bfloat16*>(d->mask);
  pr.grad0 = d->grad0;
  pr.grad1 = d->grad1;
  pr.c_in = d->c_in;
  pr.c_out = d->c_out;
  pr.h_out = static_cast<__nv_bfloat16*>(d->h_out);
  pr.gates_out = static_cast<__nv_bfloat16*>(d->gates_out);
  return 0;
}

void build_wgrad_jobs(const std::vector<WgSource>& srcs, int kb_per_src, int taps, const std::vector<WgChunk>& chunks,
                      int n_total, bool with_bias, long long dw_off, long long db_off, std::vector<WgJob>* out,
                      int contig_cols) {
  std::vector<WgUnit> units;
  int kb = 0;
  for (size_t s = 0; s < srcs.size(); ++s)
    for (int ti = 0; ti < taps; ++ti) {
      const int tap = taps == 9 ? ti : 4;
      for (int cb = 0; cb < kb_per_src; ++cb, ++kb) {
        WgUnit u{};
        u.kind = 0;
        u.view = srcs[s].view;
        u.view.ch0 += cb * 64;
        u.dx = tap % 3 - 1;
        u.dy = tap / 3 - 1;
        u.out_kb = kb;
        units.push_back(u);
      }
    }
  if (with_bias) {
    WgUnit u{};
    u.kind = 1;
    units.push_back(u);
  }
  // Column groups: what the two CTAs of a pair stage of dY.  contig_cols > 0: chunks[0] is the first 64-channel view of
  // ONE tensor with contig_cols consecutive gradient channels - it is cut into two equal halves (any multiple of 8
  // columns each, e.g. 72 + 72 for the 144-channel refine conv1), so no MMA column is wasted.  Otherwise the chunks
  // are independent 64-column views (pixel-unshuffled phases of a head gradient), two per CTA.
  struct ColGroup { int n_half; int n_slots[2]; SrcView dy[2][kWgSlots]; int col0[2][kWgSlots], ncols[2][kWgSlots]; };
  std::vector<ColGroup> groups;
  if (contig_cols > 0) {
    for (int c0 = 0; c0 < contig_cols; c0 += 256) {
      const int cols = std::min(256, contig_cols - c0);
      ColGroup g{};
      g.n_half = ((cols + 15) / 16) * 8;
      for (int r = 0; r < 2; ++r) {
        const int first = c0 + r * g.n_half, last = std::min(c0 + cols, first + g.n_half);
        g.n_slots[r] = 0;
        for (int c = first; c < last; c += 64) {
          const int sl = g.n_slots[r]++;
          g.dy[r][sl] = chunks[0].view;
          g.dy[r][sl].ch0 += c;
          g.col0[r][sl] = chunks[0].col0 + c;
          g.ncols[r][sl] = std::min(64, last - c);
        }
      }
      groups.push_back(g);
    }
  } else {
    for (size_t c0 = 0; c0 < chunks.size(); c0 += 2 * kWgSlots) {
      const int n = static_cast<int>(std::min<size_t>(2 * kWgSlots, chunks.size() - c0));
      ColGroup g{};
      const int per = (n + 1) / 2;
      g.n_half = 64 * per;
      for (int r = 0; r < 2; ++r) {
        g.n_slots[r] = 0;
        for (int i = r * per; i < std::min(n, (r + 1) * per); ++i) {
          const int sl = g.n_slots[r]++;
          g.dy[r][sl] = chunks[c0 + i].view;
          g.col0[r][sl] = chunks[c0 + i].col0;
          g.ncols[r][sl] = 64;
        }
      }
      groups.push_back(g);
    }
  }
  // Unit groups of up to 2 * kWgUnits, split evenly over the pair (the leader gets the odd one).
  for (size_t u0 = 0; u0 < units.size(); u0 += 2 * kWgUnits) {
    const int n = static_cast<int>(std::min<size_t>(2 * kWgUnits, units.size() - u0));
    const int n0 = (n + 1) / 2;
    for (const ColGroup& g : groups) {
      WgJob j{};
      j.n_units[0] = n0;
      j.n_units[1] = n - n0;
      for (int i = 0; i < n0; ++i) j.unit[0][i] = units[u0 + i];
      for (int i = n0; i < n; ++i) j.unit[1][i - n0] = units[u0 + i];
      j.n_half = g.n_half;
      for (int r = 0; r < 2; ++r) {
        j.n_slots[r] = g.n_slots[r];
        for (int c = 0; c < kWgSlots; ++c) { j.dy[r][c] = g.dy[r][c]; j.col0[r][c] = g.col0[r][c]; j.ncols[r][c] = g.ncols[r][c]; }
      }
      j.n_total = n_total;
      j.dw_off = dw_off;
      j.db_off = db_off;
      out->push_back(j);
    }
  }
}

}  // namespace pvsr

using namespace pvsr;

extern "C" {

int pvsr_version(void) { return PVSR_VERSION; }
const char* pvsr_last_error(void) { return g_err; }

int pvsr_device_check(void) {
  int dev = 0, major = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return set_error(-10, "no CUDA device");
  cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
  if (major != 10) return set_error(-11, "compute capability %d.x is not sm_100", major);
  return 0;
}

int pvsr_set_cta_pair(int enable) {
  set_cta_pair(enable);
  return 0;
}
int pvsr_get_cta_pair(void) { return get_cta_pair(); }
int pvsr_set_halo_mode(int mode) {
  if (mode < 0 || mode > 1) return set_error(-2, "halo mode must be 0 or 1");
  set_halo_mode(mode);
  return 0;
}
int pvsr_get_halo_mode(void) { return get_halo_mode(); }
int pvsr_set_pdl(int enable) {
  set_pdl(enable);
  set_pdl_explicit();
  return 0;
}
int pvsr_get_pdl(void) { return get_pdl(); }
int pvsr_set_w_resident(int enable) {
  set_w_resident(enable);
  return 0;
}
int pvsr_get_w_resident(void) { return get_w_resident(); }
int pvsr_set_pack_table(int enable) {
  set_pack_table(enable);
  return 0;
}
int pvsr_get_pack_table(void) { return get_pack_table(); }
int pvsr_set_tail_rank1(int enable) {
  set_tail_rank1(enable);
  return 0;
}
int pvsr_get_tail_rank1(void) { return get_tail_rank1(); }
int64_t pvsr_head_tail_scratch_bytes(void) { return static_cast<int64_t>(tail_scratch_bytes()); }
int pvsr_head_tail_bwd(const float* dout, const void* x_bf16, const float* w2, const float* b2, const float* w3,
                       void* dx_bf16, float* dw2, float* db2, float* dw3, float* db3, void* scratch, int64_t n_img,
                       int H1, int W1, float sign_scale, void* stream) {
  if (!dout || !x_bf16 || !w2 || !b2 || !w3 || !dx_bf16 || !scratch) return set_error(-2, "null argument");
  if (n_img < 0 || H1 < 1 || W1 < 1) return set_error(-2, "bad shape");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int sms = device_num_sms();
  int e = launch_tail_tables(w2, w3, scratch, s);
  if (!e) e = launch_tail_zero_sums(scratch, s);
  if (!e) e = launch_tail_dx(dout, scratch, dx_bf16, n_img, H1, W1, sms, s, sign_scale);
  if (!e) e = launch_tail_corr(dout, x_bf16, scratch, n_img, H1, W1, sms, s, sign_scale);
  if (!e) e = launch_tail_finish(scratch, w2, b2, w3, dw2, db2, dw3, db3, s);
  return check_cuda(e, "head_tail_bwd");
}
int pvsr_set_tail_fwd(int enable) {
  set_tail_fwd(enable);
  return 0;
}
int pvsr_get_tail_fwd(void) { return get_tail_fwd(); }
int64_t pvsr_head_tail_fwd_table_bytes(void) { return static_cast<int64_t>(tail_fwd_table_bytes()); }
int pvsr_head_tail_fwd_tables(const float* w2, const float* b2, const float* w3, const float* b3, void* tables,
                              void* stream) {
  if (!w2 || !b2 || !w3 || !b3 || !tables) return set_error(-2, "null argument");
  return check_cuda(launch_tail_fwd_tables(w2, b2, w3, b3, tables, static_cast<cudaStream_t>(stream)), "tail_fwd_tables");
}
int pvsr_head_tail_fwd(const void* x_bf16, const void* tables, float* out, int64_t n_img, int H1, int W1, void* stream) {
  if (!x_bf16 || !tables || !out) return set_error(-2, "null argument");
  if (n_img < 0 || H1 < 1 || W1 < 1) return set_error(-2, "bad shape");
  return check_cuda(launch_tail_fwd(x_bf16, tables, out, n_img, H1, W1, device_num_sms(), static_cast<cudaStream_t>(stream)),
                    "tail_fwd");
}
int pvsr_set_two_branch(int enable) {
  set_two_branch(enable);
  return 0;
}
int pvsr_get_two_branch(void) { return get_two_branch(); }
int pvsr_set_head_tma(int enable) {
  set_head_tma(enable);
  return 0;
}
int pvsr_get_head_tma(void) { return get_head_tma(); }

int pvsr_choose_tile(int H, int W, int* tw_log2_out) {
  if (H <= 0 || W <= 0 || !tw_log2_out) return set_error(-2, "bad image size");
  choose_tile(H, W, tw_log2_out);
  return 0;
}

int64_t pvsr_pack_index_count(const pvsr_pack_spec* s) {
  return static_cast<int64_t>(s->n_src) * s->taps * s->kb_per_src * s->n_total * 64;
}

static int spec_out_channel(const pvsr_pack_spec* s, int n) {
  if (s->ps_r == 0) return n;
  const int r2 = s->ps_r * s->ps_r;
  const int pc = s->ps_ch > 0 ? s->ps_ch : 64;
  const int q = n / pc, c = n % pc;
  if (q >= r2) return -1;
  return c * r2 + q;
}

int pvsr_pack_index_host(const pvsr_pack_spec* s, int32_t* idx) {
  if (s->taps != 9 && s->taps != 1) return set_error(-2, "taps must be 9 or 1");
  if (s->taps == 9 && (s->kh != 3 || s->kw != 3)) return set_error(-2, "taps=9 needs a 3x3 parameter");
  if (s->n_src < 1 || s->n_src > PVSR_MAX_SRC) return set_error(-2, "n_src out of range");
  const int centre = (s->kh * s->kw == 1) ? 0 : (s->kh / 2) * s->kw + s->kw / 2;
  int64_t e = 0;
  for (int src = 0; src < s->n_src; ++src)
    for (int ti = 0; ti < s->taps; ++ti) {
      const int tap = s->taps == 9 ? ti : centre;
      int ky = tap / s->kw, kx = tap % s->kw;
      if (s->transpose_flip) { ky = s->kh - 1 - ky; kx = s->kw - 1 - kx; }
      for (int cb = 0; cb < s->kb_per_src; ++cb)
        for (int n = 0; n < s->n_total; ++n) {
          const int col = spec_out_channel(s, n);
          for (int c = 0; c < 64; ++c, ++e) {
            const int ic = cb * 64 + c;
            int v = -1;
            if (ic < s->src_ch && col >= 0) {
              // GEMM K channel -> parameter channel
              const int gin = (s->transpose_flip && s->k_ps_r > 0) ? ic * s->k_ps_r * s->k_ps_r + src
                                                                     : s->src_ch_off[src] + ic;
              const int o = s->transpose_flip ? gin : col;
              const int i = s->transpose_flip ? col + s->src_col_off[src] : gin;
              if (o < s->c_out && i < s->c_in) v = ((o * s->c_in + i) * s->kh + ky) * s->kw + kx;
            }
            idx[e] = v;
          }
        }
    }
  return 0;
}

int pvsr_pack_bias_index_host(const pvsr_pack_spec* s, int32_t* idx) {
  const int lim = s->transpose_flip ? s->c_in : s->c_out;
  for (int n = 0; n < s->n_total; ++n) {
    const int col = spec_out_channel(s, n);
    idx[n] = (col >= 0 && col < lim) ? col : -1;
  }
  return 0;
}

int pvsr_pack_weights(const float* w, const int32_t* idx, const int32_t* idx2, void* out, int64_t n, void* stream) {
  return check_cuda(launch_pack_weights(w, idx, idx2, out, n, static_cast<cudaStream_t>(stream)), "pack_weights");
}
int pvsr_gather_f32(const float* src, const int32_t* idx, float* out, int64_t n, void* stream) {
  return check_cuda(launch_gather_f32(src, idx, out, n, static_cast<cudaStream_t>(stream)), "gather_f32");
}

int pvsr_in_conv_prelu_fwd(const float* x, const float* w, const float* b, const float* slope, void* out,
                           int64_t n_img, int H, int W, void* stream) {
  return check_cuda(launch_in_conv_prelu(x, w, b, slope, out, n_img, H, W, static_cast<cudaStream_t>(stream)),
                    "in_conv_prelu");
}

int pvsr_conv3x3_fwd(const pvsr_conv_desc* d, void* stream) {
  ConvParams p;
  int rc = fill_conv_params(d, &p);
  if (rc) return rc;
  ConvMaps maps;
  memset(&maps, 0, sizeof(maps));
  const int tw = 1 << p.tw_log2, th = kTileM >> p.tw_log2;
  int max_mul = 1;
  for (int v = 0; v < d->n_views; ++v) max_mul = d->views[v].mul > max_mul ? d->views[v].mul : max_mul;
  // slab (padded-raster) kernel whenever the geometry allows it (ConvLSTM state tensors follow the same tile map:
  // conv.h lstm_tile_geometry)
  PrGeom g{0, 0, 0};
  const bool slab = get_halo_mode() != 0 && d->taps == 9 && choose_pr(d->H, d->W, max_mul, &g);
  if (d->epi == EPI_LSTM && max_mul != 1) return set_error(-2, "ConvLSTM sources must be plain views");
  int bw = tw, bh = th;
  if (slab) {
    p.halo = 1; p.pr_wp = g.wp; p.pr_rows = g.rows; p.tiles_x = 1; p.tiles_y = g.tiles;
    p.w_resident = get_w_resident();
    bw = g.wp; bh = g.rows;
  }
  for (int v = 0; v < d->n_views; ++v) {
    const pvsr_act_view& a = d->views[v];
    if (a.channels % 8 != 0) return set_error(-2, "view channels must be a multiple of 8");
    rc = make_act_tmap(&maps.act[v], a.ptr, a.channels, a.W, a.H, a.images, bw, bh, a.mul);
    if (rc) return set_error(rc, "activation tensor map encode failed (%d)", rc);
  }
  rc = make_weight_tmap(&maps.w, d->w_packed, d->w_rows, d->bn);
  if (rc) return set_error(rc, "weight tensor map encode failed (%d)", rc);
  return check_cuda(launch_conv3x3(d->bn, d->epi, maps, p, device_num_sms(), static_cast<cudaStream_t>(stream)),
                    "conv3x3");
}

constexpr int kMaxWgJobs = 1024;
int64_t pvsr_wgrad_scratch_bytes(void) { return kMaxWgJobs * static_cast<int64_t>(sizeof(WgJob)); }

int pvsr_conv3x3_wgrad(const pvsr_wgrad_desc* d, void* stream) {
  int rc = pvsr_conv3x3_wgrad_staged(d, 1, stream);
  return rc ? rc : pvsr_conv3x3_wgrad_staged(d, 0, stream);
}

int pvsr_frame_scores(const float* sr, const float* hr, const int32_t* rects, int64_t n, int H, int W, float mean,
                      float std, const float* window11, float value_range, double* sums, void* stream) {
  if (H < 1 || W < 1 || !window11 || !sums) return set_error(-2, "frame_scores: bad arguments");
  if (n > 65535) return set_error(-2, "frame_scores: at most 65535 frames per call");
  return check_cuda(launch_frame_scores(sr, hr, rects, n, H, W, mean, std, window11, value_range, sums,
                                        static_cast<cudaStream_t>(stream)), "frame_scores");
}
int pvsr_bicubic_upsample(const float* in, float* out, int64_t n_img, int h, int w, int scale, void* stream) {
  if (h < 1 || w < 1 || scale < 1) return set_error(-2, "bicubic: bad size %dx%d x%d", h, w, scale);
  return check_cuda(launch_bicubic(in, out, n_img, h, w, scale, static_cast<cudaStream_t>(stream)), "bicubic");
}
int pvsr_pad_channel_bf16(const float* x, void* out, int64_t n, void* stream) {
  return check_cuda(launch_pad_channel_bf16(x, out, n, static_cast<cudaStream_t>(stream)), "pad_channel_bf16");
}
int pvsr_take_channel0_f32(const float* in, int stride, float* out, int64_t n, void* stream) {
  if (stride < 1) return set_error(-2, "stride must be positive");
  return check_cuda(launch_take_channel0(in, stride, out, n, static_cast<cudaStream_t>(stream)), "take_channel0");
}

int pvsr_conv3x3_wgrad_staged(const pvsr_wgrad_desc* d, int upload, void* stream) {
  return pvsr_conv3x3_wgrad_multi(d, 1, upload, stream);
}

int pvsr_conv3x3_wgrad_multi(const pvsr_wgrad_desc* descs, int n_desc, int upload, void* stream) {
  if (n_desc < 1) return set_error(-2, "no wgrad descriptors");
  const pvsr_wgrad_desc* d = &descs[0];
  if (d->n_views < 1 || d->n_views > kMaxMaps) return set_error(-2, "n_views out of range");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  WgParams p{};
  p.H = d->H; p.W = d->W;
  int max_mul = 1;
  for (int v = 0; v < d->n_views; ++v) max_mul = d->views[v].mul > max_mul ? d->views[v].mul : max_mul;
  choose_tile(d->H, d->W, &p.tw_log2, max_mul <= 2 ? 128 : 64);
  const int tw = 1 << p.tw_log2, th = kTileM >> p.tw_log2;
  p.tiles_x = (d->W + tw - 1) / tw;
  p.tiles_y = (d->H + th - 1) / th;
  p.n_img = static_cast<int>(d->n_img);
  ConvMaps maps;
  memset(&maps, 0, sizeof(maps));
  for (int v = 0; v < d->n_views; ++v) {
    const pvsr_act_view& a = d->views[v];
    int rc = make_act_tmap(&maps.act[v], a.ptr, a.channels, a.W, a.H, a.images, tw, th, a.mul);
    if (rc) return set_error(rc, "tensor map encode failed (%d)", rc);
  }
  // every descriptor = one conv; all share descs[0]'s views / geometry, results are addressed relative to
  // descs[0].dw_packed (one launch reduces the weight gradients of many layers: few jobs per layer would otherwise
  // force many pixel splits, each paying a full red.add of its accumulator tile)
  std::vector<WgJob> jobs;
  for (int k = 0; k < n_desc; ++k) {
    const pvsr_wgrad_desc* e = &descs[k];
    if (e->n_src < 1 || e->n_src > kMaxSrc || e->n_dy < 1 || e->n_dy > PVSR_MAX_DY)
      return set_error(-2, "bad source count in wgrad descriptor %d", k);
    if (e->H != d->H || e->W != d->W || e->n_img != d->n_img)
      return set_error(-2, "wgrad descriptor %d differs in geometry from descriptor 0", k);
    std::vector<WgSource> srcs;
    for (int i = 0; i < e->n_src; ++i) {
      const int v = e->src_view[i];
      if (v < 0 || v >= d->n_views) return set_error(-2, "src_view out of range");
      srcs.push_back(WgSource{SrcView{v, e->src_img_base[i], e->src_ch0[i], d->views[v].mul, e->src_off_x[i], e->src_off_y[i]}});
    }
    std::vector<WgChunk> chunks;
    for (int i = 0; i < e->n_dy; ++i) {
      const int v = e->dy_view[i];
      if (v < 0 || v >= d->n_views) return set_error(-2, "dy_view out of range");
      chunks.push_back(WgChunk{SrcView{v, e->dy_img_base[i], e->dy_ch0[i], d->views[v].mul, e->dy_off_x[i], e->dy_off_y[i]}, 64 * i});
    }
    build_wgrad_jobs(srcs, e->kb_per_src, e->taps, chunks, e->n_total, e->with_bias != 0,
                     static_cast<long long>(e->dw_packed - d->dw_packed),
                     e->with_bias ? static_cast<long long>(e->db_packed - d->dw_packed) : 0, &jobs);
  }
  if (jobs.size() > static_cast<size_t>(kMaxWgJobs)) return set_error(-2, "too many wgrad jobs (%d)", static_cast<int>(jobs.size()));
  p.n_heavy = sort_wgrad_jobs(jobs.data(), static_cast<int>(jobs.size()));
  if (upload) {
    int e = cudaMemcpyAsync(d->job_scratch, jobs.data(), jobs.size() * sizeof(WgJob), cudaMemcpyHostToDevice, s);
    if (e) return check_cuda(e, "job upload");
    e = cudaStreamSynchronize(s);   // the host vector dies at return (the RefineNet plan pre-uploads its own jobs)
    return check_cuda(e, "job upload sync");
  }
  p.n_jobs = static_cast<int>(jobs.size());
  const long long total_tiles = static_cast<long long>(p.n_img) * p.tiles_x * p.tiles_y;
  if (d->n_splits > 0) {
    p.n_splits = p.n_splits_light = static_cast<int>(d->n_splits > total_tiles ? total_tiles : d->n_splits);
  } else {
    choose_wgrad_splits(p.n_heavy, p.n_jobs - p.n_heavy, total_tiles, device_num_sms(), &p.n_splits, &p.n_splits_light);
  }
  p.jobs = static_cast<const WgJob*>(d->job_scratch);
  p.grad = d->dw_packed;
  return check_cuda(launch_wgrad(maps, p, s), "wgrad");
}

int pvsr_run_table(const pvsr_table_job* jobs_dev, int n_jobs, int64_t max_n, void* stream) {
  if (n_jobs < 0 || n_jobs > 65535) return set_error(-2, "table job count out of range");
  return check_cuda(launch_table(jobs_dev, n_jobs, max_n, static_cast<cudaStream_t>(stream)), "run_table");
}

int pvsr_scatter_add(float* param_grad, const int32_t* idx, const int32_t* idx2, const float* packed, int64_t n,
                     void* stream) {
  return check_cuda(launch_scatter_add(param_grad, idx, idx2, packed, n, static_cast<cudaStream_t>(stream)),
                    "scatter_add");
}

int pvsr_scatter_add_scaled(float* param_grad, const int32_t* idx, const float* packed, int64_t n, float scale,
                            void* stream) {
  return check_cuda(launch_scatter_add(param_grad, idx, nullptr, packed, n, static_cast<cudaStream_t>(stream), scale),
                    "scatter_add_scaled");
}

int64_t pvsr_lstm_state_elems(int64_t n_img, int H, int W) {
  // large enough for either tile map (the A/B switch pvsr_set_halo_mode may change between plan creation and run)
  int l;
  choose_tile(H, W, &l);
  const int tw = 1 << l, th = kTileM >> l;
  long long tiles = static_cast<long long>((W + tw - 1) / tw) * ((H + th - 1) / th);
  PrGeom g{0, 0, 0};
  if (choose_pr(H, W, 1, &g) && g.tiles > tiles) tiles = g.tiles;
  return n_img * tiles * 64 * kTileM;
}

int pvsr_lstm_tile_geometry(int H, int W, int* wp, int* tiles_per_img) {
  if (H <= 0 || W <= 0 || !wp || !tiles_per_img) return set_error(-2, "bad image size");
  lstm_tile_geometry(H, W, wp, tiles_per_img);
  return 0;
}

int pvsr_refine_posterm(const float* w1, const float* b1, const float* pos, float* table, int n_frames_out, int B,
                        int L, int window, int c_out, int c_in, int feat2, int n_total, void* stream) {
  return check_cuda(launch_posterm(w1, b1, pos, table, n_frames_out, B, L, window, c_out, c_in, feat2, n_total,
                                   static_cast<cudaStream_t>(stream)),
                    "posterm");
}

int pvsr_head_conv_last_fwd(const void* in, const float* w, const float* b, float* out, const float* target,
                            float* l1_partial, int64_t n_img, int H, int W, void* stream) {
  return check_cuda(
      launch_head_conv_last(in, w, b, out, target, l1_partial, n_img, H, W, static_cast<cudaStream_t>(stream)),
      "head_conv_last");
}

int pvsr_cine_gather(const void* volumes, int vol_dtype, const pvsr_cine_sample* samples, int n_samples, int n_frames,
                     int h, int w, double mean, double std, float* out, const float* pos_codes, float* pos_out,
                     void* stream) {
  if (vol_dtype < PVSR_DT_F32 || vol_dtype > PVSR_DT_F64) return set_error(-2, "unknown volume dtype %d", vol_dtype);
  if (static_cast<long long>(n_samples) * n_frames > 65535)
    return set_error(-2, "cine_gather: n_samples * n_frames = %lld exceeds 65535", static_cast<long long>(n_samples) * n_frames);
  if (std == 0.0) return set_error(-2, "cine_gather: std must be non-zero");
  return check_cuda(launch_cine_gather(volumes, vol_dtype, samples, n_samples, n_frames, h, w, mean, std, out,
                                       pos_codes, pos_out, static_cast<cudaStream_t>(stream)), "cine_gather");
}

int pvsr_add_bf16(const void* a, const void* b, void* out, int64_t n_elems, void* stream) {
  return check_cuda(launch_add_bf16(a, b, out, n_elems, static_cast<cudaStream_t>(stream)), "add_bf16");
}

int pvsr_lstm_cell_bwd_pointwise(const float* dh, const void* gates, const float* c, const float* c_prev, float* dc,
                                 int dc_zero, void* dgates, int64_t n_img, int H, int W, void* stream) {
  LstmBwdParams p{};
  p.H = H; p.W = W;
  choose_tile(H, W, &p.tw_log2);
  const int tw = 1 << p.tw_log2, th = kTileM >> p.tw_log2;
  p.tiles_x = (W + tw - 1) / tw;
  p.tiles_y = (H + th - 1) / th;
  lstm_tile_geometry(H, W, &p.wp, &p.tiles_per_img);
  p.n_img = static_cast<int>(n_img);
  p.n_prob = 1;
  p.prob[0] = LstmBwdProb{dh, gates, c, c_prev, dc, dc_zero, dgates};
  return check_cuda(launch_lstm_bwd_pointwise(p, static_cast<cudaStream_t>(stream)), "lstm_bwd_pointwise");
}

int pvsr_l1_multistage(const float* out, const float* target, const float* w, int n_lists, int64_t n_per_list,
                       float* loss, float* dout, void* stream) {
  return check_cuda(launch_l1_multistage(out, target, w, n_lists, n_per_list, loss, dout, device_num_sms(),
                                         static_cast<cudaStream_t>(stream)), "l1_multistage");
}

int pvsr_head_conv_last_bwd_data(const float* dout, const float* w, void* din, int64_t n_img, int H, int W,
                                 void* stream) {
  return check_cuda(launch_head_last_bwd_data(dout, w, din, n_img, H, W, static_cast<cudaStream_t>(stream)),
                    "head_last_bwd_data");
}

int pvsr_head_conv_last_bwd_weight(const void* in, const float* dout, float* dw, float* db, int64_t n_img, int H,
                                   int W, void* stream) {
  return check_cuda(launch_head_last_bwd_weight(in, dout, dw, db, n_img, H, W, device_num_sms(),
                                                static_cast<cudaStream_t>(stream)), "head_last_bwd_weight");
}

int pvsr_in_conv_prelu_bwd(const float* x, const float* w, const float* b, const float* slope, const float* g,
                           float* dw, float* db, float* dslope, int64_t n_img, int H, int W, void* stream) {
  return check_cuda(launch_in_conv_prelu_bwd(x, w, b, slope, g, dw, db, dslope, n_img, H, W, device_num_sms(),
                                             static_cast<cudaStream_t>(stream)), "in_conv_prelu_bwd");
}

int pvsr_refine_posterm_bwd(const void* g, const float* pos, float* sums, float* dw1, int n_frames, int B, int L,
                            int frame0, int window, int H, int W, int c_out, int c_in, int feat2, int ch,
                            void* stream) {
  if (ch > 192) return set_error(-2, "posterm_bwd supports at most 192 channels");
  return check_cuda(launch_posterm_bwd(g, pos, sums, dw1, n_frames, B, L, frame0, window, H, W, c_out, c_in, feat2, ch,
                                       static_cast<cudaStream_t>(stream)), "posterm_bwd");
}

int pvsr_prelu_fwd_bf16(const void* z, const float* slope, void* y, int64_t n, void* stream) {
  if (n % 8 != 0) return set_error(-2, "prelu: n must be a multiple of 8");
  return check_cuda(launch_prelu_fwd(z, slope, y, n, device_num_sms(), static_cast<cudaStream_t>(stream)), "prelu_fwd");
}
int pvsr_prelu_bwd_bf16(const void* g, const void* z, const float* slope, void* dz, float* dslope, int64_t n,
                        void* stream) {
  if (n % 8 != 0) return set_error(-2, "prelu: n must be a multiple of 8");
  return check_cuda(launch_prelu_bwd(g, z, slope, dz, dslope, n, device_num_sms(), static_cast<cudaStream_t>(stream)),
                    "prelu_bwd");
}
int pvsr_cast_f32_bf16(const float* in, void* out, int64_t n, void* stream) {
  if (n % 8 != 0) return set_error(-2, "cast needs a multiple of 8 elements");
  return check_cuda(launch_cast_f32_bf16(in, out, n, static_cast<cudaStream_t>(stream)), "cast_f32_bf16");
}

int pvsr_adam_step(float* p, const float* g, float* m, float* v, int64_t n, float lr, float beta1, float beta2,
                   float eps, float weight_decay, float grad_scale, float* state, void* stream) {
  if (n % 4 != 0) return set_error(-2, "adam needs a multiple of 4 elements (pad the flat buffer)");
  return check_cuda(launch_adam(p, g, m, v, n, lr, beta1, beta2, eps, weight_decay, grad_scale, state,
                                device_num_sms(), static_cast<cudaStream_t>(stream)), "adam");
}

}  // extern "C"
