// Whole-network plan: RefineNet.forward (reference src/model/nets/refine_net.py:61-135) as a fixed schedule of
// kernel launches over a caller-provided workspace.  Creation is host-only (geometry, workspace layout, packing
// indices); the forward enqueues the schedule on the caller's stream, optionally as a replayed CUDA graph.
//
// Schedule per stage:
//   ConvLSTM   : wavefront launches d = 0..L+NL-2; launch d runs every cell (dir, layer l, step t = d - l) - up to
//                2*NL independent cells - as ONE tcgen05 launch (ConvParams.prob[]).       refine_net.py:81-93
//   refine     : conv1 over all (L-4) windows in one launch, sources gathered by TMA straight from the hidden
//                stacks (the 645-channel concat of :166-177 is never materialised), pos code as border-class
//                table; conv2 + bias + residual writes the next-stage features.             refine_net.py:94,132
//   heads      : conv+PixelShuffle launches over all T frames of a list at once, last conv 64->1. refine_net.py:100-113
#include <cuda.h>
#include <cuda_runtime.h>

#include <cmath>
#include <cstdint>
#include <cstring>
#include <map>
#include <vector>

#include "conv.h"
#include "internal.h"
#include "simt.h"

using namespace pvsr;

namespace {

constexpr int kFeat = 64;
constexpr int kNumClasses = 7;
enum LaunchClass { CLS_IN = 0, CLS_LSTM = 1, CLS_CONV1 = 2, CLS_CONV2 = 3, CLS_HEAD_PS = 4, CLS_HEAD_LAST = 5, CLS_MISC = 6 };

inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

struct Tiling {
  int tw_log2, tw, th, tiles_x, tiles_y;
  void set(int H, int W) {
    choose_tile(H, W, &tw_log2);
    tw = 1 << tw_log2;
    th = kTileM >> tw_log2;
    tiles_x = (W + tw - 1) / tw;
    tiles_y = (H + th - 1) / th;
  }
};

struct GraphKey {
  const void* p[6];
  bool operator<(const GraphKey& o) const { return memcmp(p, o.p, sizeof(p)) < 0; }
};

}  // namespace

struct pvsr_plan {
  pvsr_net_config cfg;
  int B, L, U, T, h, w, S, Wn, half, NL, scale;
  int n_win;             // frames with a refine map = L - 2*half
  int n_lists;
  Tiling lr;             // LR tiling
  int n_ps;              // conv+PixelShuffle layers in the head
  int ps_r[PVSR_MAX_HEAD_CONVS];
  int ps_bn[PVSR_MAX_HEAD_CONVS], ps_nt[PVSR_MAX_HEAD_CONVS];
  int ps_h[PVSR_MAX_HEAD_CONVS + 1], ps_w[PVSR_MAX_HEAD_CONVS + 1];  // resolution at the input of head conv k
  Tiling ps_tile[PVSR_MAX_HEAD_CONVS];
  int Hs, Ws;
  int lstm_src;          // sources per cell: 2 (memory) or 1
  int lstm_rows_per_cell;

  // ---- workspace layout (bytes)
  size_t img_bytes;      // one LR image of 64 bf16 channels
  long long act_images;  // images in the LR 64-channel region
  long long img_x[8];    // image base of X[s] (s = 0..S)
  long long img_h[2][PVSR_MAX_LAYERS];
  long long img_sum[2];
  size_t off_act, off_mid, off_c, off_posterm, off_head[PVSR_MAX_HEAD_CONVS], ws_bytes;
  size_t c_elems;
  int mid_ch;

  // ---- packed-parameter layout (bytes)
  size_t pk_lstm_w, pk_lstm_b, pk_c1_w, pk_c2_w, pk_c2_b, pk_head_w[PVSR_MAX_HEAD_CONVS], pk_head_b[PVSR_MAX_HEAD_CONVS];
  size_t pk_idx, pk_bytes;
  long long c1_rows, c2_rows, head_rows[PVSR_MAX_HEAD_CONVS];
  // gather jobs: (param slot, element count, dst offset, idx offset, optional idx2 offset)
  struct PackJob { int kind, a, b; long long n; size_t dst, idx, idx2; bool has2, is_bias; };
  std::vector<PackJob> jobs;
  std::vector<int32_t> idx_host;
  const void* idx_uploaded_for = nullptr;

  // ---- device-dependent state
  int num_sms = 0;
  const void* maps_ws = nullptr;
  const void* maps_pk = nullptr;
  ConvMaps maps_lstm, maps_c1, maps_c2, maps_head[PVSR_MAX_HEAD_CONVS];   // act[0] + packed weights of each launch kind
  std::map<GraphKey, cudaGraphExec_t> graphs;
  cudaStream_t cap_stream = nullptr;   // capture happens here (the caller's stream may be the legacy stream)

  // ---- accounting (filled by a dry run at creation)
  long long launches[kNumClasses];
  double flops[kNumClasses];
};

namespace {

// ------------------------------------------------------------------------------------------------ creation helpers
void add_pack_job(pvsr_plan* p, int kind, int a, int b, const pvsr_pack_spec& spec, const pvsr_pack_spec* spec2,
                  size_t dst_w, size_t dst_b) {
  pvsr_plan::PackJob j{};
  j.kind = kind; j.a = a; j.b = b;
  j.n = pvsr_pack_index_count(&spec);
  j.dst = dst_w;
  j.idx = p->idx_host.size();
  p->idx_host.resize(p->idx_host.size() + j.n);
  pvsr_pack_index_host(&spec, p->idx_host.data() + j.idx);
  j.has2 = spec2 != nullptr;
  if (spec2) {
    j.idx2 = p->idx_host.size();
    p->idx_host.resize(p->idx_host.size() + j.n);
    pvsr_pack_index_host(spec2, p->idx_host.data() + j.idx2);
  }
  j.is_bias = false;
  p->jobs.push_back(j);
  if (dst_b != static_cast<size_t>(-1)) {
    pvsr_plan::PackJob bj{};
    bj.kind = kind; bj.a = a; bj.b = b;
    bj.n = spec.n_total;
    bj.dst = dst_b;
    bj.idx = p->idx_host.size();
    p->idx_host.resize(p->idx_host.size() + bj.n);
    pvsr_pack_bias_index_host(&spec, p->idx_host.data() + bj.idx);
    bj.is_bias = true;
    p->jobs.push_back(bj);
  }
}

pvsr_pack_spec make_spec(int c_out, int c_in, int k, int n_src, const int* offs, int src_ch, int kb, int taps,
                         int n_total, int ps_r) {
  pvsr_pack_spec s{};
  s.c_out = c_out; s.c_in = c_in; s.kh = k; s.kw = k; s.n_src = n_src;
  for (int i = 0; i < n_src; ++i) s.src_ch_off[i] = offs[i];
  s.src_ch = src_ch; s.kb_per_src = kb; s.taps = taps; s.n_total = n_total; s.ps_r = ps_r; s.transpose_flip = 0;
  s.k_ps_r = 0;
  return s;
}

// Parameter slots for PackJob.kind
enum { PK_LSTM = 0, PK_C1 = 1, PK_C2 = 2, PK_HEAD = 3 };

const float* job_weight(const pvsr_plan::PackJob& j, const pvsr_net_params* P, bool bias) {
  switch (j.kind) {
    case PK_LSTM: return bias ? P->lstm_b[j.a][j.b] : P->lstm_w[j.a][j.b];
    case PK_C1: return bias ? P->ref_b1 : P->ref_w1;
    case PK_C2: return bias ? P->ref_b2 : P->ref_w2;
    case PK_HEAD: return bias ? P->head_b[j.a] : P->head_w[j.a];
  }
  return nullptr;
}

// ------------------------------------------------------------------------------------------------ launch context
struct Ctx {
  pvsr_plan* p;
  bool dry;                 // count launches / FLOPs only
  const pvsr_net_params* P;
  uint8_t* ws;
  const uint8_t* pk;
  const float* lr;
  const float* pos;
  float* out;
  cudaStream_t stream;
  // optional per-launch timing
  std::vector<cudaEvent_t>* events = nullptr;
  std::vector<int>* event_cls = nullptr;
  int rc = 0;

  void begin(int cls) {
    (void)cls;
    if (events && !dry) {
      cudaEvent_t e;
      cudaEventCreate(&e);
      cudaEventRecord(e, stream);
      events->push_back(e);
    }
  }
  void end(int cls, double fl) {
    p->launches[cls] += dry ? 1 : 0;
    p->flops[cls] += dry ? fl : 0.0;
    if (events && !dry) {
      cudaEvent_t e;
      cudaEventCreate(&e);
      cudaEventRecord(e, stream);
      events->push_back(e);
      event_cls->push_back(cls);
    }
  }
};

__nv_bfloat16* act_img(const Ctx& c, long long img) {
  return reinterpret_cast<__nv_bfloat16*>(c.ws + c.p->off_act + static_cast<size_t>(img) * c.p->img_bytes);
}

void base_params(const pvsr_plan* p, const Tiling& t, int H, int W, ConvParams* cp) {
  memset(cp, 0, sizeof(*cp));
  cp->H = H; cp->W = W;
  cp->tw_log2 = t.tw_log2; cp->tiles_x = t.tiles_x; cp->tiles_y = t.tiles_y;
  cp->taps = 9; cp->kb_per_src = 1; cp->k16_last = 4; cp->n_tiles_n = 1;
  (void)p;
}

inline SrcView view0(long long img_base) { return SrcView{0, static_cast<int>(img_base), 0, 1, 0, 0}; }

void run_conv(Ctx& c, int cls, int bn, int epi, const ConvMaps& maps, const ConvParams& cp, double fl) {
  if (c.rc) return;
  c.begin(cls);
  if (!c.dry) {
    int e = launch_conv3x3(bn, epi, maps, cp, c.p->num_sms, c.stream);
    if (e) c.rc = check_cuda(e, "conv3x3 launch");
  }
  c.end(cls, fl);
}

void run_add(Ctx& c, long long img_a, long long img_b, long long img_out, long long n_images) {
  if (c.rc) return;
  c.begin(CLS_MISC);
  if (!c.dry) {
    int e = launch_add_bf16(act_img(c, img_a), act_img(c, img_b), act_img(c, img_out),
                            n_images * static_cast<long long>(c.p->img_bytes / 2), c.stream);
    if (e) c.rc = check_cuda(e, "add_bf16 launch");
  }
  c.end(CLS_MISC, 0.0);
}

// The whole forward schedule.  With c.dry it only counts.
void schedule(Ctx& c) {
  pvsr_plan* p = c.p;
  const int B = p->B, L = p->L, U = p->U, T = p->T, S = p->S, NL = p->NL, half = p->half;
  const long long px = static_cast<long long>(p->h) * p->w;
  const double lstm_fl = 2.0 * 9 * 2 * kFeat * 4 * kFeat;  // per pixel per cell (K = 1152 incl. the zero h at t=0)

  // in_block over every frame (refine_net.py:66-67,74-79)
  c.begin(CLS_IN);
  if (!c.dry) {
    int e = launch_in_conv_prelu(c.lr, c.P->in_w, c.P->in_b, c.P->in_slope, act_img(c, p->img_x[0]),
                                 static_cast<long long>(L) * B, p->h, p->w, c.stream);
    if (e) c.rc = check_cuda(e, "in_conv launch");
  }
  c.end(CLS_IN, 2.0 * 9 * kFeat * px * L * B);

  float* posterm = reinterpret_cast<float*>(c.ws + p->off_posterm);
  if (p->cfg.pos_enc) {
    c.begin(CLS_MISC);
    if (!c.dry) {
      int e = launch_posterm(c.P->ref_w1, c.P->ref_b1, c.pos, posterm, p->n_win, B, L, p->Wn, 2 * kFeat + 1,
                             (2 * kFeat + 1) * p->Wn, 2 * kFeat, 144, c.stream);
      if (e) c.rc = check_cuda(e, "posterm launch");
    }
    c.end(CLS_MISC, 0.0);
  }

  for (int s = 0; s < S && !c.rc; ++s) {
    // ---------------------------------------------------------------- bidirectional ConvLSTM wavefront
    for (int d = 0; d < L + NL - 1; ++d) {
      ConvParams cp;
      base_params(p, p->lr, p->h, p->w, &cp);
      cp.n_img = B;
      cp.n_total = 256; cp.n_store = 256; cp.out_ch = kFeat;
      int np = 0;
      for (int dir = 0; dir < 2; ++dir)
        for (int l = 0; l < NL; ++l) {
          const int t = d - l;
          if (t < 0 || t >= L) continue;
          const int j = dir == 0 ? t : L - 1 - t;
          const int jp = dir == 0 ? j - 1 : j + 1;
          ConvProblem& pr = cp.prob[np++];
          const long long xin = (l == 0 ? p->img_x[s] : p->img_h[dir][l - 1]) + static_cast<long long>(j) * B;
          pr.src[0] = view0(xin);
          pr.n_src = 1;
          if (p->lstm_src == 2 && t > 0) {
            pr.src[1] = view0(p->img_h[dir][l] + static_cast<long long>(jp) * B);
            pr.n_src = 2;
          }
          const int ci = dir * NL + l;
          pr.w_row_base = ci * p->lstm_rows_per_cell;
          pr.bias = reinterpret_cast<const float*>(c.pk + p->pk_lstm_b) + ci * 256;
          float* cst = reinterpret_cast<float*>(c.ws + p->off_c) + static_cast<size_t>(ci) * p->c_elems;
          pr.c_in = t > 0 ? cst : nullptr;
          pr.c_out = cst;
          pr.h_out = act_img(c, p->img_h[dir][l] + static_cast<long long>(j) * B);
        }
      cp.n_prob = np;
      run_conv(c, CLS_LSTM, 256, EPI_LSTM, p->maps_lstm, cp, lstm_fl * px * B * np);
    }

    // ---------------------------------------------------------------- refine block -> next-stage features
    const long long hf_top = p->img_h[0][NL - 1], hb_top = p->img_h[1][NL - 1];
    {
      ConvParams cp;
      base_params(p, p->lr, p->h, p->w, &cp);
      cp.n_img = p->n_win * B;
      ConvProblem& pr = cp.prob[0];
      cp.n_prob = 1;
      pr.n_src = 2 * p->Wn;
      for (int jw = 0; jw < p->Wn; ++jw) {
        pr.src[2 * jw] = view0(hf_top + static_cast<long long>(jw) * B);
        pr.src[2 * jw + 1] = view0(hb_top + static_cast<long long>(jw) * B);
      }
      const __nv_bfloat16* res = act_img(c, p->img_x[s] + static_cast<long long>(half) * B);
      __nv_bfloat16* xnext = act_img(c, p->img_x[s + 1] + static_cast<long long>(half) * B);
      if (p->cfg.pos_enc) {
        cp.n_total = 144; cp.n_store = 144; cp.out_ch = 144;
        pr.posterm = posterm;
        pr.out_bf16 = reinterpret_cast<__nv_bfloat16*>(c.ws + p->off_mid);
        run_conv(c, CLS_CONV1, 144, EPI_STORE, p->maps_c1, cp,
                 2.0 * 9 * (2 * kFeat + 1) * p->Wn * (2 * kFeat + 1) * px * cp.n_img);
        ConvParams c2;
        base_params(p, p->lr, p->h, p->w, &c2);
        c2.n_img = p->n_win * B;
        c2.n_prob = 1;
        c2.kb_per_src = 3; c2.k16_last = 1;
        c2.n_total = 64; c2.n_store = 64; c2.out_ch = 64;
        ConvProblem& p2 = c2.prob[0];
        p2.n_src = 1;
        p2.src[0] = view0(0);
        p2.bias = reinterpret_cast<const float*>(c.pk + p->pk_c2_b);
        p2.res = res;
        p2.out_bf16 = xnext;
        run_conv(c, CLS_CONV2, 64, EPI_STORE, p->maps_c2, c2,
                 2.0 * 9 * (2 * kFeat + 1) * kFeat * px * c2.n_img);
      } else {
        cp.taps = 1;
        cp.n_total = 64; cp.n_store = 64; cp.out_ch = 64;
        pr.bias = reinterpret_cast<const float*>(c.pk + p->pk_c2_b);
        pr.res = res;
        pr.out_bf16 = xnext;
        run_conv(c, CLS_CONV1, 64, EPI_STORE, p->maps_c1, cp,
                 2.0 * (2 * kFeat) * p->Wn * kFeat * px * cp.n_img);
      }
    }

    // ---------------------------------------------------------------- heads (refine_net.py:100-113)
    for (int k = 0; k < 3; ++k) {
      int list;
      if (p->cfg.all_heads) list = 3 * s + k;
      else if (s == S - 1 && k == 2) list = 0;
      else continue;
      long long in_img;
      if (k == 2) {
        in_img = p->img_x[s + 1] + static_cast<long long>(U) * B;   // x_j + r_j == next-stage feature
      } else {
        in_img = p->img_sum[k];
        run_add(c, p->img_x[s] + static_cast<long long>(U) * B,
                (k == 0 ? hf_top : hb_top) + static_cast<long long>(U) * B, in_img, static_cast<long long>(T) * B);
      }
      const long long n_head = static_cast<long long>(T) * B;
      for (int q = 0; q < p->n_ps; ++q) {
        ConvParams cp;
        base_params(p, q == 0 ? p->lr : p->ps_tile[q], p->ps_h[q], p->ps_w[q], &cp);
        cp.n_img = static_cast<int>(n_head);
        cp.n_prob = 1;
        cp.n_tiles_n = p->ps_nt[q];
        cp.n_total = p->ps_bn[q] * p->ps_nt[q];
        cp.n_store = p->ps_bn[q];
        cp.out_ch = kFeat;
        cp.ps_r = p->ps_r[q];
        ConvProblem& pr = cp.prob[0];
        pr.n_src = 1;
        pr.src[0] = view0(q == 0 ? in_img : 0);
        pr.bias = reinterpret_cast<const float*>(c.pk + p->pk_head_b[q]);
        pr.out_bf16 = reinterpret_cast<__nv_bfloat16*>(c.ws + p->off_head[q]);
        run_conv(c, CLS_HEAD_PS, p->ps_bn[q], EPI_PS, p->maps_head[q], cp,
                 2.0 * 9 * kFeat * (kFeat * p->ps_r[q] * p->ps_r[q]) * p->ps_h[q] * p->ps_w[q] * n_head);
      }
      c.begin(CLS_HEAD_LAST);
      if (!c.dry && !c.rc) {
        float* o = c.out + static_cast<size_t>(list) * T * B * p->Hs * p->Ws;
        int e = launch_head_conv_last(c.ws + p->off_head[p->n_ps - 1], c.P->head_w[p->n_ps], c.P->head_b[p->n_ps], o,
                                      nullptr, nullptr, n_head, p->Hs, p->Ws, c.stream);
        if (e) c.rc = check_cuda(e, "head_conv_last launch");
      }
      c.end(CLS_HEAD_LAST, 2.0 * 9 * kFeat * static_cast<double>(p->Hs) * p->Ws * n_head);
    }

    // ---------------------------------------------------------------- edge-frame feature updates (:120-131)
    if (s + 1 < S) {
      run_add(c, p->img_x[s], hf_top, p->img_x[s + 1], static_cast<long long>(half) * B);
      run_add(c, p->img_x[s] + static_cast<long long>(L - half) * B, hb_top + static_cast<long long>(L - half) * B,
              p->img_x[s + 1] + static_cast<long long>(L - half) * B, static_cast<long long>(half) * B);
    }
  }
}

int build_maps(pvsr_plan* p, const void* ws, const void* pk) {
  if (p->maps_ws == ws && p->maps_pk == pk) return 0;
  const uint8_t* w = static_cast<const uint8_t*>(ws);
  const uint8_t* k = static_cast<const uint8_t*>(pk);
  int rc = 0;
  CUtensorMap tm_act;
  rc |= make_act_tmap(&tm_act, w + p->off_act, kFeat, p->w, p->h, p->act_images, p->lr.tw, p->lr.th);
  p->maps_lstm.act[0] = tm_act;
  p->maps_c1.act[0] = tm_act;
  p->maps_head[0].act[0] = tm_act;
  if (p->cfg.pos_enc)
    rc |= make_act_tmap(&p->maps_c2.act[0], w + p->off_mid, 144, p->w, p->h, static_cast<long long>(p->n_win) * p->B,
                        p->lr.tw, p->lr.th);
  for (int q = 1; q < p->n_ps; ++q)
    rc |= make_act_tmap(&p->maps_head[q].act[0], w + p->off_head[q - 1], kFeat, p->ps_w[q], p->ps_h[q],
                        static_cast<long long>(p->T) * p->B, p->ps_tile[q].tw, p->ps_tile[q].th);
  rc |= make_weight_tmap(&p->maps_lstm.w, k + p->pk_lstm_w, static_cast<long long>(2 * p->NL) * p->lstm_rows_per_cell,
                         256);
  rc |= make_weight_tmap(&p->maps_c1.w, k + p->pk_c1_w, p->c1_rows, p->cfg.pos_enc ? 144 : 64);
  if (p->cfg.pos_enc) rc |= make_weight_tmap(&p->maps_c2.w, k + p->pk_c2_w, p->c2_rows, 64);
  for (int q = 0; q < p->n_ps; ++q)
    rc |= make_weight_tmap(&p->maps_head[q].w, k + p->pk_head_w[q], p->head_rows[q], p->ps_bn[q]);
  if (rc) return set_error(-20, "tensor map encode failed");
  p->maps_ws = ws;
  p->maps_pk = pk;
  for (auto& g : p->graphs) cudaGraphExecDestroy(g.second);
  p->graphs.clear();
  return 0;
}

}  // namespace

extern "C" {

int pvsr_plan_create(const pvsr_net_config* cfg, pvsr_plan** out) {
  if (!cfg || !out) return set_error(-2, "null argument");
  if (cfg->scale != 2 && cfg->scale != 3 && cfg->scale != 4 && cfg->scale != 8)
    return set_error(-3, "The upscale factor should be 2, 3, 4 or 8. Got %d.", cfg->scale);
  if (cfg->n_updated <= 0) return set_error(-3, "num_updated_frames must be > 0 (reference crashes on 0)");
  if (cfg->window < 1 || cfg->window % 2 == 0 || 2 * cfg->window > PVSR_MAX_SRC)
    return set_error(-3, "refine_window_size must be odd and <= %d", PVSR_MAX_SRC / 2);
  if (cfg->n_layers < 1 || cfg->n_layers > 3) return set_error(-3, "1..3 ConvLSTM layers supported");
  if (cfg->n_frames - 2 * cfg->n_updated < 1) return set_error(-3, "need at least one target frame");
  if (cfg->n_frames - 2 * (cfg->window / 2) < 1) return set_error(-3, "sequence shorter than the refine window");
  if (cfg->n_stages < 1 || cfg->n_stages > 7) return set_error(-3, "1..7 stages supported");
  if (cfg->batch < 1 || cfg->h < 1 || cfg->w < 1) return set_error(-3, "bad batch / frame size");
  if (cfg->save_for_backward) return set_error(-3, "save_for_backward: use the training plan API");

  pvsr_plan* p = new pvsr_plan();
  p->cfg = *cfg;
  p->B = cfg->batch; p->L = cfg->n_frames; p->U = cfg->n_updated; p->T = p->L - 2 * p->U;
  p->h = cfg->h; p->w = cfg->w; p->S = cfg->n_stages; p->Wn = cfg->window; p->half = p->Wn / 2;
  p->NL = cfg->n_layers; p->scale = cfg->scale;
  p->n_win = p->L - 2 * p->half;
  p->n_lists = cfg->all_heads ? 3 * p->S : 1;
  p->lr.set(p->h, p->w);
  if (p->scale == 3) { p->n_ps = 1; p->ps_r[0] = 3; }
  else { p->n_ps = static_cast<int>(std::lround(std::log2(p->scale))); for (int q = 0; q < p->n_ps; ++q) p->ps_r[q] = 2; }
  p->ps_h[0] = p->h; p->ps_w[0] = p->w;
  for (int q = 0; q < p->n_ps; ++q) {
    p->ps_h[q + 1] = p->ps_h[q] * p->ps_r[q];
    p->ps_w[q + 1] = p->ps_w[q] * p->ps_r[q];
    p->ps_tile[q].set(p->ps_h[q], p->ps_w[q]);
    if (p->ps_r[q] == 2) { p->ps_bn[q] = 256; p->ps_nt[q] = 1; } else { p->ps_bn[q] = 192; p->ps_nt[q] = 3; }
  }
  p->Hs = p->ps_h[p->n_ps]; p->Ws = p->ps_w[p->n_ps];
  p->lstm_src = cfg->memory ? 2 : 1;
  p->lstm_rows_per_cell = p->lstm_src * 9 * 256;

  // ---- workspace layout
  const long long LB = static_cast<long long>(p->L) * p->B, TB = static_cast<long long>(p->T) * p->B;
  p->img_bytes = static_cast<size_t>(p->h) * p->w * kFeat * 2;
  long long img = 0;
  for (int s = 0; s <= p->S; ++s) { p->img_x[s] = img; img += LB; }
  for (int d = 0; d < 2; ++d) for (int l = 0; l < p->NL; ++l) { p->img_h[d][l] = img; img += LB; }
  for (int k = 0; k < 2; ++k) { p->img_sum[k] = img; img += cfg->all_heads ? TB : 0; }
  p->act_images = img;
  size_t off = 0;
  p->off_act = off; off = align_up(off + static_cast<size_t>(img) * p->img_bytes, 1024);
  p->mid_ch = 144;
  p->off_mid = off;
  if (cfg->pos_enc) off = align_up(off + static_cast<size_t>(p->n_win) * p->B * p->h * p->w * 144 * 2, 1024);
  p->c_elems = static_cast<size_t>(pvsr_lstm_state_elems(p->B, p->h, p->w));
  p->off_c = off; off = align_up(off + p->c_elems * 4 * 2 * p->NL, 1024);
  p->off_posterm = off;
  if (cfg->pos_enc) off = align_up(off + static_cast<size_t>(p->n_win) * p->B * 16 * 144 * 4, 1024);
  for (int q = 0; q < p->n_ps; ++q) {
    p->off_head[q] = off;
    off = align_up(off + static_cast<size_t>(TB) * p->ps_h[q + 1] * p->ps_w[q + 1] * kFeat * 2, 1024);
  }
  p->ws_bytes = off;

  // ---- packed parameter layout + gather indices
  size_t pk = 0;
  const int lstm_offs[2] = {0, kFeat};
  p->pk_lstm_w = pk; pk = align_up(pk + static_cast<size_t>(2 * p->NL) * p->lstm_rows_per_cell * 128, 1024);
  p->pk_lstm_b = pk; pk = align_up(pk + static_cast<size_t>(2 * p->NL) * 256 * 4, 1024);
  for (int d = 0; d < 2; ++d)
    for (int l = 0; l < p->NL; ++l) {
      const int ci = d * p->NL + l;
      const size_t dw = p->pk_lstm_w + static_cast<size_t>(ci) * p->lstm_rows_per_cell * 128;
      const size_t db = p->pk_lstm_b + static_cast<size_t>(ci) * 256 * 4;
      if (cfg->memory) {
        pvsr_pack_spec sp = make_spec(256, 128, 3, 2, lstm_offs, kFeat, 1, 9, 256, 0);
        add_pack_job(p, PK_LSTM, d, l, sp, nullptr, dw, db);
      } else {
        // cat([x, x]) (refine_net.py:255): one source whose operand is W[:, :64] + W[:, 64:]
        const int o0[1] = {0}, o1[1] = {kFeat};
        pvsr_pack_spec sa = make_spec(256, 128, 3, 1, o0, kFeat, 1, 9, 256, 0);
        pvsr_pack_spec sb = make_spec(256, 128, 3, 1, o1, kFeat, 1, 9, 256, 0);
        add_pack_job(p, PK_LSTM, d, l, sa, &sb, dw, db);
      }
    }
  {
    int offs[PVSR_MAX_SRC];
    if (cfg->pos_enc) {
      const int per = 2 * kFeat + 1;
      for (int s = 0; s < 2 * p->Wn; ++s) offs[s] = per * (s / 2) + kFeat * (s % 2);
      pvsr_pack_spec s1 = make_spec(per, per * p->Wn, 3, 2 * p->Wn, offs, kFeat, 1, 9, 144, 0);
      p->c1_rows = static_cast<long long>(2 * p->Wn) * 9 * 144;
      p->pk_c1_w = pk; pk = align_up(pk + static_cast<size_t>(p->c1_rows) * 128, 1024);
      add_pack_job(p, PK_C1, 0, 0, s1, nullptr, p->pk_c1_w, static_cast<size_t>(-1));
      const int o0[1] = {0};
      pvsr_pack_spec s2 = make_spec(kFeat, per, 3, 1, o0, per, 3, 9, kFeat, 0);
      p->c2_rows = 27LL * 64;
      p->pk_c2_w = pk; pk = align_up(pk + static_cast<size_t>(p->c2_rows) * 128, 1024);
      p->pk_c2_b = pk; pk = align_up(pk + 64 * 4, 1024);
      add_pack_job(p, PK_C2, 0, 0, s2, nullptr, p->pk_c2_w, p->pk_c2_b);
    } else {
      for (int s = 0; s < 2 * p->Wn; ++s) offs[s] = 2 * kFeat * (s / 2) + kFeat * (s % 2);
      pvsr_pack_spec s1 = make_spec(kFeat, 2 * kFeat * p->Wn, 1, 2 * p->Wn, offs, kFeat, 1, 1, kFeat, 0);
      p->c1_rows = static_cast<long long>(2 * p->Wn) * 64;
      p->pk_c1_w = pk; pk = align_up(pk + static_cast<size_t>(p->c1_rows) * 128, 1024);
      p->pk_c2_w = p->pk_c1_w; p->c2_rows = 0;
      p->pk_c2_b = pk; pk = align_up(pk + 64 * 4, 1024);   // bias of the 1x1 conv
      add_pack_job(p, PK_C1, 0, 0, s1, nullptr, p->pk_c1_w, p->pk_c2_b);
    }
  }
  for (int q = 0; q < p->n_ps; ++q) {
    const int r = p->ps_r[q];
    const int o0[1] = {0};
    pvsr_pack_spec sh = make_spec(kFeat * r * r, kFeat, 3, 1, o0, kFeat, 1, 9, kFeat * r * r, r);
    p->head_rows[q] = 9LL * kFeat * r * r;
    p->pk_head_w[q] = pk; pk = align_up(pk + static_cast<size_t>(p->head_rows[q]) * 128, 1024);
    p->pk_head_b[q] = pk; pk = align_up(pk + static_cast<size_t>(kFeat) * r * r * 4, 1024);
    add_pack_job(p, PK_HEAD, q, 0, sh, nullptr, p->pk_head_w[q], p->pk_head_b[q]);
  }
  p->pk_idx = pk;
  pk = align_up(pk + p->idx_host.size() * 4, 1024);
  p->pk_bytes = pk;

  // ---- accounting via a dry run of the schedule
  memset(p->launches, 0, sizeof(p->launches));
  for (double& f : p->flops) f = 0.0;
  Ctx c{};
  c.p = p; c.dry = true;
  schedule(c);
  *out = p;
  return 0;
}

void pvsr_plan_destroy(pvsr_plan* p) {
  if (!p) return;
  for (auto& g : p->graphs) cudaGraphExecDestroy(g.second);
  if (p->cap_stream) cudaStreamDestroy(p->cap_stream);
  delete p;
}

int64_t pvsr_plan_workspace_bytes(const pvsr_plan* p) { return static_cast<int64_t>(p->ws_bytes); }
int64_t pvsr_plan_packed_bytes(const pvsr_plan* p) { return static_cast<int64_t>(p->pk_bytes); }
int64_t pvsr_plan_output_elems(const pvsr_plan* p) {
  return static_cast<int64_t>(p->n_lists) * p->T * p->B * p->Hs * p->Ws;
}
int pvsr_plan_num_lists(const pvsr_plan* p) { return p->n_lists; }
int64_t pvsr_plan_num_launches(const pvsr_plan* p) {
  long long n = 0;
  for (long long v : p->launches) n += v;
  return n;
}
double pvsr_plan_flops(const pvsr_plan* p) {
  double f = 0;
  for (double v : p->flops) f += v;
  return f;
}
int pvsr_plan_class_stats(const pvsr_plan* p, int64_t* launches, double* flops) {
  for (int i = 0; i < kNumClasses; ++i) { launches[i] = p->launches[i]; flops[i] = p->flops[i]; }
  return kNumClasses;
}

int pvsr_plan_pack(pvsr_plan* p, const pvsr_net_params* P, void* packed, void* stream) {
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  uint8_t* pk = static_cast<uint8_t*>(packed);
  if (p->idx_uploaded_for != packed) {
    int e = cudaMemcpyAsync(pk + p->pk_idx, p->idx_host.data(), p->idx_host.size() * 4, cudaMemcpyHostToDevice, s);
    if (e) return check_cuda(e, "index upload");
    p->idx_uploaded_for = packed;
  }
  const int32_t* idx = reinterpret_cast<const int32_t*>(pk + p->pk_idx);
  for (const auto& j : p->jobs) {
    const float* src = job_weight(j, P, j.is_bias);
    if (!src) return set_error(-4, "missing parameter pointer (kind %d)", j.kind);
    int e;
    if (j.is_bias) e = launch_gather_f32(src, idx + j.idx, reinterpret_cast<float*>(pk + j.dst), j.n, s);
    else e = launch_pack_weights(src, idx + j.idx, j.has2 ? idx + j.idx2 : nullptr, pk + j.dst, j.n, s);
    if (e) return check_cuda(e, "pack launch");
  }
  return 0;
}

static int forward_eager(pvsr_plan* p, const pvsr_net_params* P, const void* packed, const float* lr,
                         const float* pos, float* out, void* ws, cudaStream_t s, std::vector<cudaEvent_t>* ev,
                         std::vector<int>* ev_cls) {
  Ctx c{};
  c.p = p; c.dry = false; c.P = P;
  c.ws = static_cast<uint8_t*>(ws);
  c.pk = static_cast<const uint8_t*>(packed);
  c.lr = lr; c.pos = pos; c.out = out; c.stream = s;
  c.events = ev; c.event_cls = ev_cls;
  schedule(c);
  return c.rc;
}

int pvsr_plan_forward(pvsr_plan* p, const pvsr_net_params* P, const void* packed, const float* lr, const float* pos,
                      float* out, void* ws, int use_graph, void* stream) {
  if (!p || !P || !packed || !lr || !out || !ws) return set_error(-2, "null argument");
  if (p->cfg.pos_enc && !pos) return set_error(-2, "pos codes required");
  if (p->num_sms == 0) {
    int rc = pvsr_device_check();
    if (rc) return rc;
    p->num_sms = device_num_sms();
  }
  int rc = build_maps(p, ws, packed);
  if (rc) return rc;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (!use_graph) return forward_eager(p, P, packed, lr, pos, out, ws, s, nullptr, nullptr);

  GraphKey key{{ws, packed, lr, pos, out, P->in_w}};
  auto it = p->graphs.find(key);
  if (it == p->graphs.end()) {
    // First use of this pointer tuple: run eagerly on the caller's stream (this call's result; also sets the
    // kernel attributes), then capture the same schedule on an internal stream for later replays.
    rc = forward_eager(p, P, packed, lr, pos, out, ws, s, nullptr, nullptr);
    if (rc) return rc;
    int e = 0;
    if (!p->cap_stream) {
      e = cudaStreamCreateWithFlags(&p->cap_stream, cudaStreamNonBlocking);
      if (e) return check_cuda(e, "capture stream");
    }
    cudaGraph_t graph = nullptr;
    e = cudaStreamBeginCapture(p->cap_stream, cudaStreamCaptureModeThreadLocal);
    if (e) return check_cuda(e, "begin capture");
    rc = forward_eager(p, P, packed, lr, pos, out, ws, p->cap_stream, nullptr, nullptr);
    e = cudaStreamEndCapture(p->cap_stream, &graph);
    if (rc) { if (graph) cudaGraphDestroy(graph); return rc; }
    if (e) return check_cuda(e, "end capture");
    cudaGraphExec_t exec = nullptr;
    e = cudaGraphInstantiate(&exec, graph, 0);
    cudaGraphDestroy(graph);
    if (e) return check_cuda(e, "graph instantiate");
    p->graphs.emplace(key, exec);
    return 0;
  }
  return check_cuda(cudaGraphLaunch(it->second, s), "graph launch");
}

// One eager forward with CUDA events around every launch; returns summed milliseconds per launch class.
// Synchronises the stream (profiling aid for bench.py's roofline figures, not a hot path).
int pvsr_plan_profile(pvsr_plan* p, const pvsr_net_params* P, const void* packed, const float* lr, const float* pos,
                      float* out, void* ws, double* ms_by_class, void* stream) {
  if (p->num_sms == 0) {
    int rc = pvsr_device_check();
    if (rc) return rc;
    p->num_sms = device_num_sms();
  }
  int rc = build_maps(p, ws, packed);
  if (rc) return rc;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  std::vector<cudaEvent_t> ev;
  std::vector<int> cls;
  rc = forward_eager(p, P, packed, lr, pos, out, ws, s, &ev, &cls);
  int e = cudaStreamSynchronize(s);
  for (int i = 0; i < kNumClasses; ++i) ms_by_class[i] = 0.0;
  for (size_t i = 0; i < cls.size(); ++i) {
    float ms = 0.f;
    cudaEventElapsedTime(&ms, ev[2 * i], ev[2 * i + 1]);
    ms_by_class[cls[i]] += ms;
  }
  for (cudaEvent_t x : ev) cudaEventDestroy(x);
  if (rc) return rc;
  return check_cuda(e, "profile sync");
}

}  // extern "C"
