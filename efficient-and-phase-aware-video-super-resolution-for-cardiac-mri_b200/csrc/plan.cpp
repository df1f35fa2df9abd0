// Whole-network plan: RefineNet.forward (reference src/model/nets/refine_net.py:61-135) and its backward pass
// (autograd through the same lines, driven by loss.backward() at
// src/runner/trainers/acdc_vsr_refinenet_trainer.py:44-47) as fixed schedules of kernel launches over a
// caller-provided workspace.  Creation is host-only (geometry, workspace layout, packing / scatter indices,
// weight-gradient job lists); forward / backward enqueue the schedule on the caller's stream, optionally as a
// replayed CUDA graph.
//
// Forward schedule per stage:
//   ConvLSTM   : wavefront launches d = 0..L+NL-2; launch d runs every cell (dir, layer l, step t = d - l) - up to
//                2*NL independent cells - as ONE tcgen05 launch (ConvParams.prob[]).       refine_net.py:81-93
//   refine     : conv1 over all (L-4) windows in one launch, sources gathered by TMA straight from the hidden
//                stacks (the 645-channel concat of :166-177 is never materialised), pos code as border-class
//                table; conv2 + bias + residual writes the next-stage features.             refine_net.py:94,132
//   heads      : conv+PixelShuffle launches over all T frames of a list at once, last conv 64->1. refine_net.py:100-113
//
// Backward schedule per stage, last stage first (only the T gradient frames U <= j < L-U carry gradients; everything
// produced at a warm-up frame is a constant - the reference's no_grad blocks, refine_net.py:74-79,82-93,179-183):
//   heads      : last-conv adjoint (SIMT), then per conv+PixelShuffle layer a tcgen05 weight-gradient launch and a
//                tcgen05 data-gradient launch whose A operand is the pixel-UNshuffled view of the HR gradient (TMA
//                element strides).  The LR data gradient accumulates (fp32) into gx (d/dx_j) and the top-layer dh.
//   refine     : conv2 wgrad + dgrad, conv1 wgrad (+ positional-code channels by border-class sums) and the conv1
//                data gradient as a 5-frame GATHER over the zero-padded conv1-output gradient (mirror of the forward
//                window gather) accumulated into the top-layer dh of both directions.
//   ConvLSTM   : reverse wavefront; per step one pointwise launch (gate adjoints, dc recurrence) and one tcgen05
//                launch (d[x | h_prev] of all active cells, routed to the layer below / gx and to the previous step);
//                one weight-gradient launch per stage over all T frames of all cells.
//   in_block   : weight / bias / slope gradients (SIMT reduction), then the packed weight gradients are scattered
//                into the parameter-layout gradient buffers.
#include <cuda.h>
#include <cuda_runtime.h>

#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <vector>

#include "conv.h"
#include "internal.h"
#include "simt.h"

using namespace pvsr;

namespace {

constexpr int kFeat = 64;
constexpr int kNumClasses = PVSR_NUM_CLASSES;
enum LaunchClass { CLS_IN = 0, CLS_LSTM = 1, CLS_CONV1 = 2, CLS_CONV2 = 3, CLS_HEAD_PS = 4, CLS_HEAD_LAST = 5, CLS_MISC = 6 };
constexpr int kNumClassesBwd = PVSR_NUM_CLASSES_BWD;
enum LaunchClassBwd {
  BCLS_HEAD_LAST = 0, BCLS_HEAD_DGRAD = 1, BCLS_HEAD_WGRAD = 2, BCLS_REFINE_DGRAD = 3, BCLS_REFINE_WGRAD = 4,
  BCLS_LSTM_POINT = 5, BCLS_LSTM_DGRAD = 6, BCLS_LSTM_WGRAD = 7, BCLS_MISC = 8
};
constexpr int kMaxStages = 8;
constexpr int kMaxClassesAny = kNumClassesBwd > kNumClasses ? kNumClassesBwd : kNumClasses;

inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

struct Tiling {
  int tw_log2, tw, th, tiles_x, tiles_y;
  void set(int H, int W, int max_tw = 128) {
    choose_tile(H, W, &tw_log2, max_tw);
    tw = 1 << tw_log2;
    th = kTileM >> tw_log2;
    tiles_x = (W + tw - 1) / tw;
    tiles_y = (H + th - 1) / th;
  }
};

struct GraphKey {
  const void* p[8];
  bool operator<(const GraphKey& o) const { return memcmp(p, o.p, sizeof(p)) < 0; }
};

const void* hash_bytes(const void* data, size_t n) {
  const uint8_t* b = static_cast<const uint8_t*>(data);
  uint64_t h = 1469598103934665603ull;
  for (size_t i = 0; i < n; ++i) { h ^= b[i]; h *= 1099511628211ull; }
  return reinterpret_cast<const void*>(static_cast<uintptr_t>(h));
}

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
  bool train;            // cfg.save_for_backward: per-stage copies of everything the backward pass reads
  int n_slots;           // per-stage copies kept (S in training, 1 otherwise)
  int n_list_slots;      // head intermediate copies kept (3S in training, 1 otherwise)

  // ---- workspace layout (bytes)
  size_t img_bytes;      // one LR image of 64 bf16 channels
  long long act_images;  // images in the LR 64-channel region
  long long img_x[kMaxStages + 1];    // image base of X[s] (s = 0..S)
  long long img_h[kMaxStages][2][PVSR_MAX_LAYERS];
  long long img_sum[kMaxStages][2];
  size_t off_act, off_mid, off_c, off_gates, off_posterm, off_head[PVSR_MAX_HEAD_CONVS], ws_bytes;
  size_t mid_stride, head_stride[PVSR_MAX_HEAD_CONVS];
  size_t c_elems;
  int mid_ch;
  // backward-only regions
  size_t off_gx, off_dh, off_dc, off_dgates, off_gr, off_gm, off_dhead[PVSR_MAX_HEAD_CONVS], off_wg, off_sums, off_jobs;
  size_t off_tail;       // scratch of the rank-1 adjoint of the head's tail (tail_rank1.cu)
  // pvsr_plan_set_sign_gradient: dL/d(out) of list k is sign_scale[k] * {-1, 0, +1} (fused L1 path); 0 = arbitrary
  float sign_scale[3 * kMaxStages] = {0};
  size_t dh_stride;      // bytes of one [T*B] stack of fp32 64-channel LR gradient images
  Tiling bw_tile[PVSR_MAX_HEAD_CONVS];   // tiling of the data/weight-gradient launches of head conv q

  // ---- packed-parameter layout (bytes)
  size_t pk_lstm_w, pk_lstm_b, pk_c1_w, pk_c2_w, pk_c2_b, pk_head_w[PVSR_MAX_HEAD_CONVS], pk_head_b[PVSR_MAX_HEAD_CONVS];
  size_t pk_idx, pk_bytes;
  // one table_kernel launch packs every operand (pvsr_table_job rows in the packed buffer behind the index region);
  // the backward pass scatters every packed gradient the same way.  Used when no job needs a second index.
  size_t pk_table = 0, pk_sc_table = 0;
  size_t pk_tail_fwd = 0;               // tables of the composite forward of the head's tail (tail_rank1.cu)
  size_t pk_t36_w = 0, pk_t36_b = 0;    // 36-channel form: packed operand [9][48][64] bf16, bias fp32 [48]
  size_t off_t36 = 0;                   // workspace: B [T*B images][H1][W1][48] of one list (bf16; fp32 in training plans)
  ConvMaps maps_t36;
  bool table_ok = false;
  const void* table_key = nullptr;      // hash of the parameter pointers the uploaded pack table was built for
  const void* sc_table_key = nullptr;   // same for the scatter table (gradient pointers)
  long long c1_rows, c2_rows, head_rows[PVSR_MAX_HEAD_CONVS];
  // data-gradient operands (training)
  size_t pk_lstm_dg, pk_c1_dg, pk_c2_dg, pk_head_dg[PVSR_MAX_HEAD_CONVS];
  long long lstm_dg_rows_per_cell, c1_dg_rows, c2_dg_rows, head_dg_rows[PVSR_MAX_HEAD_CONVS];
  int lstm_dg_bn;
  // gather jobs: (param slot, element count, dst offset, idx offset, optional idx2 offset)
  struct PackJob { int kind, a, b; long long n; size_t dst, idx, idx2; bool has2, is_bias; };
  std::vector<PackJob> jobs;
  std::vector<int32_t> idx_host;
  const void* idx_uploaded_for = nullptr;
  // packed fp32 weight-gradient buffer (element offsets inside off_wg) and its scatter jobs
  long long wg_lstm_w[2 * PVSR_MAX_LAYERS], wg_lstm_b[2 * PVSR_MAX_LAYERS], wg_c1_w, wg_c1_b, wg_c2_w, wg_c2_b;
  long long wg_head_w[PVSR_MAX_HEAD_CONVS], wg_head_b[PVSR_MAX_HEAD_CONVS], wg_elems;
  int c1_wg_ntotal;
  struct ScatterJob { int kind, a, b; long long n, src; size_t idx, idx2; bool has2, is_bias; };
  std::vector<ScatterJob> sc_jobs;
  // weight-gradient launches: job lists pre-built at creation, uploaded on the first backward
  struct WgLaunch { int job_begin, n_jobs, n_heavy, n_img; Tiling tile; };
  std::vector<WgJob> wg_jobs;
  WgLaunch wl_lstm[kMaxStages], wl_c1[kMaxStages], wl_c2[kMaxStages], wl_head[kMaxStages][PVSR_MAX_HEAD_CONVS];
  const void* jobs_uploaded_for = nullptr;

  // ---- device-dependent state
  int num_sms = 0;
  const void* maps_ws = nullptr;
  const void* maps_pk = nullptr;
  int maps_cg = -1;      // CTA-pair setting the tensor maps (weight box height) were built for
  int maps_halo = -1;    // halo mode the activation maps (box rows) were built for
  // Slab (padded-raster) geometry per launch kind; on = false: nine shifted boxes (conv.h: PrGeom)
  struct Geo { bool on = false; PrGeom g{0, 0, 0}; };
  Geo geo_lstm;                           // ConvLSTM cells (= geo_lr; their state tensors are laid out by this tile map)
  Geo geo_lr;                             // other 3x3 launches at LR resolution (refine convs, their data gradients, ...)
  Geo geo_ps[PVSR_MAX_HEAD_CONVS];        // head conv q (its input resolution)
  Geo geo_hdg[PVSR_MAX_HEAD_CONVS];       // data gradient of head conv q (sources: pixel-unshuffled views, mul = r)
  ConvMaps maps_lstm, maps_c1, maps_c2, maps_head[PVSR_MAX_HEAD_CONVS];   // act[0] + packed weights of each launch kind
  ConvMaps bm_lstm_dg, bm_lstm_wg, bm_c1_dg, bm_c1_wg, bm_c2_dg, bm_c2_wg, bm_head_dg[PVSR_MAX_HEAD_CONVS],
      bm_head_wg[PVSR_MAX_HEAD_CONVS];
  std::map<GraphKey, cudaGraphExec_t> graphs;
  cudaStream_t cap_stream = nullptr;   // capture happens here (the caller's stream may be the legacy stream)
  // Second stream of the two-branch schedules (Ctx::side): weight-gradient launches and the HBM-bound 64 <-> 1 channel
  // head kernels leave the critical path and overlap with it; `side_stream` serves eager runs, `cap_side` captures.
  cudaStream_t side_stream = nullptr, cap_side = nullptr;
  std::vector<cudaEvent_t> ev_pool;    // fork / join events of the two-branch schedules (timing disabled)

  // ---- accounting (filled by dry runs at creation)
  long long launches[kNumClasses];
  double flops[kNumClasses];
  long long launches_bwd[kNumClassesBwd];
  double flops_bwd[kNumClassesBwd];
};

namespace {

// ------------------------------------------------------------------------------------------------ creation helpers
size_t add_index(pvsr_plan* p, const pvsr_pack_spec& spec, bool bias) {
  const size_t at = p->idx_host.size();
  const long long n = bias ? spec.n_total : pvsr_pack_index_count(&spec);
  p->idx_host.resize(at + n);
  if (bias) pvsr_pack_bias_index_host(&spec, p->idx_host.data() + at);
  else pvsr_pack_index_host(&spec, p->idx_host.data() + at);
  return at;
}

void add_pack_job(pvsr_plan* p, int kind, int a, int b, const pvsr_pack_spec& spec, const pvsr_pack_spec* spec2,
                  size_t dst_w, size_t dst_b) {
  pvsr_plan::PackJob j{};
  j.kind = kind; j.a = a; j.b = b;
  j.n = pvsr_pack_index_count(&spec);
  j.dst = dst_w;
  j.idx = add_index(p, spec, false);
  j.has2 = spec2 != nullptr;
  if (spec2) j.idx2 = add_index(p, *spec2, false);
  j.is_bias = false;
  p->jobs.push_back(j);
  if (dst_b != static_cast<size_t>(-1)) {
    pvsr_plan::PackJob bj{};
    bj.kind = kind; bj.a = a; bj.b = b;
    bj.n = spec.n_total;
    bj.dst = dst_b;
    bj.idx = add_index(p, spec, true);
    bj.is_bias = true;
    p->jobs.push_back(bj);
  }
}

// packed fp32 gradient (layout of `spec`) -> parameter gradient; bias gradient likewise.
void add_scatter_job(pvsr_plan* p, int kind, int a, int b, const pvsr_pack_spec& spec, const pvsr_pack_spec* spec2,
                     long long src_w, long long src_b) {
  pvsr_plan::ScatterJob j{};
  j.kind = kind; j.a = a; j.b = b;
  j.n = pvsr_pack_index_count(&spec);
  j.src = src_w;
  j.idx = add_index(p, spec, false);
  j.has2 = spec2 != nullptr;
  if (spec2) j.idx2 = add_index(p, *spec2, false);
  j.is_bias = false;
  p->sc_jobs.push_back(j);
  pvsr_plan::ScatterJob bj{};
  bj.kind = kind; bj.a = a; bj.b = b;
  bj.n = spec.n_total;
  bj.src = src_b;
  bj.idx = add_index(p, spec, true);
  bj.is_bias = true;
  p->sc_jobs.push_back(bj);
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

// Parameter slots for PackJob.kind / ScatterJob.kind
enum { PK_LSTM = 0, PK_C1 = 1, PK_C2 = 2, PK_HEAD = 3 };

const float* job_weight(int kind, int a, int b, const pvsr_net_params* P, bool bias) {
  switch (kind) {
    case PK_LSTM: return bias ? P->lstm_b[a][b] : P->lstm_w[a][b];
    case PK_C1: return bias ? P->ref_b1 : P->ref_w1;
    case PK_C2: return bias ? P->ref_b2 : P->ref_w2;
    case PK_HEAD: return bias ? P->head_b[a] : P->head_w[a];
  }
  return nullptr;
}
float* job_grad(int kind, int a, int b, const pvsr_net_grads* G, bool bias) {
  switch (kind) {
    case PK_LSTM: return bias ? G->lstm_b[a][b] : G->lstm_w[a][b];
    case PK_C1: return bias ? G->ref_b1 : G->ref_w1;
    case PK_C2: return bias ? G->ref_b2 : G->ref_w2;
    case PK_HEAD: return bias ? G->head_b[a] : G->head_w[a];
  }
  return nullptr;
}

// ------------------------------------------------------------------------------------------------ launch trace
// PVSR_TRACE_LAUNCH=1 (debug aid): every launch of an eager schedule run is followed by an event on its stream;
// pvsr_debug_dump_trace() then reports, per branch, the first launch that has not completed (hang hunting).
struct TraceEntry { cudaEvent_t ev; int cls, seq, side, bwd; };
std::vector<TraceEntry> g_trace;
bool trace_enabled() {
  static int on = -1;
  if (on < 0) { const char* e = getenv("PVSR_TRACE_LAUNCH"); on = (e && e[0] == '1') ? 1 : 0; }
  return on == 1;
}
bool env_flag(const char* name) {
  const char* e = getenv(name);
  return e && e[0] == '1';
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
  // backward only
  const float* dout = nullptr;
  const pvsr_net_grads* G = nullptr;
  long long* cnt_launches = nullptr;   // accounting arrays in use (forward or backward classes)
  double* cnt_flops = nullptr;
  // optional per-launch timing
  std::vector<cudaEvent_t>* events = nullptr;
  std::vector<int>* event_cls = nullptr;
  int rc = 0;
  // Two-branch schedule: launches bracketed by to_side() / to_main() go to `side` (ordered after everything enqueued on
  // `main` so far); main waits for side work only where it is about to overwrite a buffer the side branch reads
  // (main_wait) and at the end (join).  side == nullptr (profiling runs, dry runs): a single stream, same results.
  cudaStream_t main = nullptr, side = nullptr;
  int ev_next = 0;
  bool on_side = false;
  int trace_seq = 0, trace_bwd = 0;

  cudaEvent_t next_event() {
    if (ev_next >= static_cast<int>(p->ev_pool.size())) {
      cudaEvent_t e;
      cudaEventCreateWithFlags(&e, cudaEventDisableTiming);
      p->ev_pool.push_back(e);
    }
    return p->ev_pool[ev_next++];
  }
  void to_side() {
    if (dry || !side || on_side) return;
    cudaEvent_t e = next_event();
    cudaEventRecord(e, main);
    cudaStreamWaitEvent(side, e, 0);
    stream = side;
    on_side = true;
  }
  void to_main() {
    if (!on_side) return;
    stream = main;
    on_side = false;
  }
  // Event marking "everything enqueued on the side branch so far" (nullptr when there is no side branch)
  cudaEvent_t side_mark() {
    if (dry || !side) return nullptr;
    cudaEvent_t e = next_event();
    cudaEventRecord(e, side);
    return e;
  }
  void main_wait(cudaEvent_t e) {
    if (e) cudaStreamWaitEvent(main, e, 0);
  }
  void join() {
    to_main();
    main_wait(side_mark());
  }

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
    if (dry) {
      cnt_launches[cls] += 1;
      cnt_flops[cls] += fl;
    }
    if (!dry && trace_enabled()) {
      cudaStreamCaptureStatus st = cudaStreamCaptureStatusNone;
      cudaStreamIsCapturing(stream, &st);
      if (st == cudaStreamCaptureStatusNone) {
        TraceEntry t{};
        cudaEventCreateWithFlags(&t.ev, cudaEventDisableTiming);
        cudaEventRecord(t.ev, stream);
        t.cls = cls; t.seq = trace_seq; t.side = on_side ? 1 : 0; t.bwd = trace_bwd;
        g_trace.push_back(t);
      }
      ++trace_seq;
    }
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
float* c_state(const Ctx& c, int s, int dir, int l, int j) {
  const pvsr_plan* p = c.p;
  size_t idx = p->train ? ((static_cast<size_t>(s) * 2 + dir) * p->NL + l) * p->L + j : static_cast<size_t>(dir) * p->NL + l;
  return reinterpret_cast<float*>(c.ws + p->off_c) + idx * p->c_elems;
}
__nv_bfloat16* gates_buf(const Ctx& c, int s, int dir, int l, int t) {   // t = j - U, training only
  const pvsr_plan* p = c.p;
  const size_t idx = ((static_cast<size_t>(s) * 2 + dir) * p->NL + l) * p->T + t;
  return reinterpret_cast<__nv_bfloat16*>(c.ws + p->off_gates) + idx * p->c_elems * 4;
}
float* grad_stack(const Ctx& c, size_t off, long long img) {   // fp32 NHWC 64-channel LR gradient image
  return reinterpret_cast<float*>(c.ws + off) + static_cast<size_t>(img) * c.p->h * c.p->w * kFeat;
}

void base_params(const Tiling& t, int H, int W, ConvParams* cp) {
  memset(cp, 0, sizeof(*cp));
  cp->H = H; cp->W = W;
  cp->tw_log2 = t.tw_log2; cp->tiles_x = t.tiles_x; cp->tiles_y = t.tiles_y;
  cp->taps = 9; cp->kb_per_src = 1; cp->k16_last = 4; cp->n_tiles_n = 1;
}

void set_slab(ConvParams* cp, const pvsr_plan::Geo& g) {
  if (!g.on) return;
  cp->halo = 1;
  cp->w_resident = get_w_resident();
  cp->pr_wp = g.g.wp;
  cp->pr_rows = g.g.rows;
  cp->tiles_x = 1;
  cp->tiles_y = g.g.tiles;
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
    const int slot = p->train ? s : 0;
    // ---------------------------------------------------------------- bidirectional ConvLSTM wavefront
    // Last stage: nothing reads the hidden maps of the trailing U - half steps of either direction (no next-stage
    // feature update; the heads and the refine windows of the T output frames stop `half` frames past them), so the
    // recurrence stops there - same results, (U - half) / L less ConvLSTM work in that stage.
    const int dead = (s == S - 1 && U > half) ? U - half : 0;
    const int Lrun = L - dead;
    for (int d = 0; d < Lrun + NL - 1; ++d) {
      ConvParams cp;
      base_params(p->lr, p->h, p->w, &cp);
      cp.n_img = B;
      cp.n_total = 256; cp.n_store = 256; cp.out_ch = kFeat;
      int np = 0;
      for (int dir = 0; dir < 2; ++dir)
        for (int l = 0; l < NL; ++l) {
          const int t = d - l;
          if (t < 0 || t >= Lrun) continue;
          const int j = dir == 0 ? t : L - 1 - t;
          const int jp = dir == 0 ? j - 1 : j + 1;
          ConvProblem& pr = cp.prob[np++];
          const long long xin = (l == 0 ? p->img_x[s] : p->img_h[slot][dir][l - 1]) + static_cast<long long>(j) * B;
          pr.src[0] = view0(xin);
          pr.n_src = 1;
          if (p->lstm_src == 2 && t > 0) {
            pr.src[1] = view0(p->img_h[slot][dir][l] + static_cast<long long>(jp) * B);
            pr.n_src = 2;
          }
          const int ci = dir * NL + l;
          pr.w_row_base = ci * p->lstm_rows_per_cell;
          pr.bias = reinterpret_cast<const float*>(c.pk + p->pk_lstm_b) + ci * 256;
          pr.c_in = t > 0 ? c_state(c, s, dir, l, p->train ? jp : 0) : nullptr;
          pr.c_out = c_state(c, s, dir, l, p->train ? j : 0);
          pr.h_out = act_img(c, p->img_h[slot][dir][l] + static_cast<long long>(j) * B);
          if (p->train && j >= U && j < L - U) pr.gates_out = gates_buf(c, s, dir, l, j - U);
        }
      cp.n_prob = np;
      set_slab(&cp, p->geo_lstm);
      run_conv(c, CLS_LSTM, 256, EPI_LSTM, p->maps_lstm, cp, lstm_fl * px * B * np);
    }

    // ---------------------------------------------------------------- refine block -> next-stage features
    const long long hf_top = p->img_h[slot][0][NL - 1], hb_top = p->img_h[slot][1][NL - 1];
    {
      // windows computed: all n_win of them, except in the last stage where only the T output frames are read
      const int w0 = dead;                                  // first window (frame half + w0)
      const int n_run = p->n_win - 2 * dead;
      const long long wimg = static_cast<long long>(w0) * B;
      ConvParams cp;
      base_params(p->lr, p->h, p->w, &cp);
      cp.n_img = n_run * B;
      ConvProblem& pr = cp.prob[0];
      cp.n_prob = 1;
      pr.n_src = 2 * p->Wn;
      for (int jw = 0; jw < p->Wn; ++jw) {
        pr.src[2 * jw] = view0(hf_top + static_cast<long long>(jw) * B + wimg);
        pr.src[2 * jw + 1] = view0(hb_top + static_cast<long long>(jw) * B + wimg);
      }
      const __nv_bfloat16* res = act_img(c, p->img_x[s] + static_cast<long long>(half) * B + wimg);
      __nv_bfloat16* xnext = act_img(c, p->img_x[s + 1] + static_cast<long long>(half) * B + wimg);
      if (p->cfg.pos_enc) {
        cp.n_total = 144; cp.n_store = 144; cp.out_ch = 144;
        pr.posterm = posterm + static_cast<size_t>(wimg) * 16 * 144;
        pr.out_bf16 = reinterpret_cast<__nv_bfloat16*>(c.ws + p->off_mid + slot * p->mid_stride) +
                      static_cast<size_t>(wimg) * px * 144;
        set_slab(&cp, p->geo_lr);
        run_conv(c, CLS_CONV1, 144, EPI_STORE, p->maps_c1, cp,
                 2.0 * 9 * (2 * kFeat + 1) * p->Wn * (2 * kFeat + 1) * px * cp.n_img);
        ConvParams c2;
        base_params(p->lr, p->h, p->w, &c2);
        c2.n_img = n_run * B;
        c2.n_prob = 1;
        c2.kb_per_src = 3; c2.k16_last = 1;
        c2.n_total = 64; c2.n_store = 64; c2.out_ch = 64;
        ConvProblem& p2 = c2.prob[0];
        p2.n_src = 1;
        p2.src[0] = view0(static_cast<long long>(slot) * p->n_win * B + wimg);
        p2.bias = reinterpret_cast<const float*>(c.pk + p->pk_c2_b);
        p2.res = res;
        p2.out_bf16 = xnext;
        set_slab(&c2, p->geo_lr);
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
      const int lslot = p->train ? list : 0;
      long long in_img;
      if (k == 2) {
        in_img = p->img_x[s + 1] + static_cast<long long>(U) * B;   // x_j + r_j == next-stage feature
      } else {
        in_img = p->img_sum[slot][k];
        run_add(c, p->img_x[s] + static_cast<long long>(U) * B,
                (k == 0 ? hf_top : hb_top) + static_cast<long long>(U) * B, in_img, static_cast<long long>(T) * B);
      }
      const long long n_head = static_cast<long long>(T) * B;
      // x4 / x8: the last conv + PixelShuffle(2) and the final 64 -> 1 conv run as ONE composite 5x5 conv (tail_rank1.cu);
      // training plans only together with the rank-1 backward (the conv-by-conv backward reads the 64-channel HR map)
      const bool tailf = get_tail_fwd() != 0 && p->n_ps >= 2 && p->ps_r[p->n_ps - 1] == 2 &&
                         (!p->train || get_tail_rank1() != 0);
      const int n_ps_run = tailf ? p->n_ps - 1 : p->n_ps;
      for (int q = 0; q < n_ps_run; ++q) {
        ConvParams cp;
        base_params(q == 0 ? p->lr : p->ps_tile[q], p->ps_h[q], p->ps_w[q], &cp);
        cp.n_img = static_cast<int>(n_head);
        cp.n_prob = 1;
        cp.n_tiles_n = p->ps_nt[q];
        cp.n_total = p->ps_bn[q] * p->ps_nt[q];
        cp.n_store = p->ps_bn[q];
        cp.out_ch = kFeat;
        cp.ps_r = p->ps_r[q];
        ConvProblem& pr = cp.prob[0];
        pr.n_src = 1;
        pr.src[0] = view0(q == 0 ? in_img : static_cast<long long>(lslot) * n_head);
        pr.bias = reinterpret_cast<const float*>(c.pk + p->pk_head_b[q]);
        pr.out_bf16 = reinterpret_cast<__nv_bfloat16*>(c.ws + p->off_head[q] + lslot * p->head_stride[q]);
        set_slab(&cp, p->geo_ps[q]);
        run_conv(c, CLS_HEAD_PS, p->ps_bn[q], EPI_PS, p->maps_head[q], cp,
                 2.0 * 9 * kFeat * (kFeat * p->ps_r[q] * p->ps_r[q]) * p->ps_h[q] * p->ps_w[q] * n_head);
      }
      // Training plans keep one head-intermediate slot per list, so the HBM-bound 64 -> 1 conv of list k can run on the
      // side branch underneath the tensor-bound launches that follow (next list's head convs, next stage's ConvLSTM).
      if (p->train) c.to_side();
      if (tailf && get_tail_fwd() == 2) {
        // 36-channel form: one N = 48 tcgen05 launch (B records) + the 9-tap gather
        const int last = p->n_ps - 1;
        const double fl = (2.0 * 9 * kFeat * (kFeat * 4) * p->ps_h[last] * p->ps_w[last] +
                           2.0 * 9 * kFeat * static_cast<double>(p->Hs) * p->Ws) * n_head;              // algorithmic FLOPs
        // training plans keep the fp32 accumulator in the records: the loss gradient is sign(out - target), and bf16
        // records moved the LSTM bias gradients of the x4_nopos fixture from 0.03 to 0.05 rel-L2 of the reference's
        void* Bbuf = c.ws + p->off_t36;
        ConvParams cp;
        base_params(last == 0 ? p->lr : p->ps_tile[last], p->ps_h[last], p->ps_w[last], &cp);
        cp.n_img = static_cast<int>(n_head);
        cp.n_prob = 1;
        cp.n_total = 48; cp.n_store = 48; cp.out_ch = 48;
        ConvProblem& pr = cp.prob[0];
        pr.n_src = 1;
        pr.src[0] = view0(static_cast<long long>(lslot) * n_head);
        pr.bias = reinterpret_cast<const float*>(c.pk + p->pk_t36_b);
        if (p->train) pr.out_f32 = static_cast<float*>(Bbuf);
        else pr.out_bf16 = static_cast<__nv_bfloat16*>(Bbuf);
        set_slab(&cp, p->geo_ps[last]);
        run_conv(c, CLS_HEAD_PS, 48, EPI_STORE, p->maps_t36, cp, fl);
        c.begin(CLS_HEAD_LAST);
        if (!c.dry && !c.rc) {
          float* o = c.out + static_cast<size_t>(list) * T * B * p->Hs * p->Ws;
          int e = launch_tail36_gather(Bbuf, p->train ? 1 : 0, c.P->head_b[p->n_ps], o, n_head, p->ps_h[last], p->ps_w[last], c.stream);
          if (e) c.rc = check_cuda(e, "tail36 gather launch");
        }
        c.end(CLS_HEAD_LAST, 0.0);
        c.to_main();
        continue;
      }
      if (tailf) {
        const int last = p->n_ps - 1;
        c.begin(CLS_HEAD_PS);
        if (!c.dry && !c.rc) {
          float* o = c.out + static_cast<size_t>(list) * T * B * p->Hs * p->Ws;
          int e = launch_tail_fwd(c.ws + p->off_head[last - 1] + lslot * p->head_stride[last - 1], c.pk + p->pk_tail_fwd, o,
                                  n_head, p->ps_h[last], p->ps_w[last], p->num_sms, c.stream);
          if (e) c.rc = check_cuda(e, "tail_fwd launch");
        }
        c.end(CLS_HEAD_PS, (2.0 * 9 * kFeat * (kFeat * 4) * p->ps_h[last] * p->ps_w[last] +
                            2.0 * 9 * kFeat * static_cast<double>(p->Hs) * p->Ws) * n_head);     // algorithmic FLOPs
        c.to_main();
        continue;
      }
      c.begin(CLS_HEAD_LAST);
      if (!c.dry && !c.rc) {
        float* o = c.out + static_cast<size_t>(list) * T * B * p->Hs * p->Ws;
        int e = launch_head_conv_last(c.ws + p->off_head[p->n_ps - 1] + lslot * p->head_stride[p->n_ps - 1],
                                      c.P->head_w[p->n_ps], c.P->head_b[p->n_ps], o, nullptr, nullptr, n_head, p->Hs,
                                      p->Ws, c.stream);
        if (e) c.rc = check_cuda(e, "head_conv_last launch");
      }
      c.end(CLS_HEAD_LAST, 2.0 * 9 * kFeat * static_cast<double>(p->Hs) * p->Ws * n_head);
      c.to_main();
    }

    // ---------------------------------------------------------------- edge-frame feature updates (:120-131)
    if (s + 1 < S) {
      run_add(c, p->img_x[s], hf_top, p->img_x[s + 1], static_cast<long long>(half) * B);
      run_add(c, p->img_x[s] + static_cast<long long>(L - half) * B, hb_top + static_cast<long long>(L - half) * B,
              p->img_x[s + 1] + static_cast<long long>(L - half) * B, static_cast<long long>(half) * B);
    }
  }
  c.join();
}

// ------------------------------------------------------------------------------------------------ backward helpers
void run_wgrad(Ctx& c, int cls, const ConvMaps& maps, const pvsr_plan::WgLaunch& wl, int H, int W, double fl) {
  if (c.rc) return;
  c.begin(cls);
  if (!c.dry && wl.n_jobs > 0) {
    WgParams wp{};
    wp.H = H; wp.W = W;
    wp.tw_log2 = wl.tile.tw_log2; wp.tiles_x = wl.tile.tiles_x; wp.tiles_y = wl.tile.tiles_y;
    wp.n_img = wl.n_img;
    wp.n_jobs = wl.n_jobs;
    const long long total_tiles = static_cast<long long>(wl.n_img) * wl.tile.tiles_x * wl.tile.tiles_y;
    wp.n_heavy = wl.n_heavy;
    choose_wgrad_splits(wl.n_heavy, wl.n_jobs - wl.n_heavy, total_tiles, c.p->num_sms, &wp.n_splits, &wp.n_splits_light);
    wp.jobs = reinterpret_cast<const WgJob*>(c.ws + c.p->off_jobs) + wl.job_begin;
    wp.grad = reinterpret_cast<float*>(c.ws + c.p->off_wg);
    int e = launch_wgrad(maps, wp, c.stream);
    if (e) c.rc = check_cuda(e, "wgrad launch");
  }
  c.end(cls, fl);
}

void run_memset(Ctx& c, size_t off, size_t bytes) {
  if (c.rc || bytes == 0) return;
  c.begin(BCLS_MISC);
  if (!c.dry) {
    int e = cudaMemsetAsync(c.ws + off, 0, bytes, c.stream);
    if (e) c.rc = check_cuda(e, "memset");
  }
  c.end(BCLS_MISC, 0.0);
}

template <typename F>
void run_simt(Ctx& c, int cls, const char* what, F&& f) {
  if (c.rc) return;
  c.begin(cls);
  if (!c.dry) {
    int e = f();
    if (e) c.rc = check_cuda(e, what);
  }
  c.end(cls, 0.0);
}

// The whole backward schedule (see the file header).  With c.dry it only counts.
void schedule_backward(Ctx& c) {
  pvsr_plan* p = c.p;
  const int B = p->B, L = p->L, U = p->U, T = p->T, S = p->S, NL = p->NL, half = p->half, Wn = p->Wn;
  const long long TB = static_cast<long long>(T) * B;
  const long long px = static_cast<long long>(p->h) * p->w;
  const size_t lr_img_f32 = static_cast<size_t>(px) * kFeat * 4;
  const bool pos = p->cfg.pos_enc != 0;

  run_memset(c, p->off_wg, static_cast<size_t>(p->wg_elems) * 4);
  run_memset(c, p->off_gx, static_cast<size_t>(TB) * lr_img_f32);
  // zero frames around the refine gradients (sources of the conv1 gather that fall on warm-up frames)
  {
    const size_t gr_img = static_cast<size_t>(px) * kFeat * 2, gm_img = static_cast<size_t>(px) * 144 * 2;
    run_memset(c, p->off_gr, static_cast<size_t>(half) * B * gr_img);
    run_memset(c, p->off_gr + static_cast<size_t>(half + T) * B * gr_img, static_cast<size_t>(half) * B * gr_img);
    if (pos) {
      run_memset(c, p->off_gm, static_cast<size_t>(half) * B * gm_img);
      run_memset(c, p->off_gm + static_cast<size_t>(half + T) * B * gm_img, static_cast<size_t>(half) * B * gm_img);
    }
  }
  float* wg = reinterpret_cast<float*>(c.ws + p->off_wg);
  __nv_bfloat16* gr = reinterpret_cast<__nv_bfloat16*>(c.ws + p->off_gr);
  // Rank-1 adjoint of [last conv + PixelShuffle(2), final conv] (tail_rank1.cu): replaces head_last_bwd_data /
  // _weight and the dgrad / wgrad launches of the last 64 -> 256 conv when the head has >= 2 shuffle stages (x4, x8).
  const bool tail = get_tail_rank1() != 0 && p->n_ps >= 2 && p->ps_r[p->n_ps - 1] == 2;
  void* tail_ws = c.ws + p->off_tail;
  if (tail) {
    run_simt(c, BCLS_MISC, "tail tables", [&] {
      int e = launch_tail_tables(c.P->head_w[p->n_ps - 1], c.P->head_w[p->n_ps], tail_ws, c.stream);
      return e ? e : launch_tail_zero_sums(tail_ws, c.stream);
    });
  }

  // Two-branch schedule (Ctx::side).  Main branch = the dependent chain of data gradients: head adjoints -> refine
  // dgrad -> reverse ConvLSTM wavefront (gate adjoint + dgrad per diagonal).  Side branch = everything that only
  // produces parameter gradients: every tcgen05 weight-gradient launch and the 64 -> 1 head weight adjoint.  The
  // HBM-bound main-branch kernels (head_last_bwd_data, casts, the gate adjoints) then run underneath tensor-bound
  // side-branch work instead of leaving the tensor pipe idle.  The side branch of stage s may still be reading the
  // per-stage gradient buffers when the main branch reaches stage s - 1: ev_* mark the three places it must wait.
  cudaEvent_t ev_head = nullptr, ev_refine = nullptr, ev_lstm = nullptr;
  for (int s = S - 1; s >= 0 && !c.rc; --s) {
    run_memset(c, p->off_dh, static_cast<size_t>(2 * NL) * p->dh_stride);
    const size_t dh_top_f = p->off_dh + static_cast<size_t>(0 * NL + NL - 1) * p->dh_stride;
    const size_t dh_top_b = p->off_dh + static_cast<size_t>(1 * NL + NL - 1) * p->dh_stride;

    // ---------------------------------------------------------------- heads of lists 3s, 3s+1, 3s+2
    const int last = p->n_ps - 1;
    const float* dout_s = c.dout ? c.dout + static_cast<size_t>(3 * s) * TB * p->Hs * p->Ws : nullptr;
    const uint8_t* head_last_in = c.ws + p->off_head[last] + static_cast<size_t>(3 * s) * p->head_stride[last];
    c.main_wait(ev_head);     // head weight gradients of stage s + 1 read dhead[*]
    int q_first = last;
    if (tail) {
      // d(input of the last 64 -> 256 conv) straight from dL/d(out); the correlation sums on the side branch
      const double tail_fl = 2.0 * 9 * kFeat * (kFeat * 4) * p->ps_h[last] * p->ps_w[last] * 3.0 * TB;   // algorithmic
      const float sgn = (p->sign_scale[3 * s] != 0.f && p->sign_scale[3 * s] == p->sign_scale[3 * s + 1] &&
                         p->sign_scale[3 * s] == p->sign_scale[3 * s + 2]) ? p->sign_scale[3 * s] : 0.f;
      c.begin(BCLS_HEAD_DGRAD);
      if (!c.dry && !c.rc) {
        int e = launch_tail_dx(dout_s, tail_ws, c.ws + p->off_dhead[last - 1], 3 * TB, p->ps_h[last], p->ps_w[last],
                               p->num_sms, c.stream, sgn);
        if (e) c.rc = check_cuda(e, "tail_dx launch");
      }
      c.end(BCLS_HEAD_DGRAD, tail_fl);
      // Measured (profiles/stress_train.py): with this launch on the side branch the captured two-branch graph hangs after
      // 20-30 replays (eager streams and the single-chain graph run 600+ steps); PVSR_TAIL_CORR_SIDE=1 reproduces it
      if (env_flag("PVSR_TAIL_CORR_SIDE")) c.to_side();
      c.begin(BCLS_HEAD_WGRAD);
      if (!c.dry && !c.rc) {
        const uint8_t* x_last = c.ws + p->off_head[last - 1] + static_cast<size_t>(3 * s) * p->head_stride[last - 1];
        int e = launch_tail_corr(dout_s, x_last, tail_ws, 3 * TB, p->ps_h[last], p->ps_w[last], p->num_sms, c.stream, sgn);
        if (e) c.rc = check_cuda(e, "tail_corr launch");
      }
      c.end(BCLS_HEAD_WGRAD, tail_fl);
      c.to_main();
      q_first = last - 1;
    } else {
      run_simt(c, BCLS_HEAD_LAST, "head_last_bwd_data", [&] {
        return launch_head_last_bwd_data(dout_s, c.P->head_w[p->n_ps], c.ws + p->off_dhead[last], 3 * TB, p->Hs, p->Ws,
                                         c.stream);
      });
      if (c.dry || (c.G->head_w[p->n_ps] && c.G->head_b[p->n_ps])) {
        c.to_side();
        run_simt(c, BCLS_HEAD_LAST, "head_last_bwd_weight", [&] {
          return launch_head_last_bwd_weight(head_last_in, dout_s, c.G->head_w[p->n_ps], c.G->head_b[p->n_ps], 3 * TB,
                                             p->Hs, p->Ws, p->num_sms, c.stream);
        });
        c.to_main();
      }
    }
    for (int q = q_first; q >= 0; --q) {
      const int r = p->ps_r[q];
      const double conv_fl = 2.0 * 9 * kFeat * (kFeat * r * r) * p->ps_h[q] * p->ps_w[q] * 3.0 * TB;
      c.to_side();
      run_wgrad(c, BCLS_HEAD_WGRAD, p->bm_head_wg[q], p->wl_head[s][q], p->ps_h[q], p->ps_w[q], conv_fl);
      c.to_main();
      ConvParams cp;
      base_params(p->bw_tile[q], p->ps_h[q], p->ps_w[q], &cp);
      set_slab(&cp, p->geo_hdg[q]);
      cp.n_total = 64; cp.n_store = 64; cp.out_ch = 64;
      if (q > 0) {
        cp.n_img = static_cast<int>(3 * TB);
        cp.n_prob = 1;
        ConvProblem& pr = cp.prob[0];
        pr.n_src = r * r;
        for (int u = 0; u < r * r; ++u) pr.src[u] = SrcView{0, 0, 0, r, u % r, u / r};
        pr.out_bf16 = reinterpret_cast<__nv_bfloat16*>(c.ws + p->off_dhead[q - 1]);
        run_conv(c, BCLS_HEAD_DGRAD, 64, EPI_STORE, p->bm_head_dg[q], cp, conv_fl);
      } else {
        // k = 2 (fused head): d/dx^{s+1}_j.  gx then equals dL/dx^{s+1} = dL/dr, and is frozen as the bf16 operand
        // of the refine backward before the k = 0, 1 heads add their share of dL/dx^s.
        cp.n_img = static_cast<int>(TB);
        cp.n_prob = 1;
        cp.grad_split = 0;
        {
          ConvProblem& pr = cp.prob[0];
          pr.n_src = r * r;
          for (int u = 0; u < r * r; ++u) pr.src[u] = SrcView{0, static_cast<int>(2 * TB), 0, r, u % r, u / r};
          pr.grad0 = grad_stack(c, p->off_gx, 0);
          pr.grad1 = nullptr;
        }
        run_conv(c, BCLS_HEAD_DGRAD, 64, EPI_GRAD, p->bm_head_dg[0], cp, conv_fl / 3);
        c.main_wait(ev_refine);   // refine weight gradients of stage s + 1 read gr / gm
        run_simt(c, BCLS_MISC, "cast gx", [&] {
          return launch_cast_f32_bf16(grad_stack(c, p->off_gx, 0), gr + static_cast<size_t>(half) * B * px * kFeat,
                                      TB * px * kFeat, c.stream);
        });
        cp.n_prob = 2;
        for (int k = 0; k < 2; ++k) {
          ConvProblem& pr = cp.prob[k];
          pr.n_src = r * r;
          for (int u = 0; u < r * r; ++u) pr.src[u] = SrcView{0, static_cast<int>(k * TB), 0, r, u % r, u / r};
          pr.grad0 = grad_stack(c, p->off_gx, 0);
          pr.grad1 = grad_stack(c, k == 0 ? dh_top_f : dh_top_b, 0);
        }
        run_conv(c, BCLS_HEAD_DGRAD, 64, EPI_GRAD, p->bm_head_dg[0], cp, conv_fl * 2 / 3);
      }
    }
    ev_head = c.side_mark();

    // ---------------------------------------------------------------- refine block
    if (pos) {
      const double c2_fl = 2.0 * 9 * (2 * kFeat + 1) * kFeat * px * TB;
      const double c1_fl = 2.0 * 9 * (2 * kFeat + 1) * Wn * (2 * kFeat + 1) * px * TB;
      c.to_side();
      run_wgrad(c, BCLS_REFINE_WGRAD, p->bm_c2_wg, p->wl_c2[s], p->h, p->w, c2_fl);
      c.to_main();
      {
        ConvParams cp;
        base_params(p->lr, p->h, p->w, &cp);
        cp.n_img = static_cast<int>(TB);
        cp.n_prob = 1;
        cp.n_total = 144; cp.n_store = 144; cp.out_ch = 144;
        ConvProblem& pr = cp.prob[0];
        pr.n_src = 1;
        pr.src[0] = view0(static_cast<long long>(half) * B);
        pr.out_bf16 = reinterpret_cast<__nv_bfloat16*>(c.ws + p->off_gm) + static_cast<size_t>(half) * B * px * 144;
        set_slab(&cp, p->geo_lr);
        run_conv(c, BCLS_REFINE_DGRAD, 144, EPI_STORE, p->bm_c2_dg, cp, c2_fl);
      }
      if (c.dry || c.G->ref_w1)
        run_simt(c, BCLS_MISC, "posterm_bwd", [&] {
          return launch_posterm_bwd(reinterpret_cast<__nv_bfloat16*>(c.ws + p->off_gm) + static_cast<size_t>(half) * B * px * 144,
                                    c.pos, reinterpret_cast<float*>(c.ws + p->off_sums), c.G->ref_w1, T, B, L, U - half,
                                    Wn, p->h, p->w, 2 * kFeat + 1, (2 * kFeat + 1) * Wn, 2 * kFeat, 144, c.stream);
        });
      c.to_side();
      run_wgrad(c, BCLS_REFINE_WGRAD, p->bm_c1_wg, p->wl_c1[s], p->h, p->w, c1_fl);
      c.to_main();
      {
        ConvParams cp;
        base_params(p->lr, p->h, p->w, &cp);
        cp.n_img = static_cast<int>(TB);
        cp.n_prob = 1;
        cp.kb_per_src = 3; cp.k16_last = 1;
        cp.n_total = 128; cp.n_store = 128; cp.out_ch = 64;
        cp.grad_split = 1;
        ConvProblem& pr = cp.prob[0];
        pr.n_src = Wn;
        for (int sd = 0; sd < Wn; ++sd) pr.src[sd] = view0(static_cast<long long>(sd) * B);
        pr.grad0 = grad_stack(c, dh_top_f, 0);
        pr.grad1 = grad_stack(c, dh_top_b, 0);
        set_slab(&cp, p->geo_lr);
        run_conv(c, BCLS_REFINE_DGRAD, 128, EPI_GRAD, p->bm_c1_dg, cp, c1_fl);
      }
    } else {
      const double c1_fl = 2.0 * (2 * kFeat) * Wn * kFeat * px * TB;
      c.to_side();
      run_wgrad(c, BCLS_REFINE_WGRAD, p->bm_c1_wg, p->wl_c1[s], p->h, p->w, c1_fl);
      c.to_main();
      ConvParams cp;
      base_params(p->lr, p->h, p->w, &cp);
      cp.n_img = static_cast<int>(TB);
      cp.n_prob = 1;
      cp.taps = 1;
      cp.n_total = 128; cp.n_store = 128; cp.out_ch = 64;
      cp.grad_split = 1;
      ConvProblem& pr = cp.prob[0];
      pr.n_src = Wn;
      for (int sd = 0; sd < Wn; ++sd) pr.src[sd] = view0(static_cast<long long>(sd) * B);
      pr.grad0 = grad_stack(c, dh_top_f, 0);
      pr.grad1 = grad_stack(c, dh_top_b, 0);
      run_conv(c, BCLS_REFINE_DGRAD, 128, EPI_GRAD, p->bm_c1_dg, cp, c1_fl);
    }

    ev_refine = c.side_mark();
    // ---------------------------------------------------------------- ConvLSTM: reverse wavefront over the T frames
    c.main_wait(ev_lstm);       // the ConvLSTM weight gradient of stage s + 1 reads dgates
    const double lstm_fl = 2.0 * 9 * 2 * kFeat * 4 * kFeat;
    for (int d = 0; d < T + NL - 1; ++d) {
      LstmBwdParams lp{};
      lp.H = p->h; lp.W = p->w;
      lp.tw_log2 = p->lr.tw_log2; lp.tiles_x = p->lr.tiles_x; lp.tiles_y = p->lr.tiles_y;
      lp.wp = p->geo_lstm.on ? p->geo_lstm.g.wp : 0;
      lp.tiles_per_img = p->geo_lstm.on ? p->geo_lstm.g.tiles : p->lr.tiles_x * p->lr.tiles_y;
      lp.n_img = B;
      ConvParams cp;
      base_params(p->lr, p->h, p->w, &cp);
      cp.n_img = B;
      cp.kb_per_src = 4;
      cp.n_total = p->lstm_dg_bn; cp.n_store = p->lstm_dg_bn; cp.out_ch = 64;
      cp.grad_split = p->lstm_src == 2 ? 1 : 0;
      int np = 0;
      for (int dir = 0; dir < 2; ++dir)
        for (int l = 0; l < NL; ++l) {
          const int tp = d - (NL - 1 - l);
          if (tp < 0 || tp >= T) continue;
          const int j = dir == 0 ? (L - U - 1) - tp : U + tp;   // frame of this backward step
          const int jp = dir == 0 ? j - 1 : j + 1;              // frame whose (h, c) entered the cell
          const int ci = dir * NL + l;
          const long long gimg = static_cast<long long>(j - U) * B;
          LstmBwdProb& lb = lp.prob[np];
          lb.dh = grad_stack(c, p->off_dh + static_cast<size_t>(ci) * p->dh_stride, gimg);
          lb.gates = gates_buf(c, s, dir, l, j - U);
          lb.c = c_state(c, s, dir, l, j);
          lb.c_prev = (jp >= 0 && jp < L) ? c_state(c, s, dir, l, jp) : nullptr;
          lb.dc = reinterpret_cast<float*>(c.ws + p->off_dc) + static_cast<size_t>(ci) * p->c_elems;
          lb.dc_zero = tp == 0 ? 1 : 0;
          __nv_bfloat16* dg = reinterpret_cast<__nv_bfloat16*>(c.ws + p->off_dgates) +
                              (static_cast<size_t>(ci) * TB + gimg) * px * 256;
          lb.dgates = dg;
          ConvProblem& pr = cp.prob[np];
          pr.n_src = 1;
          pr.src[0] = view0(static_cast<long long>(ci) * TB + gimg);
          pr.w_row_base = static_cast<int>(ci * p->lstm_dg_rows_per_cell);
          pr.grad0 = l > 0 ? grad_stack(c, p->off_dh + static_cast<size_t>(ci - 1) * p->dh_stride, gimg)
                           : grad_stack(c, p->off_gx, gimg);
          const bool prev_has_grad = jp >= U && jp < L - U;
          pr.grad1 = (p->lstm_src == 2 && prev_has_grad)
                         ? grad_stack(c, p->off_dh + static_cast<size_t>(ci) * p->dh_stride, static_cast<long long>(jp - U) * B)
                         : nullptr;
          ++np;
        }
      lp.n_prob = np;
      cp.n_prob = np;
      set_slab(&cp, p->geo_lr);
      run_simt(c, BCLS_LSTM_POINT, "lstm_bwd_pointwise", [&] { return launch_lstm_bwd_pointwise(lp, c.stream); });
      run_conv(c, BCLS_LSTM_DGRAD, p->lstm_dg_bn, EPI_GRAD, p->bm_lstm_dg, cp, lstm_fl * px * B * np);
    }
    c.to_side();
    run_wgrad(c, BCLS_LSTM_WGRAD, p->bm_lstm_wg, p->wl_lstm[s], p->h, p->w, lstm_fl * px * TB * 2 * NL);
    c.to_main();
    ev_lstm = c.side_mark();
  }

  // ---------------------------------------------------------------- in_block (refine_net.py:188-192), gradient frames
  if (c.dry || (c.G->in_w && c.G->in_b && c.G->in_slope))
    run_simt(c, BCLS_MISC, "in_conv_prelu_bwd", [&] {
      return launch_in_conv_prelu_bwd(c.lr + static_cast<size_t>(U) * B * px, c.P->in_w, c.P->in_b, c.P->in_slope,
                                      grad_stack(c, p->off_gx, 0), c.G->in_w, c.G->in_b, c.G->in_slope, TB, p->h, p->w,
                                      p->num_sms, c.stream);
    });

  if (tail) {
    // parameter gradients of the tail from the accumulated correlation sums (all stages, all lists)
    c.to_side();
    run_simt(c, BCLS_MISC, "tail finish", [&] {
      return launch_tail_finish(tail_ws, c.P->head_w[p->n_ps - 1], c.P->head_b[p->n_ps - 1], c.P->head_w[p->n_ps],
                                c.G->head_w[p->n_ps - 1], c.G->head_b[p->n_ps - 1], c.G->head_w[p->n_ps],
                                c.G->head_b[p->n_ps], c.stream);
    });
    c.to_main();
  }
  // ---------------------------------------------------------------- packed gradients -> parameter layout
  c.join();
  const int32_t* idx = reinterpret_cast<const int32_t*>(c.pk + p->pk_idx);
  if (p->table_ok && get_pack_table()) {
    // one table launch (uploaded by upload_scatter_table before the schedule runs / is captured)
    run_simt(c, BCLS_MISC, "scatter table", [&] {
      long long max_n = 0;
      int n = 0;
      for (const auto& j : p->sc_jobs) {
        if (!job_grad(j.kind, j.a, j.b, c.G, j.is_bias)) continue;
        max_n = j.n > max_n ? j.n : max_n;
        ++n;
      }
      return launch_table(reinterpret_cast<const pvsr_table_job*>(c.pk + p->pk_sc_table), n, max_n, c.stream);
    });
    return;
  }
  for (const auto& j : p->sc_jobs) {
    float* dst = c.dry ? nullptr : job_grad(j.kind, j.a, j.b, c.G, j.is_bias);
    if (!c.dry && !dst) continue;
    run_simt(c, BCLS_MISC, "scatter_add", [&] {
      return launch_scatter_add(dst, idx + j.idx, j.has2 ? idx + j.idx2 : nullptr, wg + j.src, j.n, c.stream);
    });
  }
}

int build_maps(pvsr_plan* p, const void* ws, const void* pk) {
  if (p->maps_ws == ws && p->maps_pk == pk && p->maps_cg == get_cta_pair() && p->maps_halo == get_halo_mode())
    return 0;
  const uint8_t* w = static_cast<const uint8_t*>(ws);
  const uint8_t* k = static_cast<const uint8_t*>(pk);
  int rc = 0;
  const long long TB = static_cast<long long>(p->T) * p->B;
  // Slab launches read (rows x Wp)-position boxes (conv.h: PrGeom); box launches read TH x TW tiles.
  const bool slab = get_halo_mode() != 0;
  p->geo_lr.on = slab && choose_pr(p->h, p->w, 1, &p->geo_lr.g);
  p->geo_lstm = p->geo_lr;   // the state tensors follow the same tile -> pixel map (conv.h: lstm_tile_geometry)
  for (int q = 0; q < p->n_ps; ++q) {
    if (q == 0) p->geo_ps[q] = p->geo_lr;
    else p->geo_ps[q].on = slab && choose_pr(p->ps_h[q], p->ps_w[q], 1, &p->geo_ps[q].g);
    p->geo_hdg[q].on = slab && p->train && choose_pr(p->ps_h[q], p->ps_w[q], p->ps_r[q], &p->geo_hdg[q].g);
  }
  auto box_w = [](const pvsr_plan::Geo& g, const Tiling& t) { return g.on ? g.g.wp : t.tw; };
  auto box_h = [](const pvsr_plan::Geo& g, const Tiling& t) { return g.on ? g.g.rows : t.th; };
  CUtensorMap tm_act, tm_act_conv, tm_act_lstm;
  rc |= make_act_tmap(&tm_act, w + p->off_act, kFeat, p->w, p->h, p->act_images, p->lr.tw, p->lr.th);
  rc |= make_act_tmap(&tm_act_conv, w + p->off_act, kFeat, p->w, p->h, p->act_images, box_w(p->geo_lr, p->lr),
                      box_h(p->geo_lr, p->lr));
  rc |= make_act_tmap(&tm_act_lstm, w + p->off_act, kFeat, p->w, p->h, p->act_images, box_w(p->geo_lstm, p->lr),
                      box_h(p->geo_lstm, p->lr));
  p->maps_lstm.act[0] = tm_act_lstm;
  p->maps_c1.act[0] = p->cfg.pos_enc ? tm_act_conv : tm_act;     // the 1x1 variant has a single tap
  p->maps_head[0].act[0] = tm_act_conv;
  if (p->cfg.pos_enc)
    rc |= make_act_tmap(&p->maps_c2.act[0], w + p->off_mid, 144, p->w, p->h,
                        static_cast<long long>(p->n_slots) * p->n_win * p->B, box_w(p->geo_lr, p->lr),
                        box_h(p->geo_lr, p->lr));
  for (int q = 1; q < p->n_ps; ++q)
    rc |= make_act_tmap(&p->maps_head[q].act[0], w + p->off_head[q - 1], kFeat, p->ps_w[q], p->ps_h[q],
                        static_cast<long long>(p->n_list_slots) * TB, box_w(p->geo_ps[q], p->ps_tile[q]),
                        box_h(p->geo_ps[q], p->ps_tile[q]));
  rc |= make_weight_tmap(&p->maps_lstm.w, k + p->pk_lstm_w, static_cast<long long>(2 * p->NL) * p->lstm_rows_per_cell,
                         256);
  rc |= make_weight_tmap(&p->maps_c1.w, k + p->pk_c1_w, p->c1_rows, p->cfg.pos_enc ? 144 : 64);
  if (p->cfg.pos_enc) rc |= make_weight_tmap(&p->maps_c2.w, k + p->pk_c2_w, p->c2_rows, 64);
  for (int q = 0; q < p->n_ps; ++q)
    rc |= make_weight_tmap(&p->maps_head[q].w, k + p->pk_head_w[q], p->head_rows[q], p->ps_bn[q]);
  if (p->n_ps >= 2 && p->ps_r[p->n_ps - 1] == 2) {
    p->maps_t36.act[0] = p->maps_head[p->n_ps - 1].act[0];      // the input of the last shuffle conv
    rc |= make_weight_tmap(&p->maps_t36.w, k + p->pk_t36_w, 9LL * 48, 48);
  }

  if (p->train) {
    const int n_pad = p->T + 2 * p->half;
    CUtensorMap tm_dgates, tm_gr, tm_gm, tm_dgates_conv, tm_gr_conv, tm_gm_conv;
    rc |= make_act_tmap(&tm_dgates, w + p->off_dgates, 256, p->w, p->h, 2LL * p->NL * TB, p->lr.tw, p->lr.th);
    const int lr_bw = box_w(p->geo_lr, p->lr), lr_bh = box_h(p->geo_lr, p->lr);
    rc |= make_act_tmap(&tm_dgates_conv, w + p->off_dgates, 256, p->w, p->h, 2LL * p->NL * TB, lr_bw, lr_bh);
    rc |= make_act_tmap(&tm_gr, w + p->off_gr, kFeat, p->w, p->h, static_cast<long long>(n_pad) * p->B, p->lr.tw,
                        p->lr.th);
    rc |= make_act_tmap(&tm_gr_conv, w + p->off_gr, kFeat, p->w, p->h, static_cast<long long>(n_pad) * p->B, lr_bw,
                        lr_bh);
    if (p->cfg.pos_enc) {
      rc |= make_act_tmap(&tm_gm, w + p->off_gm, 144, p->w, p->h, static_cast<long long>(n_pad) * p->B, p->lr.tw,
                          p->lr.th);
      rc |= make_act_tmap(&tm_gm_conv, w + p->off_gm, 144, p->w, p->h, static_cast<long long>(n_pad) * p->B, lr_bw,
                          lr_bh);
    }
    // ConvLSTM
    p->bm_lstm_dg.act[0] = tm_dgates_conv;
    rc |= make_weight_tmap(&p->bm_lstm_dg.w, k + p->pk_lstm_dg, 2LL * p->NL * p->lstm_dg_rows_per_cell, p->lstm_dg_bn);
    p->bm_lstm_wg.act[0] = tm_act;
    p->bm_lstm_wg.act[1] = tm_dgates;
    // refine
    if (p->cfg.pos_enc) {
      p->bm_c2_dg.act[0] = tm_gr_conv;
      rc |= make_weight_tmap(&p->bm_c2_dg.w, k + p->pk_c2_dg, p->c2_dg_rows, 144);
      rc |= make_act_tmap(&p->bm_c2_wg.act[0], w + p->off_mid, 144, p->w, p->h,
                          static_cast<long long>(p->n_slots) * p->n_win * p->B, p->lr.tw, p->lr.th);
      p->bm_c2_wg.act[1] = tm_gr;
      p->bm_c1_dg.act[0] = tm_gm_conv;
      p->bm_c1_wg.act[0] = tm_act;
      p->bm_c1_wg.act[1] = tm_gm;
    } else {
      p->bm_c1_dg.act[0] = tm_gr;
      p->bm_c1_wg.act[0] = tm_act;
      p->bm_c1_wg.act[1] = tm_gr;
    }
    rc |= make_weight_tmap(&p->bm_c1_dg.w, k + p->pk_c1_dg, p->c1_dg_rows, 128);
    // heads: gradient wrt the output of conv q lives at resolution q+1 and is read pixel-unshuffled (mul = r)
    for (int q = 0; q < p->n_ps; ++q) {
      CUtensorMap tm_dy, tm_dy_dg;
      rc |= make_act_tmap(&tm_dy, w + p->off_dhead[q], kFeat, p->ps_w[q + 1], p->ps_h[q + 1], 3 * TB, p->bw_tile[q].tw,
                          p->bw_tile[q].th, p->ps_r[q]);
      rc |= make_act_tmap(&tm_dy_dg, w + p->off_dhead[q], kFeat, p->ps_w[q + 1], p->ps_h[q + 1], 3 * TB,
                          box_w(p->geo_hdg[q], p->bw_tile[q]), box_h(p->geo_hdg[q], p->bw_tile[q]), p->ps_r[q]);
      p->bm_head_dg[q].act[0] = tm_dy_dg;
      rc |= make_weight_tmap(&p->bm_head_dg[q].w, k + p->pk_head_dg[q], p->head_dg_rows[q], 64);
      if (q == 0)
        rc |= make_act_tmap(&p->bm_head_wg[q].act[0], w + p->off_act, kFeat, p->w, p->h, p->act_images,
                            p->bw_tile[q].tw, p->bw_tile[q].th);
      else
        rc |= make_act_tmap(&p->bm_head_wg[q].act[0], w + p->off_head[q - 1], kFeat, p->ps_w[q], p->ps_h[q],
                            static_cast<long long>(p->n_list_slots) * TB, p->bw_tile[q].tw, p->bw_tile[q].th);
      p->bm_head_wg[q].act[1] = tm_dy;
    }
  }
  if (rc) return set_error(-20, "tensor map encode failed");
  p->maps_ws = ws;
  p->maps_pk = pk;
  p->maps_cg = get_cta_pair();
  p->maps_halo = get_halo_mode();
  for (auto& g : p->graphs) cudaGraphExecDestroy(g.second);
  p->graphs.clear();
  return 0;
}

void push_wg_launch(pvsr_plan* p, pvsr_plan::WgLaunch* wl, const std::vector<WgJob>& jobs, int n_img, const Tiling& t) {
  wl->job_begin = static_cast<int>(p->wg_jobs.size());
  wl->n_jobs = static_cast<int>(jobs.size());
  wl->n_img = n_img;
  wl->tile = t;
  std::vector<WgJob> sorted(jobs);
  wl->n_heavy = sort_wgrad_jobs(sorted.data(), static_cast<int>(sorted.size()));   // heavy-first (launch_wgrad)
  p->wg_jobs.insert(p->wg_jobs.end(), sorted.begin(), sorted.end());
}

// Weight-gradient job lists of every stage (image bases are fixed by the workspace layout).
void build_wg_jobs(pvsr_plan* p) {
  const int B = p->B, U = p->U, T = p->T, NL = p->NL, half = p->half, Wn = p->Wn;
  const int TB = T * B;
  for (int s = 0; s < p->S; ++s) {
    // ---- ConvLSTM cells: X = [x_in | h_prev] over the T gradient frames, dY = dgates (4 chunks)
    std::vector<WgJob> jobs;
    for (int dir = 0; dir < 2; ++dir)
      for (int l = 0; l < NL; ++l) {
        const int ci = dir * NL + l;
        std::vector<WgSource> srcs;
        const long long xin = (l == 0 ? p->img_x[s] : p->img_h[s][dir][l - 1]) + static_cast<long long>(U) * B;
        srcs.push_back(WgSource{SrcView{0, static_cast<int>(xin), 0, 1, 0, 0}});
        if (p->lstm_src == 2) {
          const long long hp = p->img_h[s][dir][l] + static_cast<long long>(dir == 0 ? U - 1 : U + 1) * B;
          srcs.push_back(WgSource{SrcView{0, static_cast<int>(hp), 0, 1, 0, 0}});
        }
        std::vector<WgChunk> chunks;
        chunks.push_back(WgChunk{SrcView{1, ci * TB, 0, 1, 0, 0}, 0});
        build_wgrad_jobs(srcs, 1, 9, chunks, 256, true, p->wg_lstm_w[ci], p->wg_lstm_b[ci], &jobs, 256);
      }
    push_wg_launch(p, &p->wl_lstm[s], jobs, TB, p->lr);

    // ---- refine conv1: X = 2*Wn hidden-map sources of the window, dY = d(conv1 out) (pos) or dr (1x1 variant)
    jobs.clear();
    {
      std::vector<WgSource> srcs;
      for (int jw = 0; jw < Wn; ++jw)
        for (int dir = 0; dir < 2; ++dir) {
          const long long base = p->img_h[s][dir][NL - 1] + static_cast<long long>(U - half + jw) * B;
          srcs.push_back(WgSource{SrcView{0, static_cast<int>(base), 0, 1, 0, 0}});
        }
      std::vector<WgChunk> chunks;
      if (p->cfg.pos_enc) {
        chunks.push_back(WgChunk{SrcView{1, half * B, 0, 1, 0, 0}, 0});
        build_wgrad_jobs(srcs, 1, 9, chunks, p->c1_wg_ntotal, true, p->wg_c1_w, p->wg_c1_b, &jobs, 144);
      } else {
        chunks.push_back(WgChunk{SrcView{1, half * B, 0, 1, 0, 0}, 0});
        build_wgrad_jobs(srcs, 1, 1, chunks, p->c1_wg_ntotal, true, p->wg_c1_w, p->wg_c1_b, &jobs, 64);
      }
    }
    push_wg_launch(p, &p->wl_c1[s], jobs, TB, p->lr);

    // ---- refine conv2: X = conv1 output (144 stored channels) at the gradient frames, dY = dr
    jobs.clear();
    if (p->cfg.pos_enc) {
      std::vector<WgSource> srcs;
      srcs.push_back(WgSource{SrcView{0, (s * p->n_win + (U - half)) * B, 0, 1, 0, 0}});
      std::vector<WgChunk> chunks;
      chunks.push_back(WgChunk{SrcView{1, half * B, 0, 1, 0, 0}, 0});
      build_wgrad_jobs(srcs, 3, 9, chunks, 64, true, p->wg_c2_w, p->wg_c2_b, &jobs, 64);
    }
    push_wg_launch(p, &p->wl_c2[s], jobs, TB, p->lr);

    // ---- head conv q + PixelShuffle: X = conv input of the 3 lists, dY = r*r pixel-unshuffled chunk views
    for (int q = 0; q < p->n_ps; ++q) {
      jobs.clear();
      const int r = p->ps_r[q];
      for (int k = 0; k < 3; ++k) {
        long long xin;
        if (q > 0) xin = static_cast<long long>(3 * s + k) * TB;
        else if (k == 2) xin = p->img_x[s + 1] + static_cast<long long>(U) * B;
        else xin = p->img_sum[s][k];
        std::vector<WgSource> srcs;
        srcs.push_back(WgSource{SrcView{0, static_cast<int>(xin), 0, 1, 0, 0}});
        std::vector<WgChunk> chunks;
        for (int u = 0; u < r * r; ++u) chunks.push_back(WgChunk{SrcView{1, k * TB, 0, r, u % r, u / r}, 64 * u});
        build_wgrad_jobs(srcs, 1, 9, chunks, kFeat * r * r, true, p->wg_head_w[q], p->wg_head_b[q], &jobs);
      }
      push_wg_launch(p, &p->wl_head[s][q], jobs, TB, p->bw_tile[q]);
    }
  }
}

// Runs `eager(stream)` directly, or - with use_graph - eagerly on first use of `key` (this call's result; also sets
// the kernel attributes) followed by a capture of the same schedule on an internal stream for later replays.
template <typename F>
static int run_or_replay(pvsr_plan* p, const GraphKey& key, int use_graph, cudaStream_t s, F&& eager) {
  int e = 0;
  const bool two = get_two_branch() != 0 && p->train;
  if (two && !p->side_stream) {
    e = cudaStreamCreateWithFlags(&p->side_stream, cudaStreamNonBlocking);
    if (!e) e = cudaStreamCreateWithFlags(&p->cap_side, cudaStreamNonBlocking);
    if (e) return check_cuda(e, "side stream");
  }
  if (!use_graph) return eager(s, two ? p->side_stream : nullptr);
  auto it = p->graphs.find(key);
  if (it != p->graphs.end()) return check_cuda(cudaGraphLaunch(it->second, s), "graph launch");
  int rc = eager(s, two ? p->side_stream : nullptr);
  if (rc) return rc;
  if (!p->cap_stream) {
    e = cudaStreamCreateWithFlags(&p->cap_stream, cudaStreamNonBlocking);
    if (e) return check_cuda(e, "capture stream");
  }
  cudaGraph_t graph = nullptr;
  e = cudaStreamBeginCapture(p->cap_stream, cudaStreamCaptureModeThreadLocal);
  if (e) return check_cuda(e, "begin capture");
  // small inference plans are launch-latency bound: their graph is captured with programmatic dependent launch
  const bool auto_pdl = !p->train && !pdl_is_explicit() && get_pdl() == 0 &&
                        static_cast<long long>(p->B) * p->h * p->w <= kPdlAutoPixels;
  if (auto_pdl) set_pdl(1);
  rc = eager(p->cap_stream, two ? p->cap_side : nullptr);
  if (auto_pdl) set_pdl(0);
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

}  // namespace

extern "C" {

int pvsr_debug_dump_trace(void) {
  int first[2][2] = {{-1, -1}, {-1, -1}}, done = 0;
  for (size_t i = 0; i < g_trace.size(); ++i) {
    const TraceEntry& t = g_trace[i];
    if (cudaEventQuery(t.ev) == cudaSuccess) { ++done; continue; }
    if (first[t.bwd][t.side] < 0) first[t.bwd][t.side] = static_cast<int>(i);
  }
  fprintf(stderr, "[pvsr trace] %zu launches recorded, %d complete\n", g_trace.size(), done);
  for (int b = 0; b < 2; ++b)
    for (int sd = 0; sd < 2; ++sd)
      if (first[b][sd] >= 0) {
        const TraceEntry& t = g_trace[first[b][sd]];
        fprintf(stderr, "[pvsr trace] first incomplete: %s schedule, %s branch, launch #%d, class %d\n",
                b ? "backward" : "forward", sd ? "side" : "main", t.seq, t.cls);
      }
  return 0;
}
void pvsr_debug_clear_trace(void) {
  for (auto& t : g_trace) cudaEventDestroy(t.ev);
  g_trace.clear();
}

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
  if (cfg->save_for_backward && !cfg->all_heads) return set_error(-3, "training plans need all_heads = 1");
  if (cfg->save_for_backward && cfg->n_updated < cfg->window / 2)
    return set_error(-3, "training needs num_updated_frames >= refine_window_size // 2");

  pvsr_plan* p = new pvsr_plan();
  p->cfg = *cfg;
  p->B = cfg->batch; p->L = cfg->n_frames; p->U = cfg->n_updated; p->T = p->L - 2 * p->U;
  p->h = cfg->h; p->w = cfg->w; p->S = cfg->n_stages; p->Wn = cfg->window; p->half = p->Wn / 2;
  p->NL = cfg->n_layers; p->scale = cfg->scale;
  p->n_win = p->L - 2 * p->half;
  p->n_lists = cfg->all_heads ? 3 * p->S : 1;
  p->train = cfg->save_for_backward != 0;
  p->n_slots = p->train ? p->S : 1;
  p->n_list_slots = p->train ? 3 * p->S : 1;
  p->lr.set(p->h, p->w);
  if (p->scale == 3) { p->n_ps = 1; p->ps_r[0] = 3; }
  else { p->n_ps = static_cast<int>(std::lround(std::log2(p->scale))); for (int q = 0; q < p->n_ps; ++q) p->ps_r[q] = 2; }
  p->ps_h[0] = p->h; p->ps_w[0] = p->w;
  for (int q = 0; q < p->n_ps; ++q) {
    p->ps_h[q + 1] = p->ps_h[q] * p->ps_r[q];
    p->ps_w[q + 1] = p->ps_w[q] * p->ps_r[q];
    p->ps_tile[q].set(p->ps_h[q], p->ps_w[q]);
    p->bw_tile[q].set(p->ps_h[q], p->ps_w[q], p->ps_r[q] <= 2 ? 128 : 64);
    if (p->ps_r[q] == 2) { p->ps_bn[q] = 256; p->ps_nt[q] = 1; } else { p->ps_bn[q] = 192; p->ps_nt[q] = 3; }
  }
  p->Hs = p->ps_h[p->n_ps]; p->Ws = p->ps_w[p->n_ps];
  p->lstm_src = cfg->memory ? 2 : 1;
  p->lstm_rows_per_cell = p->lstm_src * 9 * 256;

  // ---- workspace layout
  const long long LB = static_cast<long long>(p->L) * p->B, TB = static_cast<long long>(p->T) * p->B;
  const size_t px = static_cast<size_t>(p->h) * p->w;
  p->img_bytes = px * kFeat * 2;
  long long img = 0;
  for (int s = 0; s <= p->S; ++s) { p->img_x[s] = img; img += LB; }
  for (int s = 0; s < p->n_slots; ++s)
    for (int d = 0; d < 2; ++d) for (int l = 0; l < p->NL; ++l) { p->img_h[s][d][l] = img; img += LB; }
  for (int s = 0; s < p->n_slots; ++s)
    for (int k = 0; k < 2; ++k) { p->img_sum[s][k] = img; img += cfg->all_heads ? TB : 0; }
  p->act_images = img;
  size_t off = 0;
  p->off_act = off; off = align_up(off + static_cast<size_t>(img) * p->img_bytes, 1024);
  p->mid_ch = 144;
  p->off_mid = off;
  p->mid_stride = static_cast<size_t>(p->n_win) * p->B * px * 144 * 2;
  if (cfg->pos_enc) off = align_up(off + p->mid_stride * p->n_slots, 1024);
  p->c_elems = static_cast<size_t>(pvsr_lstm_state_elems(p->B, p->h, p->w));
  p->off_c = off;
  off = align_up(off + p->c_elems * 4 * 2 * p->NL * (p->train ? static_cast<size_t>(p->S) * p->L : 1), 1024);
  p->off_gates = off;
  if (p->train) off = align_up(off + p->c_elems * 4 * 2 * static_cast<size_t>(2 * p->NL) * p->S * p->T, 1024);
  p->off_posterm = off;
  if (cfg->pos_enc) off = align_up(off + static_cast<size_t>(p->n_win) * p->B * 16 * 144 * 4, 1024);
  for (int q = 0; q < p->n_ps; ++q) {
    p->off_head[q] = off;
    p->head_stride[q] = static_cast<size_t>(TB) * p->ps_h[q + 1] * p->ps_w[q + 1] * kFeat * 2;
    off = align_up(off + p->head_stride[q] * p->n_list_slots, 1024);
  }

  // ---- packed parameter layout + gather indices
  size_t pk = 0;
  const int lstm_offs[2] = {0, kFeat};
  const int o0[1] = {0}, o1[1] = {kFeat};
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
        pvsr_pack_spec sa = make_spec(256, 128, 3, 1, o0, kFeat, 1, 9, 256, 0);
        pvsr_pack_spec sb = make_spec(256, 128, 3, 1, o1, kFeat, 1, 9, 256, 0);
        add_pack_job(p, PK_LSTM, d, l, sa, &sb, dw, db);
      }
    }
  const int per = 2 * kFeat + 1;
  int c1_offs[PVSR_MAX_SRC];
  {
    if (cfg->pos_enc) {
      for (int s = 0; s < 2 * p->Wn; ++s) c1_offs[s] = per * (s / 2) + kFeat * (s % 2);
      pvsr_pack_spec s1 = make_spec(per, per * p->Wn, 3, 2 * p->Wn, c1_offs, kFeat, 1, 9, 144, 0);
      p->c1_rows = static_cast<long long>(2 * p->Wn) * 9 * 144;
      p->pk_c1_w = pk; pk = align_up(pk + static_cast<size_t>(p->c1_rows) * 128, 1024);
      add_pack_job(p, PK_C1, 0, 0, s1, nullptr, p->pk_c1_w, static_cast<size_t>(-1));
      pvsr_pack_spec s2 = make_spec(kFeat, per, 3, 1, o0, per, 3, 9, kFeat, 0);
      p->c2_rows = 27LL * 64;
      p->pk_c2_w = pk; pk = align_up(pk + static_cast<size_t>(p->c2_rows) * 128, 1024);
      p->pk_c2_b = pk; pk = align_up(pk + 64 * 4, 1024);
      add_pack_job(p, PK_C2, 0, 0, s2, nullptr, p->pk_c2_w, p->pk_c2_b);
    } else {
      for (int s = 0; s < 2 * p->Wn; ++s) c1_offs[s] = 2 * kFeat * (s / 2) + kFeat * (s % 2);
      pvsr_pack_spec s1 = make_spec(kFeat, 2 * kFeat * p->Wn, 1, 2 * p->Wn, c1_offs, kFeat, 1, 1, kFeat, 0);
      p->c1_rows = static_cast<long long>(2 * p->Wn) * 64;
      p->pk_c1_w = pk; pk = align_up(pk + static_cast<size_t>(p->c1_rows) * 128, 1024);
      p->pk_c2_w = p->pk_c1_w; p->c2_rows = 0;
      p->pk_c2_b = pk; pk = align_up(pk + 64 * 4, 1024);   // bias of the 1x1 conv
      add_pack_job(p, PK_C1, 0, 0, s1, nullptr, p->pk_c1_w, p->pk_c2_b);
    }
  }
  for (int q = 0; q < p->n_ps; ++q) {
    const int r = p->ps_r[q];
    pvsr_pack_spec sh = make_spec(kFeat * r * r, kFeat, 3, 1, o0, kFeat, 1, 9, kFeat * r * r, r);
    p->head_rows[q] = 9LL * kFeat * r * r;
    p->pk_head_w[q] = pk; pk = align_up(pk + static_cast<size_t>(p->head_rows[q]) * 128, 1024);
    p->pk_head_b[q] = pk; pk = align_up(pk + static_cast<size_t>(kFeat) * r * r * 4, 1024);
    add_pack_job(p, PK_HEAD, q, 0, sh, nullptr, p->pk_head_w[q], p->pk_head_b[q]);
  }

  if (p->train) {
    // ---- backward workspace
    p->dh_stride = static_cast<size_t>(TB) * px * kFeat * 4;
    p->off_gx = off; off = align_up(off + p->dh_stride, 1024);
    p->off_dh = off; off = align_up(off + p->dh_stride * 2 * p->NL, 1024);
    p->off_dc = off; off = align_up(off + p->c_elems * 4 * 2 * p->NL, 1024);
    p->off_dgates = off; off = align_up(off + static_cast<size_t>(2 * p->NL) * TB * px * 256 * 2, 1024);
    const size_t n_pad = static_cast<size_t>(p->T + 2 * p->half) * p->B;
    p->off_gr = off; off = align_up(off + n_pad * px * kFeat * 2, 1024);
    p->off_gm = off;
    if (cfg->pos_enc) off = align_up(off + n_pad * px * 144 * 2, 1024);
    for (int q = 0; q < p->n_ps; ++q) {
      p->off_dhead[q] = off;
      off = align_up(off + 3 * static_cast<size_t>(TB) * p->ps_h[q + 1] * p->ps_w[q + 1] * kFeat * 2, 1024);
    }
    p->off_sums = off; off = align_up(off + static_cast<size_t>(p->Wn) * 16 * 144 * 4, 1024);
    p->off_tail = off; off = align_up(off + tail_scratch_bytes(), 1024);

    // ---- data-gradient operands: transposed, spatially flipped weights
    {
      // ConvLSTM gate conv: K = 256 gate channels (4 blocks), N = [d x | d h_prev] (128) or d x with summed halves
      p->lstm_dg_bn = cfg->memory ? 128 : 64;
      p->lstm_dg_rows_per_cell = 36LL * p->lstm_dg_bn;
      p->pk_lstm_dg = pk; pk = align_up(pk + static_cast<size_t>(2 * p->NL) * p->lstm_dg_rows_per_cell * 128, 1024);
      for (int d = 0; d < 2; ++d)
        for (int l = 0; l < p->NL; ++l) {
          const int ci = d * p->NL + l;
          const size_t dst = p->pk_lstm_dg + static_cast<size_t>(ci) * p->lstm_dg_rows_per_cell * 128;
          pvsr_pack_spec sa = make_spec(256, 128, 3, 1, o0, 256, 4, 9, p->lstm_dg_bn, 0);
          sa.transpose_flip = 1;
          if (cfg->memory) {
            add_pack_job(p, PK_LSTM, d, l, sa, nullptr, dst, static_cast<size_t>(-1));
          } else {
            pvsr_pack_spec sb = sa;
            sb.src_col_off[0] = kFeat;
            add_pack_job(p, PK_LSTM, d, l, sa, &sb, dst, static_cast<size_t>(-1));
          }
        }
      if (cfg->pos_enc) {
        // conv2: K = 64 (dr), N = 144 (129 real conv1-output channels)
        pvsr_pack_spec s2 = make_spec(kFeat, per, 3, 1, o0, kFeat, 1, 9, 144, 0);
        s2.transpose_flip = 1;
        p->c2_dg_rows = 9LL * 144;
        p->pk_c2_dg = pk; pk = align_up(pk + static_cast<size_t>(p->c2_dg_rows) * 128, 1024);
        add_pack_job(p, PK_C2, 0, 0, s2, nullptr, p->pk_c2_dg, static_cast<size_t>(-1));
        // conv1 as a gather over the window: source sd reads d(conv1 out) of frame j' + sd - half, whose window
        // position of frame j' is Wn-1-sd; K = 129 conv1-output channels (3 blocks), N = [d hf | d hb]
        int zero_offs[PVSR_MAX_SRC] = {0};
        pvsr_pack_spec s1 = make_spec(per, per * p->Wn, 3, p->Wn, zero_offs, per, 3, 9, 128, 0);
        s1.transpose_flip = 1;
        for (int sd = 0; sd < p->Wn; ++sd) s1.src_col_off[sd] = per * (p->Wn - 1 - sd);
        p->c1_dg_rows = static_cast<long long>(p->Wn) * 9 * 3 * 128;
        p->pk_c1_dg = pk; pk = align_up(pk + static_cast<size_t>(p->c1_dg_rows) * 128, 1024);
        add_pack_job(p, PK_C1, 0, 0, s1, nullptr, p->pk_c1_dg, static_cast<size_t>(-1));
      } else {
        int zero_offs[PVSR_MAX_SRC] = {0};
        pvsr_pack_spec s1 = make_spec(kFeat, 2 * kFeat * p->Wn, 1, p->Wn, zero_offs, kFeat, 1, 1, 128, 0);
        s1.transpose_flip = 1;
        for (int sd = 0; sd < p->Wn; ++sd) s1.src_col_off[sd] = 2 * kFeat * (p->Wn - 1 - sd);
        p->c1_dg_rows = static_cast<long long>(p->Wn) * 128;
        p->pk_c1_dg = pk; pk = align_up(pk + static_cast<size_t>(p->c1_dg_rows) * 128, 1024);
        add_pack_job(p, PK_C1, 0, 0, s1, nullptr, p->pk_c1_dg, static_cast<size_t>(-1));
        p->pk_c2_dg = p->pk_c1_dg; p->c2_dg_rows = 0;
      }
      for (int q = 0; q < p->n_ps; ++q) {
        const int r = p->ps_r[q];
        int zero_offs[PVSR_MAX_SRC] = {0};
        pvsr_pack_spec sh = make_spec(kFeat * r * r, kFeat, 3, r * r, zero_offs, kFeat, 1, 9, kFeat, 0);
        sh.transpose_flip = 1;
        sh.k_ps_r = r;
        p->head_dg_rows[q] = static_cast<long long>(r * r) * 9 * 64;
        p->pk_head_dg[q] = pk; pk = align_up(pk + static_cast<size_t>(p->head_dg_rows[q]) * 128, 1024);
        add_pack_job(p, PK_HEAD, q, 0, sh, nullptr, p->pk_head_dg[q], static_cast<size_t>(-1));
      }
    }

    // ---- packed weight-gradient buffer + scatter jobs
    long long we = 0;
    for (int d = 0; d < 2; ++d)
      for (int l = 0; l < p->NL; ++l) {
        const int ci = d * p->NL + l;
        p->wg_lstm_w[ci] = we; we += static_cast<long long>(p->lstm_src) * 9 * 256 * 64;
        p->wg_lstm_b[ci] = we; we += 256;
        if (cfg->memory) {
          pvsr_pack_spec sp = make_spec(256, 128, 3, 2, lstm_offs, kFeat, 1, 9, 256, 0);
          add_scatter_job(p, PK_LSTM, d, l, sp, nullptr, p->wg_lstm_w[ci], p->wg_lstm_b[ci]);
        } else {
          pvsr_pack_spec sa = make_spec(256, 128, 3, 1, o0, kFeat, 1, 9, 256, 0);
          pvsr_pack_spec sb = make_spec(256, 128, 3, 1, o1, kFeat, 1, 9, 256, 0);
          add_scatter_job(p, PK_LSTM, d, l, sa, &sb, p->wg_lstm_w[ci], p->wg_lstm_b[ci]);
        }
      }
    if (cfg->pos_enc) {
      p->c1_wg_ntotal = 192;
      pvsr_pack_spec s1 = make_spec(per, per * p->Wn, 3, 2 * p->Wn, c1_offs, kFeat, 1, 9, 192, 0);
      p->wg_c1_w = we; we += static_cast<long long>(2 * p->Wn) * 9 * 192 * 64;
      p->wg_c1_b = we; we += 192;
      add_scatter_job(p, PK_C1, 0, 0, s1, nullptr, p->wg_c1_w, p->wg_c1_b);
      pvsr_pack_spec s2 = make_spec(kFeat, per, 3, 1, o0, per, 3, 9, kFeat, 0);
      p->wg_c2_w = we; we += 27LL * 64 * 64;
      p->wg_c2_b = we; we += 64;
      add_scatter_job(p, PK_C2, 0, 0, s2, nullptr, p->wg_c2_w, p->wg_c2_b);
    } else {
      p->c1_wg_ntotal = 64;
      pvsr_pack_spec s1 = make_spec(kFeat, 2 * kFeat * p->Wn, 1, 2 * p->Wn, c1_offs, kFeat, 1, 1, kFeat, 0);
      p->wg_c1_w = we; we += static_cast<long long>(2 * p->Wn) * 64 * 64;
      p->wg_c1_b = we; we += 64;
      add_scatter_job(p, PK_C1, 0, 0, s1, nullptr, p->wg_c1_w, p->wg_c1_b);
      p->wg_c2_w = p->wg_c2_b = 0;
    }
    for (int q = 0; q < p->n_ps; ++q) {
      const int r = p->ps_r[q];
      pvsr_pack_spec sh = make_spec(kFeat * r * r, kFeat, 3, 1, o0, kFeat, 1, 9, kFeat * r * r, r);
      p->wg_head_w[q] = we; we += 9LL * kFeat * r * r * 64;
      p->wg_head_b[q] = we; we += kFeat * r * r;
      add_scatter_job(p, PK_HEAD, q, 0, sh, nullptr, p->wg_head_w[q], p->wg_head_b[q]);
    }
    p->wg_elems = (we + 3) / 4 * 4;
    p->off_wg = off; off = align_up(off + static_cast<size_t>(p->wg_elems) * 4, 1024);
    build_wg_jobs(p);
    p->off_jobs = off; off = align_up(off + p->wg_jobs.size() * sizeof(WgJob), 1024);
  }
  if (p->n_ps >= 2 && p->ps_r[p->n_ps - 1] == 2) {
    p->off_t36 = off;
    off = align_up(off + static_cast<size_t>(TB) * p->ps_h[p->n_ps - 1] * p->ps_w[p->n_ps - 1] * 48 * (p->train ? 4 : 2), 1024);
  }
  p->ws_bytes = off;
  p->pk_idx = pk;
  pk = align_up(pk + p->idx_host.size() * 4, 1024);
  p->table_ok = true;
  for (const auto& j : p->jobs) if (j.has2) p->table_ok = false;
  for (const auto& j : p->sc_jobs) if (j.has2) p->table_ok = false;
  p->pk_table = pk; pk = align_up(pk + p->jobs.size() * sizeof(pvsr_table_job), 1024);
  p->pk_sc_table = pk; pk = align_up(pk + (p->sc_jobs.size() + 1) * sizeof(pvsr_table_job), 1024);
  p->pk_tail_fwd = pk; pk = align_up(pk + tail_fwd_table_bytes(), 1024);
  p->pk_t36_w = pk; pk = align_up(pk + tail36_weight_bytes(), 1024);
  p->pk_t36_b = pk; pk = align_up(pk + 48 * 4, 1024);
  p->pk_bytes = pk;

  // ---- accounting via dry runs of the schedules
  memset(p->launches, 0, sizeof(p->launches));
  memset(p->launches_bwd, 0, sizeof(p->launches_bwd));
  for (double& f : p->flops) f = 0.0;
  for (double& f : p->flops_bwd) f = 0.0;
  Ctx c{};
  c.p = p; c.dry = true;
  c.cnt_launches = p->launches; c.cnt_flops = p->flops;
  schedule(c);
  if (p->train) {
    Ctx b{};
    b.p = p; b.dry = true;
    b.cnt_launches = p->launches_bwd; b.cnt_flops = p->flops_bwd;
    schedule_backward(b);
  }
  *out = p;
  return 0;
}

void pvsr_plan_destroy(pvsr_plan* p) {
  if (!p) return;
  for (auto& g : p->graphs) cudaGraphExecDestroy(g.second);
  if (p->cap_stream) cudaStreamDestroy(p->cap_stream);
  if (p->side_stream) cudaStreamDestroy(p->side_stream);
  if (p->cap_side) cudaStreamDestroy(p->cap_side);
  for (cudaEvent_t e : p->ev_pool) cudaEventDestroy(e);
  delete p;
}

int pvsr_plan_set_sign_gradient(pvsr_plan* p, const float* scales, int n_lists) {
  if (!p) return set_error(-2, "null argument");
  for (float& v : p->sign_scale) v = 0.f;
  if (scales) {
    if (n_lists != p->n_lists || n_lists > 3 * kMaxStages) return set_error(-2, "sign-gradient scales: expected %d lists", p->n_lists);
    for (int k = 0; k < n_lists; ++k) p->sign_scale[k] = scales[k];
  }
  return 0;
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
int64_t pvsr_plan_num_launches_bwd(const pvsr_plan* p) {
  long long n = 0;
  for (long long v : p->launches_bwd) n += v;
  return n;
}
double pvsr_plan_flops_bwd(const pvsr_plan* p) {
  double f = 0;
  for (double v : p->flops_bwd) f += v;
  return f;
}
int pvsr_plan_class_stats(const pvsr_plan* p, int64_t* launches, double* flops) {
  for (int i = 0; i < kNumClasses; ++i) { launches[i] = p->launches[i]; flops[i] = p->flops[i]; }
  return kNumClasses;
}
int pvsr_plan_class_stats_bwd(const pvsr_plan* p, int64_t* launches, double* flops) {
  for (int i = 0; i < kNumClassesBwd; ++i) { launches[i] = p->launches_bwd[i]; flops[i] = p->flops_bwd[i]; }
  return kNumClassesBwd;
}

// Composite-forward tables of the head's tail (x4 / x8 heads): functions of the last two convs' parameters only.
static bool plan_has_tail(const pvsr_plan* p) { return p->n_ps >= 2 && p->ps_r[p->n_ps - 1] == 2; }
static int pack_tail_tables(pvsr_plan* p, const pvsr_net_params* P, uint8_t* pk, cudaStream_t s) {
  if (!plan_has_tail(p)) return 0;
  const int last = p->n_ps - 1;
  if (!P->head_w[last] || !P->head_b[last] || !P->head_w[p->n_ps] || !P->head_b[p->n_ps])
    return set_error(-4, "missing head parameter pointer");
  int e = launch_tail36_weights(P->head_w[last], P->head_b[last], P->head_w[p->n_ps], pk + p->pk_t36_w,
                                reinterpret_cast<float*>(pk + p->pk_t36_b), s);
  if (e) return check_cuda(e, "tail36 weights");
  return check_cuda(launch_tail_fwd_tables(P->head_w[last], P->head_b[last], P->head_w[p->n_ps], P->head_b[p->n_ps],
                                           pk + p->pk_tail_fwd, s), "tail forward tables");
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
  if (p->table_ok && get_pack_table()) {
    // one launch for all operands (forward packs, bias gathers, transposed data-gradient packs)
    std::vector<pvsr_table_job> tab(p->jobs.size());
    long long max_n = 0;
    for (size_t i = 0; i < p->jobs.size(); ++i) {
      const auto& j = p->jobs[i];
      const float* src = job_weight(j.kind, j.a, j.b, P, j.is_bias);
      if (!src) return set_error(-4, "missing parameter pointer (kind %d)", j.kind);
      tab[i].src = src; tab[i].idx = idx + j.idx; tab[i].dst = pk + j.dst; tab[i].n = j.n; tab[i].scale = 1.f;
      tab[i].kind = j.is_bias ? PVSR_TJ_GATHER : PVSR_TJ_PACK;
      max_n = j.n > max_n ? j.n : max_n;
    }
    const void* key = hash_bytes(tab.data(), tab.size() * sizeof(pvsr_table_job));
    if (p->table_key != key) {
      int e = cudaMemcpyAsync(pk + p->pk_table, tab.data(), tab.size() * sizeof(pvsr_table_job), cudaMemcpyHostToDevice, s);
      if (!e) e = cudaStreamSynchronize(s);      // `tab` dies at return; happens once per set of parameter pointers
      if (e) return check_cuda(e, "pack table upload");
      p->table_key = key;
    }
    int e = launch_table(reinterpret_cast<const pvsr_table_job*>(pk + p->pk_table), static_cast<int>(tab.size()), max_n, s);
    if (e) return check_cuda(e, "pack table launch");
    return pack_tail_tables(p, P, pk, s);
  }
  for (const auto& j : p->jobs) {
    const float* src = job_weight(j.kind, j.a, j.b, P, j.is_bias);
    if (!src) return set_error(-4, "missing parameter pointer (kind %d)", j.kind);
    int e;
    if (j.is_bias) e = launch_gather_f32(src, idx + j.idx, reinterpret_cast<float*>(pk + j.dst), j.n, s);
    else e = launch_pack_weights(src, idx + j.idx, j.has2 ? idx + j.idx2 : nullptr, pk + j.dst, j.n, s);
    if (e) return check_cuda(e, "pack launch");
  }
  return pack_tail_tables(p, P, pk, s);
}

static int ensure_device(pvsr_plan* p) {
  if (p->num_sms == 0) {
    int rc = pvsr_device_check();
    if (rc) return rc;
    p->num_sms = device_num_sms();
  }
  return 0;
}

static Ctx make_ctx(pvsr_plan* p, const pvsr_net_params* P, const void* packed, const float* lr, const float* pos,
                    void* ws, cudaStream_t s, cudaStream_t side, std::vector<cudaEvent_t>* ev, std::vector<int>* ev_cls) {
  Ctx c{};
  c.p = p; c.dry = false; c.P = P;
  c.ws = static_cast<uint8_t*>(ws);
  c.pk = static_cast<const uint8_t*>(packed);
  c.lr = lr; c.pos = pos; c.stream = s;
  c.main = s; c.side = ev ? nullptr : side;     // per-launch timing runs are single-stream
  c.events = ev; c.event_cls = ev_cls;
  return c;
}

static int forward_eager(pvsr_plan* p, const pvsr_net_params* P, const void* packed, const float* lr,
                         const float* pos, float* out, void* ws, cudaStream_t s, cudaStream_t side,
                         std::vector<cudaEvent_t>* ev, std::vector<int>* ev_cls) {
  Ctx c = make_ctx(p, P, packed, lr, pos, ws, s, side, ev, ev_cls);
  c.out = out;
  c.cnt_launches = p->launches; c.cnt_flops = p->flops;
  schedule(c);
  return c.rc;
}

static int backward_eager(pvsr_plan* p, const pvsr_net_params* P, const void* packed, const float* lr,
                          const float* pos, const float* dout, const pvsr_net_grads* G, void* ws, cudaStream_t s,
                          cudaStream_t side, std::vector<cudaEvent_t>* ev, std::vector<int>* ev_cls) {
  Ctx c = make_ctx(p, P, packed, lr, pos, ws, s, side, ev, ev_cls);
  c.dout = dout; c.G = G;
  c.cnt_launches = p->launches_bwd; c.cnt_flops = p->flops_bwd;
  c.trace_bwd = 1;
  schedule_backward(c);
  return c.rc;
}

int pvsr_plan_forward(pvsr_plan* p, const pvsr_net_params* P, const void* packed, const float* lr, const float* pos,
                      float* out, void* ws, int use_graph, void* stream) {
  if (!p || !P || !packed || !lr || !out || !ws) return set_error(-2, "null argument");
  if (p->cfg.pos_enc && !pos) return set_error(-2, "pos codes required");
  int rc = ensure_device(p);
  if (rc) return rc;
  rc = build_maps(p, ws, packed);
  if (rc) return rc;
  GraphKey key{{ws, packed, lr, pos, out, hash_bytes(P, sizeof(*P)), nullptr, nullptr}};
  return run_or_replay(p, key, use_graph, static_cast<cudaStream_t>(stream), [&](cudaStream_t s, cudaStream_t side) {
    return forward_eager(p, P, packed, lr, pos, out, ws, s, side, nullptr, nullptr);
  });
}

// Scatter table of the backward pass: packed fp32 gradient regions -> parameter-layout gradient buffers.
static int upload_scatter_table(pvsr_plan* p, const void* packed, const pvsr_net_grads* G, void* ws, cudaStream_t s) {
  if (!(p->table_ok && get_pack_table())) return 0;
  const uint8_t* pk = static_cast<const uint8_t*>(packed);
  const int32_t* idx = reinterpret_cast<const int32_t*>(pk + p->pk_idx);
  const float* wg = reinterpret_cast<const float*>(static_cast<const uint8_t*>(ws) + p->off_wg);
  std::vector<pvsr_table_job> tab;
  for (const auto& j : p->sc_jobs) {
    float* dst = job_grad(j.kind, j.a, j.b, G, j.is_bias);
    if (!dst) continue;
    pvsr_table_job t{};
    t.src = wg + j.src; t.idx = idx + j.idx; t.dst = dst; t.n = j.n; t.scale = 1.f; t.kind = PVSR_TJ_SCATTER;
    tab.push_back(t);
  }
  if (tab.empty()) return 0;
  const void* key = hash_bytes(tab.data(), tab.size() * sizeof(pvsr_table_job));
  if (p->sc_table_key == key) return 0;
  int e = cudaMemcpyAsync(const_cast<uint8_t*>(pk) + p->pk_sc_table, tab.data(), tab.size() * sizeof(pvsr_table_job),
                          cudaMemcpyHostToDevice, s);
  if (!e) e = cudaStreamSynchronize(s);
  if (e) return check_cuda(e, "scatter table upload");
  p->sc_table_key = key;
  return 0;
}

static int upload_jobs(pvsr_plan* p, void* ws, cudaStream_t s) {
  if (p->jobs_uploaded_for == ws || p->wg_jobs.empty()) return 0;
  // synchronous w.r.t. the host vector (owned by the plan, so it outlives the copy anyway)
  int e = cudaMemcpyAsync(static_cast<uint8_t*>(ws) + p->off_jobs, p->wg_jobs.data(), p->wg_jobs.size() * sizeof(WgJob),
                          cudaMemcpyHostToDevice, s);
  if (e) return check_cuda(e, "wgrad job upload");
  p->jobs_uploaded_for = ws;
  return 0;
}

int pvsr_plan_backward(pvsr_plan* p, const pvsr_net_params* P, const void* packed, const float* lr, const float* pos,
                       const float* dout, const pvsr_net_grads* G, void* ws, int use_graph, void* stream) {
  if (!p || !P || !packed || !lr || !dout || !G || !ws) return set_error(-2, "null argument");
  if (!p->train) return set_error(-3, "plan was created without save_for_backward");
  if (p->cfg.pos_enc && !pos) return set_error(-2, "pos codes required");
  if (p->idx_uploaded_for != packed) return set_error(-4, "pvsr_plan_pack must run before pvsr_plan_backward");
  int rc = ensure_device(p);
  if (rc) return rc;
  rc = build_maps(p, ws, packed);
  if (rc) return rc;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  rc = upload_jobs(p, ws, s);
  if (rc) return rc;
  rc = upload_scatter_table(p, packed, G, ws, s);
  if (rc) return rc;
  GraphKey key{{ws, packed, lr, pos, dout, hash_bytes(P, sizeof(*P)), hash_bytes(G, sizeof(*G)),
                hash_bytes(p->sign_scale, sizeof(p->sign_scale))}};
  return run_or_replay(p, key, use_graph, s, [&](cudaStream_t st, cudaStream_t side) {
    return backward_eager(p, P, packed, lr, pos, dout, G, ws, st, side, nullptr, nullptr);
  });
}

static int finish_profile(cudaStream_t s, std::vector<cudaEvent_t>& ev, std::vector<int>& cls, int n_cls,
                          double* ms_by_class, int rc) {
  int e = cudaStreamSynchronize(s);
  for (int i = 0; i < n_cls; ++i) ms_by_class[i] = 0.0;
  for (size_t i = 0; i < cls.size(); ++i) {
    float ms = 0.f;
    cudaEventElapsedTime(&ms, ev[2 * i], ev[2 * i + 1]);
    ms_by_class[cls[i]] += ms;
  }
  for (cudaEvent_t x : ev) cudaEventDestroy(x);
  if (rc) return rc;
  return check_cuda(e, "profile sync");
}

// One eager forward with CUDA events around every launch; returns summed milliseconds per launch class.
// Synchronises the stream (profiling aid for bench.py's roofline figures, not a hot path).
int pvsr_plan_profile(pvsr_plan* p, const pvsr_net_params* P, const void* packed, const float* lr, const float* pos,
                      float* out, void* ws, double* ms_by_class, void* stream) {
  int rc = ensure_device(p);
  if (rc) return rc;
  rc = build_maps(p, ws, packed);
  if (rc) return rc;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  std::vector<cudaEvent_t> ev;
  std::vector<int> cls;
  rc = forward_eager(p, P, packed, lr, pos, out, ws, s, nullptr, &ev, &cls);
  return finish_profile(s, ev, cls, kNumClasses, ms_by_class, rc);
}

int pvsr_plan_profile_bwd(pvsr_plan* p, const pvsr_net_params* P, const void* packed, const float* lr,
                          const float* pos, const float* dout, const pvsr_net_grads* G, void* ws, double* ms_by_class,
                          void* stream) {
  if (!p->train) return set_error(-3, "plan was created without save_for_backward");
  int rc = ensure_device(p);
  if (rc) return rc;
  rc = build_maps(p, ws, packed);
  if (rc) return rc;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  rc = upload_jobs(p, ws, s);
  if (rc) return rc;
  rc = upload_scatter_table(p, packed, G, ws, s);
  if (rc) return rc;
  std::vector<cudaEvent_t> ev;
  std::vector<int> cls;
  rc = backward_eager(p, P, packed, lr, pos, dout, G, ws, s, nullptr, &ev, &cls);
  return finish_profile(s, ev, cls, kNumClassesBwd, ms_by_class, rc);
}

}  // extern "C"
