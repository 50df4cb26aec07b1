// C ABI of libstat_b200.so (include/stat_b200.h): parameter packing, workspace
// layout and the per-step launch sequences built from the kernels in
// gemm_tf32x3.cu / att_step.cu / recurrent.cu.  Everything is enqueued on the
// caller's stream; nothing here allocates, synchronises or touches the host side
// of a tensor.
#include <stdlib.h>
#include <string.h>

#include "kernels.cuh"
#include "stat_common.cuh"

#include <vector>

namespace stat {
const char *get_error();
unsigned long long launch_count();
int gemm_set_trap_log(int *dev_ptr);
int att_group_set_trap_log(int *dev_ptr);

namespace {

// ---------------------------------------------------------------------------
// optional in-situ phase timing (stat_profile_*): CUDA events recorded on the
// caller's stream around each launch group.  Off by default; never used while a
// stream is being captured.
// ---------------------------------------------------------------------------
enum Phase { PH_INIT = 0, PH_K0_GLOBAL, PH_K0_MOTION, PH_K0_LOCAL, PH_K0_PROJ, PH_HPROJ, PH_ATT, PH_CTXPROJ,
             PH_GATES, PH_READOUT, PH_LOGITS, PH_PICK, PH_FUSED_B, PH_FUSED_C, PH_FUSED_LOGITS, PH_COMBINE, PH_COUNT };
const char *const kPhaseNames[PH_COUNT] = {"init_state", "k0_ff_global", "k0_ff_motion", "k0_ff_local",
                                           "k0_ctx_proj", "step_h_proj", "step_attention", "step_ctx_proj",
                                           "step_gates", "step_readout", "step_logits", "step_pick",
                                           "step_gates_fused", "step_queries_readout_fused", "step_logits_fused",
                                           "step_pick_combine"};
struct ProfRec { int phase; cudaEvent_t a, b; };
bool g_prof_on = false;
std::vector<ProfRec> g_prof;
std::vector<cudaEvent_t> g_event_pool;   // recycled, so the hot path never creates events

cudaEvent_t prof_event() {
  if (g_event_pool.empty()) {
    cudaEvent_t e;
    cudaEventCreate(&e);
    return e;
  }
  cudaEvent_t e = g_event_pool.back();
  g_event_pool.pop_back();
  return e;
}

struct ProfScope {
  cudaStream_t st;
  bool on;
  ProfRec r;
  ProfScope(int phase, cudaStream_t s) : st(s), on(g_prof_on) {
    if (!on) return;
    r.phase = phase;
    r.a = prof_event();
    r.b = prof_event();
    cudaEventRecord(r.a, st);
  }
  ~ProfScope() {
    if (!on) return;
    cudaEventRecord(r.b, st);
    g_prof.push_back(r);
  }
};

inline size_t up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// ---------------------------------------------------------------------------
// prepared parameters: K-major ("transposed") concatenations, float offsets
// ---------------------------------------------------------------------------
struct Prep {
  // weights, each (features, K) row-major
  size_t WaT;    // (E+8H+1, H): ff_logit_lstm_W (E) | Wdl | Wdg | Wdm | Wdlt | U (4H) | W_sel   K = H
                 //   = everything that multiplies h: rows [0,E) = WlT, rows [E, E+8H+1) = WhT
  size_t ba;     // (E+8H+1) : 0 (E) | 0 | 0 | 0 | blt | 0 | b_sel
  size_t WhT, bh, WlT;   // views into WaT / ba
  size_t WcT;    // (4H+E, H): Wc (4H) | ff_logit_ctxglm_W (E)               K = H
  size_t bz;     // (E)      : ff_logit_lstm_b (+ ff_logit_ctxglm_b)
  size_t WvT;    // (V, E)   : ff_logit_W                                    K = E
  size_t bv;     // (V)
  size_t WstT;   // (2H, Dg) : ff_state_W | ff_memory_W                      K = Dg
  size_t bst;    // (2H)
  size_t WgT;    // (H, Dg)  : ff_global_W (global_proj)                     K = Dg
  size_t bg0;    // (H)
  size_t WlocT;  // (H, Dr)                                                  K = Dr
  size_t bloc;   // (H)
  size_t WmotT;  // (H, Dm)                                                  K = Dm
  size_t bmot;   // (H)
  size_t WcgT;   // (H, H)   : Wcg_att
  size_t bcg;    // (H)
  size_t WcmT;   // (H, H)   : Wcm_att
  size_t bcm;    // (H)
  size_t WclT;   // (2H, H)  : Wcl_att | Wclt_att
  size_t bcl;    // (2H)     : bl_att | 0
  size_t WdT;    // (4H, E)  : decoder_W                                     K = E
  size_t U4;     // Ul | Ug | Um | Ult (4 x H) then cl, cg, cm, clt (4 scalars)
  size_t EW;     // (V+1, 4H): token -> emb.W + b ; row V = b
  size_t Wemb;   // (V, E) copy (prev2out add)
  // fused step (step_fused.cu)
  size_t WcI;    // (4H+E, H): rows 4u+g = Wc[:, gH+u] (gate-interleaved: the four gates of a unit adjacent);
                 //            rows 4H.. = ff_logit_ctxglm_W[:, e]                          -- multiplies ctx
  size_t WqT;    // (8H+4+E, H): Wdl | Wdg | Wdm | Wdlt | W_sel | 3 zero rows | U (gate-interleaved rows) |
                 //              ff_logit_lstm_W                                             -- multiplies h
  size_t bq;     // (8H+4)    : 0 | 0 | 0 | blt | b_sel | 0 ...
  size_t bdi;    // (4H)      : decoder_b, gate-interleaved
  // cell step (cell_step.cu): per-CTA [k][column] slabs cut from WcI / WqT
  size_t W1s, W2s;
  size_t total;  // floats
  int NH, NC;
};

Prep prep_layout(const StatDims &d) {
  Prep p;
  const size_t H = d.H, E = d.E, V = d.V, Dg = d.Dg, Dr = d.Dr, Dm = d.Dm;
  size_t o = 0;
  auto take = [&](size_t n) {
    size_t at = o;
    o = up(o + n, 64);  // 256-byte aligned regions
    return at;
  };
  p.NH = static_cast<int>(8 * H + 1);
  p.NC = static_cast<int>(4 * H + ((d.flags & STAT_CTX2OUT) ? E : 0));
  p.WaT = take((E + 8 * H + 1) * H);
  p.ba = take(E + 8 * H + 1);
  p.WlT = p.WaT;
  p.WhT = p.WaT + E * H;
  p.bh = p.ba + E;
  p.WcT = take((4 * H + E) * H);
  p.bz = take(E);
  p.WvT = take(V * E);
  p.bv = take(V);
  p.WstT = take(2 * H * Dg);
  p.bst = take(2 * H);
  p.WgT = take(H * Dg);
  p.bg0 = take(H);
  p.WlocT = take(H * Dr);
  p.bloc = take(H);
  p.WmotT = take(H * Dm);
  p.bmot = take(H);
  p.WcgT = take(H * H);
  p.bcg = take(H);
  p.WcmT = take(H * H);
  p.bcm = take(H);
  p.WclT = take(2 * H * H);
  p.bcl = take(2 * H);
  p.WdT = take(4 * H * E);
  p.U4 = take(4 * H + 4);
  p.EW = take((V + 1) * 4 * H);
  p.Wemb = take(V * E);
  p.WcI = take((4 * H + E) * H);
  p.WqT = take((8 * H + 4 + E) * H);
  p.bq = take(8 * H + 4);
  p.bdi = take(4 * H);
  {
    CellPlan cp;
    const bool ok = cell_plan(d.H, d.E, &cp);
    p.W1s = take(ok ? cp.w1_floats : 0);
    p.W2s = take(ok ? cp.w2_floats : 0);
  }
  p.total = o;
  return p;
}

// ---------------------------------------------------------------------------
// workspace
// ---------------------------------------------------------------------------
struct Ws {
  size_t ctxg0, pctxg, ctxm0, pctxm, ctxl0, pctxl, qctxl;
  size_t gbar, h0c0;
  size_t h, c, hpz, ctx, pre_c, zadd, hd, zpre, z, logits;
  int ksa, ksc;                      // k-slices (= output planes) of the h- and ctx-projections
  size_t hpz_plane, pc_plane;        // floats between planes
  size_t rec_vec, rec_ms, att_scores, alpha_l;
  size_t counters, tok_prev, alive;  // byte-typed regions (still float offsets)
  size_t hb, cb, cand_cost, cand_word, hist, hist_len, bscore, src_row, row_clip, dead_k, bdone;   // beam search
  size_t lpc;                        // Kahan compensation of the teacher-forced log-prob accumulation
  size_t hq, part, tgt;              // fused step: [queries | selector logit | h.U] rows, vocabulary partials
  int ldq2, npart;
  size_t ctxT, hT, hq2, cell_bar;    // cell step: transposed ctx / h ([chunk][H][64]), second hq (beam), barrier words
  size_t total;                      // floats
  int ldhp, ldpc, ldl, S, Tc;
  int att_impl;      // 2 = att_group_kernel (bulk-copy streaming), 0 = the generic att_step_kernel
};

// most rows of a product that dense() issues as a 128-row and a narrow skinny launch (beam search: 32 clips x 5 slots)
constexpr int SPLIT_ROWS_MAX = 192;

// k-slices for a skinny (rows <= 128, or 128 + a few: see dense()) projection: enough CTAs to cover the SMs once
int pick_ksplit(int nfeat, int rows, int K) {
  if (rows > SPLIT_ROWS_MAX) return 1;
  if (rows > 128) rows = 128;
  const int bq = rows > 64 ? 128 : (rows > 32 ? 64 : 32);
  const int tiles = ((nfeat + 127) / 128) * ((rows + bq - 1) / bq);
  const int nk = (K + 31) / 32;
  int ks = 148 / tiles;
  if (ks > 4) ks = 4;      // the consumers sum at most 4 planes in one batch of loads
  if (ks > nk) ks = nk;
  if (ks < 1) ks = 1;
  return ks;
}

void pick_segments(const StatDims &d, int rows, int *S, int *Tc) {
  // enough CTAs to cover the 148 SMs twice, frames split as evenly as possible
  int s = (2 * 148 + rows - 1) / rows;
  if (s < 1) s = 1;
  if (s > d.T) s = d.T;
  int tc = (d.T + s - 1) / s;
  const int tc_max = static_cast<int>((160 * 1024) / (static_cast<size_t>(d.H) * 4));
  if (tc > tc_max) tc = tc_max;
  if (tc < 1) tc = 1;
  s = (d.T + tc - 1) / tc;
  *S = s;
  *Tc = tc;
}

Ws ws_layout(const StatDims &d, int rows) {
  Ws w;
  const size_t B = d.B, T = d.T, R = d.R, H = d.H, E = d.E, V = d.V;
  const size_t n = rows;
  size_t o = 0;
  auto take = [&](size_t k) {
    size_t at = o;
    o = up(o + k, 64);
    return at;
  };
  int nchunks = 0, nparts = 0, nstages = 0;
  // (the streaming kernel reads the h-projection rows as float4: their offset E must be 16-byte aligned)
  w.att_impl = 0;
  if (d.E % 4 == 0) {
    const char *impl = getenv("STAT_ATT_IMPL");      // "generic": force the plain kernel (cross-checks)
    if (!(impl && !strcmp(impl, "generic")) && att_group_plan(rows, d.T, d.R, d.H, &nchunks, &nstages, &nparts))
      w.att_impl = 2;
  }
  if (w.att_impl) {
    w.S = nparts;
    w.Tc = 0;
  } else {
    pick_segments(d, rows, &w.S, &w.Tc);
  }
  w.ldhp = static_cast<int>(up(E + 8 * H + 1, 4));
  w.ksa = pick_ksplit(static_cast<int>(E + 8 * H + 1), rows, d.H);
  w.ksc = pick_ksplit(static_cast<int>(4 * H + E), rows, d.H);
  w.hpz_plane = up(n * w.ldhp, 64);
  w.pc_plane = up(n * (4 * H + E), 64);
  w.ldpc = static_cast<int>(4 * H + E);
  w.ldl = static_cast<int>(up(V, 4));
  w.ctxg0 = take(B * T * H);
  w.pctxg = take(B * T * H);
  w.ctxm0 = take(B * T * H);
  w.pctxm = take(B * T * H);
  w.ctxl0 = take(B * T * R * H);
  w.pctxl = take(B * T * R * H);
  w.qctxl = take(B * T * R * H);
  w.gbar = take(B * d.Dg);
  w.h0c0 = take(B * 2 * H);
  w.h = take(n * H);
  w.c = take(n * H);
  w.hpz = take(w.hpz_plane * w.ksa);
  w.ctx = take(n * H);
  w.pre_c = take(w.pc_plane * w.ksc);
  w.zadd = take(n * E);
  w.hd = take(n * H);
  w.zpre = take(n * E);
  w.z = take(n * E);
  w.logits = take(n * w.ldl);
  w.rec_vec = take(n * w.S * 3 * H);
  w.rec_ms = take(n * w.S * 6);
  w.att_scores = take(3 * n * T);
  w.alpha_l = take(n * T * R);
  w.counters = take(n);
  w.tok_prev = take(2 * n);
  w.alive = take(n);
  w.hb = take(n * H);
  w.cb = take(n * H);
  w.cand_cost = take(n * BEAM_KMAX);
  w.cand_word = take(n * BEAM_KMAX);
  w.hist = take(n * BEAM_LMAX);
  w.hist_len = take(n);
  w.bscore = take(n);
  w.src_row = take(n);
  w.row_clip = take(n);
  w.dead_k = take(n);
  w.bdone = take(n);
  w.ldq2 = static_cast<int>(8 * H + 4);
  w.npart = 2 * static_cast<int>((V + 127) / 128);
  w.hq = take(n * w.ldq2);
  w.part = take(n * w.npart * 4);
  w.tgt = take(n);
  w.lpc = take(n);
  {
    const size_t chunks = (n + 63) / 64;
    w.ctxT = take(chunks * H * 64);
    w.hT = take(chunks * H * 64);
    w.hq2 = take(n * w.ldq2);
    w.cell_bar = take(64);
  }
  w.total = o;
  return w;
}

int check_dims(const StatDims *d) {
  STAT_REQUIRE(d != nullptr, STAT_EINVAL, "dims is NULL");
  STAT_REQUIRE(d->B >= 1 && d->T >= 1 && d->R >= 1 && d->R <= 16, STAT_EINVAL,
               "dims: need B>=1, T>=1, 1<=R<=16 (B=%d T=%d R=%d)", d->B, d->T, d->R);
  STAT_REQUIRE(d->H >= 1 && d->H <= 1024 && d->E >= 1 && d->V >= 2, STAT_EINVAL,
               "dims: need 1<=H<=1024, E>=1, V>=2 (H=%d E=%d V=%d)", d->H, d->E, d->V);
  STAT_REQUIRE(d->Dg >= 1 && d->Dm >= 1 && d->Dr >= 1, STAT_EINVAL, "dims: feature widths must be >= 1");
  STAT_REQUIRE((d->flags & STAT_GLOBAL_PROJ) || d->Dg == d->H, STAT_EINVAL,
               "dims: the reference graph needs ctxg_dim == dim (Dg=%d H=%d) unless STAT_GLOBAL_PROJ is set",
               d->Dg, d->H);
  return STAT_OK;
}

int check_device() {
  int dev = -1;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) {
    set_error("no CUDA device: %s (libstat_b200 has no CPU path)", cudaGetErrorString(e));
    return STAT_ECUDA;
  }
  return STAT_OK;
}

// out (rows, ldc)[r][f] = post*act(alpha * x[r,:].Wt[f,:] + bias[f] + addend[r][f]); skinny activations
// ride the tensor-core column axis ("swap"), wide ones the 128-lane axis.
int dense(const float *x, int ldx, int rows, const float *Wt, int K, int nfeat, const float *bias, float *out,
          int ldc, int act, float alpha, float post, const float *addend, int ld_add, cudaStream_t st,
          int ksplit = 1, size_t plane = 0) {
  if (rows > 128 && rows <= SPLIT_ROWS_MAX) {
    // A few rows more than one 128-row tile (beam search: 32 clips x k = 5 = 160 rows): with the rows on the 128-lane
    // axis the second row tile would be mostly padding and the tile count (2 x 99 for the vocabulary) a second,
    // one-third-full wave.  Two skinny launches instead: 128 rows, then the rest with a narrow tile (N = 32 / 64 per
    // MMA instruction: about half the cost), which run side by side (programmatic dependent launch).  Both keep the
    // k-split of the skinny form (same planes, `plane` floats apart).
    STAT_TRY(dense(x, ldx, 128, Wt, K, nfeat, bias, out, ldc, act, alpha, post, addend, ld_add, st, ksplit, plane));
    return dense(x + static_cast<size_t>(128) * ldx, ldx, rows - 128, Wt, K, nfeat, bias,
                 out + static_cast<size_t>(128) * ldc, ldc, act, alpha, post,
                 addend ? addend + static_cast<size_t>(128) * ld_add : nullptr, ld_add, st, ksplit, plane);
  }
  GemmArgs g;
  memset(&g, 0, sizeof(g));
  g.K = K;
  g.nseg = 1;
  g.ksplit = ksplit;
  g.plane = plane;
  g.seg[0] = GemmSeg{out, ldc, bias, addend, ld_add, alpha, post, act, 0, nfeat};
  if (rows <= 128) {
    g.P = Wt; g.ldp = K; g.NP = nfeat;
    g.Q = x; g.ldq = ldx; g.NQ = rows;
    g.feat_on_p = 1;
  } else {
    g.P = x; g.ldp = ldx; g.NP = rows;
    g.Q = Wt; g.ldq = K; g.NQ = nfeat;
    g.feat_on_p = 0;
  }
  return gemm_launch(g, st);
}

// Side streams and events of the internal fork/join points: one set per host thread and device, created
// on first use and never destroyed (they may be part of captured graphs).
constexpr int MAX_DEVICES = 64;
struct SideSet {
  cudaStream_t k0[2] = {nullptr, nullptr};      // K0 chains
  cudaStream_t readout = nullptr;               // readout chain of a decode step
  cudaEvent_t fork = nullptr, join[2] = {nullptr, nullptr}, state = nullptr, done = nullptr;
  bool ready = false;
};
int side_set(SideSet **out) {
  thread_local SideSet tab[MAX_DEVICES];
  int dev = 0;
  STAT_CUDA_CHECK(cudaGetDevice(&dev));
  STAT_REQUIRE(dev >= 0 && dev < MAX_DEVICES, STAT_EINVAL, "device ordinal %d out of range", dev);
  SideSet &t = tab[dev];
  if (!t.ready) {
    for (int i = 0; i < 2; ++i) {
      STAT_CUDA_CHECK(cudaStreamCreateWithFlags(&t.k0[i], cudaStreamNonBlocking));
      STAT_CUDA_CHECK(cudaEventCreateWithFlags(&t.join[i], cudaEventDisableTiming));
    }
    {
      // STAT_SIDE_PRIO=1: the readout chain gets the highest stream priority (its blocks are placed before
      // those of the next attention when both are ready)
      const char *e = getenv("STAT_SIDE_PRIO");
      int lo = 0, hi = 0;
      STAT_CUDA_CHECK(cudaDeviceGetStreamPriorityRange(&lo, &hi));
      STAT_CUDA_CHECK(cudaStreamCreateWithPriority(&t.readout, cudaStreamNonBlocking, (e && e[0] == '1') ? hi : 0));
    }
    STAT_CUDA_CHECK(cudaEventCreateWithFlags(&t.fork, cudaEventDisableTiming));
    STAT_CUDA_CHECK(cudaEventCreateWithFlags(&t.state, cudaEventDisableTiming));
    STAT_CUDA_CHECK(cudaEventCreateWithFlags(&t.done, cudaEventDisableTiming));
    t.ready = true;
  }
  *out = &t;
  return STAT_OK;
}

// Fork / join of two side streams around the caller's stream (event based, so it also works while
// the caller's stream is being captured into a CUDA graph).
struct Fork {
  cudaStream_t main_ = nullptr;
  bool active_ = false;
  SideSet *s_ = nullptr;
  int open(cudaStream_t st, bool enable) {
    main_ = st;
    active_ = false;
    if (!enable) return STAT_OK;
    STAT_TRY(side_set(&s_));
    STAT_CUDA_CHECK(cudaEventRecord(s_->fork, st));
    for (int i = 0; i < 2; ++i) STAT_CUDA_CHECK(cudaStreamWaitEvent(s_->k0[i], s_->fork, 0));
    active_ = true;
    return STAT_OK;
  }
  cudaStream_t side(int i) const { return active_ ? s_->k0[i] : main_; }
  int join() {
    if (!active_) return STAT_OK;
    for (int i = 0; i < 2; ++i) {
      STAT_CUDA_CHECK(cudaEventRecord(s_->join[i], s_->k0[i]));
      STAT_CUDA_CHECK(cudaStreamWaitEvent(main_, s_->join[i], 0));
    }
    active_ = false;
    return STAT_OK;
  }
};

// P2, P3: mean-pooled global feature -> h0 | c0   (:618,649,657-660)
int init_state(const StatDims &d, const Prep &p, const float *P, const Ws &w, float *W, const float *ctxg,
               const float *mask_ctxg, cudaStream_t st) {
  ProfScope ps(PH_INIT, st);
  STAT_TRY(meanpool_launch(ctxg, mask_ctxg, W + w.gbar, d.B, d.T, d.Dg, st));
  return dense(W + w.gbar, d.Dg, d.B, P + p.WstT, d.Dg, 2 * d.H, P + p.bst, W + w.h0c0, 2 * d.H, 1, 1.f, 1.f,
               nullptr, 0, st);
}

AttArgs att_args(const StatDims &d, const Prep &p, const float *P, const Ws &w, float *W, int rows,
                 const int32_t *row_clip, float *att_scores, float *alpha_l) {
  const int H = d.H;
  AttArgs a;
  memset(&a, 0, sizeof(a));
  a.pctxl = W + w.pctxl; a.ctxl0 = W + w.ctxl0; a.qctxl = W + w.qctxl;
  a.pctxg = W + w.pctxg; a.ctxg0 = W + w.ctxg0; a.pctxm = W + w.pctxm; a.ctxm0 = W + w.ctxm0;
  const int E = d.E;
  a.hp = W + w.hpz; a.ldhp = w.ldhp; a.hp_parts = w.ksa; a.hp_plane = w.hpz_plane;
  a.off_sl = E; a.off_sg = E + H; a.off_sm = E + 2 * H; a.off_slt = E + 3 * H; a.off_sel = E + 8 * H;
  a.Ul = P + p.U4; a.Ug = P + p.U4 + H; a.Um = P + p.U4 + 2 * H; a.Ult = P + p.U4 + 3 * H;
  a.cl = P + p.U4 + 4 * H; a.cg = a.cl + 1; a.cm = a.cl + 2; a.clt = a.cl + 3;
  a.row_clip = row_clip;
  a.rows = rows; a.T = d.T; a.R = d.R; a.H = H;
  a.S = w.S; a.Tc = w.Tc;
  a.selector = (d.flags & STAT_SELECTOR) ? 1 : 0;
  a.ctx = W + w.ctx;
  a.rec_vec = W + w.rec_vec; a.rec_ms = W + w.rec_ms;
  a.counters = reinterpret_cast<unsigned int *>(W + w.counters);
  a.att_scores = att_scores;
  a.alpha_l = alpha_l;
  return a;
}

int att_launch(const Ws &w, const AttArgs &a, cudaStream_t st) {
  if (w.att_impl == 2) return att_group_launch(a, st);
  return att_step_launch(a, st);
}

struct StepIO {
  int rows;
  const int32_t *row_clip;
  const int64_t *tok_prev;  // (rows) or null
  const float *mask;        // (rows) or null
  const float *dp_gates, *dp_h, *dp_z;
  const float *h_in, *c_in;
  float *h_out, *c_out;
  float *h_all;
  float *alpha_l;     // (rows,T,R) or null
  float *att_scores;  // (3,rows,T) or null
  int reverse;        // attention walks the frames backwards (odd decode steps: L2 reuse across steps)
  int rows_per_clip;  // beam search: row = clip * rows_per_clip + slot (0: row_clip / identity)
};

// what multiplies a hidden state: which = 1 the attention queries / h.U / selector logit of the NEXT
// cell (:371,389,402,415,433,437), which = 2 the readout term h.ff_logit_lstm_W (:684-688), 3 = both
// in one pass over h.  Results go to the k-slice planes of the hpz region.
int h_proj(const StatDims &d, const Prep &p, const float *P, const Ws &w, float *W, const float *h, int rows,
           int which, cudaStream_t st) {
  const int H = d.H, E = d.E;
  const int nh = (d.flags & STAT_SELECTOR) ? 8 * H + 1 : 8 * H;
  ProfScope ps(PH_HPROJ, st);
  if (which == 3)
    return dense(h, H, rows, P + p.WaT, H, E + nh, P + p.ba, W + w.hpz, w.ldhp, 0, 1.f, 1.f, nullptr, 0, st, w.ksa,
                 w.hpz_plane);
  if (which == 1)
    return dense(h, H, rows, P + p.WhT, H, nh, P + p.bh, W + w.hpz + E, w.ldhp, 0, 1.f, 1.f, nullptr, 0, st, w.ksa,
                 w.hpz_plane);
  return dense(h, H, rows, P + p.WlT, H, E, nullptr, W + w.hpz, w.ldhp, 0, 1.f, 1.f, nullptr, 0, st, w.ksa,
               w.hpz_plane);
}

// One decode step given the h-projections of h_in in hpz, in three launch groups:
//   step_att   S1-S9 and the context projections (ctx.Wc for the gates, ctx.ff_logit_ctxglm_W)
//   step_gates S10-S13, then every product of the new hidden state: the readout term and, when
//              next_hproj, the attention queries / h.U / selector logit of the NEXT step
//   step_out   R1-R3 without the soft-max: readout activation and the vocabulary logits
// step_out(t) neither reads anything step_att(t+1) writes nor writes anything it reads, so the
// decode loops run the two on different streams (Overlap below).
int step_att(const StatDims &d, const Prep &p, const float *P, const Ws &w, float *W, const StepIO &io,
             cudaStream_t st) {
  const int H = d.H, rows = io.rows;
  AttArgs a = att_args(d, p, P, w, W, rows, io.row_clip, io.att_scores, io.alpha_l);
  a.reverse = io.reverse;
  a.rows_per_clip = io.rows_per_clip;
  {
    ProfScope ps(PH_ATT, st);
    STAT_TRY(att_launch(w, a, st));
  }
  // ctx.Wc (gates, :439) and ctx.ff_logit_ctxglm_W (:691-693) in one pass
  ProfScope ps(PH_CTXPROJ, st);
  return dense(W + w.ctx, H, rows, P + p.WcT, H, p.NC, nullptr, W + w.pre_c, w.ldpc, 0, 1.f, 1.f, nullptr, 0, st,
               w.ksc, w.pc_plane);
}

int step_gates(const StatDims &d, const Prep &p, const float *P, const Ws &w, float *W, const StepIO &io,
               bool next_hproj, cudaStream_t st, bool gates_only = false) {
  const int H = d.H, E = d.E, V = d.V, rows = io.rows;
  const bool ctx2out = (d.flags & STAT_CTX2OUT) != 0;
  GateArgs g;
  memset(&g, 0, sizeof(g));
  g.rows = rows; g.H = H; g.E = E; g.V = V;
  g.pre_c = W + w.pre_c; g.ldpc = w.ldpc; g.zc_off = ctx2out ? 4 * H : -1;
  g.pc_parts = w.ksc; g.pc_plane = w.pc_plane;
  g.hp = W + w.hpz; g.ldhp = w.ldhp; g.off_u = E + 4 * H;
  g.hp_parts = w.ksa; g.hp_plane = w.hpz_plane;
  g.EW = P + p.EW; g.Wemb = P + p.Wemb;
  g.tok_prev = io.tok_prev; g.mask = io.mask;
  g.dp_gates = io.dp_gates; g.dp_h = io.dp_h;
  g.h_in = io.h_in; g.c_in = io.c_in; g.h_out = io.h_out; g.c_out = io.c_out;
  g.hd_out = W + w.hd;
  g.bz = P + p.bz; g.zadd = W + w.zadd;
  g.prev2out = (d.flags & STAT_PREV2OUT) ? 1 : 0;
  g.h_all = io.h_all;
  {
    ProfScope ps(PH_GATES, st);
    STAT_TRY(gates_launch(g, st));
  }
  if (gates_only) return STAT_OK;
  if (io.dp_h) {
    // explicit dropout mask on h (use_noise=1): the readout multiplies h*mask, the next cell h itself
    {
      ProfScope ps(PH_READOUT, st);
      STAT_TRY(dense(W + w.hd, H, rows, P + p.WlT, H, E, nullptr, W + w.zpre, E, 0, 1.f, 1.f, nullptr, 0, st));
    }
    if (next_hproj) STAT_TRY(h_proj(d, p, P, w, W, io.h_out, rows, 1, st));
    return STAT_OK;
  }
  return h_proj(d, p, P, w, W, io.h_out, rows, next_hproj ? 3 : 2, st);
}

int step_out(const StatDims &d, const Prep &p, const float *P, const Ws &w, float *W, const StepIO &io,
             cudaStream_t st) {
  const int E = d.E, V = d.V, rows = io.rows;
  // z = tanh(dp_h*h . Wl + b + emb + ctx.Wctx) * dp_z   (:684-696)
  ZactArgs z;
  memset(&z, 0, sizeof(z));
  z.rows = rows; z.E = E;
  z.zadd = W + w.zadd; z.dp_z = io.dp_z; z.z = W + w.z;
  if (io.dp_h) {
    z.zpre = W + w.zpre; z.ldz = E; z.parts = 1; z.plane = 0; z.alpha = 1.0f;
  } else {
    z.zpre = W + w.hpz; z.ldz = w.ldhp; z.parts = w.ksa; z.plane = w.hpz_plane; z.alpha = 0.5f;
  }
  {
    ProfScope ps(PH_READOUT, st);
    STAT_TRY(zact_launch(z, st));
  }
  // logits = z . ff_logit_W + b  (:704-705)
  ProfScope pl(PH_LOGITS, st);
  return dense(W + w.z, E, rows, P + p.WvT, E, V, P + p.bv, W + w.logits, w.ldl, 0, 1.f, 1.f, nullptr, 0, st);
}


// ---------------------------------------------------------------------------
// fused decode step (step_fused.cu): attention -> B (ctx.Wc, gates fused) -> C (everything that multiplies the new
// h: next queries, selector logit, h.U of the next cell, readout activation) on the caller's stream, logits (partial
// vocabulary reduction fused) -> combine beside the next attention.
// ---------------------------------------------------------------------------
// step implementation: 0 = separate kernels (k-split products + gates / readout / vocabulary kernels), 1 = the fused
// tile kernels, 2 = the cell step (cell_step.cu, default).  stat_set_step_impl / STAT_STEP.
int g_step_impl = -1;
bool g_step_impl_set = false;      // stat_set_step_impl was called with an explicit implementation
int step_impl() {
  if (g_step_impl < 0) {
    const char *e = getenv("STAT_STEP");
    const char *f = getenv("STAT_FUSED");
    g_step_impl = e ? atoi(e) : ((f && f[0] == '1') ? 1 : 0);
    if (g_step_impl < 0 || g_step_impl > 2) g_step_impl = 0;
  }
  return g_step_impl;
}
bool fused_enabled(const StatDims &d, int rows) {
  return step_impl() == 1 && fused_supported(d.H, d.E) && rows <= 128;
}

AttArgs att_args_fused(const StatDims &d, const Prep &p, const float *P, const Ws &w, float *W, int rows,
                       const int32_t *row_clip, float *att_scores, float *alpha_l) {
  AttArgs a = att_args(d, p, P, w, W, rows, row_clip, att_scores, alpha_l);
  const int H = d.H;
  a.hp = W + w.hq; a.ldhp = w.ldq2; a.hp_parts = 1; a.hp_plane = 0;
  a.off_sl = 0; a.off_sg = H; a.off_sm = 2 * H; a.off_slt = 3 * H; a.off_sel = 4 * H;
  return a;
}

// B: gate pre-activations ctx_t.Wc + h_{t-1}.U (from C of the previous step) + EW[word], S10-S13 in the epilogue;
// the readout addend bz + ctx.Wctx (+ Wemb[word]) rides in the same launch.  h, c updated in place.
int fstep_gates(const StatDims &d, const Prep &p, const float *P, const Ws &w, float *W, const StepIO &io,
                cudaStream_t st) {
  const int H = d.H, E = d.E;
  FusedPhase f;
  memset(&f, 0, sizeof(f));
  f.swap = 1;
  f.W = P + p.WcI; f.wrows = 4 * H + E; f.wK = H; f.ldw = H;
  f.X[0] = W + w.ctx; f.xK[0] = H; f.ldx[0] = H;
  f.rows = io.rows;
  f.nseg = 2;
  f.seg[0] = FusedSegment{FE_GATES, 0, 4 * H, H, 0};
  f.seg[1] = FusedSegment{FE_ZC, 4 * H, E, H, 0};
  FusedEpi &e = f.e;
  e.H = H; e.V = d.V; e.E = E;
  e.EWi = P + p.EW;
  e.hu = W + w.hq + 4 * H + 4; e.ld_hu = w.ldq2;
  e.tok_prev = io.tok_prev; e.mask = io.mask; e.dp_gates = io.dp_gates;
  e.c_in = W + w.c; e.c_out = W + w.c;
  e.h_in = W + w.h; e.ld_hin = H;
  e.h_out = W + w.h; e.ld_hout = H;
  e.h_copy = nullptr;
  e.h_all = io.h_all;
  e.hd_out = io.dp_h ? W + w.hd : nullptr; e.dp_h = io.dp_h;
  e.zadd = W + w.zadd; e.bz = P + p.bz; e.Wemb = P + p.Wemb; e.prev2out = (d.flags & STAT_PREV2OUT) ? 1 : 0;
  ProfScope ps(PH_FUSED_B, st);
  return fused_phase_launch(f, st);
}

// C: what multiplies h: the attention queries / selector logit / h.U of the next cell (want_q) and the readout
// activation z of this step (want_z; with explicit dropout on h the readout reads h * dp_h instead)
int fstep_hidden(const StatDims &d, const Prep &p, const float *P, const Ws &w, float *W, const StepIO &io,
                 bool want_q, bool want_z, cudaStream_t st) {
  const int H = d.H, E = d.E;
  FusedPhase f;
  memset(&f, 0, sizeof(f));
  f.swap = 1;
  f.W = P + p.WqT; f.wrows = 8 * H + 4 + E; f.wK = H; f.ldw = H;
  f.X[0] = W + w.h; f.xK[0] = H; f.ldx[0] = H;
  if (io.dp_h) { f.X[1] = W + w.hd; f.xK[1] = H; f.ldx[1] = H; }
  f.rows = io.rows;
  int n = 0;
  if (want_q) f.seg[n++] = FusedSegment{FE_STORE, 0, 8 * H + 4, H, 0};
  if (want_z) f.seg[n++] = FusedSegment{FE_Z, 8 * H + 4, E, H, io.dp_h ? 1 : 0};
  f.nseg = n;
  FusedEpi &e = f.e;
  e.E = E;
  e.out = W + w.hq; e.ldo = w.ldq2; e.bias = P + p.bq;
  e.z = W + w.z; e.zadd = W + w.zadd; e.z_alpha = io.dp_h ? 1.0f : 0.5f; e.dp_z = io.dp_z;
  ProfScope ps(PH_FUSED_C, st);
  return fused_phase_launch(f, st);
}

// logits tiles with the per-tile vocabulary reduction, then the combine + bookkeeping of PickArgs k
int fstep_vocab(const StatDims &d, const Prep &p, const float *P, const Ws &w, float *W, int rows, const PickArgs &k,
                cudaStream_t st) {
  FusedPhase f;
  memset(&f, 0, sizeof(f));
  f.swap = 0;
  f.W = P + p.WvT; f.wrows = d.V; f.wK = d.E; f.ldw = d.E;
  f.X[0] = W + w.z; f.xK[0] = d.E; f.ldx[0] = d.E;
  f.rows = rows;
  f.nseg = 1;
  f.seg[0] = FusedSegment{FE_PICK, 0, d.V, d.E, 0};
  FusedEpi &e = f.e;
  e.V = d.V; e.bv = P + p.bv; e.part = W + w.part; e.npart = w.npart; e.x_t = k.x_t; e.tgt = W + w.tgt;
  {
    ProfScope ps(PH_FUSED_LOGITS, st);
    STAT_TRY(fused_phase_launch(f, st));
  }
  ProfScope ps(PH_COMBINE, st);
  return pick_combine_launch(k, W + w.part, w.npart, W + w.tgt, st);
}

// initial state of the rows -> h, c; queries / h.U of the first cell
int fused_begin(const StatDims &d, const Prep &p, const float *P, const Ws &w, float *W, int rows, cudaStream_t st) {
  const int H = d.H;
  STAT_CUDA_CHECK(cudaMemcpy2DAsync(W + w.h, sizeof(float) * H, W + w.h0c0, sizeof(float) * 2 * H,
                                    sizeof(float) * H, rows, cudaMemcpyDeviceToDevice, st));
  STAT_CUDA_CHECK(cudaMemcpy2DAsync(W + w.c, sizeof(float) * H, W + w.h0c0 + H, sizeof(float) * 2 * H,
                                    sizeof(float) * H, rows, cudaMemcpyDeviceToDevice, st));
  StepIO io;
  memset(&io, 0, sizeof(io));
  io.rows = rows;
  return fstep_hidden(d, p, P, w, W, io, true, false, st);
}

// ---------------------------------------------------------------------------
// cell step (cell_step.cu): attention -> cell on the caller's stream (2 dependent launches per step), logits -> pick
// beside the next attention.  Default whenever the shape allows it (H % 32 == 0, 160 <= H <= 512, streaming
// attention kernel, no explicit dropout mask on h); STAT_STEP=0 forces the separate kernels.
// ---------------------------------------------------------------------------
bool cell_enabled(const StatDims &d, const Ws &w, bool dp_h) {
  return step_impl() == 2 && !dp_h && w.att_impl == 2 && cell_plan(d.H, d.E, nullptr);
}

// Beam search (rows = B*k): beyond 128 rows the tensor-core products lose their skinny k-split form (two row tiles,
// full-K loops) and the cell step is the faster chain (measured): default there unless
// STAT_STEP / stat_set_step_impl says otherwise.
// STAT_SPLIT_CHAIN=0: beam searches over 129..192 rows take the cell step (the default before the products of such a
// step were issued as two skinny k-split launches)
bool split_chain() {
  static int on = -1;
  if (on < 0) {
    const char *e = getenv("STAT_SPLIT_CHAIN");
    on = (e && e[0] == '0') ? 0 : 1;
  }
  return on != 0;
}
bool cell_enabled_rows(const StatDims &d, const Ws &w, int rows) {
  static int forced = -2;
  if (forced == -2) forced = getenv("STAT_STEP") ? 1 : 0;
  const bool ok = w.att_impl == 2 && cell_plan(d.H, d.E, nullptr);
  if (!ok) return false;
  if (step_impl() == 2) return true;
  return !forced && !g_step_impl_set && rows > (split_chain() ? SPLIT_ROWS_MAX : 128);
}

CellLaunch cell_args(const StatDims &d, const Prep &p, const float *P, const Ws &w, float *W, const StepIO &io) {
  CellLaunch c;
  memset(&c, 0, sizeof(c));
  c.rows = io.rows; c.H = d.H; c.E = d.E; c.V = d.V;
  c.prev2out = (d.flags & STAT_PREV2OUT) ? 1 : 0;
  c.W1 = P + p.W1s; c.W2 = P + p.W2s;
  c.ctxT = W + w.ctxT; c.hT = W + w.hT;
  c.hq = W + w.hq; c.ldq = w.ldq2;
  c.EW = P + p.EW; c.Wemb = P + p.Wemb; c.bz = P + p.bz; c.bq = P + p.bq;
  c.tok_prev = io.tok_prev; c.mask = io.mask; c.dp_gates = io.dp_gates; c.dp_z = io.dp_z;
  c.h_in = io.h_in; c.c_in = io.c_in; c.h_out = io.h_out; c.c_out = io.c_out; c.h_all = io.h_all;
  c.zadd = W + w.zadd; c.z = W + w.z;
  c.bar = reinterpret_cast<unsigned int *>(W + w.cell_bar);
  return c;
}

// h (rows, H) -> hT, then the queries / selector logit / h.U of the first cell (phase 2 alone)
int cell_begin(const StatDims &d, const Prep &p, const float *P, const Ws &w, float *W, const float *h, int rows,
               cudaStream_t st) {
  const int H = d.H;
  STAT_CUDA_CHECK(cudaMemsetAsync(W + w.cell_bar, 0, 64 * sizeof(float), st));
  for (int r0 = 0; r0 < rows; r0 += 64)
    STAT_TRY(transpose_launch(h + static_cast<size_t>(r0) * H, rows - r0 < 64 ? rows - r0 : 64, H,
                              W + w.hT + static_cast<size_t>(r0 / 64) * H * 64, 64, 0, st));
  StepIO io;
  memset(&io, 0, sizeof(io));
  io.rows = rows;
  CellLaunch c = cell_args(d, p, P, w, W, io);
  c.do1 = 0; c.do2 = 1; c.want_q = 1; c.want_z = 0;
  ProfScope ps(PH_HPROJ, st);
  return cell_launch(c, st);
}

// The readout chain of step t (step_out + the vocabulary reduction) on a side stream, next to the
// attention of step t+1 on the caller's stream; the two meet again at the gates of step t+1, which
// need the token picked in step t.  Event based, so it also works under stream capture.
struct Overlap {
  cudaStream_t main_ = nullptr;
  bool on_ = false, side_busy_ = false;
  SideSet *s_ = nullptr;
  int open(cudaStream_t st, bool enable) {
    main_ = st;
    on_ = false;
    side_busy_ = false;
    if (!enable) return STAT_OK;
    STAT_TRY(side_set(&s_));
    on_ = true;
    return STAT_OK;
  }
  cudaStream_t side() const { return on_ ? s_->readout : main_; }
  // the new hidden state and its products are enqueued on the caller's stream: the side stream may go on
  int state_ready() {
    if (!on_) return STAT_OK;
    STAT_CUDA_CHECK(cudaEventRecord(s_->state, main_));
    STAT_CUDA_CHECK(cudaStreamWaitEvent(s_->readout, s_->state, 0));
    return STAT_OK;
  }
  // the readout chain of this step is enqueued on the side stream
  int side_enqueued() {
    if (!on_) return STAT_OK;
    STAT_CUDA_CHECK(cudaEventRecord(s_->done, s_->readout));
    side_busy_ = true;
    return STAT_OK;
  }
  // the caller's stream needs what the side stream produced (picked token, free readout buffers)
  int join() {
    if (!on_ || !side_busy_) return STAT_OK;
    STAT_CUDA_CHECK(cudaStreamWaitEvent(main_, s_->done, 0));
    side_busy_ = false;
    return STAT_OK;
  }
};

// Serpentine frame order: the attention of odd steps walks each slice backwards, so that what step t read last
// (still in L2) is what step t+1 reads first.   STAT_ATT_SERP=0 turns it off.
int serpentine(int t) {
  static int on = -1;
  if (on < 0) {
    const char *e = getenv("STAT_ATT_SERP");
    on = (e && e[0] == '0') ? 0 : 1;
  }
  return on ? (t & 1) : 0;
}

// attention (writes ctx transposed) -> cell; `ov` = wait for the side stream before the cell (it needs the word
// picked by the previous step's readout chain)
int cell_step(const StatDims &d, const Prep &p, const float *P, const Ws &w, float *W, const StepIO &io, bool want_q,
              float *hq_out, Overlap *ov, cudaStream_t st) {
  AttArgs a = att_args_fused(d, p, P, w, W, io.rows, io.row_clip, io.att_scores, io.alpha_l);
  a.reverse = io.reverse;
  a.rows_per_clip = io.rows_per_clip;
  a.ctx_t = W + w.ctxT;
  {
    ProfScope ps(PH_ATT, st);
    STAT_TRY(att_launch(w, a, st));
  }
  if (ov) STAT_TRY(ov->join());
  CellLaunch c = cell_args(d, p, P, w, W, io);
  c.do1 = 1; c.do2 = 1; c.want_q = want_q ? 1 : 0; c.want_z = 1;
  c.hq_out = hq_out;
  ProfScope ps(PH_FUSED_B, st);
  return cell_launch(c, st);
}

// logits of the readout activation z (written by the cell kernel)
int cell_logits(const StatDims &d, const Prep &p, const float *P, const Ws &w, float *W, int rows, cudaStream_t st) {
  ProfScope pl(PH_LOGITS, st);
  return dense(W + w.z, d.E, rows, P + p.WvT, d.E, d.V, P + p.bv, W + w.logits, w.ldl, 0, 1.f, 1.f, nullptr, 0, st);
}

// Greedy decoding forks the readout chain right after the gates: h.Wl (4 feature tiles x k-split) goes out on the side
// stream while the products the next attention needs run on the caller's stream, so the readout activation does not
// wait for the larger product.  Measured at B = 64 in the captured graph: 43.50 -> 42.93 us per step (two alternating
// runs each).  STAT_SPLIT_HPROJ=0: one product for both (h.[Wl | Wd* | U | W_sel]) on the caller's stream.
bool split_hproj() {
  static int on = -1;
  if (on < 0) {
    const char *e = getenv("STAT_SPLIT_HPROJ");
    on = (e && e[0] == '0') ? 0 : 1;
  }
  return on != 0;
}

bool overlap_enabled() {
  static int on = -1;
  if (on < 0) {
    const char *e = getenv("STAT_OVERLAP");
    on = (e && e[0] == '0') ? 0 : 1;
  }
  return on && !g_prof_on;
}

}  // namespace
}  // namespace stat

using namespace stat;

extern "C" {

int stat_version(void) { return STAT_ABI_VERSION; }

const char *stat_last_error(void) { return stat::get_error(); }

unsigned long long stat_launch_count(void) { return stat::launch_count(); }

int stat_profile_enable(int on) {
  for (auto &r : g_prof) {
    g_event_pool.push_back(r.a);
    g_event_pool.push_back(r.b);
  }
  g_prof.clear();
  if (on) {
    while (g_event_pool.size() < 1024) {
      cudaEvent_t e;
      if (cudaEventCreate(&e) != cudaSuccess) break;
      g_event_pool.push_back(e);
    }
  }
  g_prof_on = on != 0;
  return STAT_OK;
}

int stat_profile_phases(void) { return PH_COUNT; }

const char *stat_profile_phase_name(int phase) {
  return (phase >= 0 && phase < PH_COUNT) ? kPhaseNames[phase] : nullptr;
}

int stat_profile_collect(float *ms_by_phase, int *count_by_phase, int nphase) {
  STAT_REQUIRE(ms_by_phase && count_by_phase && nphase >= PH_COUNT, STAT_EINVAL,
               "profile_collect: need arrays of %d entries", PH_COUNT);
  for (int i = 0; i < nphase; ++i) {
    ms_by_phase[i] = 0.f;
    count_by_phase[i] = 0;
  }
  for (auto &r : g_prof) {
    STAT_CUDA_CHECK(cudaEventSynchronize(r.b));
    float ms = 0.f;
    STAT_CUDA_CHECK(cudaEventElapsedTime(&ms, r.a, r.b));
    ms_by_phase[r.phase] += ms;
    count_by_phase[r.phase] += 1;
    g_event_pool.push_back(r.a);
    g_event_pool.push_back(r.b);
  }
  g_prof.clear();
  return STAT_OK;
}

// fault hunting: 4 ints in host-mapped memory [site, blockIdx.x, threadIdx.x, blockIdx.y<<16|z], written by the
// spin-wait that gives up (its __trap() kills the context; the host word survives).  Returns the HOST pointer.
int *stat_debug_trap_log(void) {
  static int *host = nullptr;
  if (!host) {
    int *dev = nullptr;
    if (cudaHostAlloc(reinterpret_cast<void **>(&host), 64, cudaHostAllocMapped) != cudaSuccess) return nullptr;
    memset(host, 0, 64);
    if (cudaHostGetDevicePointer(reinterpret_cast<void **>(&dev), host, 0) != cudaSuccess) return nullptr;
    if (stat::gemm_set_trap_log(dev) != STAT_OK || stat::att_group_set_trap_log(dev) != STAT_OK) return nullptr;
  }
  return host;
}

int stat_debug_gemm_trace(void *dev_buffer_64_int64) {
  gemm_set_trace(static_cast<long long *>(dev_buffer_64_int64));
  att_group_set_trace(static_cast<long long *>(dev_buffer_64_int64));
  return STAT_OK;
}

long long stat_set_l2_persist(long long bytes) {
  if (check_device() != STAT_OK) return STAT_ECUDA;
  int dev = 0, mx = 0;
  STAT_CUDA_CHECK(cudaGetDevice(&dev));
  STAT_CUDA_CHECK(cudaDeviceGetAttribute(&mx, cudaDevAttrMaxPersistingL2CacheSize, dev));
  size_t want = (bytes < 0 || bytes > mx) ? static_cast<size_t>(mx) : static_cast<size_t>(bytes);
  STAT_CUDA_CHECK(cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, want));
  return static_cast<long long>(want);
}

size_t stat_clip_scratch_bytes(void) { return clip_scratch_bytes(); }

int stat_grad_clip(float *grads, size_t n, float clip_c, void *scratch, float *out_g2, void *stream) {
  STAT_TRY(check_device());
  STAT_REQUIRE(grads && scratch && n > 0, STAT_EINVAL, "grad_clip: bad argument");
  STAT_REQUIRE((reinterpret_cast<uintptr_t>(grads) & 15) == 0 && (reinterpret_cast<uintptr_t>(scratch) & 15) == 0,
               STAT_EALIGN, "grad_clip: buffers must be 16-byte aligned");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  STAT_TRY(grad_clip_launch(grads, n, clip_c, scratch, st));
  if (out_g2)
    STAT_CUDA_CHECK(cudaMemcpyAsync(out_g2, static_cast<char *>(scratch) + clip_scratch_bytes() - 16, 2 * sizeof(float),
                                    cudaMemcpyDeviceToDevice, st));
  return STAT_OK;
}

int stat_alpha_coverage(const float *alphas, int L, int rows, int n, void *scratch, float *out, void *stream) {
  STAT_TRY(check_device());
  STAT_REQUIRE(alphas && scratch && out && L >= 1 && rows >= 1 && n >= 1, STAT_EINVAL, "alpha_coverage: bad argument");
  return coverage_launch(alphas, L, rows, n, scratch, out, static_cast<cudaStream_t>(stream));
}

int stat_adam_step(float *params, const float *grads, float *m, float *v, size_t n, int step, void *stream) {
  STAT_TRY(check_device());
  STAT_REQUIRE(params && grads && m && v && n > 0 && step >= 1, STAT_EINVAL, "adam_step: bad argument (step is 1-based)");
  STAT_REQUIRE(((reinterpret_cast<uintptr_t>(params) | reinterpret_cast<uintptr_t>(grads) | reinterpret_cast<uintptr_t>(m) |
                 reinterpret_cast<uintptr_t>(v)) & 15) == 0,
               STAT_EALIGN, "adam_step: buffers must be 16-byte aligned");
  return adam_launch(params, grads, m, v, n, step, static_cast<cudaStream_t>(stream));
}

int stat_adadelta_step(float *params, const float *grads, float *rg2, float *ru2, size_t n, int phase, void *stream) {
  STAT_TRY(check_device());
  STAT_REQUIRE(params && grads && rg2 && ru2 && n > 0 && (phase == 0 || phase == 1), STAT_EINVAL,
               "adadelta_step: bad argument");
  return adadelta_launch(params, grads, rg2, ru2, n, phase, static_cast<cudaStream_t>(stream));
}

int stat_set_step_impl(int impl) {
  STAT_REQUIRE(impl >= -1 && impl <= 2, STAT_EINVAL,
               "step impl must be 0 (separate kernels), 1 (fused tile kernels), 2 (cell step) or -1 (default)");
  g_step_impl = impl;
  g_step_impl_set = impl >= 0;
  return STAT_OK;
}

int stat_set_beam_share(int on) {
  STAT_REQUIRE(on >= -1 && on <= 1, STAT_EINVAL, "beam share must be 0, 1 or -1 (default)");
  att_group_set_share(on);
  return STAT_OK;
}

int stat_set_gemm_impl(int impl) {
  STAT_REQUIRE(impl >= 0 && impl <= 2, STAT_EINVAL,
               "gemm impl must be 0 (tcgen05 3xTF32, A from TMEM), 1 (fp32 SIMT) or 2 (tcgen05 3xTF32, A from smem)");
  gemm_set_impl(impl);
  return STAT_OK;
}

size_t stat_prepared_bytes(const StatDims *d) {
  if (check_dims(d) != STAT_OK) return 0;
  return prep_layout(*d).total * sizeof(float);
}

int stat_prepare_params(const StatDims *d, const StatParams *sp, void *prepared, void *stream) {
  STAT_TRY(check_dims(d));
  STAT_TRY(check_device());
  STAT_REQUIRE(sp != nullptr && prepared != nullptr, STAT_EINVAL, "prepare: NULL argument");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const Prep p = prep_layout(*d);
  float *P = static_cast<float *>(prepared);
  const int H = d->H, E = d->E, V = d->V;
  const bool sel = d->flags & STAT_SELECTOR, c2o = d->flags & STAT_CTX2OUT, gp = d->flags & STAT_GLOBAL_PROJ;
#define NEED(f) STAT_REQUIRE(sp->f != nullptr, STAT_EINVAL, "prepare: parameter %s is NULL", #f)
  NEED(Wemb); NEED(ff_state_W); NEED(ff_state_b); NEED(ff_memory_W); NEED(ff_memory_b);
  NEED(ff_local_W); NEED(ff_local_b); NEED(ff_motion_W); NEED(ff_motion_b);
  NEED(decoder_W); NEED(decoder_U); NEED(decoder_b); NEED(decoder_Wc);
  NEED(decoder_Wcg_att); NEED(decoder_Wcm_att); NEED(decoder_Wclt_att);
  NEED(decoder_Wdg_att); NEED(decoder_Wdm_att); NEED(decoder_Wdlt_att);
  NEED(decoder_bg_att); NEED(decoder_bm_att); NEED(decoder_blt_att);
  NEED(decoder_Wcl_att); NEED(decoder_Wdl_att); NEED(decoder_bl_att);
  NEED(decoder_Ug_att); NEED(decoder_cg_att); NEED(decoder_Um_att); NEED(decoder_cm_att);
  NEED(decoder_Ult_att); NEED(decoder_clt_att); NEED(decoder_Ul_att); NEED(decoder_cl_att);
  NEED(ff_logit_lstm_W); NEED(ff_logit_lstm_b); NEED(ff_logit_W); NEED(ff_logit_b);
  if (sel) { NEED(decoder_W_sel); NEED(decoder_b_sel); }
  if (c2o) { NEED(ff_logit_ctxglm_W); NEED(ff_logit_ctxglm_b); }
  if (gp) { NEED(ff_global_W); NEED(ff_global_b); }
#undef NEED
  STAT_CUDA_CHECK(cudaMemsetAsync(P, 0, p.total * sizeof(float), st));
  auto cp = [&](size_t dst, const float *src, size_t n) {
    return cudaMemcpyAsync(P + dst, src, n * sizeof(float), cudaMemcpyDeviceToDevice, st);
  };
  // hidden-state projections
  STAT_TRY(transpose_launch(sp->decoder_Wdl_att, H, H, P + p.WhT, H, 0, st));
  STAT_TRY(transpose_launch(sp->decoder_Wdg_att, H, H, P + p.WhT, H, H, st));
  STAT_TRY(transpose_launch(sp->decoder_Wdm_att, H, H, P + p.WhT, H, 2 * H, st));
  STAT_TRY(transpose_launch(sp->decoder_Wdlt_att, H, H, P + p.WhT, H, 3 * H, st));
  STAT_TRY(transpose_launch(sp->decoder_U, H, 4 * H, P + p.WhT, H, 4 * H, st));
  STAT_CUDA_CHECK(cp(p.bh + 3 * H, sp->decoder_blt_att, H));
  if (sel) {
    STAT_CUDA_CHECK(cp(p.WhT + static_cast<size_t>(8) * H * H, sp->decoder_W_sel, H));
    STAT_CUDA_CHECK(cp(p.bh + 8 * H, sp->decoder_b_sel, 1));
  }
  // context -> gates / readout
  STAT_TRY(transpose_launch(sp->decoder_Wc, H, 4 * H, P + p.WcT, H, 0, st));
  if (c2o) STAT_TRY(transpose_launch(sp->ff_logit_ctxglm_W, H, E, P + p.WcT, H, 4 * H, st));
  STAT_TRY(transpose_launch(sp->ff_logit_lstm_W, H, E, P + p.WlT, H, 0, st));
  STAT_TRY(add_vec_launch(P + p.bz, sp->ff_logit_lstm_b, c2o ? sp->ff_logit_ctxglm_b : nullptr, E, st));
  STAT_TRY(transpose_launch(sp->ff_logit_W, E, V, P + p.WvT, E, 0, st));
  STAT_CUDA_CHECK(cp(p.bv, sp->ff_logit_b, V));
  // prologue
  STAT_TRY(transpose_launch(sp->ff_state_W, d->Dg, H, P + p.WstT, d->Dg, 0, st));
  STAT_TRY(transpose_launch(sp->ff_memory_W, d->Dg, H, P + p.WstT, d->Dg, H, st));
  STAT_CUDA_CHECK(cp(p.bst, sp->ff_state_b, H));
  STAT_CUDA_CHECK(cp(p.bst + H, sp->ff_memory_b, H));
  if (gp) {
    STAT_TRY(transpose_launch(sp->ff_global_W, d->Dg, H, P + p.WgT, d->Dg, 0, st));
    STAT_CUDA_CHECK(cp(p.bg0, sp->ff_global_b, H));
  }
  STAT_TRY(transpose_launch(sp->ff_local_W, d->Dr, H, P + p.WlocT, d->Dr, 0, st));
  STAT_CUDA_CHECK(cp(p.bloc, sp->ff_local_b, H));
  STAT_TRY(transpose_launch(sp->ff_motion_W, d->Dm, H, P + p.WmotT, d->Dm, 0, st));
  STAT_CUDA_CHECK(cp(p.bmot, sp->ff_motion_b, H));
  STAT_TRY(transpose_launch(sp->decoder_Wcg_att, H, H, P + p.WcgT, H, 0, st));
  STAT_CUDA_CHECK(cp(p.bcg, sp->decoder_bg_att, H));
  STAT_TRY(transpose_launch(sp->decoder_Wcm_att, H, H, P + p.WcmT, H, 0, st));
  STAT_CUDA_CHECK(cp(p.bcm, sp->decoder_bm_att, H));
  STAT_TRY(transpose_launch(sp->decoder_Wcl_att, H, H, P + p.WclT, H, 0, st));
  STAT_TRY(transpose_launch(sp->decoder_Wclt_att, H, H, P + p.WclT, H, H, st));
  STAT_CUDA_CHECK(cp(p.bcl, sp->decoder_bl_att, H));
  // score vectors and their scalar biases
  STAT_CUDA_CHECK(cp(p.U4, sp->decoder_Ul_att, H));
  STAT_CUDA_CHECK(cp(p.U4 + H, sp->decoder_Ug_att, H));
  STAT_CUDA_CHECK(cp(p.U4 + 2 * H, sp->decoder_Um_att, H));
  STAT_CUDA_CHECK(cp(p.U4 + 3 * H, sp->decoder_Ult_att, H));
  STAT_CUDA_CHECK(cp(p.U4 + 4 * H, sp->decoder_cl_att, 1));
  STAT_CUDA_CHECK(cp(p.U4 + 4 * H + 1, sp->decoder_cg_att, 1));
  STAT_CUDA_CHECK(cp(p.U4 + 4 * H + 2, sp->decoder_cm_att, 1));
  STAT_CUDA_CHECK(cp(p.U4 + 4 * H + 3, sp->decoder_clt_att, 1));
  // token -> gate input table: EW[x] = Wemb[x].W + b ; EW[V] = b (no previous word)
  // (columns gate-interleaved, 4*unit + gate: the four gate inputs of a hidden unit are adjacent)
  STAT_CUDA_CHECK(cp(p.Wemb, sp->Wemb, static_cast<size_t>(V) * E));
  STAT_TRY(transpose_il_launch(sp->decoder_W, E, H, P + p.WdT, E, 0, st));
  STAT_TRY(interleave4_launch(sp->decoder_b, P + p.bdi, H, st));
  STAT_TRY(dense(P + p.Wemb, E, V, P + p.WdT, E, 4 * H, P + p.bdi, P + p.EW, 4 * H, 0, 1.f, 1.f, nullptr, 0,
                 st));
  STAT_CUDA_CHECK(cp(p.EW + static_cast<size_t>(V) * 4 * H, P + p.bdi, 4 * H));
  // fused step: ctx . Wc on gate-interleaved rows, the ctx -> readout rows behind them
  STAT_TRY(transpose_il_launch(sp->decoder_Wc, H, H, P + p.WcI, H, 0, st));
  if (c2o) STAT_TRY(transpose_launch(sp->ff_logit_ctxglm_W, H, E, P + p.WcI, H, 4 * H, st));
  // everything that multiplies the new hidden state: next step's attention queries, selector logit, h.U of the
  // next cell (gate-interleaved like ctx.Wc), readout
  STAT_TRY(transpose_launch(sp->decoder_Wdl_att, H, H, P + p.WqT, H, 0, st));
  STAT_TRY(transpose_launch(sp->decoder_Wdg_att, H, H, P + p.WqT, H, H, st));
  STAT_TRY(transpose_launch(sp->decoder_Wdm_att, H, H, P + p.WqT, H, 2 * H, st));
  STAT_TRY(transpose_launch(sp->decoder_Wdlt_att, H, H, P + p.WqT, H, 3 * H, st));
  if (sel) STAT_CUDA_CHECK(cp(p.WqT + static_cast<size_t>(4) * H * H, sp->decoder_W_sel, H));
  STAT_TRY(transpose_il_launch(sp->decoder_U, H, H, P + p.WqT + static_cast<size_t>(4 * H + 4) * H, H, 0, st));
  STAT_TRY(transpose_launch(sp->ff_logit_lstm_W, H, E, P + p.WqT, H, 8 * H + 4, st));
  STAT_CUDA_CHECK(cp(p.bq + 3 * H, sp->decoder_blt_att, H));
  if (sel) STAT_CUDA_CHECK(cp(p.bq + 4 * H, sp->decoder_b_sel, 1));
  STAT_TRY(cell_pack_launch(P + p.WcI, P + p.WqT, P + p.W1s, P + p.W2s, H, E, c2o ? 1 : 0, st));
  return STAT_OK;
}

size_t stat_workspace_bytes(const StatDims *d, int rows) {
  if (check_dims(d) != STAT_OK || rows < 1) return 0;
  return ws_layout(*d, rows).total * sizeof(float);
}

int stat_workspace_region(const StatDims *d, int rows, const char *name, size_t *offset, size_t *bytes) {
  STAT_TRY(check_dims(d));
  STAT_REQUIRE(rows >= 1 && name && offset && bytes, STAT_EINVAL, "workspace_region: bad argument");
  const Ws w = ws_layout(*d, rows);
  const size_t B = d->B, T = d->T, R = d->R, H = d->H, n = rows;
  struct { const char *nm; size_t off, cnt; } tab[] = {
      {"ctxg0", w.ctxg0, B * T * H}, {"pctxg", w.pctxg, B * T * H}, {"ctxm0", w.ctxm0, B * T * H},
      {"pctxm", w.pctxm, B * T * H}, {"ctxl0", w.ctxl0, B * T * R * H}, {"pctxl", w.pctxl, B * T * R * H},
      {"qctxl", w.qctxl, B * T * R * H}, {"h0", w.h0c0, B * 2 * H}, {"c0", w.h0c0 + H, B * 2 * H - H},
      {"h", w.h, n * H}, {"c", w.c, n * H}, {"hp", w.hpz, n * w.ldhp}, {"ctx", w.ctx, n * H},
      {"logits", w.logits, n * w.ldl}, {"att_scores", w.att_scores, 3 * n * T},
      {"alpha_l", w.alpha_l, n * T * R}, {"z", w.z, n * static_cast<size_t>(d->E)},
  };
  for (const auto &e : tab) {
    if (strcmp(e.nm, name) == 0) {
      *offset = e.off * sizeof(float);
      *bytes = e.cnt * sizeof(float);
      return STAT_OK;
    }
  }
  set_error("workspace_region: unknown region '%s'", name);
  return STAT_EINVAL;
}

int stat_precompute(const StatDims *d, const void *prepared, const float *ctxg, const float *mask_ctxg,
                    const float *ctxl, const float *ctxm, void *ws, void *stream) {
  STAT_TRY(check_dims(d));
  STAT_TRY(check_device());
  STAT_REQUIRE(prepared && ctxg && mask_ctxg && ctxl && ctxm && ws, STAT_EINVAL, "precompute: NULL argument");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const Prep p = prep_layout(*d);
  const Ws w = ws_layout(*d, d->B);  // context blocks sit first; their offsets do not depend on rows
  const float *P = static_cast<const float *>(prepared);
  float *W = static_cast<float *>(ws);
  const int B = d->B, T = d->T, R = d->R, H = d->H;
  // Three independent chains: local (main stream), global + init state, motion.  The two short
  // ones run on side streams forked from / joined to the caller's stream (also under graph
  // capture); with phase timing on, everything stays on the caller's stream.
  Fork fk;
  STAT_TRY(fk.open(st, !g_prof_on));
  cudaStream_t sg = fk.side(0), sm = fk.side(1);
  STAT_TRY(init_state(*d, p, P, w, W, ctxg, mask_ctxg, sg));
  {
    ProfScope ps(PH_K0_GLOBAL, sg);
    if (d->flags & STAT_GLOBAL_PROJ) {
      STAT_TRY(dense(ctxg, d->Dg, B * T, P + p.WgT, d->Dg, H, P + p.bg0, W + w.ctxg0, H, 1, 1.f, 1.f, nullptr, 0,
                     sg));
    } else {
      STAT_CUDA_CHECK(cudaMemcpyAsync(W + w.ctxg0, ctxg, sizeof(float) * B * T * H, cudaMemcpyDeviceToDevice, sg));
    }
    STAT_TRY(dense(W + w.ctxg0, H, B * T, P + p.WcgT, H, H, P + p.bcg, W + w.pctxg, H, 0, 1.f, 1.f, nullptr, 0,
                   sg));
  }
  {
    ProfScope ps(PH_K0_MOTION, sm);
    STAT_TRY(dense(ctxm, d->Dm, B * T, P + p.WmotT, d->Dm, H, P + p.bmot, W + w.ctxm0, H, 1, 1.f, 1.f, nullptr, 0,
                   sm));
    STAT_TRY(dense(W + w.ctxm0, H, B * T, P + p.WcmT, H, H, P + p.bcm, W + w.pctxm, H, 0, 1.f, 1.f, nullptr, 0,
                   sm));
  }
  {
    ProfScope ps(PH_K0_LOCAL, st);
    STAT_TRY(dense(ctxl, d->Dr, B * T * R, P + p.WlocT, d->Dr, H, P + p.bloc, W + w.ctxl0, H, 1, 1.f, 1.f, nullptr,
                   0, st));
  }
  // P5 for the local block, and Q = ctxl0.Wclt_att (the :416 product made step-invariant)
  {
    ProfScope ps5(PH_K0_PROJ, st);
    const int nlr = B * T * R;
    if (H % 128 == 0 && nlr > 128) {
      GemmArgs g;
      memset(&g, 0, sizeof(g));
      g.P = W + w.ctxl0; g.ldp = H; g.NP = nlr;
      g.Q = P + p.WclT; g.ldq = H; g.NQ = 2 * H;
      g.K = H; g.feat_on_p = 0; g.nseg = 2;
      g.seg[0] = GemmSeg{W + w.pctxl, H, P + p.bcl, nullptr, 0, 1.f, 1.f, 0, 0, H};
      g.seg[1] = GemmSeg{W + w.qctxl, H, nullptr, nullptr, 0, 1.f, 1.f, 0, H, 2 * H};
      STAT_TRY(gemm_launch(g, st));
    } else {
      STAT_TRY(dense(W + w.ctxl0, H, nlr, P + p.WclT, H, H, P + p.bcl, W + w.pctxl, H, 0, 1.f, 1.f, nullptr, 0,
                     st));
      STAT_TRY(dense(W + w.ctxl0, H, nlr, P + p.WclT + static_cast<size_t>(H) * H, H, H, nullptr, W + w.qctxl, H,
                     0, 1.f, 1.f, nullptr, 0, st));
    }
  }
  return fk.join();
}

int stat_init_state(const StatDims *d, const void *prepared, const float *ctxg, const float *mask_ctxg, void *ws,
                    float *out_h0, float *out_c0, void *stream) {
  STAT_TRY(check_dims(d));
  STAT_TRY(check_device());
  STAT_REQUIRE(prepared && ctxg && mask_ctxg && ws && out_h0 && out_c0, STAT_EINVAL, "init_state: NULL argument");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const Prep p = prep_layout(*d);
  const Ws w = ws_layout(*d, d->B);
  const float *P = static_cast<const float *>(prepared);
  float *W = static_cast<float *>(ws);
  const int H = d->H;
  STAT_TRY(init_state(*d, p, P, w, W, ctxg, mask_ctxg, st));
  STAT_CUDA_CHECK(cudaMemcpy2DAsync(out_h0, sizeof(float) * H, W + w.h0c0, sizeof(float) * 2 * H,
                                    sizeof(float) * H, d->B, cudaMemcpyDeviceToDevice, st));
  STAT_CUDA_CHECK(cudaMemcpy2DAsync(out_c0, sizeof(float) * H, W + w.h0c0 + H, sizeof(float) * 2 * H,
                                    sizeof(float) * H, d->B, cudaMemcpyDeviceToDevice, st));
  return STAT_OK;
}

int stat_forward_teacher(const StatDims *d, const void *prepared, void *ws, int L, const int64_t *x,
                         const float *mask, const float *dp_gates, const float *dp_h, const float *dp_z,
                         float *out_logprob, float *out_alpha_l, float *out_alpha_g, float *out_alpha_m,
                         float *out_alpha_lt, float *out_h, void *stream) {
  STAT_TRY(check_dims(d));
  STAT_TRY(check_device());
  STAT_REQUIRE(prepared && ws && x && mask && out_logprob && L >= 1, STAT_EINVAL, "forward_teacher: bad argument");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const Prep p = prep_layout(*d);
  const Ws w = ws_layout(*d, d->B);
  const float *P = static_cast<const float *>(prepared);
  float *W = static_cast<float *>(ws);
  const int B = d->B, T = d->T, R = d->R, H = d->H, E = d->E;
  const bool want_t = out_alpha_g || out_alpha_m || out_alpha_lt;
  STAT_CUDA_CHECK(cudaMemsetAsync(out_logprob, 0, sizeof(float) * B, st));
  STAT_CUDA_CHECK(cudaMemsetAsync(W + w.lpc, 0, sizeof(float) * B, st));
  STAT_CUDA_CHECK(cudaMemsetAsync(W + w.counters, 0, sizeof(float) * B, st));
  Overlap ov;
  if (cell_enabled(*d, w, dp_h != nullptr)) {
    STAT_CUDA_CHECK(cudaMemcpy2DAsync(W + w.h, sizeof(float) * H, W + w.h0c0, sizeof(float) * 2 * H,
                                      sizeof(float) * H, B, cudaMemcpyDeviceToDevice, st));
    STAT_CUDA_CHECK(cudaMemcpy2DAsync(W + w.c, sizeof(float) * H, W + w.h0c0 + H, sizeof(float) * 2 * H,
                                      sizeof(float) * H, B, cudaMemcpyDeviceToDevice, st));
    STAT_TRY(cell_begin(*d, p, P, w, W, W + w.h, B, st));
    STAT_TRY(ov.open(st, overlap_enabled()));
    for (int t = 0; t < L; ++t) {
      StepIO io;
      memset(&io, 0, sizeof(io));
      io.rows = B;
      io.tok_prev = t > 0 ? x + static_cast<size_t>(t - 1) * B : nullptr;   // emb shifted by one step (:613-617)
      io.mask = mask + static_cast<size_t>(t) * B;
      io.dp_gates = dp_gates ? dp_gates + static_cast<size_t>(t) * B * 3 * H : nullptr;
      io.dp_z = dp_z ? dp_z + static_cast<size_t>(t) * B * E : nullptr;
      io.h_in = W + w.h; io.c_in = W + w.c; io.h_out = W + w.h; io.c_out = W + w.c;
      io.h_all = out_h ? out_h + static_cast<size_t>(t) * B * H : nullptr;
      io.alpha_l = out_alpha_l ? out_alpha_l + static_cast<size_t>(t) * B * T * R : nullptr;
      io.att_scores = want_t ? W + w.att_scores : nullptr;
      io.reverse = serpentine(t);
      STAT_TRY(cell_step(*d, p, P, w, W, io, t + 1 < L, nullptr, &ov, st));
      if (want_t) {
        const size_t n = static_cast<size_t>(B) * T;
        STAT_TRY(softmax_rows3_launch(W + w.att_scores, out_alpha_g ? out_alpha_g + t * n : nullptr,
                                      out_alpha_m ? out_alpha_m + t * n : nullptr,
                                      out_alpha_lt ? out_alpha_lt + t * n : nullptr, B, T, st));
      }
      STAT_TRY(ov.state_ready());
      STAT_TRY(cell_logits(*d, p, P, w, W, B, ov.side()));
      PickArgs k;
      memset(&k, 0, sizeof(k));
      k.rows = B; k.V = d->V; k.ldl = w.ldl; k.logits = W + w.logits;
      k.x_t = x + static_cast<size_t>(t) * B;
      k.mask_t = mask + static_cast<size_t>(t) * B;
      k.logprob = out_logprob; k.logprob_comp = W + w.lpc;
      {
        ProfScope ps(PH_PICK, ov.side());
        STAT_TRY(pick_launch(k, ov.side()));
      }
      STAT_TRY(ov.side_enqueued());
    }
    return ov.join();
  }
  if (fused_enabled(*d, B)) {
    STAT_TRY(fused_begin(*d, p, P, w, W, B, st));
    STAT_TRY(ov.open(st, overlap_enabled()));
    for (int t = 0; t < L; ++t) {
      StepIO io;
      memset(&io, 0, sizeof(io));
      io.rows = B;
      io.tok_prev = t > 0 ? x + static_cast<size_t>(t - 1) * B : nullptr;   // emb shifted by one step (:613-617)
      io.mask = mask + static_cast<size_t>(t) * B;
      io.dp_gates = dp_gates ? dp_gates + static_cast<size_t>(t) * B * 3 * H : nullptr;
      io.dp_h = dp_h ? dp_h + static_cast<size_t>(t) * B * H : nullptr;
      io.dp_z = dp_z ? dp_z + static_cast<size_t>(t) * B * E : nullptr;
      io.h_all = out_h ? out_h + static_cast<size_t>(t) * B * H : nullptr;
      io.alpha_l = out_alpha_l ? out_alpha_l + static_cast<size_t>(t) * B * T * R : nullptr;
      io.att_scores = want_t ? W + w.att_scores : nullptr;
      {
        AttArgs a = att_args_fused(*d, p, P, w, W, B, nullptr, io.att_scores, io.alpha_l);
        ProfScope ps(PH_ATT, st);
        STAT_TRY(att_launch(w, a, st));
      }
      if (want_t) {
        const size_t n = static_cast<size_t>(B) * T;
        STAT_TRY(softmax_rows3_launch(W + w.att_scores, out_alpha_g ? out_alpha_g + t * n : nullptr,
                                      out_alpha_m ? out_alpha_m + t * n : nullptr,
                                      out_alpha_lt ? out_alpha_lt + t * n : nullptr, B, T, st));
      }
      STAT_TRY(ov.join());                  // zadd / z / partial buffers of the previous step's readout are free
      STAT_TRY(fstep_gates(*d, p, P, w, W, io, st));
      STAT_TRY(fstep_hidden(*d, p, P, w, W, io, t + 1 < L, true, st));
      STAT_TRY(ov.state_ready());
      PickArgs k;
      memset(&k, 0, sizeof(k));
      k.rows = B; k.V = d->V;
      k.x_t = x + static_cast<size_t>(t) * B;
      k.mask_t = mask + static_cast<size_t>(t) * B;
      k.logprob = out_logprob; k.logprob_comp = W + w.lpc;
      STAT_TRY(fstep_vocab(*d, p, P, w, W, B, k, ov.side()));
      STAT_TRY(ov.side_enqueued());
    }
    return ov.join();
  }
  STAT_CUDA_CHECK(cudaMemcpy2DAsync(W + w.h, sizeof(float) * H, W + w.h0c0, sizeof(float) * 2 * H,
                                    sizeof(float) * H, B, cudaMemcpyDeviceToDevice, st));
  STAT_CUDA_CHECK(cudaMemcpy2DAsync(W + w.c, sizeof(float) * H, W + w.h0c0 + H, sizeof(float) * 2 * H,
                                    sizeof(float) * H, B, cudaMemcpyDeviceToDevice, st));
  STAT_TRY(h_proj(*d, p, P, w, W, W + w.h, B, 1, st));
  STAT_TRY(ov.open(st, overlap_enabled()));
  for (int t = 0; t < L; ++t) {
    StepIO io;
    memset(&io, 0, sizeof(io));
    io.rows = B;
    io.tok_prev = t > 0 ? x + static_cast<size_t>(t - 1) * B : nullptr;   // emb shifted by one step (:613-617)
    io.mask = mask + static_cast<size_t>(t) * B;
    io.dp_gates = dp_gates ? dp_gates + static_cast<size_t>(t) * B * 3 * H : nullptr;
    io.dp_h = dp_h ? dp_h + static_cast<size_t>(t) * B * H : nullptr;
    io.dp_z = dp_z ? dp_z + static_cast<size_t>(t) * B * E : nullptr;
    io.h_in = W + w.h; io.c_in = W + w.c; io.h_out = W + w.h; io.c_out = W + w.c;
    io.reverse = serpentine(t);
    io.h_all = out_h ? out_h + static_cast<size_t>(t) * B * H : nullptr;
    io.alpha_l = out_alpha_l ? out_alpha_l + static_cast<size_t>(t) * B * T * R : nullptr;
    io.att_scores = want_t ? W + w.att_scores : nullptr;
    STAT_TRY(step_att(*d, p, P, w, W, io, st));
    STAT_TRY(ov.join());
    STAT_TRY(step_gates(*d, p, P, w, W, io, t + 1 < L, st));
    STAT_TRY(ov.state_ready());
    STAT_TRY(step_out(*d, p, P, w, W, io, ov.side()));
    if (want_t) {
      const size_t n = static_cast<size_t>(B) * T;
      STAT_TRY(softmax_rows3_launch(W + w.att_scores, out_alpha_g ? out_alpha_g + t * n : nullptr,
                                    out_alpha_m ? out_alpha_m + t * n : nullptr,
                                    out_alpha_lt ? out_alpha_lt + t * n : nullptr, B, T, st));
    }
    PickArgs k;
    memset(&k, 0, sizeof(k));
    k.rows = B; k.V = d->V; k.ldl = w.ldl; k.logits = W + w.logits;
    k.x_t = x + static_cast<size_t>(t) * B;
    k.mask_t = mask + static_cast<size_t>(t) * B;
    k.logprob = out_logprob; k.logprob_comp = W + w.lpc;
    {
      ProfScope ps(PH_PICK, ov.side());
      STAT_TRY(pick_launch(k, ov.side()));
    }
    STAT_TRY(ov.side_enqueued());
  }
  return ov.join();
}

int stat_decode_greedy(const StatDims *d, const void *prepared, void *ws, int maxlen, int64_t *out_tokens,
                       int32_t *out_lengths, float *out_scores, void *stream) {
  STAT_TRY(check_dims(d));
  STAT_TRY(check_device());
  STAT_REQUIRE(prepared && ws && out_tokens && out_lengths && out_scores && maxlen >= 1, STAT_EINVAL,
               "decode_greedy: bad argument");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const Prep p = prep_layout(*d);
  const Ws w = ws_layout(*d, d->B);
  const float *P = static_cast<const float *>(prepared);
  float *W = static_cast<float *>(ws);
  const int B = d->B, H = d->H;
  int64_t *tok_prev = reinterpret_cast<int64_t *>(W + w.tok_prev);
  int32_t *alive = reinterpret_cast<int32_t *>(W + w.alive);
  STAT_CUDA_CHECK(cudaMemsetAsync(W + w.counters, 0, sizeof(float) * B, st));
  STAT_TRY(init_rows_launch(B, tok_prev, alive, out_lengths, out_scores, out_tokens, maxlen, st));
  Overlap ov;
  if (cell_enabled(*d, w, false)) {
    STAT_CUDA_CHECK(cudaMemcpy2DAsync(W + w.h, sizeof(float) * H, W + w.h0c0, sizeof(float) * 2 * H,
                                      sizeof(float) * H, B, cudaMemcpyDeviceToDevice, st));
    STAT_CUDA_CHECK(cudaMemcpy2DAsync(W + w.c, sizeof(float) * H, W + w.h0c0 + H, sizeof(float) * 2 * H,
                                      sizeof(float) * H, B, cudaMemcpyDeviceToDevice, st));
    STAT_TRY(cell_begin(*d, p, P, w, W, W + w.h, B, st));
    STAT_TRY(ov.open(st, overlap_enabled()));
    for (int t = 0; t < maxlen; ++t) {
      StepIO io;
      memset(&io, 0, sizeof(io));
      io.rows = B;
      io.tok_prev = tok_prev;   // -1 on the first step: no previous word (:893, :803-804)
      io.h_in = W + w.h; io.c_in = W + w.c; io.h_out = W + w.h; io.c_out = W + w.c;
      io.reverse = serpentine(t);
      STAT_TRY(cell_step(*d, p, P, w, W, io, t + 1 < maxlen, nullptr, &ov, st));
      STAT_TRY(ov.state_ready());
      STAT_TRY(cell_logits(*d, p, P, w, W, B, ov.side()));
      PickArgs k;
      memset(&k, 0, sizeof(k));
      k.rows = B; k.V = d->V; k.ldl = w.ldl; k.logits = W + w.logits;
      k.tokens = out_tokens; k.maxlen = maxlen; k.t = t;
      k.lengths = out_lengths; k.scores = out_scores; k.alive = alive; k.tok_prev = tok_prev;
      {
        ProfScope ps(PH_PICK, ov.side());
        STAT_TRY(pick_launch(k, ov.side()));
      }
      STAT_TRY(ov.side_enqueued());
    }
    return ov.join();
  }
  if (fused_enabled(*d, B)) {
    STAT_TRY(fused_begin(*d, p, P, w, W, B, st));
    STAT_TRY(ov.open(st, overlap_enabled()));
    for (int t = 0; t < maxlen; ++t) {
      StepIO io;
      memset(&io, 0, sizeof(io));
      io.rows = B;
      io.tok_prev = tok_prev;   // -1 on the first step: no previous word (:893, :803-804)
      {
        AttArgs a = att_args_fused(*d, p, P, w, W, B, nullptr, nullptr, nullptr);
        ProfScope ps(PH_ATT, st);
        STAT_TRY(att_launch(w, a, st));
      }
      STAT_TRY(ov.join());                  // the gates need the word picked in the previous step
      STAT_TRY(fstep_gates(*d, p, P, w, W, io, st));
      STAT_TRY(fstep_hidden(*d, p, P, w, W, io, t + 1 < maxlen, true, st));
      STAT_TRY(ov.state_ready());
      PickArgs k;
      memset(&k, 0, sizeof(k));
      k.rows = B; k.V = d->V;
      k.tokens = out_tokens; k.maxlen = maxlen; k.t = t;
      k.lengths = out_lengths; k.scores = out_scores; k.alive = alive; k.tok_prev = tok_prev;
      STAT_TRY(fstep_vocab(*d, p, P, w, W, B, k, ov.side()));
      STAT_TRY(ov.side_enqueued());
    }
    return ov.join();
  }
  STAT_CUDA_CHECK(cudaMemcpy2DAsync(W + w.h, sizeof(float) * H, W + w.h0c0, sizeof(float) * 2 * H,
                                    sizeof(float) * H, B, cudaMemcpyDeviceToDevice, st));
  STAT_CUDA_CHECK(cudaMemcpy2DAsync(W + w.c, sizeof(float) * H, W + w.h0c0 + H, sizeof(float) * 2 * H,
                                    sizeof(float) * H, B, cudaMemcpyDeviceToDevice, st));
  STAT_TRY(h_proj(*d, p, P, w, W, W + w.h, B, 1, st));
  STAT_TRY(ov.open(st, overlap_enabled()));
  for (int t = 0; t < maxlen; ++t) {
    StepIO io;
    memset(&io, 0, sizeof(io));
    io.rows = B;
    io.tok_prev = tok_prev;   // -1 on the first step: no previous word (:893, :803-804)
    io.h_in = W + w.h; io.c_in = W + w.c; io.h_out = W + w.h; io.c_out = W + w.c;
    io.reverse = serpentine(t);
    STAT_TRY(step_att(*d, p, P, w, W, io, st));
    STAT_TRY(ov.join());
    if (split_hproj() && ov.on_) {
      // the readout chain forks right after the gates: h.Wl (4 feature tiles) on the side stream, the products the
      // next attention needs on the caller's stream -- the readout activation does not wait for the larger product
      STAT_TRY(step_gates(*d, p, P, w, W, io, false, st, true));
      STAT_TRY(ov.state_ready());
      if (t + 1 < maxlen) STAT_TRY(h_proj(*d, p, P, w, W, io.h_out, B, 1, st));
      STAT_TRY(h_proj(*d, p, P, w, W, io.h_out, B, 2, ov.side()));
    } else {
      STAT_TRY(step_gates(*d, p, P, w, W, io, t + 1 < maxlen, st));
      STAT_TRY(ov.state_ready());
    }
    STAT_TRY(step_out(*d, p, P, w, W, io, ov.side()));
    PickArgs k;
    memset(&k, 0, sizeof(k));
    k.rows = B; k.V = d->V; k.ldl = w.ldl; k.logits = W + w.logits;
    k.tokens = out_tokens; k.maxlen = maxlen; k.t = t;
    k.lengths = out_lengths; k.scores = out_scores; k.alive = alive; k.tok_prev = tok_prev;
    {
      ProfScope ps(PH_PICK, ov.side());
      STAT_TRY(pick_launch(k, ov.side()));
    }
    STAT_TRY(ov.side_enqueued());
  }
  return ov.join();
}

int stat_decode_beam(const StatDims *d, const void *prepared, void *ws, int k, int maxlen, int64_t *out_tokens,
                     int32_t *out_lengths, float *out_scores, int32_t *out_count, void *stream) {
  STAT_TRY(check_dims(d));
  STAT_TRY(check_device());
  STAT_REQUIRE(prepared && ws && out_tokens && out_lengths && out_scores && out_count, STAT_EINVAL,
               "decode_beam: bad argument");
  STAT_REQUIRE(k >= 1 && k <= BEAM_KMAX && maxlen >= 1 && maxlen <= BEAM_LMAX, STAT_EINVAL,
               "decode_beam: need 1 <= k <= %d and 1 <= maxlen <= %d (k=%d maxlen=%d)", BEAM_KMAX, BEAM_LMAX, k, maxlen);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int B = d->B, rows = B * k;
  const Prep p = prep_layout(*d);
  const Ws w = ws_layout(*d, rows);
  const float *P = static_cast<const float *>(prepared);
  float *W = static_cast<float *>(ws);
  int32_t *row_clip = reinterpret_cast<int32_t *>(W + w.row_clip);
  BeamArgs b;
  memset(&b, 0, sizeof(b));
  b.B = B; b.k = k; b.V = d->V; b.maxlen = maxlen;
  b.cand_cost = W + w.cand_cost;
  b.cand_word = reinterpret_cast<int32_t *>(W + w.cand_word);
  b.alive = reinterpret_cast<int32_t *>(W + w.alive);
  b.score = W + w.bscore;
  b.hist = reinterpret_cast<int32_t *>(W + w.hist);
  b.hist_len = reinterpret_cast<int32_t *>(W + w.hist_len);
  b.src_row = reinterpret_cast<int32_t *>(W + w.src_row);
  b.tok_prev = reinterpret_cast<int64_t *>(W + w.tok_prev);
  b.dead_k = reinterpret_cast<int32_t *>(W + w.dead_k);
  b.done = reinterpret_cast<int32_t *>(W + w.bdone);
  b.out_tokens = out_tokens; b.out_lengths = out_lengths; b.out_scores = out_scores; b.out_count = out_count;
  STAT_CUDA_CHECK(cudaMemsetAsync(W + w.counters, 0, sizeof(float) * rows, st));
  STAT_TRY(beam_init_launch(b, W + w.h0c0, W + w.h, W + w.c, d->H, row_clip, st));
  if (cell_enabled_rows(*d, w, rows)) {
    // attention -> cell (gates | barrier | every product of the new state, for the rows as they are) -> logits ->
    // pick -> select; the rows of the next step are gathered by src_row: h, c and the h-products (no second pass
    // over the weights for the re-ordered rows)
    STAT_TRY(cell_begin(*d, p, P, w, W, W + w.h, rows, st));
    for (int t = 0; t < maxlen; ++t) {
      StepIO io;
      memset(&io, 0, sizeof(io));
      io.rows = rows;
      io.row_clip = row_clip;
      io.rows_per_clip = k;
      io.tok_prev = b.tok_prev;
      io.h_in = W + w.h; io.c_in = W + w.c; io.h_out = W + w.hb; io.c_out = W + w.cb;
      io.reverse = serpentine(t);
      STAT_TRY(cell_step(*d, p, P, w, W, io, t + 1 < maxlen, W + w.hq2, nullptr, st));
      STAT_TRY(cell_logits(*d, p, P, w, W, rows, st));
      PickArgs pk;
      memset(&pk, 0, sizeof(pk));
      pk.rows = rows; pk.V = d->V; pk.ldl = w.ldl; pk.logits = W + w.logits;
      pk.beam_k = k; pk.row_alive = b.alive; pk.row_score = b.score;
      pk.cand_cost = W + w.cand_cost; pk.cand_word = reinterpret_cast<int32_t *>(W + w.cand_word);
      {
        ProfScope ps(PH_PICK, st);
        STAT_TRY(pick_launch(pk, st));
        b.t = t;
        // the selection also gathers the state and the h-products of the rows the new slots continue
        const bool more = t + 1 < maxlen;
        b.src_h = W + w.hb; b.src_c = W + w.cb; b.H = d->H;
        b.dst_h = more ? W + w.h : nullptr; b.dst_c = more ? W + w.c : nullptr;
        b.src_q = W + w.hq2; b.dst_q = more ? W + w.hq : nullptr; b.ldq = w.ldq2;
        STAT_TRY(beam_select_launch(b, st));
      }
    }
    return STAT_OK;
  }
  STAT_TRY(h_proj(*d, p, P, w, W, W + w.h, rows, 1, st));
  for (int t = 0; t < maxlen; ++t) {
    // the k row slots of a clip share its context blocks; slot order changes every step, so the new
    // state is written next to the old one and gathered back in the new order after the selection
    StepIO io;
    memset(&io, 0, sizeof(io));
    io.rows = rows;
    io.row_clip = row_clip;
    io.rows_per_clip = k;
    io.tok_prev = b.tok_prev;
    io.h_in = W + w.h; io.c_in = W + w.c; io.h_out = W + w.hb; io.c_out = W + w.cb;
    io.reverse = serpentine(t);
    STAT_TRY(step_att(*d, p, P, w, W, io, st));
    STAT_TRY(step_gates(*d, p, P, w, W, io, false, st));
    STAT_TRY(step_out(*d, p, P, w, W, io, st));
    PickArgs pk;
    memset(&pk, 0, sizeof(pk));
    pk.rows = rows; pk.V = d->V; pk.ldl = w.ldl; pk.logits = W + w.logits;
    pk.beam_k = k; pk.row_alive = b.alive; pk.row_score = b.score;
    pk.cand_cost = W + w.cand_cost; pk.cand_word = reinterpret_cast<int32_t *>(W + w.cand_word);
    {
      ProfScope ps(PH_PICK, st);
      STAT_TRY(pick_launch(pk, st));
      b.t = t;
      const bool more = t + 1 < maxlen;
      b.src_h = W + w.hb; b.src_c = W + w.cb; b.H = d->H;
      b.dst_h = more ? W + w.h : nullptr; b.dst_c = more ? W + w.c : nullptr;
      STAT_TRY(beam_select_launch(b, st));
    }
    if (t + 1 < maxlen) {
      STAT_TRY(h_proj(*d, p, P, w, W, W + w.h, rows, 1, st));
    }
  }
  return STAT_OK;
}

int stat_step(const StatDims *d, const void *prepared, void *ws, int rows, const int32_t *row_clip,
              const int64_t *x, const float *h_in, const float *c_in, float *out_probs, float *out_h,
              float *out_c, void *stream) {
  STAT_TRY(check_dims(d));
  STAT_TRY(check_device());
  STAT_REQUIRE(prepared && ws && x && h_in && c_in && out_probs && out_h && out_c && rows >= 1, STAT_EINVAL,
               "step: bad argument");
  STAT_REQUIRE(row_clip != nullptr || rows <= d->B, STAT_EINVAL, "step: rows=%d > B=%d needs row_clip", rows, d->B);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const Prep p = prep_layout(*d);
  const Ws w = ws_layout(*d, rows);
  const float *P = static_cast<const float *>(prepared);
  float *W = static_cast<float *>(ws);
  STAT_CUDA_CHECK(cudaMemsetAsync(W + w.counters, 0, sizeof(float) * rows, st));
  StepIO io;
  memset(&io, 0, sizeof(io));
  io.rows = rows;
  io.row_clip = row_clip;
  io.tok_prev = x;
  io.h_in = h_in; io.c_in = c_in; io.h_out = out_h; io.c_out = out_c;
  STAT_TRY(h_proj(*d, p, P, w, W, h_in, rows, 1, st));
  STAT_TRY(step_att(*d, p, P, w, W, io, st));
  STAT_TRY(step_gates(*d, p, P, w, W, io, false, st));
  STAT_TRY(step_out(*d, p, P, w, W, io, st));
  PickArgs k;
  memset(&k, 0, sizeof(k));
  k.rows = rows; k.V = d->V; k.ldl = w.ldl; k.logits = W + w.logits;
  k.probs = out_probs;
  {
    ProfScope ps(PH_PICK, st);
    STAT_TRY(pick_launch(k, st));
  }
  return STAT_OK;
}

int stat_attention(const StatDims *d, const void *prepared, void *ws, int rows, const int32_t *row_clip,
                   void *stream) {
  STAT_TRY(check_dims(d));
  STAT_TRY(check_device());
  STAT_REQUIRE(prepared && ws && rows >= 1, STAT_EINVAL, "attention: bad argument");
  STAT_REQUIRE(row_clip != nullptr || rows <= d->B, STAT_EINVAL, "attention: rows=%d > B=%d needs row_clip", rows,
               d->B);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const Prep p = prep_layout(*d);
  const Ws w = ws_layout(*d, rows);
  AttArgs a = att_args(*d, p, static_cast<const float *>(prepared), w, static_cast<float *>(ws), rows, row_clip,
                       nullptr, nullptr);
  ProfScope ps(PH_ATT, st);
  return att_launch(w, a, st);
}

int stat_gemm(const float *A, int lda, const float *Bt, int ldb, float *C, int ldc, int M, int N, int K,
              const float *bias, float alpha, float post, int act, int swap, void *stream) {
  STAT_TRY(check_device());
  STAT_REQUIRE(A && Bt && C, STAT_EINVAL, "gemm: NULL operand");
  GemmArgs g;
  memset(&g, 0, sizeof(g));
  g.K = K;
  g.nseg = 1;
  g.seg[0] = GemmSeg{C, ldc, bias, nullptr, 0, alpha, post, act, 0, N};
  if (swap) {
    g.P = Bt; g.ldp = ldb; g.NP = N;
    g.Q = A; g.ldq = lda; g.NQ = M;
    g.feat_on_p = 1;
  } else {
    g.P = A; g.ldp = lda; g.NP = M;
    g.Q = Bt; g.ldq = ldb; g.NQ = N;
    g.feat_on_p = 0;
  }
  return gemm_launch(g, static_cast<cudaStream_t>(stream));
}

}  // extern "C"
