// Per-phase kernels of the fused decode step (one tile of fused_tile.cuh per CTA) and the small vocabulary
// combine that follows the logits tiles.  A decode step on this path is
//     attention -> B (gates fused) -> C (queries + readout activation)        dependent chain: 3 launches
//     logits (partial arg-max / log-sum-exp fused) -> combine                  beside the next attention
// instead of attention -> ctx_proj -> gates -> h_proj and readout -> logits -> pick (7 launches, 4 dependent).
//   B:  [ctx | h_].[Wc ; U] on gate-interleaved rows (one product instead of two k-split ones, :437-439), epilogue
//       = S10-S13 (the gates kernel is gone); the ctx.ff_logit_ctxglm_W rows ride in the same launch (:691-693)
//   C:  h.[Wdl | Wdg | Wdm | Wdlt | W_sel] (next step's attention queries, :371,389,402,415,433) and
//       z = 0.5 tanh(0.5 h.ff_logit_lstm_W + zadd) (:684-696) in one launch
#include <stdlib.h>
#include <string.h>

#include "fused_tile.cuh"

namespace stat {
namespace fused {

struct SegDev {
  int kind;       // FE_*
  int prow0;      // swap: first weight row of the segment in W
  int nfeat;      // features of the segment
  int nk;         // k-atoms
  int xsel;       // which activation map
};

struct PhaseDev {
  int swap, mp, bq, nacc, nseg, qtiles, ntile0, ks;    // ks: k-slices = CTAs of a cluster that share a tile
  SegDev seg[2];
  unsigned long long pol_w, pol_x;
  long long *trace;     // debug: clock64 stamps of CTA 0 (stat_debug_gemm_trace), or null
  EpiParams e;
};

__global__ void __launch_bounds__(NROLE, 1)
    fused_phase_kernel(const __grid_constant__ CUtensorMap tmW, const __grid_constant__ CUtensorMap tmX0,
                       const __grid_constant__ CUtensorMap tmX1, const PhaseDev a) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  __shared__ uint32_t tmem_slot;
  if (a.trace && blockIdx.x == 0 && threadIdx.x == 0) a.trace[0] = clock64();
  Cta c;
  cta_setup(c, smem_raw, &tmem_slot);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int ks = a.ks;
  const int rank = ks > 1 ? static_cast<int>(cluster_ctarank()) : 0;
  const int idx = static_cast<int>(blockIdx.x) / ks;
  const int si = (a.nseg > 1 && idx >= a.ntile0) ? 1 : 0;
  const int local = idx - (si ? a.ntile0 : 0);
  const int kind = si ? a.seg[1].kind : a.seg[0].kind;
  const int prow0 = si ? a.seg[1].prow0 : a.seg[0].prow0;
  const int nfeat = si ? a.seg[1].nfeat : a.seg[0].nfeat;
  const int nk_all = si ? a.seg[1].nk : a.seg[0].nk;
  const int xsel = si ? a.seg[1].xsel : a.seg[0].xsel;
  // this CTA's K slice
  const int kb0 = (rank * nk_all) / ks, kb1 = ((rank + 1) * nk_all) / ks;
  const int nk = kb1 - kb0;
  const CUtensorMap *tmX = xsel ? &tmX1 : &tmX0;
  int f0, q0, prow, qrow;
  const CUtensorMap *tmP, *tmQ;
  unsigned long long pol_p, pol_q;
  if (a.swap) {
    const int pt = local / a.qtiles, qt = local - pt * a.qtiles;
    f0 = pt * a.mp; q0 = qt * a.bq;
    prow = prow0 + f0; qrow = q0;
    tmP = &tmW; tmQ = tmX; pol_p = a.pol_w; pol_q = a.pol_x;
  } else {
    f0 = 0; q0 = local * a.bq;
    prow = 0; qrow = q0;
    tmP = tmX; tmQ = &tmW; pol_p = a.pol_x; pol_q = a.pol_w;
  }
  // everything above overlaps the previous kernel's tail; its results are read from here on
  pdl_wait();
  pdl_trigger();
  long long *trace = (a.trace && blockIdx.x == 0) ? a.trace : nullptr;
  if (trace && threadIdx.x == 0) trace[1] = clock64();
  Ring r = {0, 0};
  const Geo g = make_geo(a.mp, a.bq);
  const int nacc = nk < NISSUE ? nk : NISSUE;
  EpiParams e = a.e;
  if (kind == FE_PICK) e.part0 = 2 * local;
  if (warp == 0) {
    if (lane == 0) produce(c, r, tmP, prow, tmQ, qrow, kb0, kb1, g, pol_p, pol_q, trace);
    __syncwarp();
  } else if (warp == 1 || warp >= 10) {
    uint32_t kt = 0, n = 0;
    issue(c, r, nk, a.bq, g, warp == 1 ? 0 : warp - 9, kt, n, true, trace);
  } else {
    uint32_t kc = 0;
    split(c, r, nk, a.mp, a.bq, g, kc, warp, lane, trace);
    mbar_wait(c.bar_acc, 0);
    tc_fence_after();
    if (trace && threadIdx.x == 64) trace[140] = clock64();
    epilogue_stage(c, e, kind, a.mp, a.bq, nacc, q0, warp, lane);
    if (trace && threadIdx.x == 64) trace[149] = clock64();
  }
  if (ks > 1) {
    // the partial tiles of all K slices are staged: every thread of the cluster passes the barrier, then the
    // items are summed over the ranks through distributed shared memory
    __syncthreads();
    cluster_sync_all();
  }
  if (warp >= 2 && warp < 10) {
    epilogue_items(c, e, kind, a.mp, a.bq, f0, nfeat, q0, ks, rank, nullptr, false);
    if (trace && threadIdx.x == 64) trace[154] = clock64();
  }
  if (ks > 1) cluster_sync_all();      // nobody leaves while a partner still reads its shared memory
  cta_teardown(c);
  if (trace && threadIdx.x == 0) trace[142] = clock64();
}

// ---- combine of the per-tile vocabulary partials of one step: one warp per decode row -----------------------
// (max, sum exp, arg-max) partials in ascending word order -> row maximum, lowest arg-max among equal maxima,
// sum exp; then the greedy bookkeeping (model_attention.py:905-973 with k = 1) or the teacher-forced term
// mask * log(p[x] + 1e-8) (:712-715), exactly as pick_kernel does it.
__device__ __forceinline__ void combine_row(const PickArgs &a, const float4 *part, int npart, const float *tgt, int row,
                                            int lane) {
  float m = -INFINITY, s = 0.f;
  int bi = 0x7fffffff;
  for (int p = lane; p < npart; p += 32) {
    const float4 v = part[static_cast<size_t>(row) * npart + p];
    const int vi = __float_as_int(v.z);
    if (v.x > m) {
      s = s * expf(m - v.x) + v.y;
      m = v.x;
      bi = vi;
    } else if (v.x > -INFINITY) {
      s += v.y * expf(v.x - m);
      if (v.x == m && vi < bi) bi = vi;
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float om = __shfl_xor_sync(0xffffffffu, m, o);
    const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
    const float os = __shfl_xor_sync(0xffffffffu, s, o);
    const float mn = fmaxf(m, om);
    const float w1 = (m == -INFINITY) ? 0.f : expf(m - mn);
    const float w2 = (om == -INFINITY) ? 0.f : expf(om - mn);
    s = s * w1 + os * w2;
    if (om > m || (om == m && oi < bi)) bi = oi;
    m = mn;
  }
  if (lane == 0) {
    const int tok = bi;
    if (a.tokens) {
      const bool live = a.alive[row] != 0;
      a.tokens[static_cast<size_t>(row) * a.maxlen + a.t] = live ? tok : -1;
      if (live) {
        a.scores[row] += logf(s);                  // -log p(argmax) = log sum exp(l - max)
        a.lengths[row] = a.t + 1;
        a.alive[row] = tok != 0;
        a.tok_prev[row] = tok;
      }
    }
    if (a.x_t) {
      const float p = expf(tgt[row] - m) / s;
      {
        // compensated accumulation: the sum of L terms of magnitude ~10 would otherwise lose ~L/2 ulp of the sum
        const float term = a.mask_t[row] * logf(p + 1e-8f);      // model_attention.py:712-715
        if (a.logprob_comp) {
          const float y = term - a.logprob_comp[row];
          const float sum = a.logprob[row];
          const float tt = sum + y;
          a.logprob_comp[row] = (tt - sum) - y;
          a.logprob[row] = tt;
        } else {
          a.logprob[row] += term;
        }
      }
    }
  }
}

__global__ void __launch_bounds__(256) pick_combine_kernel(const PickArgs a, const float4 *part, int npart,
                                                           const float *tgt) {
  pdl_wait();
  pdl_trigger();
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (row < a.rows) combine_row(a, part, npart, tgt, row, threadIdx.x & 31);
}

}  // namespace fused

int pick_combine_launch(const PickArgs &a, const float *part, int npart, const float *tgt, cudaStream_t stream) {
  return launch_pdl(fused::pick_combine_kernel, dim3((a.rows + 7) / 8), dim3(256), 0, stream, a,
                    reinterpret_cast<const float4 *>(part), npart, tgt);
}

int fused_phase_launch(const FusedPhase &p, cudaStream_t stream) {
  using namespace fused;
  STAT_REQUIRE(p.nseg == 1 || p.nseg == 2, STAT_EINVAL, "fused phase: nseg must be 1 or 2");
  STAT_REQUIRE(p.rows >= 1 && (p.swap || p.rows <= BP), STAT_EINVAL, "fused phase: bad row count %d", p.rows);
  PhaseDev d;
  memset(&d, 0, sizeof(d));
  d.swap = p.swap;
  // swap tiles: 128 features x 32 decode rows, K split over the CTAs of a cluster; normal tiles: all rows x 128
  // vocabulary words
  d.mp = BP;
  d.bq = p.swap ? 32 : 128;
  {   // accumulators the products of a k-atom rotate over (debug knob STAT_FUSED_NACC: 1, 2, 4, 8)
    static int nacc_env = -1;
    if (nacc_env < 0) {
      const char *e = getenv("STAT_FUSED_NACC");
      nacc_env = e ? atoi(e) : 1;
      if (nacc_env != 1 && nacc_env != 2 && nacc_env != 4 && nacc_env != 8) nacc_env = 1;
    }
    d.nacc = nacc_env;
    while (d.nacc * d.bq > 256) d.nacc >>= 1;
  }
  d.nseg = p.nseg;
  d.qtiles = p.swap ? (p.rows + d.bq - 1) / d.bq : 1;
  int tiles[2] = {0, 0};
  for (int i = 0; i < p.nseg; ++i) {
    const FusedSegment &s = p.seg[i];
    STAT_REQUIRE(s.nfeat >= 1 && s.K >= 1, STAT_EINVAL, "fused phase: empty segment");
    d.seg[i].kind = s.kind;
    d.seg[i].prow0 = s.wrow0;
    d.seg[i].nfeat = s.nfeat;
    d.seg[i].nk = (s.K + BK - 1) / BK;
    d.seg[i].xsel = s.xsel;
    tiles[i] = p.swap ? ((s.nfeat + d.mp - 1) / d.mp) * d.qtiles : (s.nfeat + d.bq - 1) / d.bq;
  }
  d.ntile0 = tiles[0];
  // k-split: one tcgen05.mma.kind::tf32 covers K = 8 and takes ~55-65 cycles whatever N <= 128 is, so a K = 512 tile
  // costs ~5.5 us in one CTA: spread the K range of every tile over the CTAs of a cluster (as many as keep the grid
  // inside one wave) and sum the partial tiles through distributed shared memory in the epilogue
  d.ks = 1;
  if (p.swap) {
    int nk_min = d.seg[0].nk;
    for (int i = 1; i < p.nseg; ++i) nk_min = d.seg[i].nk < nk_min ? d.seg[i].nk : nk_min;
    static int ks_env = -1;
    if (ks_env < 0) {
      const char *e = getenv("STAT_FUSED_KSPLIT");     // 0 = automatic, n = at most n
      ks_env = e ? atoi(e) : 0;
    }
    const int total = tiles[0] + tiles[1];
    int ks = 148 / total;
    if (ks > 8) ks = 8;
    if (ks_env > 0 && ks > ks_env) ks = ks_env;
    if (ks > nk_min) ks = nk_min;
    if (ks < 1) ks = 1;
    d.ks = ks;
  }
  // L2 priorities: the context blocks the attention kernel re-reads every step own the evict_last set (they do not
  // all fit anyway); the step's weights are streamed evict_first like the separate GEMMs do, the small activation
  // rows every CTA re-reads are kept.  STAT_FUSED_L2=last keeps the weights instead (measurement knob).
  static int wl = -1;
  if (wl < 0) {
    const char *e = getenv("STAT_FUSED_L2");
    wl = (e && !strcmp(e, "last")) ? 1 : 0;
  }
  d.pol_w = wl ? L2_EVICT_LAST : L2_EVICT_FIRST;
  d.pol_x = L2_EVICT_LAST;
  d.e = p.e;
  d.e.rows = p.rows;
  d.trace = gemm_get_trace();
  if (d.trace) {   // debug: STAT_TRACE_KIND selects which launches stamp the trace buffer (FE_* of the first segment)
    const char *tk = getenv("STAT_TRACE_KIND");
    if (tk && atoi(tk) != p.seg[0].kind) d.trace = nullptr;
  }
  CUtensorMap tmW, tmX0, tmX1;
  STAT_TRY(make_tensor_map(&tmW, p.W, p.wrows, p.wK, p.ldw, p.swap ? d.mp : d.bq));
  const int xbox = p.swap ? d.bq : BP;
  STAT_TRY(make_tensor_map(&tmX0, p.X[0], p.rows, p.xK[0], p.ldx[0], xbox));
  if (p.X[1]) STAT_TRY(make_tensor_map(&tmX1, p.X[1], p.rows, p.xK[1], p.ldx[1], xbox));
  else tmX1 = tmX0;
  static size_t smem_set[STAT_MAX_DEV] = {};
  STAT_TRY(ensure_dyn_smem(fused_phase_kernel, SMEM_BYTES, smem_set));
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((tiles[0] + tiles[1]) * d.ks);
  cfg.blockDim = dim3(NROLE);
  cfg.dynamicSmemBytes = SMEM_BYTES;
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  cfg.attrs = attr;
  int na = 0;
  if (d.ks > 1) {
    attr[na].id = cudaLaunchAttributeClusterDimension;
    attr[na].val.clusterDim.x = d.ks;
    attr[na].val.clusterDim.y = 1;
    attr[na].val.clusterDim.z = 1;
    ++na;
  }
  na += pdl_attr(attr + na);
  cfg.numAttrs = na;
  STAT_CUDA_CHECK(cudaLaunchKernelEx(&cfg, fused_phase_kernel, tmW, tmX0, tmX1, d));
  note_launch();
  return STAT_OK;
}

bool fused_supported(int H, int E) { return (H % 4) == 0 && (E % 4) == 0; }

}  // namespace stat
