// Tile engine of the fused decode-step products (sm_100a): one 128 x BQ output tile of
//     D = P . Q^T          P (lane operand, 128 rows) and Q (column operand, BQ rows), both K-major fp32
// with fp32-faithful 3xTF32 accuracy on tcgen05 (the scheme of gemm_tf32x3.cu: the lane operand is split
// into hi / lo in registers and stored to tensor memory, the column operand is split in shared memory;
// lo.hi + hi.lo + hi.hi accumulate in fp32 TMEM), and the epilogues that replace the separate
// gates / readout-activation / vocabulary-reduction kernels of a decode step:
//   FE_STORE  out = acc + bias                                     attention queries, selector logit
//   FE_GATES  S10-S13 of SURVEY App. A on gate-interleaved rows    (model_attention.py:437-457)
//   FE_ZC     readout addend  bz + ctx.Wctx (+ Wemb[tok])           (:689-693)
//   FE_Z      z = post * tanh(alpha * h.Wl + zadd)                  (:684-696)
//   FE_PICK   per-row partial (max, sum exp, arg-max) of the logits (:704-709), target logit
// "swap" tiles put the weights on the 128-lane axis (skinny activations: 32 decode rows per tile),
// "normal" tiles the decode rows (vocabulary product: 128 vocabulary columns per tile).
// The engine is a set of device functions over a per-CTA pipeline (STAGES-deep shared-memory ring, three
// mbarrier arrays, tensor memory) so that the per-phase kernels (one tile per CTA) and the persistent
// decode kernel (a static list of tiles per step and CTA) share it.
#pragma once

#include "kernels.cuh"
#include "stat_common.cuh"
#include "tc_ptx.cuh"

namespace stat {
namespace fused {

using namespace tcx;

constexpr int BP = 128;                    // tile rows (TMEM lanes)
constexpr int BK = 32;                     // fp32 per k-atom row = 128 B = one swizzle row
constexpr int UMMA_K = 8;                  // tf32 MMA K
constexpr int NSPLIT = 256;                // splitter / epilogue threads (warps 2..9)
constexpr int NISSUE = 1;                  // MMA-issuing warps: warp 1 and warps 10 .. 10 + NISSUE - 2 (a second issuer
                                           // was measured: no gain, the mainloop is bound by the latency of a stage round)
constexpr int NROLE = 64 + NSPLIT + 32 * (NISSUE - 1);   // + TMA producer warp (0) + MMA issuer warps
constexpr int MAX_STAGES = 8;
constexpr int RING_BYTES = 192 * 1024;
constexpr int BAR_BYTES = 256;
constexpr int SMEM_BYTES = RING_BYTES + 1024 /*align*/ + BAR_BYTES;
constexpr int TMEM_COLS = 512;

// Ring geometry of a tile shape.  A stage = raw P tile (mp x 128 B) | Q hi (bq x 128 B) | Q lo; its A operand
// (hi 32 columns | lo 32 columns) sits in tensor-memory columns a_base + 64 s, behind the accumulators.  One stage
// round (TMA issue -> arrival -> split -> MMA -> commit -> refill) takes ~2800 cycles (measured), so the mainloop
// runs at nst / 2800 k-atoms per cycle: as many stages as tensor memory (64 columns each) and shared memory allow.
struct Geo {
  int nst;            // stages
  int stage_bytes;
  int q_off;          // Q hi from the stage start
  int qlo_off;        // Q lo from Q hi
  int a_base;         // first tensor-memory column of the A stages
};
__host__ __device__ inline Geo make_geo(int mp, int bq) {
  Geo g;
  g.q_off = mp * BK * 4;
  g.qlo_off = bq * BK * 4;
  g.stage_bytes = g.q_off + 2 * g.qlo_off;
  g.a_base = (NISSUE * bq <= 128) ? 128 : 256;
  int n = (TMEM_COLS - g.a_base) / 64;
  if (n > RING_BYTES / g.stage_bytes) n = RING_BYTES / g.stage_bytes;
  if (n > MAX_STAGES) n = MAX_STAGES;
  g.nst = n;
  return g;
}

using EpiParams = FusedEpi;

struct Cta {
  uint8_t *ring;            // 1024-byte aligned
  uint32_t bar_full, bar_split, bar_empty, bar_acc;
  uint32_t tmem;
};

// ring position of one role: stage index and, per stage, the parity of the phase it is in (a bit per stage, so
// that tiles with different stage counts can follow each other on the same barriers)
struct Ring {
  int s;
  uint32_t bits;
  __device__ __forceinline__ uint32_t ph() const { return (bits >> s) & 1u; }
  __device__ __forceinline__ void next(int nst) {
    bits ^= 1u << s;
    if (++s == nst) s = 0;
  }
};

// ---- set-up / tear-down (all threads of the CTA) ------------------------------------------------------
__device__ __forceinline__ void cta_setup(Cta &c, uint8_t *smem_raw, uint32_t *tmem_slot) {
  uint8_t *smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  c.ring = smem;
  uint64_t *bars = reinterpret_cast<uint64_t *>(smem + RING_BYTES);
  c.bar_full = smem_u32(bars);
  c.bar_split = c.bar_full + 8 * MAX_STAGES;
  c.bar_empty = c.bar_split + 8 * MAX_STAGES;
  c.bar_acc = c.bar_empty + 8 * MAX_STAGES;
  if (threadIdx.x == 0) {
    for (int s = 0; s < MAX_STAGES; ++s) {
      mbar_init(c.bar_full + 8 * s, 1);
      mbar_init(c.bar_split + 8 * s, 4);          // one arrival per warp of the splitter group that owns the k-atom
      mbar_init(c.bar_empty + 8 * s, 1);
    }
    mbar_init(c.bar_acc, NISSUE);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if ((threadIdx.x >> 5) == 1) tc_alloc(smem_u32(tmem_slot), TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  c.tmem = *tmem_slot;
}
__device__ __forceinline__ void cta_teardown(const Cta &c) {
  tc_fence_before();
  __syncthreads();
  if ((threadIdx.x >> 5) == 1) {
    tc_fence_after();
    tc_dealloc(c.tmem, TMEM_COLS);
  }
}

// ---- the three mainloop roles ------------------------------------------------------------------------
// producer: one thread.  k-atoms [kb0, kb1) of the P rows [prow, prow+mp) and the Q rows [qrow, qrow+bq)
// (mp = rows of the P box: 128, or 64 / 32 for half / quarter tiles -- the remaining lanes then hold garbage)
__device__ __forceinline__ void produce(const Cta &c, Ring &r, const CUtensorMap *tmP, int prow, const CUtensorMap *tmQ,
                                        int qrow, int kb0, int kb1, const Geo &g, unsigned long long pol_p,
                                        unsigned long long pol_q, long long *trace = nullptr) {
  for (int kb = kb0; kb < kb1; ++kb) {
    mbar_wait(c.bar_empty + 8 * r.s, r.ph() ^ 1u);
    if (trace && kb - kb0 < 32) trace[100 + kb - kb0] = clock64();
    const uint32_t stage = smem_u32(c.ring + r.s * g.stage_bytes);
    mbar_expect_tx(c.bar_full + 8 * r.s, static_cast<uint32_t>(g.q_off + g.qlo_off));
    tma_load_2d(stage, tmP, kb * BK, prow, c.bar_full + 8 * r.s, pol_p);
    tma_load_2d(stage + g.q_off, tmQ, kb * BK, qrow, c.bar_full + 8 * r.s, pol_q);
    r.next(g.nst);
  }
}

// MMA issuers: NISSUE warps (warp 1 and warps 10..), issuer ii takes the k-atoms kt with kt % NISSUE == ii of a
// tile (kt = k-atoms of the tile so far) and accumulates them in its own accumulator, columns [ii*bq, (ii+1)*bq);
// the epilogue adds the accumulators in order.  One tcgen05.mma.kind::tf32 covers K = 8, so a k-atom is 12
// instructions whatever N is, and a single thread issues them at ~58 cycles each (measured: 700 cycles per
// k-atom at N = 32, 1030 at N = 128): with skinny tiles the issue rate of one thread, not the tensor pipe, bounds
// the mainloop -- hence several issuing threads.
// Every lane computes the (warp-uniform) descriptors so that they live in uniform registers, and the MMA / commit
// instructions run in the lane chosen by elect.sync: inside an `if (lane == 0)` block ptxas wraps each
// tcgen05.mma in an ELECT / R2UR / branch loop (measured 110-170 cycles per MMA).
// `n` = products this issuer has issued for the tile (0 = its first one overwrites its accumulator); `last`: the
// tile is complete after these k-atoms (every issuer commits to bar_acc, which expects NISSUE arrivals).
__device__ __forceinline__ void issue(const Cta &c, Ring &r, int nk, int bq, const Geo &g, int ii, uint32_t &kt,
                                      uint32_t &n, bool last, long long *trace = nullptr) {
  const uint32_t idesc = make_idesc_tf32(BP, bq);
  const uint32_t tmem = __shfl_sync(0xffffffffu, c.tmem, 0);
  const uint32_t ring0 = __shfl_sync(0xffffffffu, smem_u32(c.ring), 0);
  const uint32_t acc = tmem + static_cast<uint32_t>(ii * bq);
  for (int kb = 0; kb < nk; ++kb, ++kt) {
    if (kt % NISSUE == static_cast<uint32_t>(ii)) {
      mbar_wait(c.bar_split + 8 * r.s, r.ph());
      tc_fence_after();
      __syncwarp();
      const bool leader = elect_one();
      if (trace && leader && kb < 32) trace[180 + kb] = clock64();
      const uint32_t stage = ring0 + static_cast<uint32_t>(r.s * g.stage_bytes);
      const uint64_t dQh = make_desc(stage + g.q_off);
      const uint64_t dQl = make_desc(stage + g.q_off + g.qlo_off);
      const uint32_t a_hi = tmem + g.a_base + 64 * r.s;
      if (leader) {
        tc_mma_tf32_katom(acc, a_hi, dQh, dQl, idesc, n ? 1u : 0u);
        tc_commit(c.bar_empty + 8 * r.s);
      }
      n += 12;
      __syncwarp();
    }
    r.next(g.nst);
  }
  if (last) {
    __syncwarp();
    if (elect_one()) tc_commit(c.bar_acc);
    __syncwarp();
  }
}

// splitter: warps 2..9 in two groups of four warps (one warp per TMEM lane quadrant) that take alternate k-atoms,
// so that the latencies of one k-atom's split (shared-memory loads, tensor-memory stores and their wait, the proxy
// fence) overlap the other group's.  `kc` = k-atoms this CTA has split so far (both groups count all of them).
// P: a thread owns tile row prow (= its TMEM lane) and all 32 k-columns, in two halves of 16; the TMA tile is
// SWIZZLE_128B (16-byte chunk c of row r sits at chunk c ^ (r & 7)).  Q: hi in place, lo beside it.
__device__ __forceinline__ void split(const Cta &c, Ring &r, int nk, int mp, int bq, const Geo &g, uint32_t &kc,
                                      int warp, int lane, long long *trace = nullptr) {
  const int grp = (warp - 2) >> 2;
  const int tg = ((warp - 2) & 3) * 32 + lane;       // 0..127 inside the group
  const int prow = (warp & 3) * 32 + lane;
  const int nq4 = bq * (BK / 4);                     // 16-byte chunks of the Q tile
  for (int kb = 0; kb < nk; ++kb, ++kc) {
    // both groups observe every phase of a stage's barrier (a parity wait is only valid one phase behind; with an
    // odd ring depth a group would otherwise skip the other group's phase of a stage -- see gemm_tf32x3.cu)
    mbar_wait(c.bar_full + 8 * r.s, r.ph());
    if ((kc & 1u) == static_cast<uint32_t>(grp)) {
      if (trace && tg == 0 && kb < 32) trace[2 + kb] = clock64();
      uint8_t *stage = c.ring + r.s * g.stage_bytes;
      const uint8_t *rowp = stage + prow * 128;
      float4 *Qh = reinterpret_cast<float4 *>(stage + g.q_off);
      float4 *Ql = reinterpret_cast<float4 *>(stage + g.q_off + g.qlo_off);
      if ((warp & 3) * 32 < mp) {
        const uint32_t ta = c.tmem + (static_cast<uint32_t>((warp & 3) * 32) << 16) + g.a_base + 64 * r.s;
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          float4 xp[4];
#pragma unroll
          for (int i = 0; i < 4; ++i)
            xp[i] = *reinterpret_cast<const float4 *>(rowp + (((4 * half + i) ^ (prow & 7)) << 4));
          uint32_t hi[16], lo[16];
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const float xs[4] = {xp[i].x, xp[i].y, xp[i].z, xp[i].w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const float h = rna_tf32(xs[e]);
              hi[4 * i + e] = __float_as_uint(h);
              lo[4 * i + e] = __float_as_uint(xs[e] - h);   // exact; the tensor core reads its top 19 bits
            }
          }
          tc_st16(ta + 16 * half, hi);
          tc_st16(ta + 32 + 16 * half, lo);
        }
      }
      for (int idx = tg; idx < nq4; idx += 128) {
        const float4 x = Qh[idx];
        float4 h, l;
        h.x = rna_tf32(x.x); h.y = rna_tf32(x.y); h.z = rna_tf32(x.z); h.w = rna_tf32(x.w);
        l.x = x.x - h.x; l.y = x.y - h.y; l.z = x.z - h.z; l.w = x.w - h.w;
        Qh[idx] = h;
        Ql[idx] = l;
      }
      tc_wait_st();
      tc_fence_before();
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(c.bar_split + 8 * r.s);
      if (trace && tg == 0 && kb < 32) trace[36 + kb] = clock64();
    }
    r.next(g.nst);
  }
}

// ---- epilogues (warps 2..9, after bar_acc) --------------------------------------------------------------
// Branch-free activations (ex2.approx + rcp.approx, relative error ~3e-7; the attention kernel's tanh is the same).
__device__ __forceinline__ float sigmoid_f(float x) {
  return rcp_approx(1.0f + ex2_approx(-1.4426950408889634f * x));
}
__device__ __forceinline__ float tanh_f(float x) { return tanh_fast(fminf(fmaxf(x, -15.0f), 15.0f)); }
__device__ __forceinline__ float exp_f(float x) { return ex2_approx(1.4426950408889634f * x); }

__device__ __forceinline__ void epi_sync() { asm volatile("bar.sync 1, 256;" ::: "memory"); }

// v[i] = sum over the accumulators (in order) of columns [col, col + 16) of this thread's lane
__device__ __forceinline__ void acc_ld16(uint32_t trow, int col, int bq, int nacc, uint32_t (&v)[16]) {
  tc_ld16(trow + col, v);
  for (int a = bq; a < nacc * bq; a += bq) {
    uint32_t w[16];
    tc_ld16(trow + a + col, w);
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __float_as_uint(__uint_as_float(v[i]) + __uint_as_float(w[i]));
  }
}

// Swap tiles.  The accumulator tile (lane = feature, column = decode row) is first transposed through shared memory
// into S[row][feature] (the ring is free once bar_acc has completed), then all 256 epilogue threads walk (row,
// feature group) items in SHORT ROLLED loops with coalesced global accesses.  A first version kept everything in
// registers -- 16 rows per thread, fully unrolled -- and ran at the speed of cold instruction fetches: straight-line
// code that every warp executes once per launch (ncu: `no_instructions` the top stall; 17 000 cycles for 32 stores).
constexpr int SPAD = 4;     // S row pitch = mp + SPAD floats (16-byte aligned rows)

__device__ __forceinline__ void epi_stage(const Cta &c, float *S, int mp, int bq, int nacc, int warp, int lane) {
  const int wq = warp & 3, chalf = (warp - 2) >> 2, ch = bq >> 1, SP = mp + SPAD;
  if (wq * 32 < mp) {
    const uint32_t trow = c.tmem + (static_cast<uint32_t>(wq * 32) << 16);
#pragma unroll 1
    for (int cc = chalf * ch; cc < (chalf + 1) * ch; cc += 16) {
      uint32_t v[16];
      acc_ld16(trow, cc, bq, nacc, v);
#pragma unroll
      for (int i = 0; i < 16; ++i) S[(cc + i) * SP + wq * 32 + lane] = __uint_as_float(v[i]);
    }
  }
  epi_sync();
}

// FE_GATES: S10-S13 (model_attention.py:437-457).  Feature j = 4*unit + gate (0 input, 1 forget, 2 output,
// 3 candidate): an item = (row, unit) reads its four pre-activations as one float4.  creg != null: the cell state
// of this thread's items (item k of the thread -> creg[k]) lives in registers across decode steps.
// k-split tiles: the ks CTAs of a cluster each hold the partial tile of their K slice in their own S; an item is
// summed over the ranks in rank order (deterministic) through distributed shared memory, and the items are dealt
// out over the CTAs of the cluster (item it -> CTA (it / NSPLIT) % ks).
__device__ __forceinline__ float4 s_sum4(const float *p, int ks) {
  if (ks == 1) return *reinterpret_cast<const float4 *>(p);
  float4 a = ld_dsmem4(dsmem_addr(p, 0));
  for (int r = 1; r < ks; ++r) {
    const float4 b = ld_dsmem4(dsmem_addr(p, r));
    a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
  }
  return a;
}
__device__ __forceinline__ float s_sum1(const float *p, int ks) {
  if (ks == 1) return *p;
  float a = ld_dsmem(dsmem_addr(p, 0));
  for (int r = 1; r < ks; ++r) a += ld_dsmem(dsmem_addr(p, r));
  return a;
}

__device__ __forceinline__ void epi_gates(const EpiParams &e, const float *S, int mp, int bq, int f0, int nfeat, int q0,
                                          int t, int ks, int rank, float *creg, bool creg_load) {
  const int H = e.H, SP = mp + SPAD, nu = mp >> 2, u0 = f0 >> 2;
  int k = 0;
#pragma unroll 1
  for (int it = t + NSPLIT * rank; it < bq * nu; it += NSPLIT * ks, ++k) {
    const int row = it / nu, ul = it - row * nu;
    const int r = q0 + row, u = u0 + ul;
    if (r >= e.rows || 4 * u >= nfeat) continue;
    const float4 a = s_sum4(S + row * SP + 4 * ul, ks);
    const long long tok = e.tok_prev ? e.tok_prev[r] : -1;
    const float4 ew = __ldg(reinterpret_cast<const float4 *>(e.EWi + static_cast<size_t>(tok >= 0 ? tok : e.V) * 4 * H) + u);
    float4 hu = make_float4(0.f, 0.f, 0.f, 0.f);
    if (e.hu) hu = *reinterpret_cast<const float4 *>(e.hu + static_cast<size_t>(r) * e.ld_hu + 4 * u);
    const size_t at = static_cast<size_t>(r) * H + u;
    const float c_ = (creg && !creg_load) ? creg[k] : e.c_in[at];
    float di = 0.5f, df = 0.5f, dO = 0.5f;
    if (e.dp_gates) {
      const float *dp = e.dp_gates + static_cast<size_t>(r) * 3 * H + u;
      di = dp[0]; df = dp[H]; dO = dp[2 * H];
    }
    const float ig = sigmoid_f(((a.x + hu.x) + ew.x) * di);
    const float fg = sigmoid_f(((a.y + hu.y) + ew.y) * df);
    const float og = sigmoid_f(((a.z + hu.z) + ew.z) * dO);
    const float gg = tanh_f((a.w + hu.w) + ew.w);
    float c = fmaf(fg, c_, ig * gg);
    float m = 1.0f;
    if (e.mask) {
      m = e.mask[r];
      c = m * c + (1.0f - m) * c_;                                         // :454
    }
    float h = og * tanh_f(c);                                              // :456 (the masked c)
    if (e.mask) h = m * h + (1.0f - m) * e.h_in[static_cast<size_t>(r) * e.ld_hin + u];   // :457
    if (creg) creg[k] = c;
    if (e.c_out) e.c_out[at] = c;
    e.h_out[static_cast<size_t>(r) * e.ld_hout + u] = h;
    if (e.h_copy) e.h_copy[at] = h;
    if (e.h_all) e.h_all[at] = h;
    if (e.dp_h) e.hd_out[at] = h * e.dp_h[at];
  }
}

// FE_STORE / FE_ZC / FE_Z: an item = (row, feature); consecutive threads take consecutive features of a row
__device__ __forceinline__ void epi_rows(const EpiParams &e, int kind, const float *S, int mp, int bq, int f0, int nfeat,
                                         int q0, int t, int ks, int rank) {
  const int SP = mp + SPAD;
#pragma unroll 1
  for (int it = t + NSPLIT * rank; it < bq * mp; it += NSPLIT * ks) {
    const int row = it / mp, fl = it - row * mp;
    const int r = q0 + row, j = f0 + fl;
    if (r >= e.rows || j >= nfeat) continue;
    const float acc = s_sum1(S + row * SP + fl, ks);
    if (kind == FE_STORE) {
      e.out[static_cast<size_t>(r) * e.ldo + j] = acc + (e.bias ? __ldg(e.bias + j) : 0.f);
    } else if (kind == FE_ZC) {
      float z = acc + __ldg(e.bz + j);
      if (e.prev2out && e.tok_prev) {
        const long long tok = e.tok_prev[r];
        if (tok >= 0) z += __ldg(e.Wemb + static_cast<size_t>(tok) * e.E + j);
      }
      e.zadd[static_cast<size_t>(r) * e.E + j] = z;
    } else {
      const size_t at = static_cast<size_t>(r) * e.E + j;
      e.z[at] = tanh_f(fmaf(e.z_alpha, acc, e.zadd[at])) * (e.dp_z ? e.dp_z[at] : 0.5f);
    }
  }
}

// normal tiles (FE_PICK): lane = decode row, TMEM column = vocabulary word q0 + col.  Each thread folds its
// columns into a (max, sum exp, arg-max) partial; the first maximum wins ties (ascending words).
struct PickAcc {
  float m, s;
  int bi;
  float tv;        // target logit when it falls into this thread's columns
  int thit;
};
__device__ __forceinline__ void pick_fold16(const EpiParams &e, const uint32_t (&v)[16], int w0, int V, long long tgt_word,
                                            PickAcc &a) {
  // branch-free: chunk maximum first, one rescale of the running sum, then the 16 exponentials
  float x[16];
  float cm = -INFINITY;
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    const int w = w0 + i;
    x[i] = (w < V) ? __uint_as_float(v[i]) + __ldg(e.bv + w) : -INFINITY;
    cm = fmaxf(cm, x[i]);
  }
  if (cm == -INFINITY) return;                       // a chunk entirely beyond the vocabulary
  int ci = 0x7fffffff;
#pragma unroll
  for (int i = 15; i >= 0; --i) ci = (x[i] == cm) ? w0 + i : ci;       // first word that attains the chunk maximum
#pragma unroll
  for (int i = 0; i < 16; ++i)
    if (static_cast<long long>(w0 + i) == tgt_word) { a.tv = x[i]; a.thit = 1; }
  const float mn = fmaxf(a.m, cm);
  float s = a.s * exp_f(a.m - mn);                   // exp(-inf) = 0 on the first chunk
#pragma unroll
  for (int i = 0; i < 16; ++i) s += exp_f(x[i] - mn);                  // words beyond V contribute exp(-inf) = 0
  a.s = s;
  if (cm > a.m) a.bi = ci;                           // strictly greater: the earlier chunk keeps ties
  a.m = mn;
}

// Tile epilogue in two parts (kind / geometry are uniform over the CTA; warps 2..9):
//   epilogue_stage: FE_PICK does everything here (normal tiles: q0 = first vocabulary word, e.part0 = partial slot);
//                   swap tiles transpose their accumulators into S (f0 = first feature of the tile inside its
//                   segment, nfeat = features of the segment, q0 = first decode row)
//   epilogue_items: swap tiles only, after the cluster barrier when the tile is k-split over ks CTAs
__device__ __forceinline__ void epilogue_stage(const Cta &c, const EpiParams &e, int kind, int mp, int bq, int nacc,
                                               int q0, int warp, int lane) {
  if (kind == FE_PICK) {
    const int wq = warp & 3, chalf = (warp - 2) >> 2, ch = bq >> 1;
    const uint32_t trow = c.tmem + (static_cast<uint32_t>(wq * 32) << 16);
    const int r = wq * 32 + lane;                       // decode row (single row tile: rows <= 128)
    PickAcc a;
    a.m = -INFINITY; a.s = 0.f; a.bi = 0x7fffffff; a.tv = 0.f; a.thit = 0;
    const long long tgt_word = (e.x_t && r < e.rows) ? e.x_t[r] : -1;
#pragma unroll 1
    for (int cc = chalf * ch; cc < (chalf + 1) * ch; cc += 16) {
      uint32_t v[16];
      acc_ld16(trow, cc, bq, nacc, v);
      pick_fold16(e, v, q0 + cc, e.V, tgt_word, a);
    }
    if (r < e.rows) {
      float4 *dst = reinterpret_cast<float4 *>(e.part) + static_cast<size_t>(r) * e.npart + e.part0 + chalf;
      *dst = make_float4(a.m, a.s, __int_as_float(a.bi), 0.f);
      if (a.thit) e.tgt[r] = a.tv;
    }
    return;
  }
  epi_stage(c, reinterpret_cast<float *>(c.ring), mp, bq, nacc, warp, lane);
}

__device__ __forceinline__ void epilogue_items(const Cta &c, const EpiParams &e, int kind, int mp, int bq, int f0,
                                               int nfeat, int q0, int ks, int rank, float *creg, bool creg_load) {
  if (kind == FE_PICK) return;
  const float *S = reinterpret_cast<const float *>(c.ring);
  const int t = threadIdx.x - 64;
  if (kind == FE_GATES) epi_gates(e, S, mp, bq, f0, nfeat, q0, t, ks, rank, creg, creg_load);
  else epi_rows(e, kind, S, mp, bq, f0, nfeat, q0, t, ks, rank);
}

}  // namespace fused
}  // namespace stat
