// att_group_kernel: the four soft-attentions of one decode step (model_attention.py:370-435,
// SURVEY App. A S1-S9) -- the HBM/L2-bound kernel of the path.
//
// Work split: one thread-block CLUSTER per decode row; the cs CTAs of a cluster (cs = 1, 2, 4
// or 8, chosen so that rows*cs fills the SMs once) take contiguous slices of the row's T frames.
// Inside a CTA, G (<= 4) independent GROUPS of four warps each own whole frames (frame j of the
// slice goes to group j mod G): a group never synchronises with another group while it streams,
// so the latencies of one frame (shuffle reductions, group barriers, bulk-copy waits) are hidden
// by the other groups' frames.
//
// Per frame a group receives two bulk-copy chunks in its private shared-memory slots
//   P = pctxl[t] (R*H) | pctxg[t] | pctxm[t]   -> scores         (S1, S4, S5)
//   V = ctxl0[t] (R*H) | qctxl[t] (R*H)        -> weighted sums  (S3, S6-S7)
// issued by the group's own thread 0 (cp.async.bulk + mbarrier complete_tx; L2 priorities: qctxl is
// streamed evict_first, the other blocks evict_last -- they are re-read every step and together with
// the weights of the step do not all fit in the L2).  The group barrier that follows the last read
// of a slot is what frees it, so the copy of P(j+1) flies during the second half of frame j and
// V(j+1) during the first half of frame j+1; ctxg0[t] / ctxm0[t] (2 x H floats) are read straight from
// global memory at the top of phase C.  Every byte of the seven blocks is read once per step.
// Groups 2 and 3 request their first frame when the first chunk of groups 0 / 1 has landed, so the
// groups work out of phase rather than in lock-step.  Each thread owns a float4 of columns:
//   A. sum_h tanh(pctxl + h.Wdl) * Ul for the R regions and the g / m scores (four tanh share one
//      reciprocal); one butterfly reduction for all R (+2) values, group barrier, soft-max over R
//   C. alpha-weighted sums cL = sum_r a_r ctxl0_r (S3) and pLT = sum_r a_r qctxl_r + h.Wdlt + blt
//      (S6: the :416 GEMM folded by linearity), the lt score, group barrier            (S7)
//   D. fold the frame into the running (max, sum, weighted vector) states of the g / m / lt
//      temporal soft-maxes; the weighted vectors live in registers.
// At the end the G group states are merged in shared memory; the other CTAs of the cluster push
// their state from registers into landing pads of rank 0 (st.async + mbarrier complete_tx: no
// fence, no cluster barrier on the way out) and exit; rank 0 waits on its mbarrier, merges in rank
// order, applies the selector gate and writes ctx (S8, S9).  No global scratch, no atomics; the
// merge order is fixed, so results are bit-reproducible.
// The kernel calls griddepcontrol.wait before it reads the h-projections, so it may be launched with
// programmatic stream serialization (STAT_PDL_ATT=1); by default it is not, see launch().
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "kernels.cuh"
#include "stat_common.cuh"

namespace stat {
namespace {

// L2 eviction priorities of the four bulk-copy streams of a frame
struct AttPolicies {
  uint64_t pl, gm, cl, q;      // pctxl | pctxg, pctxm | ctxl0 | qctxl
};

constexpr int GMAX = 4;                   // groups per CTA
constexpr int GT = 128;                   // threads per group
constexpr int NTHREADS = GMAX * GT;       // 512
constexpr int RMAX = 16;

__device__ __forceinline__ uint32_t smem_u32(const void *p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive_local(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
// fault hunting (stat_debug_trap_log): a host-mapped word that survives the fault receives the site of the trap
__device__ volatile int *g_trap_log = nullptr;
__device__ __forceinline__ void trap_at(int site) {
  if (g_trap_log) {
    g_trap_log[0] = site;
    g_trap_log[1] = static_cast<int>(blockIdx.x);
    g_trap_log[2] = static_cast<int>(threadIdx.x);
    g_trap_log[3] = static_cast<int>(blockIdx.y) * 65536 + static_cast<int>(blockIdx.z);
    __threadfence_system();
  }
  __trap();
}
#define mbar_wait(bar, parity) mbar_wait_((bar), (parity), 2000000 + __LINE__)
__device__ __forceinline__ void mbar_wait_(uint32_t bar, uint32_t parity, int site) {
  uint32_t ok;
  uint32_t spins = 0;
  do {
    if (++spins > (1u << 26)) trap_at(site);
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
  } while (!ok);
}
// bulk copy global -> shared with an L2 eviction-priority policy
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar,
                                         uint64_t policy) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::
          "r"(dst),
      "l"(src), "r"(bytes), "r"(bar), "l"(policy)
      : "memory");
}
__device__ __forceinline__ void group_sync(int g) { asm volatile("bar.sync %0, 128;" ::"r"(g + 1) : "memory"); }

// e^x through the SFU (rel. error 2^-22): soft-max numerators
__device__ __forceinline__ float exp_fast(float x) { return ex2_approx(x * 1.4426950408889634f); }

// acc += u * tanh(x + s): with C = 2 log2(e), sc = s*C and m2u = -2u precomputed,
// u*tanh(x+s) = u - 2u / (1 + 2^(x*C + sc)); the "+u" terms are pre-summed into acc.
// Four of them with ONE reciprocal: sum_i m2u_i / d_i = (sum_i m2u_i prod_{j != i} d_j) / prod_j d_j with
// d_i = 1 + 2^(arg_i).  The exponent arguments are clamped to 30 (tanh is 1 to fp32 precision far
// below that), so the product of the four denominators stays below 2^121.  5 MUFU per 4 tanh.
__device__ __forceinline__ float tanh_acc4(const float4 x, const float4 sc, const float4 m2u, float acc) {
  // packed fp32 (FFMA2 / FADD2 / FMUL2: one issue slot for two lanes of work -- the kernel is issue-bound when the
  // blocks hit L2): a = (d0, d1), b = (d2, d3);  sum_i m_i / d_i = N.x / D.x + N.y / D.y  with
  // D = a * b,  N = (m0, m1) * b + (m2, m3) * a
  constexpr float C2 = 2.885390081777927f;
  const float2 c2 = make_float2(C2, C2), one = make_float2(1.0f, 1.0f);
  float2 a = __ffma2_rn(make_float2(x.x, x.y), c2, make_float2(sc.x, sc.y));
  float2 b = __ffma2_rn(make_float2(x.z, x.w), c2, make_float2(sc.z, sc.w));
  a.x = ex2_approx(fminf(a.x, 30.0f)); a.y = ex2_approx(fminf(a.y, 30.0f));
  b.x = ex2_approx(fminf(b.x, 30.0f)); b.y = ex2_approx(fminf(b.y, 30.0f));
  a = __fadd2_rn(a, one);
  b = __fadd2_rn(b, one);
  const float2 D = __fmul2_rn(a, b);
  float2 N = __fmul2_rn(make_float2(m2u.x, m2u.y), b);
  N = __ffma2_rn(make_float2(m2u.z, m2u.w), a, N);
  const float num = fmaf(N.y, D.x, N.x * D.y);
  return fmaf(num, rcp_approx(D.x * D.y), acc);
}
__device__ __forceinline__ float4 ld4(const float *p) { return *reinterpret_cast<const float4 *>(p); }
__device__ __forceinline__ float sum4(const float4 v) { return (v.x + v.y) + (v.z + v.w); }
__device__ __forceinline__ void fma4(float4 &acc, float s, const float4 v) {
  const float2 ss = make_float2(s, s);
  const float2 lo = __ffma2_rn(ss, make_float2(v.x, v.y), make_float2(acc.x, acc.y));
  const float2 hi = __ffma2_rn(ss, make_float2(v.z, v.w), make_float2(acc.z, acc.w));
  acc = make_float4(lo.x, lo.y, hi.x, hi.y);
}
__device__ __forceinline__ void scale4(float4 &v, float s) {
  const float2 ss = make_float2(s, s);
  const float2 lo = __fmul2_rn(make_float2(v.x, v.y), ss), hi = __fmul2_rn(make_float2(v.z, v.w), ss);
  v = make_float4(lo.x, lo.y, hi.x, hi.y);
}

// Butterfly reduction of N (power of two, <= 16) per-lane values over the warp: halving steps over
// the lane bits 4, 3, ... fold pairs of values, plain xor steps finish.  Afterwards v[0] of lane l
// holds the warp total of value  idx(l) = sum_i bit_{4-i}(l) << i  (i < log2 N).
template <int N>
__device__ __forceinline__ void warp_multi_reduce(float (&v)[N], int lane) {
  int o = 16;
#pragma unroll
  for (int n = N; n > 1; n >>= 1, o >>= 1) {
    const bool up = lane & o;
#pragma unroll
    for (int k = 0; k < n / 2; ++k) {
      const float keep = up ? v[2 * k + 1] : v[2 * k];
      const float send = up ? v[2 * k] : v[2 * k + 1];
      v[k] = keep + __shfl_xor_sync(0xffffffffu, send, o);
    }
  }
#pragma unroll
  for (; o > 0; o >>= 1) v[0] += __shfl_xor_sync(0xffffffffu, v[0], o);
}
template <int N>
__device__ __forceinline__ int multi_reduce_index(int lane) {
  int idx = 0, o = 16, sh = 0;
#pragma unroll
  for (int n = N; n > 1; n >>= 1, o >>= 1, ++sh) idx |= ((lane & o) ? 1 : 0) << sh;
  return idx;
}

__device__ __forceinline__ uint32_t cluster_rank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_arrive() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
}
__device__ __forceinline__ void cluster_wait() {
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// address of `local` (a shared-memory location of this CTA) in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t dsmem_addr(const void *local, uint32_t rank) {
  uint32_t ra;
  asm("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(smem_u32(local)), "r"(rank));
  return ra;
}
// 16 bytes from registers into another CTA's shared memory; the write signals that CTA's mbarrier
// (complete_tx), so no fence or cluster barrier is needed on either side
__device__ __forceinline__ void st_async_v4(uint32_t remote_dst, float x, float y, float z, float w,
                                            uint32_t remote_bar) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.b32 [%0], {%1, %2, %3, %4}, [%5];" ::
                   "r"(remote_dst),
               "r"(__float_as_uint(x)), "r"(__float_as_uint(y)), "r"(__float_as_uint(z)), "r"(__float_as_uint(w)),
               "r"(remote_bar)
               : "memory");
}
__device__ __forceinline__ void mbar_wait_cluster(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  uint32_t spins = 0;
  do {
    if (++spins > (1u << 26)) trap_at(2999999);
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
  } while (!ok);
}

// NV = ceil(H / 512) float4 column chunks per thread (columns 4*gt + 512*j);
// RT = compile-time R (8) or 0 for a runtime R <= 16;  HT = compile-time H (512) or 0.
template <int NV, int RT, int HT>
__global__ void __launch_bounds__(NTHREADS, 1)
    att_group_kernel(const AttArgs a, const int G, const int cs, const AttPolicies pol, const int stagger) {
  extern __shared__ __align__(128) uint8_t smem_raw[];
  constexpr int RU = RT ? RT : RMAX;              // unroll bound of the region loops
  const int H = HT ? HT : a.H, T = a.T;
  const int R = RT ? RT : a.R;
  const int RH = R * H;
  const int p_floats = RH + 2 * H, v_floats = 2 * RH;
  const int slot_floats = p_floats + v_floats;
  float *slots = reinterpret_cast<float *>(smem_raw);
  float *s_red = slots + static_cast<size_t>(G) * slot_floats;   // [G][RMAX + 2][4] per-warp score partials
  float *s_lt = s_red + GMAX * (RMAX + 2) * 4;                    // [G][4]
  float *s_gms = s_lt + GMAX * 4;                                 // [G][8] group (max, sum) x 3
  uint64_t *bars = reinterpret_cast<uint64_t *>(s_gms + GMAX * 8);   // fullP[G], fullV[G], merge
  // landing pads of rank 0 for the partial states of the other CTAs of the cluster:
  // [cs - 1] x ([NV * 512][4] weighted sums of a column (g, m, lt, -) | [8] (max, sum) x 3)
  float *s_pad = reinterpret_cast<float *>(bars + 2 * GMAX + 2);
  constexpr int PAD_FLOATS = NV * 512 * 4 + 8;

  // cluster <-> decode row; CTA rank <-> slice [t0, t0 + nframes) of its T frames
  const int row = static_cast<int>(blockIdx.x) / cs;
  const int rank = cs > 1 ? static_cast<int>(cluster_rank()) : 0;
  const int t0 = (rank * T) / cs;
  const int nframes = ((rank + 1) * T) / cs - t0;
  const int tid = threadIdx.x, lane = tid & 31;
  const int g = tid >> 7, gt = tid & (GT - 1), gw = gt >> 5;
  const bool active = g < G;
  const uint32_t bar_p = smem_u32(bars + g), bar_v = smem_u32(bars + GMAX + g);
  float *slot_p = slots + static_cast<size_t>(active ? g : 0) * slot_floats;
  float *slot_v = slot_p + p_floats;
  const int clip = a.row_clip ? a.row_clip[row] : row;
  // debug trace (stat_debug_gemm_trace): 16 stamps per (CTA, group), written by the group's thread 0
  long long *tr = (a.trace && gt == 0) ? a.trace + (static_cast<size_t>(blockIdx.x) * GMAX + g) * 16 : nullptr;
  if (tr) {
    unsigned long long gtm;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gtm));
    tr[0] = static_cast<long long>(gtm);
    tr[1] = clock64();
  }

  if (tid < 2 * GMAX) mbar_init(smem_u32(bars + tid), 1);
  if (tid == 0) {
    if (cs > 1 && rank == 0) {
      // every other CTA of the cluster sends H x 16 bytes of weighted sums and 32 bytes of (max, sum)
      mbar_init(smem_u32(bars + 2 * GMAX), 1);
      mbar_expect_tx(smem_u32(bars + 2 * GMAX), static_cast<uint32_t>(cs - 1) * (static_cast<uint32_t>(H) * 16u + 32u));
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (cs > 1) cluster_arrive();     // matched by cluster_wait() before the first remote write (end of kernel)

  const uint32_t bytes_rh = static_cast<uint32_t>(RH) * 4u, bytes_h = static_cast<uint32_t>(H) * 4u;
  // odd decode steps walk the slice backwards: what the previous step read last is read first (L2 reuse)
  const int rev = a.reverse;
  auto frame_of = [&](int j) { return t0 + (rev ? nframes - 1 - j : j); };
  auto issue_p = [&](int j) {
    const size_t frame = static_cast<size_t>(clip) * T + frame_of(j);
    const uint32_t dst = smem_u32(slot_p);
    mbar_expect_tx(bar_p, static_cast<uint32_t>(p_floats) * 4u);
    bulk_g2s(dst, a.pctxl + frame * RH, bytes_rh, bar_p, pol.pl);
    bulk_g2s(dst + bytes_rh, a.pctxg + frame * H, bytes_h, bar_p, pol.gm);
    bulk_g2s(dst + bytes_rh + bytes_h, a.pctxm + frame * H, bytes_h, bar_p, pol.gm);
  };
  auto issue_v = [&](int j) {
    const size_t frame = static_cast<size_t>(clip) * T + frame_of(j);
    const uint32_t dst = smem_u32(slot_v);
    mbar_expect_tx(bar_v, static_cast<uint32_t>(v_floats) * 4u);
    bulk_g2s(dst, a.ctxl0 + frame * RH, bytes_rh, bar_v, pol.cl);
    bulk_g2s(dst + bytes_rh, a.qctxl + frame * RH, bytes_rh, bar_v, pol.q);
  };
  // The context blocks were written by the prologue of the batch, long before the kernel this launch
  // programmatically depends on: their first copies start before that kernel has finished.
  // Groups 2 and 3 ask for their first frame only when the first chunk of group 0 / 1 has landed:
  // the first copies of a CTA then arrive in two waves and the groups work out of phase from the
  // start (compute of one pair overlaps the copies of the other) instead of in lock-step.
  // Only chunk P goes out here: chunk V (needed a phase later) is requested below, after the loads of
  // the per-row constants have been issued, which would otherwise queue behind it.
  if (active && gt == 0 && g < nframes) {
    if (stagger && g >= 2) mbar_wait(smem_u32(bars + g - 2), 0);
    issue_p(g);
  }

  // ---- per-row constants of this thread's columns ------------------------------------------
  // tanh constants folded (see tanh_acc4): s -> s*C, u -> -2u; the sums of the u values are the
  // starting values of the score partials
  constexpr float C2 = 2.885390081777927f;
  const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
  float4 ul[NV], ug[NV], um[NV], ult[NV], sl[NV], sg[NV], sm[NV], slt[NV];
  float su_l = 0.f, su_g = 0.f, su_m = 0.f, su_lt = 0.f;
  float beta = 1.0f;
#pragma unroll
  for (int j = 0; j < NV; ++j) {
    const int c = 4 * gt + 512 * j;
    const bool ok = HT || c < H;
    ul[j] = ok ? ld4(a.Ul + c) : z4;
    ug[j] = ok ? ld4(a.Ug + c) : z4;
    um[j] = ok ? ld4(a.Um + c) : z4;
    ult[j] = ok ? ld4(a.Ult + c) : z4;
    sl[j] = z4; sg[j] = z4; sm[j] = z4; slt[j] = z4;
  }
  const float cl = __ldg(a.cl), cg = __ldg(a.cg), cm = __ldg(a.cm), clt = __ldg(a.clt);

  // everything below reads what the previous kernel wrote (the h-projections)
  pdl_wait();
  pdl_trigger();
  {
    // The h-projections arrive as k-slice planes, summed here in plane order.  The loads of four
    // planes are issued together before their values are used (one memory round trip).
    const float *hp = a.hp + static_cast<size_t>(row) * a.ldhp;
    float bsel = 0.f;
    for (int q0 = 0; q0 < a.hp_parts; q0 += 4) {
      float4 x[4][NV][4];
      float xb[4];
#pragma unroll
      for (int qq = 0; qq < 4; ++qq) {
        const bool okq = q0 + qq < a.hp_parts;
        const float *hq = hp + static_cast<size_t>(okq ? q0 + qq : q0) * a.hp_plane;
#pragma unroll
        for (int j = 0; j < NV; ++j) {
          const int c = 4 * gt + 512 * j;
          const bool ok = okq && (HT || c < H);
          x[qq][j][0] = ok ? ld4(hq + a.off_sl + c) : z4;
          x[qq][j][1] = ok ? ld4(hq + a.off_sg + c) : z4;
          x[qq][j][2] = ok ? ld4(hq + a.off_sm + c) : z4;
          x[qq][j][3] = ok ? ld4(hq + a.off_slt + c) : z4;
        }
        xb[qq] = (okq && a.selector) ? hq[a.off_sel] : 0.f;
      }
      if (q0 == 0 && active && gt == 0 && g < nframes) issue_v(g);
      auto add4 = [](float4 &d, const float4 v) { d.x += v.x; d.y += v.y; d.z += v.z; d.w += v.w; };
#pragma unroll
      for (int qq = 0; qq < 4; ++qq) {
#pragma unroll
        for (int j = 0; j < NV; ++j) {
          add4(sl[j], x[qq][j][0]);
          add4(sg[j], x[qq][j][1]);
          add4(sm[j], x[qq][j][2]);
          add4(slt[j], x[qq][j][3]);
        }
        bsel += xb[qq];
      }
    }
    beta = a.selector ? sigmoid_acc(bsel) : 1.0f;
    auto fold = [&](float4 &u, float4 &s, float &su) {
      su += sum4(u);
      u.x *= -2.0f; u.y *= -2.0f; u.z *= -2.0f; u.w *= -2.0f;
      s.x *= C2; s.y *= C2; s.z *= C2; s.w *= C2;
    };
#pragma unroll
    for (int j = 0; j < NV; ++j) {
      fold(ul[j], sl[j], su_l);
      fold(ug[j], sg[j], su_g);
      fold(um[j], sm[j], su_m);
      fold(ult[j], slt[j], su_lt);
    }
  }
  if (tr) tr[2] = clock64();

  float4 acc[3][NV];
  float rm[3], rs[3];
#pragma unroll
  for (int q = 0; q < 3; ++q) {
    rm[q] = -INFINITY;
    rs[q] = 0.f;
#pragma unroll
    for (int j = 0; j < NV; ++j) acc[q][j] = z4;
  }

  // ================================ the frames of this group ================================
  if (active) {
    float *red = s_red + g * (RMAX + 2) * 4;
    uint32_t ph = 0;
    int fi = 0;
    for (int j = g; j < nframes; j += G, ph ^= 1, ++fi) {
      const int t = frame_of(j);
      const float *pL = slot_p, *pG = slot_p + RH, *pM = pG + H;
      const float *cL0 = slot_v, *qL = slot_v + RH;
      // the g / m values of the frame (2 x H floats) come straight from global memory: requested at the
      // top of phase C, used after its barrier -- that keeps 4 KB per group of shared memory free
      const float *G0 = a.ctxg0 + (static_cast<size_t>(clip) * T + t) * H;
      const float *M0 = a.ctxm0 + (static_cast<size_t>(clip) * T + t) * H;

      // ---- A: region scores and the g / m scores of the frame -----------------------------
      mbar_wait(bar_p, ph);
      if (tr && fi < 4) tr[3 + 3 * fi] = clock64();
      {
        float part[RU];
#pragma unroll
        for (int r = 0; r < RU; ++r) {
          part[r] = (RT || r < R) ? su_l : 0.f;
          if (RT || r < R) {
#pragma unroll
            for (int jj = 0; jj < NV; ++jj) {
              const int c = 4 * gt + 512 * jj;
              if (HT || c < H) part[r] = tanh_acc4(ld4(pL + r * H + c), sl[jj], ul[jj], part[r]);
            }
          }
        }
        float gm[2] = {su_g, su_m};
#pragma unroll
        for (int jj = 0; jj < NV; ++jj) {
          const int c = 4 * gt + 512 * jj;
          if (HT || c < H) {
            gm[0] = tanh_acc4(ld4(pG + c), sg[jj], ug[jj], gm[0]);
            gm[1] = tanh_acc4(ld4(pM + c), sm[jj], um[jj], gm[1]);
          }
        }
        warp_multi_reduce<RU>(part, lane);
        warp_multi_reduce<2>(gm, lane);
        constexpr int STEP = 32 / RU;          // lanes per region value after the reduction
        if ((lane & (STEP - 1)) == 0) red[multi_reduce_index<RU>(lane) * 4 + gw] = part[0];
        if ((lane & 15) == 0) red[(RMAX + (lane >> 4)) * 4 + gw] = gm[0];
      }
      group_sync(g);                            // scores complete; every thread is done with slot P
      if (gt == 0 && j + G < nframes) issue_p(j + G);

      // ---- soft-max over the regions ----------------------------------------------------------
      float al[RU];        // un-normalised numerators e_r; inv = 1 / sum
      float inv;
      {
        float sc[RU];
        float mx = -INFINITY;
#pragma unroll
        for (int r = 0; r < RU; ++r) {
          sc[r] = (RT || r < R) ? sum4(ld4(red + r * 4)) + cl : -INFINITY;
          mx = fmaxf(mx, sc[r]);
        }
        // lane r evaluates e_r once for the warp; the numerators are then broadcast
        float mine = -INFINITY;
#pragma unroll
        for (int r = 0; r < RU; ++r) mine = (lane == r) ? sc[r] : mine;
        const float e_mine = (lane < R) ? exp_fast(mine - mx) : 0.f;
        float den = 0.f;
#pragma unroll
        for (int r = 0; r < RU; ++r) {
          al[r] = __shfl_sync(0xffffffffu, e_mine, r);
          den += al[r];
        }
        inv = rcp_approx(den);
        if (a.alpha_l && gt < R) a.alpha_l[(static_cast<size_t>(row) * T + t) * R + gt] = e_mine * inv;
      }
      float sc3[3];
      sc3[0] = sum4(ld4(red + RMAX * 4)) + cg;
      sc3[1] = sum4(ld4(red + (RMAX + 1) * 4)) + cm;

      // ---- C: attended local context, its projection, the lt score --------------------------
      mbar_wait(bar_v, ph);
      if (tr && fi < 4) tr[4 + 3 * fi] = clock64();
      float4 cLv[NV], g0[NV], m0[NV];
#pragma unroll
      for (int jj = 0; jj < NV; ++jj) {
        const int c = 4 * gt + 512 * jj;
        const bool ok = HT || c < H;
        g0[jj] = ok ? __ldg(reinterpret_cast<const float4 *>(G0 + c)) : z4;
        m0[jj] = ok ? __ldg(reinterpret_cast<const float4 *>(M0 + c)) : z4;
      }
      {
        float plt = su_lt;
#pragma unroll
        for (int jj = 0; jj < NV; ++jj) {
          const int c = 4 * gt + 512 * jj;
          float4 c0 = z4, p0 = z4;
          if (HT || c < H) {
#pragma unroll
            for (int r = 0; r < RU; ++r) {
              if (RT || r < R) {
                fma4(c0, al[r], ld4(cL0 + r * H + c));
                fma4(p0, al[r], ld4(qL + r * H + c));
              }
            }
            scale4(c0, inv);
            scale4(p0, inv);
            plt = tanh_acc4(p0, slt[jj], ult[jj], plt);
          }
          cLv[jj] = c0;
        }
        plt = warp_sum(plt);
        if (lane == 0) s_lt[g * 4 + gw] = plt;
      }
      group_sync(g);                            // lt score complete; every thread is done with slot V
      if (gt == 0 && j + G < nframes) issue_v(j + G);
      sc3[2] = sum4(ld4(s_lt + g * 4)) + clt;
      if (a.att_scores && gt == 0) {
        const size_t plane = static_cast<size_t>(a.rows) * T, at = static_cast<size_t>(row) * T + t;
        a.att_scores[at] = sc3[0];
        a.att_scores[plane + at] = sc3[1];
        a.att_scores[2 * plane + at] = sc3[2];
      }

      // ---- D: fold the frame into the three running soft-max states ---------------------------
      // (the scores are uniform over the group, so the rescale branch never diverges)
#pragma unroll
      for (int q = 0; q < 3; ++q) {
        float e = 1.0f;
        if (sc3[q] > rm[q]) {
          const float keep = exp_fast(rm[q] - sc3[q]);      // 0 on the first frame (rm = -inf)
          rm[q] = sc3[q];
          rs[q] *= keep;
#pragma unroll
          for (int jj = 0; jj < NV; ++jj) scale4(acc[q][jj], keep);
        } else {
          e = exp_fast(sc3[q] - rm[q]);
        }
        rs[q] += e;
#pragma unroll
        for (int jj = 0; jj < NV; ++jj) fma4(acc[q][jj], e, q == 0 ? g0[jj] : (q == 1 ? m0[jj] : cLv[jj]));
      }
      if (tr && fi < 4) tr[5 + 3 * fi] = clock64();
    }

    // park the group state in the group's own (drained) V slot
    if (gt == 0) {
#pragma unroll
      for (int q = 0; q < 3; ++q) { s_gms[g * 8 + 2 * q] = rm[q]; s_gms[g * 8 + 2 * q + 1] = rs[q]; }
    }
#pragma unroll
    for (int q = 0; q < 3; ++q)
#pragma unroll
      for (int jj = 0; jj < NV; ++jj) {
        const int c = 4 * gt + 512 * jj;
        if (HT || c < H) *reinterpret_cast<float4 *>(slot_v + q * H + c) = acc[q][jj];
      }
  }
  __syncthreads();

  // ---- merge of the G group states: thread <-> columns tid + 512 k -------------------------------
  constexpr int NCOL = NV;
  float om[3], os[3], ov[3][NCOL];
#pragma unroll
  for (int q = 0; q < 3; ++q) {
    float mx = -INFINITY;
    for (int p = 0; p < G; ++p) mx = fmaxf(mx, s_gms[p * 8 + 2 * q]);
    float den = 0.f;
#pragma unroll
    for (int k = 0; k < NCOL; ++k) ov[q][k] = 0.f;
    for (int p = 0; p < G; ++p) {
      const float mp = s_gms[p * 8 + 2 * q];
      const float w = (mp == -INFINITY) ? 0.f : expf(mp - mx);
      den = fmaf(w, s_gms[p * 8 + 2 * q + 1], den);
      const float *vp = slots + static_cast<size_t>(p) * slot_floats + p_floats + q * H;
#pragma unroll
      for (int k = 0; k < NCOL; ++k) {
        const int col = tid + 512 * k;
        if (col < H) ov[q][k] = fmaf(w, vp[col], ov[q][k]);
      }
    }
    om[q] = mx;
    os[q] = den;
  }
  float *ctx = a.ctx + static_cast<size_t>(row) * (a.ldctx ? a.ldctx : H);
  if (cs == 1) {
#pragma unroll
    for (int k = 0; k < NCOL; ++k) {
      const int col = tid + 512 * k;
      if (col < H) {
        const float v = beta * (ov[0][k] / os[0] + ov[1][k] / os[1] + ov[2][k] / os[2]);
        ctx[col] = v;
        if (a.ctx_t) a.ctx_t[(static_cast<size_t>(row >> 6) * H + col) * 64 + (row & 63)] = v;
      }
    }
    return;
  }
  cluster_wait();           // every CTA of the cluster has initialised its barriers (arrive: top of the kernel)
  if (rank != 0) {
    // push this CTA's state into its landing pad in rank 0 and leave: the data travel from registers
    // and complete rank 0's merge barrier on arrival
    float *pad = s_pad + static_cast<size_t>(rank - 1) * PAD_FLOATS;
    const uint32_t rbar = dsmem_addr(bars + 2 * GMAX, 0);
#pragma unroll
    for (int k = 0; k < NCOL; ++k) {
      const int col = tid + 512 * k;
      if (col < H) st_async_v4(dsmem_addr(pad + 4 * col, 0), ov[0][k], ov[1][k], ov[2][k], 0.f, rbar);
    }
    if (tid == 0) {
      st_async_v4(dsmem_addr(pad + NV * 512 * 4, 0), om[0], os[0], om[1], os[1], rbar);
      st_async_v4(dsmem_addr(pad + NV * 512 * 4 + 4, 0), om[2], os[2], 0.f, 0.f, rbar);
    }
    if (tr) tr[15] = clock64();
    return;
  }
  mbar_wait_cluster(smem_u32(bars + 2 * GMAX), 0);
  {
    // fixed order: this CTA first, then the pads by rank
    float o[NCOL];
#pragma unroll
    for (int k = 0; k < NCOL; ++k) o[k] = 0.f;
#pragma unroll
    for (int q = 0; q < 3; ++q) {
      float mx = om[q];
      for (int p = 0; p + 1 < cs; ++p) mx = fmaxf(mx, s_pad[p * PAD_FLOATS + NV * 512 * 4 + 2 * q]);
      const float w0 = expf(om[q] - mx);
      float den = w0 * os[q];
      float num[NCOL];
#pragma unroll
      for (int k = 0; k < NCOL; ++k) num[k] = w0 * ov[q][k];
      for (int p = 0; p + 1 < cs; ++p) {
        const float *pad = s_pad + static_cast<size_t>(p) * PAD_FLOATS;
        const float w = expf(pad[NV * 512 * 4 + 2 * q] - mx);
        den = fmaf(w, pad[NV * 512 * 4 + 2 * q + 1], den);
#pragma unroll
        for (int k = 0; k < NCOL; ++k) {
          const int col = tid + 512 * k;
          if (col < H) num[k] = fmaf(w, pad[4 * col + q], num[k]);
        }
      }
      const float inv = 1.0f / den;
#pragma unroll
      for (int k = 0; k < NCOL; ++k) o[k] = fmaf(num[k], inv, o[k]);
    }
#pragma unroll
    for (int k = 0; k < NCOL; ++k) {
      const int col = tid + 512 * k;
      if (col < H) {
        ctx[col] = beta * o[k];
        if (a.ctx_t) a.ctx_t[(static_cast<size_t>(row >> 6) * H + col) * 64 + (row & 63)] = beta * o[k];
      }
    }
  }
  if (tr) tr[15] = clock64();
}

// ---------------------------------------------------------------------------------------------------------
// att_clip_kernel: the same four soft-attentions for the beam search, where the k row slots of a clip attend
// over the SAME context blocks (model_attention.py:786-788, 328-332: the reference broadcasts one clip's
// context to its k live hypotheses).  One cluster per (clip, pass of <= KSH slots); the CTAs of the cluster
// take contiguous slices of the clip's T frames as above, but a frame is brought to shared memory ONCE and
// every group of four warps computes a different ROW against it: group g <-> row slot g of the clip.  The
// frames travel through a ring of NS whole-frame slots filled by a producer warp (bulk copies, one "P" and
// one "V" barrier per slot as above); a group arrives on the slot's `empty` barrier after its last read, the
// producer refills the slot when all groups have.  k-fold fewer bytes than one cluster per row, and
// clips x cs CTAs fill the SMs once where k x clips rows needed a second, nearly empty wave (32 clips, k = 5:
// 160 one-CTA rows on 148 SMs).
// A group keeps the running soft-max states of its row in registers for the whole slice, so there is no
// merge inside the CTA; across the cluster the other ranks push their states into landing pads of rank 0 as
// above -- the pads reuse the drained ring (a cluster barrier separates the two uses).
// ---------------------------------------------------------------------------------------------------------
constexpr int KSH = 5;                         // most row slots (groups) per CTA; shared memory is laid out for KSH

// KT = most groups of a launch (3 or KSH): the register budget of a thread follows from it -- 65 536 / (KT x 128 + 32):
// 80 registers at five groups, 128 at three, where the ten tanh chains of a frame are all in flight.
template <int RT, int HT, int KT>
__global__ void __launch_bounds__(KT * GT + 32, 1)
    att_clip_kernel(const AttArgs a, const int K, const int npass, const int cs, const int NS, const AttPolicies pol) {
  extern __shared__ __align__(128) uint8_t smem_raw[];
  constexpr int RU = RT ? RT : RMAX;
  const int H = HT ? HT : a.H, T = a.T;
  const int R = RT ? RT : a.R;
  const int RH = R * H;
  const int p_floats = RH + 2 * H, v_floats = 2 * RH;
  const int slot_floats = p_floats + v_floats;
  float *ring = reinterpret_cast<float *>(smem_raw);                 // [NS][slot_floats]; later the landing pads
  float *s_red = ring + static_cast<size_t>(NS) * slot_floats;       // [KSH][RMAX + 2][4] per-warp score partials
  float *s_lt = s_red + KSH * (RMAX + 2) * 4;                        // [KSH][4]
  float *s_u = s_lt + KSH * 4;                                       // -2 Ug | -2 Um | -2 Ult  (3 x H)
  uint64_t *bars = reinterpret_cast<uint64_t *>(s_u + 3 * H);        // fullP[NS] fullV[NS] empty[NS] merge
  constexpr int NSMAX = 4;
  const int PAD_FLOATS = H * 4 + 8;          // per (rank, slot): [H][4] weighted sums of a column (g, m, lt, -) | (max, sum) x 3

  const int rpc = a.rows_per_clip;
  const int cl_id = static_cast<int>(blockIdx.x) / cs;
  const int clip = cl_id / npass, slot0 = (cl_id - clip * npass) * K;
  const int nrows = min(K, rpc - slot0);                              // row slots of this cluster
  const int rank = cs > 1 ? static_cast<int>(cluster_rank()) : 0;
  const int t0 = (rank * T) / cs;
  const int nframes = ((rank + 1) * T) / cs - t0;
  const int tid = threadIdx.x, lane = tid & 31;
  const int g = tid >> 7, gt = tid & (GT - 1), gw = gt >> 5;
  const bool producer = tid >= K * GT;
  const bool active = !producer && g < nrows;
  const int row = clip * rpc + slot0 + (active ? g : 0);
  const int c = 4 * gt;
  const bool okc = HT || c < H;

  if (tid < 2 * NS) mbar_init(smem_u32(bars + (tid < NS ? tid : NSMAX + tid - NS)), 1);
  if (tid >= 2 * NS && tid < 3 * NS) mbar_init(smem_u32(bars + 2 * NSMAX + tid - 2 * NS), static_cast<uint32_t>(nrows));
  if (tid == 0) {
    if (cs > 1 && rank == 0) {
      mbar_init(smem_u32(bars + 3 * NSMAX), 1);
      mbar_expect_tx(smem_u32(bars + 3 * NSMAX),
                     static_cast<uint32_t>(cs - 1) * static_cast<uint32_t>(nrows) * (static_cast<uint32_t>(H) * 16u + 32u));
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  // score vectors (parameters: not written by the previous kernel).  Ul stays in registers, the other three are
  // shared by all groups from shared memory, already folded (u -> -2u, see tanh_acc4)
  const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
  float4 ul = z4;
  float su_l = 0.f, su_g = 0.f, su_m = 0.f, su_lt = 0.f;
  if (!producer && okc) {
    ul = ld4(a.Ul + c);
    const float4 ug = ld4(a.Ug + c), um = ld4(a.Um + c), ult = ld4(a.Ult + c);
    su_l = sum4(ul); su_g = sum4(ug); su_m = sum4(um); su_lt = sum4(ult);
    ul.x *= -2.0f; ul.y *= -2.0f; ul.z *= -2.0f; ul.w *= -2.0f;
    if (g == 0) {
      *reinterpret_cast<float4 *>(s_u + c) = make_float4(-2.0f * ug.x, -2.0f * ug.y, -2.0f * ug.z, -2.0f * ug.w);
      *reinterpret_cast<float4 *>(s_u + H + c) = make_float4(-2.0f * um.x, -2.0f * um.y, -2.0f * um.z, -2.0f * um.w);
      *reinterpret_cast<float4 *>(s_u + 2 * H + c) = make_float4(-2.0f * ult.x, -2.0f * ult.y, -2.0f * ult.z, -2.0f * ult.w);
    }
  }
  const float cl = __ldg(a.cl), cg = __ldg(a.cg), cm = __ldg(a.cm), clt = __ldg(a.clt);
  __syncthreads();

  const int rev = a.reverse;
  auto frame_of = [&](int j) { return t0 + (rev ? nframes - 1 - j : j); };

  // ================================ producer warp ================================
  if (producer) {
    if (lane == 0) {
      // the context blocks were written by the prologue of the batch, long before the kernel this launch depends on
      const uint32_t bytes_rh = static_cast<uint32_t>(RH) * 4u, bytes_h = static_cast<uint32_t>(H) * 4u;
      for (int j = 0; j < nframes; ++j) {
        const int s = j % NS, use = j / NS;
        if (use > 0) mbar_wait(smem_u32(bars + 2 * NSMAX + s), static_cast<uint32_t>((use - 1) & 1));
        const size_t frame = static_cast<size_t>(clip) * T + frame_of(j);
        const uint32_t dst = smem_u32(ring + static_cast<size_t>(s) * slot_floats);
        const uint32_t bp = smem_u32(bars + s), bv = smem_u32(bars + NSMAX + s);
        mbar_expect_tx(bp, static_cast<uint32_t>(p_floats) * 4u);
        bulk_g2s(dst, a.pctxl + frame * RH, bytes_rh, bp, pol.pl);
        bulk_g2s(dst + bytes_rh, a.pctxg + frame * H, bytes_h, bp, pol.gm);
        bulk_g2s(dst + bytes_rh + bytes_h, a.pctxm + frame * H, bytes_h, bp, pol.gm);
        mbar_expect_tx(bv, static_cast<uint32_t>(v_floats) * 4u);
        bulk_g2s(dst + static_cast<uint32_t>(p_floats) * 4u, a.ctxl0 + frame * RH, bytes_rh, bv, pol.cl);
        bulk_g2s(dst + static_cast<uint32_t>(p_floats) * 4u + bytes_rh, a.qctxl + frame * RH, bytes_rh, bv, pol.q);
      }
    }
  }

  // ================================ one row per group ================================
  float4 acc[3] = {z4, z4, z4};
  float rm[3] = {-INFINITY, -INFINITY, -INFINITY}, rs[3] = {0.f, 0.f, 0.f};
  float beta = 1.0f;
  // everything below reads what the previous kernel wrote (the h-projections)
  pdl_wait();
  pdl_trigger();
  if (active) {
    constexpr float C2 = 2.885390081777927f;
    float4 sl = z4, sg = z4, sm = z4, slt = z4;
    {
      // the h-projections arrive as k-slice planes, summed in plane order
      const float *hp = a.hp + static_cast<size_t>(row) * a.ldhp;
      float bsel = 0.f;
      auto add4 = [](float4 &d, const float4 v) { d.x += v.x; d.y += v.y; d.z += v.z; d.w += v.w; };
      for (int q = 0; q < a.hp_parts; ++q) {
        const float *hq = hp + static_cast<size_t>(q) * a.hp_plane;
        if (okc) {
          add4(sl, ld4(hq + a.off_sl + c));
          add4(sg, ld4(hq + a.off_sg + c));
          add4(sm, ld4(hq + a.off_sm + c));
          add4(slt, ld4(hq + a.off_slt + c));
        }
        if (a.selector) bsel += hq[a.off_sel];
      }
      beta = a.selector ? sigmoid_acc(bsel) : 1.0f;
      auto fold = [&](float4 &v) { v.x *= C2; v.y *= C2; v.z *= C2; v.w *= C2; };
      fold(sl); fold(sg); fold(sm); fold(slt);
    }
    float *red = s_red + g * (RMAX + 2) * 4;
    for (int j = 0; j < nframes; ++j) {
      const int s = j % NS;
      const uint32_t ph = static_cast<uint32_t>((j / NS) & 1);
      const int t = frame_of(j);
      const float *slot_p = ring + static_cast<size_t>(s) * slot_floats, *slot_v = slot_p + p_floats;
      const float *pL = slot_p, *pG = slot_p + RH, *pM = pG + H;
      const float *cL0 = slot_v, *qL = slot_v + RH;
      const float *G0 = a.ctxg0 + (static_cast<size_t>(clip) * T + t) * H;
      const float *M0 = a.ctxm0 + (static_cast<size_t>(clip) * T + t) * H;

      // ---- A: region scores and the g / m scores of the frame
      mbar_wait(smem_u32(bars + s), ph);
      {
        float part[RU];
#pragma unroll
        for (int r = 0; r < RU; ++r) {
          part[r] = (RT || r < R) ? su_l : 0.f;
          if ((RT || r < R) && okc) part[r] = tanh_acc4(ld4(pL + r * H + c), sl, ul, part[r]);
        }
        float gm[2] = {su_g, su_m};
        if (okc) {
          gm[0] = tanh_acc4(ld4(pG + c), sg, ld4(s_u + c), gm[0]);
          gm[1] = tanh_acc4(ld4(pM + c), sm, ld4(s_u + H + c), gm[1]);
        }
        warp_multi_reduce<RU>(part, lane);
        warp_multi_reduce<2>(gm, lane);
        constexpr int STEP = 32 / RU;
        if ((lane & (STEP - 1)) == 0) red[multi_reduce_index<RU>(lane) * 4 + gw] = part[0];
        if ((lane & 15) == 0) red[(RMAX + (lane >> 4)) * 4 + gw] = gm[0];
      }
      group_sync(g);                            // scores complete

      // ---- soft-max over the regions
      float al[RU];
      float inv;
      {
        float sc[RU];
        float mx = -INFINITY;
#pragma unroll
        for (int r = 0; r < RU; ++r) {
          sc[r] = (RT || r < R) ? sum4(ld4(red + r * 4)) + cl : -INFINITY;
          mx = fmaxf(mx, sc[r]);
        }
        float mine = -INFINITY;
#pragma unroll
        for (int r = 0; r < RU; ++r) mine = (lane == r) ? sc[r] : mine;
        const float e_mine = (lane < R) ? exp_fast(mine - mx) : 0.f;
        float den = 0.f;
#pragma unroll
        for (int r = 0; r < RU; ++r) {
          al[r] = __shfl_sync(0xffffffffu, e_mine, r);
          den += al[r];
        }
        inv = rcp_approx(den);
        if (a.alpha_l && gt < R) a.alpha_l[(static_cast<size_t>(row) * T + t) * R + gt] = e_mine * inv;
      }
      float sc3[3];
      sc3[0] = sum4(ld4(red + RMAX * 4)) + cg;
      sc3[1] = sum4(ld4(red + (RMAX + 1) * 4)) + cm;

      // ---- C: attended local context, its projection, the lt score
      mbar_wait(smem_u32(bars + NSMAX + s), ph);
      const float4 g0 = okc ? __ldg(reinterpret_cast<const float4 *>(G0 + c)) : z4;
      const float4 m0 = okc ? __ldg(reinterpret_cast<const float4 *>(M0 + c)) : z4;
      float4 cLv = z4;
      {
        float plt = su_lt;
        if (okc) {
          float4 p0 = z4;
#pragma unroll
          for (int r = 0; r < RU; ++r) {
            if (RT || r < R) {
              fma4(cLv, al[r], ld4(cL0 + r * H + c));
              fma4(p0, al[r], ld4(qL + r * H + c));
            }
          }
          scale4(cLv, inv);
          scale4(p0, inv);
          plt = tanh_acc4(p0, slt, ld4(s_u + 2 * H + c), plt);
        }
        plt = warp_sum(plt);
        if (lane == 0) s_lt[g * 4 + gw] = plt;
      }
      group_sync(g);                            // lt score complete; the group is done with the slot
      if (gt == 0) mbar_arrive_local(smem_u32(bars + 2 * NSMAX + s));
      sc3[2] = sum4(ld4(s_lt + g * 4)) + clt;
      if (a.att_scores && gt == 0) {
        const size_t plane = static_cast<size_t>(a.rows) * T, at = static_cast<size_t>(row) * T + t;
        a.att_scores[at] = sc3[0];
        a.att_scores[plane + at] = sc3[1];
        a.att_scores[2 * plane + at] = sc3[2];
      }

      // ---- D: fold the frame into the three running soft-max states
#pragma unroll
      for (int q = 0; q < 3; ++q) {
        float e = 1.0f;
        if (sc3[q] > rm[q]) {
          const float keep = exp_fast(rm[q] - sc3[q]);      // 0 on the first frame (rm = -inf)
          rm[q] = sc3[q];
          rs[q] *= keep;
          scale4(acc[q], keep);
        } else {
          e = exp_fast(sc3[q] - rm[q]);
        }
        rs[q] += e;
        fma4(acc[q], e, q == 0 ? g0 : (q == 1 ? m0 : cLv));
      }
    }
  }

  float *ctx = a.ctx + static_cast<size_t>(row) * (a.ldctx ? a.ldctx : H);
  auto write_ctx = [&](const float4 v) {
    if (!okc) return;
    *reinterpret_cast<float4 *>(ctx + c) = v;
    if (a.ctx_t) {
      float *ct = a.ctx_t + (static_cast<size_t>(row >> 6) * H + c) * 64 + (row & 63);
      ct[0] = v.x; ct[64] = v.y; ct[128] = v.z; ct[192] = v.w;
    }
  };
  if (cs == 1) {
    if (active) {
      const float i0 = 1.0f / rs[0], i1 = 1.0f / rs[1], i2 = 1.0f / rs[2];
      write_ctx(make_float4(beta * (acc[0].x * i0 + acc[1].x * i1 + acc[2].x * i2),
                            beta * (acc[0].y * i0 + acc[1].y * i1 + acc[2].y * i2),
                            beta * (acc[0].z * i0 + acc[1].z * i1 + acc[2].z * i2),
                            beta * (acc[0].w * i0 + acc[1].w * i1 + acc[2].w * i2)));
    }
    return;
  }
  // every CTA of the cluster is done with its ring (and has initialised its barriers): rank 0's ring becomes the pads
  __syncthreads();
  cluster_arrive();
  cluster_wait();
  if (rank != 0) {
    if (active) {
      float *pad = ring + (static_cast<size_t>(rank - 1) * K + g) * PAD_FLOATS;
      const uint32_t rbar = dsmem_addr(bars + 3 * NSMAX, 0);
      if (okc) {
        st_async_v4(dsmem_addr(pad + 4 * c, 0), acc[0].x, acc[1].x, acc[2].x, 0.f, rbar);
        st_async_v4(dsmem_addr(pad + 4 * c + 4, 0), acc[0].y, acc[1].y, acc[2].y, 0.f, rbar);
        st_async_v4(dsmem_addr(pad + 4 * c + 8, 0), acc[0].z, acc[1].z, acc[2].z, 0.f, rbar);
        st_async_v4(dsmem_addr(pad + 4 * c + 12, 0), acc[0].w, acc[1].w, acc[2].w, 0.f, rbar);
      }
      if (gt == 0) {
        st_async_v4(dsmem_addr(pad + 4 * H, 0), rm[0], rs[0], rm[1], rs[1], rbar);
        st_async_v4(dsmem_addr(pad + 4 * H + 4, 0), rm[2], rs[2], 0.f, 0.f, rbar);
      }
    }
    return;
  }
  if (!active) return;
  mbar_wait_cluster(smem_u32(bars + 3 * NSMAX), 0);
  {
    // fixed order: this CTA first, then the pads by rank
    float4 o = z4;
#pragma unroll
    for (int q = 0; q < 3; ++q) {
      float mx = rm[q];
      for (int p = 0; p + 1 < cs; ++p) mx = fmaxf(mx, ring[(static_cast<size_t>(p) * K + g) * PAD_FLOATS + 4 * H + 2 * q]);
      const float w0 = expf(rm[q] - mx);
      float den = w0 * rs[q];
      float4 num = acc[q];
      num.x *= w0; num.y *= w0; num.z *= w0; num.w *= w0;
      for (int p = 0; p + 1 < cs; ++p) {
        const float *pad = ring + (static_cast<size_t>(p) * K + g) * PAD_FLOATS;
        const float w = expf(pad[4 * H + 2 * q] - mx);
        den = fmaf(w, pad[4 * H + 2 * q + 1], den);
        if (okc) {
          num.x = fmaf(w, pad[4 * c + q], num.x);
          num.y = fmaf(w, pad[4 * c + 4 + q], num.y);
          num.z = fmaf(w, pad[4 * c + 8 + q], num.z);
          num.w = fmaf(w, pad[4 * c + 12 + q], num.w);
        }
      }
      const float inv = 1.0f / den;
      o.x = fmaf(num.x, inv, o.x); o.y = fmaf(num.y, inv, o.y); o.z = fmaf(num.z, inv, o.z); o.w = fmaf(num.w, inv, o.w);
    }
    write_ctx(make_float4(beta * o.x, beta * o.y, beta * o.z, beta * o.w));
  }
}

int sm_count() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess ||
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0)
      n = 148;   // B200; also the answer when sizes are queried without a device
  }
  return n;
}

size_t frame_bytes(int R, int H) { return (static_cast<size_t>(3) * R * H + 2 * static_cast<size_t>(H)) * 4; }
size_t pad_bytes(int H) { return (static_cast<size_t>(H <= 512 ? 1 : 2) * 512 * 4 + 8) * 4; }
constexpr size_t SMEM_EXTRA = (GMAX * (RMAX + 2) * 4 + GMAX * 4 + GMAX * 8) * 4 + (2 * GMAX + 2) * 8 + 128;
constexpr size_t SMEM_MAX = 227 * 1024;

// L2 eviction priorities of the context-block copies.  The blocks are re-read on every decode step while the weights
// and activations between two attention launches are streamed once; evict_last lines survive inside the
// persisting-L2 carve-out (stat_set_l2_persist, 79 MB on the B200).  Marking everything evict_last does not work:
// the working set then exceeds what the carve-out holds, all lines have the same priority, the access is cyclic and
// nearly everything misses (measured: 19 % hits with 68 MB marked).  STAT_ATT_KEEP is the set of streams that are
// kept (bit 0 pctxl, bit 1 pctxg / pctxm, bit 2 ctxl0, bit 3 qctxl); the others are streamed evict_first.
//   STAT_ATT_L2 = last (default) | normal | first : the priority of the kept streams
AttPolicies l2_policies() {
  static AttPolicies pol = {0, 0, 0, 0};
  if (pol.pl == 0) {
    const char *e = getenv("STAT_ATT_L2");
    uint64_t keep = 0x14F0000000000000ull;                               // evict_last
    if (e && !strcmp(e, "normal")) keep = 0x1000000000000000ull;
    if (e && !strcmp(e, "first")) keep = 0x12F0000000000000ull;
    const uint64_t stream = 0x12F0000000000000ull;                      // evict_first
    const char *k = getenv("STAT_ATT_KEEP");
    const int mask = k ? atoi(k) : 7;
    pol.pl = (mask & 1) ? keep : stream;
    pol.gm = (mask & 2) ? keep : stream;
    pol.cl = (mask & 4) ? keep : stream;
    pol.q = (mask & 8) ? keep : stream;
  }
  return pol;
}

long long *g_group_trace = nullptr;

template <int NV, int RT, int HT>
int launch(const AttArgs &a, int cs, int G, cudaStream_t stream) {
  const size_t smem = G * frame_bytes(a.R, a.H) + SMEM_EXTRA + (cs - 1) * pad_bytes(a.H);
  static size_t smem_set[STAT_MAX_DEV] = {};
  STAT_TRY(ensure_dyn_smem(att_group_kernel<NV, RT, HT>, smem, smem_set));
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(static_cast<unsigned>(a.rows) * cs);
  cfg.blockDim = dim3(NTHREADS);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = cs;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  // An early-launched attention CTA would sit on a whole SM (all of its shared memory) while it waits
  // for the h-projections, in the way of the readout chain running on the side stream: by default
  // this kernel is launched with full stream serialization (STAT_PDL_ATT=1 to try otherwise).
  static int pdl_att = -1;
  if (pdl_att < 0) {
    const char *e = getenv("STAT_PDL_ATT");
    pdl_att = (e && e[0] == '1') ? 1 : 0;
  }
  cfg.numAttrs = 1 + (pdl_att ? pdl_attr(attr + 1) : 0);
  static int stagger = -1;
  if (stagger < 0) {
    const char *e = getenv("STAT_ATT_STAGGER");
    stagger = (e && e[0] == '0') ? 0 : 1;
  }
  set_launch_label("att_group");
  STAT_CUDA_CHECK(cudaLaunchKernelEx(&cfg, att_group_kernel<NV, RT, HT>, a, G, cs, l2_policies(), stagger));
  note_launch();
  return STAT_OK;
}

}  // namespace

// plan shared by the launcher and the workspace layout: cluster size (CTAs per decode row) and
// groups per CTA; the partial soft-max states never leave the cluster (max_parts = 1)
bool att_group_plan(int rows, int T, int R, int H, int *cluster, int *groups, int *max_parts) {
  if ((H & 3) != 0 || H > 1024 || R > RMAX || R < 1 || rows < 1 || T < 1) return false;
  int cs = 1;
  while (cs < 8 && 2 * cs <= T && static_cast<long long>(rows) * 2 * cs <= sm_count()) cs *= 2;
  {   // measurement knob: STAT_ATT_CS caps the CTAs per decode row (1 leaves SMs free for the readout chain)
    static int cap = -1;
    if (cap < 0) {
      const char *e = getenv("STAT_ATT_CS");
      cap = e ? atoi(e) : 0;
    }
    while (cap > 0 && cs > cap) cs /= 2;
  }
  // rank 0 of a cluster also holds one landing pad per partner
  while (cs > 1 && SMEM_EXTRA + (cs - 1) * pad_bytes(H) + frame_bytes(R, H) > SMEM_MAX) cs /= 2;
  if (SMEM_EXTRA + (cs - 1) * pad_bytes(H) + frame_bytes(R, H) > SMEM_MAX) return false;
  int G = static_cast<int>((SMEM_MAX - SMEM_EXTRA - (cs - 1) * pad_bytes(H)) / frame_bytes(R, H));
  if (G > GMAX) G = GMAX;
  *cluster = cs;
  *groups = G;
  *max_parts = 1;
  return true;
}

// debug: device buffer of >= ctas * 64 int64 (clock stamps per CTA and group), or null
void att_group_set_trace(long long *p) { g_group_trace = p; }
int att_group_set_trap_log(int *dev_ptr) {
  STAT_CUDA_CHECK(cudaMemcpyToSymbol(g_trap_log, &dev_ptr, sizeof(dev_ptr)));
  return STAT_OK;
}

// att_clip_kernel: plan (row slots per cluster, passes per clip, cluster size, ring depth) and launch
namespace {
size_t clip_extra_bytes(int H) {
  return (static_cast<size_t>(KSH) * (RMAX + 2) * 4 + KSH * 4 + 3 * static_cast<size_t>(H)) * 4 + (3 * 4 + 1) * 8 + 128;
}
bool att_clip_plan(int clips, int rpc, int kmax, int T, int R, int H, int *K, int *npass, int *cs, int *NS) {
  if ((H & 3) != 0 || H > 512 || R > RMAX || R < 1 || clips < 1 || rpc < 2 || T < 1) return false;
  const size_t extra = clip_extra_bytes(H);
  if (extra + 2 * frame_bytes(R, H) > SMEM_MAX) return false;
  int ns = static_cast<int>((SMEM_MAX - extra) / frame_bytes(R, H));
  if (ns > 4) ns = 4;
  const int np = (rpc + kmax - 1) / kmax, k = (rpc + np - 1) / np;
  int c = 1;
  while (c < 8 && 2 * c <= T && static_cast<long long>(clips) * np * 2 * c <= sm_count()) c *= 2;
  // the landing pads of rank 0 reuse the ring
  while (c > 1 && static_cast<size_t>(c - 1) * k * (static_cast<size_t>(H) * 4 + 8) * 4 > ns * frame_bytes(R, H)) c /= 2;
  *K = k; *npass = np; *cs = c; *NS = ns;
  return true;
}

template <int RT, int HT, int KT>
int launch_clip(const AttArgs &a, int clips, int K, int npass, int cs, int NS, cudaStream_t stream) {
  const size_t smem = NS * frame_bytes(a.R, a.H) + clip_extra_bytes(a.H);
  static size_t smem_set[STAT_MAX_DEV] = {};
  STAT_TRY(ensure_dyn_smem(att_clip_kernel<RT, HT, KT>, smem, smem_set));
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(static_cast<unsigned>(clips) * npass * cs);
  cfg.blockDim = dim3(K * GT + 32);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = cs;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  set_launch_label("att_clip");
  STAT_CUDA_CHECK(cudaLaunchKernelEx(&cfg, att_clip_kernel<RT, HT, KT>, a, K, npass, cs, NS, l2_policies()));
  note_launch();
  return STAT_OK;
}
}  // namespace

static int g_share = -1;
void att_group_set_share(int on) { g_share = on; }

int att_group_launch(const AttArgs &a_in, cudaStream_t stream) {
  AttArgs a = a_in;
  a.trace = g_group_trace;
  if (a.rows_per_clip >= 2 && a.rows % a.rows_per_clip == 0) {
    // beam search: the row slots of a clip share one pass over its frames (STAT_ATT_SHARE=0: one cluster per row)
    if (g_share < 0) {
      const char *e = getenv("STAT_ATT_SHARE");
      g_share = (e && e[0] == '0') ? 0 : 1;
    }
    const int share = g_share;
    // STAT_ATT_KSH = 3 (default) | 5: most row slots per CTA.  Three slots = 128 registers per thread and, at k = 5, two
    // passes per clip (3 + 2 slots, 64 clusters of 2 CTAs); five slots = one pass at 80 registers.  Measured at 32 clips
    // x k = 5: 39.6 against 45.7 us per launch (one cluster per row: 47.8), 4.22 against 4.41 ms per beam batch.
    static int kmax = -1;
    if (kmax < 0) {
      const char *e = getenv("STAT_ATT_KSH");
      kmax = (e && atoi(e) == KSH) ? KSH : 3;
    }
    int K, np, c, ns;
    if (share && att_clip_plan(a.rows / a.rows_per_clip, a.rows_per_clip, kmax, a.T, a.R, a.H, &K, &np, &c, &ns)) {
      const int clips = a.rows / a.rows_per_clip;
      const bool base = a.H == 512 && a.R == 8;
      if (K <= 3) return base ? launch_clip<8, 512, 3>(a, clips, K, np, c, ns, stream)
                              : launch_clip<0, 0, 3>(a, clips, K, np, c, ns, stream);
      return base ? launch_clip<8, 512, KSH>(a, clips, K, np, c, ns, stream)
                  : launch_clip<0, 0, KSH>(a, clips, K, np, c, ns, stream);
    }
  }
  int cs, G, S;
  STAT_REQUIRE(att_group_plan(a.rows, a.T, a.R, a.H, &cs, &G, &S), STAT_EINVAL,
               "att_group: unsupported shape R=%d H=%d", a.R, a.H);
  const int H = a.H;
  if (H == 512 && a.R == 8) return launch<1, 8, 512>(a, cs, G, stream);   // BASELINE shape
  if (H <= 512) return launch<1, 0, 0>(a, cs, G, stream);
  return launch<2, 0, 0>(a, cs, G, stream);
}

}  // namespace stat
