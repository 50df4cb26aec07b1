// att_stream_kernel: the four soft-attentions of one decode step (model_attention.py:370-435,
// SURVEY App. A S1-S9) as a persistent, TMA-fed streaming kernel -- the HBM-bound kernel
// of the path.
//
// Work split: one thread-block CLUSTER per decode row; the cs CTAs of a cluster (cs = 1, 2, 4
// or 8, chosen so that rows*cs fills the SMs once) take contiguous slices of the row's T frames.
// Per frame the producer warp issues seven cp.async.bulk copies (pctxl[t], ctxl0[t], qctxl[t]:
// R*H floats each; pctxg[t], pctxm[t], ctxg0[t], ctxm0[t]: H floats each) into one slot of a
// shared-memory ring guarded by full/empty mbarriers; every byte of the seven context blocks is
// read from HBM/L2 exactly once per step.  The 16 consumer warps then
//   A. score the R regions (warp = region x column half): sum_h tanh(pctxl + h.Wdl) * Ul (S1)
//      and the g / m temporal scores of the frame (column-sliced partials)          (S4, S5)
//   C. softmax over R (S2), alpha-weighted sums cL = sum_r a_r ctxl0_r (S3) and
//      pLT = sum_r a_r qctxl_r + h.Wdlt + blt (S6, the :416 GEMM folded by linearity),
//      then the lt score partial                                                     (S7)
//   D. fold the frame into flash-style running (max, sum, weighted vector) states of
//      the g / m / lt temporal soft-maxes; the weighted vectors live in registers.
// At the end every CTA parks its partial state in its own shared memory, the cluster
// synchronises, and rank 0 merges the cs parts through distributed shared memory, applies
// the selector gate and writes ctx (S8, S9).  No global scratch, no atomics; the merge
// order is fixed, so results are bit-reproducible.
#include <string.h>

#include "kernels.cuh"
#include "stat_common.cuh"

namespace stat {
namespace {

constexpr int NWARPS = 16;                      // consumer warps
constexpr int CONSUMERS = NWARPS * 32;          // 512 consumer threads
constexpr int NTHREADS = CONSUMERS + 32;        // + one producer warp
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
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  uint32_t spins = 0;
  do {
    if (++spins > (1u << 26)) __trap();
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
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
      "l"(src), "r"(bytes), "r"(bar)
      : "memory");
}
__device__ __forceinline__ uint32_t cluster_rank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// read a float from the shared memory of CTA `rank` of this cluster (same offset as `local`)
__device__ __forceinline__ float ld_dsmem(const float *local, uint32_t rank) {
  uint32_t ra;
  float v;
  asm("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(smem_u32(local)), "r"(rank));
  asm volatile("ld.shared::cluster.f32 %0, [%1];" : "=f"(v) : "r"(ra));
  return v;
}
__device__ __forceinline__ void consumer_sync() { asm volatile("bar.sync 1, 512;" ::: "memory"); }

// e^x through the SFU (rel. error 2^-22): soft-max numerators
__device__ __forceinline__ float exp_fast(float x) { return ex2_approx(x * 1.4426950408889634f); }

// acc += u * tanh(x + s) in five instructions: with C = 2 log2(e), sc = s*C and m2u = -2u
// precomputed, u*tanh(x+s) = u - 2u / (1 + 2^(x*C + sc)); the "+u" terms are pre-summed into acc.
__device__ __forceinline__ float tanh_acc(float x, float sc, float m2u, float acc) {
  const float e = ex2_approx(fmaf(x, 2.885390081777927f, sc));
  return fmaf(rcp_approx(e + 1.0f), m2u, acc);
}

// largest chunk index whose first frame is <= g, for n equal chunks of F frames:
// chunk i covers [i*F/n, (i+1)*F/n)
__host__ __device__ __forceinline__ int chunk_of(long long g, int n, long long F) {
  return static_cast<int>(((g + 1) * n - 1) / F);
}

// NCOL = ceil(H / 512) columns per thread in phases C/D (col = tid + 512 k);
// NV4  = ceil(H / 256) float4 chunks per lane in phase A (warp = (region, column half));
// RT   = compile-time R (8) or 0 for a runtime R <= 16;  HT = compile-time H (512) or 0.
template <int NCOL, int NV4, int RT, int HT>
__global__ void __launch_bounds__(NTHREADS, 1) att_stream_kernel(const AttArgs a, const int nstages, const int cs) {
  extern __shared__ __align__(128) uint8_t smem_raw[];
  constexpr int RU = RT ? RT : RMAX;              // unroll bound of the region loops
  const int H = HT ? HT : a.H, T = a.T;
  const int R = RT ? RT : a.R;
  const int RH = R * H;
  const int stage_floats = 3 * RH + 4 * H;
  float *ring = reinterpret_cast<float *>(smem_raw);
  float *s_sc = ring + static_cast<size_t>(nstages) * stage_floats;   // [2][RMAX][2] region score halves
  float *s_part = s_sc + 2 * RMAX * 2;                                  // [2][NWARPS][4] g, m, lt partials
  int *s_flag = reinterpret_cast<int *>(s_part + 2 * NWARPS * 4);
  uint64_t *bars = reinterpret_cast<uint64_t *>(s_flag + 2);           // full[ns], empty[ns]
  const uint32_t bar_full = smem_u32(bars);
  const uint32_t bar_empty = bar_full + 8 * nstages;

  // cluster <-> decode row; CTA rank <-> slice [t0, t0 + nframes) of its T frames
  const int row = static_cast<int>(blockIdx.x) / cs;
  const int rank = cs > 1 ? static_cast<int>(cluster_rank()) : 0;
  const int t0 = (rank * T) / cs;
  const int nframes = ((rank + 1) * T) / cs - t0;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  long long *ktrace = (a.trace && blockIdx.x == 0 && tid == 0) ? a.trace + 64 : nullptr;
  if (ktrace) ktrace[0] = clock64();

  if (tid == 0) {
    for (int s = 0; s < nstages; ++s) {
      mbar_init(bar_full + 8 * s, 1);
      mbar_init(bar_empty + 8 * s, NWARPS);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  if (warp == NWARPS) {
    // ============================ producer ============================
    if (lane == 0) {
      const uint32_t bytes_rh = static_cast<uint32_t>(RH) * 4u, bytes_h = static_cast<uint32_t>(H) * 4u;
      int t = t0, s = 0;
      uint32_t ph = 0;
      const int clip = a.row_clip ? a.row_clip[row] : row;
      for (int i = 0; i < nframes; ++i) {
        mbar_wait(bar_empty + 8 * s, ph ^ 1);
        const size_t frame = static_cast<size_t>(clip) * T + t;
        const uint32_t dst = smem_u32(ring + static_cast<size_t>(s) * stage_floats);
        const uint32_t fb = bar_full + 8 * s;
        mbar_expect_tx(fb, static_cast<uint32_t>(stage_floats) * 4u);
        bulk_g2s(dst, a.pctxl + frame * RH, bytes_rh, fb);
        bulk_g2s(dst + bytes_rh, a.ctxl0 + frame * RH, bytes_rh, fb);
        bulk_g2s(dst + 2 * bytes_rh, a.qctxl + frame * RH, bytes_rh, fb);
        bulk_g2s(dst + 3 * bytes_rh, a.pctxg + frame * H, bytes_h, fb);
        bulk_g2s(dst + 3 * bytes_rh + bytes_h, a.pctxm + frame * H, bytes_h, fb);
        bulk_g2s(dst + 3 * bytes_rh + 2 * bytes_h, a.ctxg0 + frame * H, bytes_h, fb);
        bulk_g2s(dst + 3 * bytes_rh + 3 * bytes_h, a.ctxm0 + frame * H, bytes_h, fb);
        ++t;
        if (++s == nstages) { s = 0; ph ^= 1; }
      }
    }
    __syncwarp();
    if (cs > 1) {          // the producer warp takes part in the two cluster barriers of the merge
      cluster_sync_all();
      cluster_sync_all();
    }
    return;
  }

  // ============================== consumers ==============================
  // phase A: warp w scores region (w & 7) [+8] over the column half (w >> 3); lane owns float4
  //          chunks c = half*H/2 + 4*lane + 128*j (j < NV4)
  // phases C/D: thread owns columns col = tid + 512*k (k < NCOL)
  const int half = warp >> 3, hw = H >> 1;
  float ul[NV4][4], sl[NV4][4];
  float ug[NCOL], um[NCOL], ult[NCOL], sg[NCOL], sm[NCOL], slt[NCOL];
  float acc[3][NCOL];
  float rm[3], rs[3];
  float beta = 1.0f;
#pragma unroll
  for (int j = 0; j < NV4; ++j) {
    const int c = 4 * lane + 128 * j;
#pragma unroll
    for (int v = 0; v < 4; ++v) ul[j][v] = (c < hw) ? __ldg(a.Ul + half * hw + c + v) : 0.f;
  }
#pragma unroll
  for (int k = 0; k < NCOL; ++k) {
    const int col = tid + 512 * k;
    const bool ok = col < H;
    ug[k] = ok ? __ldg(a.Ug + col) : 0.f;
    um[k] = ok ? __ldg(a.Um + col) : 0.f;
    ult[k] = ok ? __ldg(a.Ult + col) : 0.f;
  }
  const float cl = __ldg(a.cl), cg = __ldg(a.cg), cm = __ldg(a.cm), clt = __ldg(a.clt);

  {
    // The h-projections arrive as k-slice planes (<= 8), summed here in plane order.  The loads of
    // four planes are issued together before their values are used (one memory round trip).
    const float *hp = a.hp + static_cast<size_t>(row) * a.ldhp;
#pragma unroll
    for (int j = 0; j < NV4; ++j)
#pragma unroll
      for (int v = 0; v < 4; ++v) sl[j][v] = 0.f;
#pragma unroll
    for (int k = 0; k < NCOL; ++k) {
      sg[k] = 0.f; sm[k] = 0.f; slt[k] = 0.f;
      acc[0][k] = 0.f; acc[1][k] = 0.f; acc[2][k] = 0.f;
    }
    float bsel = 0.f;
    for (int q0 = 0; q0 < a.hp_parts; q0 += 4) {
      float4 xs[4][NV4];
      float xg[4][NCOL], xm[4][NCOL], xt[4][NCOL], xb[4];
#pragma unroll
      for (int qq = 0; qq < 4; ++qq) {
        const bool okq = q0 + qq < a.hp_parts;
        const float *hq = hp + static_cast<size_t>(okq ? q0 + qq : q0) * a.hp_plane;
#pragma unroll
        for (int j = 0; j < NV4; ++j) {
          const int c = 4 * lane + 128 * j;
          xs[qq][j] = (okq && c < hw) ? *reinterpret_cast<const float4 *>(hq + a.off_sl + half * hw + c)
                                      : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int k = 0; k < NCOL; ++k) {
          const int col = tid + 512 * k;
          const bool ok = okq && col < H;
          xg[qq][k] = ok ? hq[a.off_sg + col] : 0.f;
          xm[qq][k] = ok ? hq[a.off_sm + col] : 0.f;
          xt[qq][k] = ok ? hq[a.off_slt + col] : 0.f;
        }
        xb[qq] = (okq && a.selector) ? hq[a.off_sel] : 0.f;
      }
#pragma unroll
      for (int qq = 0; qq < 4; ++qq) {
#pragma unroll
        for (int j = 0; j < NV4; ++j) {
          sl[j][0] += xs[qq][j].x; sl[j][1] += xs[qq][j].y; sl[j][2] += xs[qq][j].z; sl[j][3] += xs[qq][j].w;
        }
#pragma unroll
        for (int k = 0; k < NCOL; ++k) { sg[k] += xg[qq][k]; sm[k] += xm[qq][k]; slt[k] += xt[qq][k]; }
        bsel += xb[qq];
      }
    }
    beta = a.selector ? sigmoid_acc(bsel) : 1.0f;
#pragma unroll
    for (int q = 0; q < 3; ++q) { rm[q] = -INFINITY; rs[q] = 0.f; }
  }
  // fold the tanh constants into the per-row state (see tanh_acc): s -> s*C, u -> -2u, and the
  // sum of this thread's u values as the starting value of each score partial
  constexpr float C2 = 2.885390081777927f;
  float su_l = 0.f, su_g = 0.f, su_m = 0.f, su_lt = 0.f;
#pragma unroll
  for (int j = 0; j < NV4; ++j)
#pragma unroll
    for (int v = 0; v < 4; ++v) {
      su_l += ul[j][v];
      ul[j][v] *= -2.0f;
      sl[j][v] *= C2;
    }
#pragma unroll
  for (int k = 0; k < NCOL; ++k) {
    su_g += ug[k]; su_m += um[k]; su_lt += ult[k];
    ug[k] *= -2.0f; um[k] *= -2.0f; ult[k] *= -2.0f;
    sg[k] *= C2; sm[k] *= C2; slt[k] *= C2;
  }
  if (ktrace) ktrace[1] = clock64();

  // phase A of a frame (ring slot st, score buffer bufA): region scores; the g / m score
  // partials of this thread's columns are returned in pg / pm (reduced later, with lt)
  float pg = 0.f, pm = 0.f;
  auto phase_a = [&](const float *st, int bufA) {
    const float *pL = st, *pG = st + 3 * RH, *pM = pG + H;
    for (int r = warp & 7; r < R; r += 8) {
      float part = su_l;
#pragma unroll
      for (int j = 0; j < NV4; ++j) {
        const int c = 4 * lane + 128 * j;
        if (HT || c < hw) {
          const float4 x = *reinterpret_cast<const float4 *>(pL + r * H + half * hw + c);
          part = tanh_acc(x.x, sl[j][0], ul[j][0], part);
          part = tanh_acc(x.y, sl[j][1], ul[j][1], part);
          part = tanh_acc(x.z, sl[j][2], ul[j][2], part);
          part = tanh_acc(x.w, sl[j][3], ul[j][3], part);
        }
      }
      part = warp_sum(part);
      if (lane == 0) s_sc[(bufA * RMAX + r) * 2 + half] = part;
    }
    pg = su_g;
    pm = su_m;
#pragma unroll
    for (int k = 0; k < NCOL; ++k) {
      const int col = tid + 512 * k;
      if (HT || col < H) {
        pg = tanh_acc(pG[col], sg[k], ug[k], pg);
        pm = tanh_acc(pM[col], sm[k], um[k], pm);
      }
    }
  };

  // Software pipeline over the frames of this CTA (two block barriers per frame):
  //   prologue  A(0)                                   | bar
  //   frame i   C(i)                                   | bar
  //             D(i) and A(i+1), independent, overlap  | bar
  int t = t0, s = 0;
  uint32_t ph = 0;
  if (nframes > 0) {
    mbar_wait(bar_full, 0);
    phase_a(ring, 0);
  }
  consumer_sync();
  for (int i = 0; i < nframes; ++i) {
    const int buf = i & 1;
    const float *st = ring + static_cast<size_t>(s) * stage_floats;
    const float *cL0 = st + RH, *qL = st + 2 * RH;
    const float *G0 = st + 3 * RH + 2 * H, *M0 = G0 + H;

    // ---- C: softmax over regions, attended local context, its projection, lt partial ---
    float al[RU];        // un-normalised soft-max numerators e_r; inv = 1 / sum
    float inv;
    {
      float sc[RU];
      float mx = -INFINITY;
#pragma unroll
      for (int r = 0; r < RU; ++r) {
        const float2 hs = *reinterpret_cast<const float2 *>(s_sc + (buf * RMAX + r) * 2);
        sc[r] = (RT || r < R) ? (hs.x + hs.y) + cl : -INFINITY;
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
      if (a.alpha_l && tid < R) a.alpha_l[(static_cast<size_t>(row) * T + t) * R + tid] = e_mine * inv;
    }
    float cLv[NCOL];
    float plt = su_lt;
#pragma unroll
    for (int k = 0; k < NCOL; ++k) {
      const int col = tid + 512 * k;
      float c0 = 0.f, p0 = 0.f;
      if (HT || col < H) {
#pragma unroll
        for (int r = 0; r < RU; ++r) {
          if (RT || r < R) {
            c0 = fmaf(al[r], cL0[r * H + col], c0);
            p0 = fmaf(al[r], qL[r * H + col], p0);
          }
        }
        c0 *= inv;
        plt = tanh_acc(p0 * inv, slt[k], ult[k], plt);
      }
      cLv[k] = c0;
    }
    {
      // one reduction for the three temporal-score partials (pg, pm from phase A of this frame):
      // xor-16 leaves (pg, pm) sums in lanes < 16 and (plt, -) in lanes >= 16, xor-8 one value per
      // lane, then xor 4, 2, 1: totals end up in lane 0 (g), 8 (m), 16 (lt)
      const bool up = lane & 16;
      const float keep0 = up ? plt : pg, keep1 = up ? 0.f : pm;
      const float send0 = up ? pg : plt, send1 = up ? pm : 0.f;
      const float v0 = keep0 + __shfl_xor_sync(0xffffffffu, send0, 16);
      const float v1 = keep1 + __shfl_xor_sync(0xffffffffu, send1, 16);
      const bool up8 = lane & 8;
      float v = (up8 ? v1 : v0) + __shfl_xor_sync(0xffffffffu, up8 ? v0 : v1, 8);
      v += __shfl_xor_sync(0xffffffffu, v, 4);
      v += __shfl_xor_sync(0xffffffffu, v, 2);
      v += __shfl_xor_sync(0xffffffffu, v, 1);
      if ((lane & 7) == 0 && lane < 24) s_part[(buf * NWARPS + warp) * 4 + (lane >> 3)] = v;
    }
    consumer_sync();

    // ---- A(i+1): next frame's scores, independent of D(i) below ------------------------
    if (i + 1 < nframes) {
      int s1 = s + 1;
      uint32_t ph1 = ph;
      if (s1 == nstages) { s1 = 0; ph1 ^= 1; }
      mbar_wait(bar_full + 8 * s1, ph1);
      phase_a(ring + static_cast<size_t>(s1) * stage_floats, buf ^ 1);
    }

    // ---- D(i): fold the frame into the three running soft-max states --------------------
    float sc3[3];
    {
      // lane l holds the partials of warp (l & 15); xor-shuffles over 8,4,2,1 sum the 16 warps
      const float4 p4 = *reinterpret_cast<const float4 *>(s_part + (buf * NWARPS + (lane & (NWARPS - 1))) * 4);
      float x = p4.x, y = p4.y, z = p4.z;
#pragma unroll
      for (int o = 8; o > 0; o >>= 1) {
        x += __shfl_xor_sync(0xffffffffu, x, o);
        y += __shfl_xor_sync(0xffffffffu, y, o);
        z += __shfl_xor_sync(0xffffffffu, z, o);
      }
      sc3[0] = x + cg;
      sc3[1] = y + cm;
      sc3[2] = z + clt;
    }
    if (a.att_scores && tid == 0) {
      const size_t plane = static_cast<size_t>(a.rows) * T, at = static_cast<size_t>(row) * T + t;
      a.att_scores[at] = sc3[0];
      a.att_scores[plane + at] = sc3[1];
      a.att_scores[2 * plane + at] = sc3[2];
    }
    float e[3], keep[3];
    {
      // lanes 0..2 evaluate the three rescale factors, lanes 3..5 the three new numerators
      float mn[3];
#pragma unroll
      for (int q = 0; q < 3; ++q) mn[q] = fmaxf(rm[q], sc3[q]);
      const int q = lane % 3;
      const float from = (lane < 3) ? (q == 0 ? rm[0] : (q == 1 ? rm[1] : rm[2]))
                                    : (q == 0 ? sc3[0] : (q == 1 ? sc3[1] : sc3[2]));
      const float to = q == 0 ? mn[0] : (q == 1 ? mn[1] : mn[2]);
      const float ex = (lane < 6) ? exp_fast(from - to) : 0.f;
#pragma unroll
      for (int qq = 0; qq < 3; ++qq) {
        keep[qq] = __shfl_sync(0xffffffffu, ex, qq);
        e[qq] = __shfl_sync(0xffffffffu, ex, 3 + qq);
        rs[qq] = fmaf(rs[qq], keep[qq], e[qq]);
        rm[qq] = mn[qq];
      }
    }
#pragma unroll
    for (int k = 0; k < NCOL; ++k) {
      const int col = tid + 512 * k;
      if (col < H) {
        acc[0][k] = fmaf(acc[0][k], keep[0], e[0] * G0[col]);
        acc[1][k] = fmaf(acc[1][k], keep[1], e[1] * M0[col]);
        acc[2][k] = fmaf(acc[2][k], keep[2], e[2] * cLv[k]);
      }
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(bar_empty + 8 * s);
    consumer_sync();
    ++t;
    if (++s == nstages) { s = 0; ph ^= 1; }
  }

  // ---- merge of the cs partial states of this row (S8, S9) -----------------------------
  if (ktrace) ktrace[2] = clock64();
  float *ctx = a.ctx + static_cast<size_t>(row) * H;
  if (cs == 1) {
#pragma unroll
    for (int k = 0; k < NCOL; ++k) {
      const int col = tid + 512 * k;
      if (col < H) ctx[col] = beta * (acc[0][k] / rs[0] + acc[1][k] / rs[1] + acc[2][k] / rs[2]);
    }
    return;
  }
  // park (m, s)[3] and the three weighted vectors at the start of the (now drained) ring
  consumer_sync();
  float *s_ms = ring;                 // [6]
  float *s_vec = ring + 8;            // [3][H]
  if (tid == 0) {
#pragma unroll
    for (int q = 0; q < 3; ++q) { s_ms[2 * q] = rm[q]; s_ms[2 * q + 1] = rs[q]; }
  }
#pragma unroll
  for (int q = 0; q < 3; ++q)
#pragma unroll
    for (int k = 0; k < NCOL; ++k) {
      const int col = tid + 512 * k;
      if (col < H) s_vec[q * H + col] = acc[q][k];
    }
  cluster_sync_all();
  if (ktrace) ktrace[3] = clock64();
  if (rank == 0) {
    // all remote values first (independent DSMEM loads), then the arithmetic
    float pm[3][8], psum[3][8], v[3][8][NCOL];
#pragma unroll
    for (int q = 0; q < 3; ++q)
#pragma unroll
      for (int p = 0; p < 8; ++p) {
        const bool okp = p < cs;
        pm[q][p] = okp ? ld_dsmem(s_ms + 2 * q, p) : -INFINITY;
        psum[q][p] = okp ? ld_dsmem(s_ms + 2 * q + 1, p) : 0.f;
#pragma unroll
        for (int k = 0; k < NCOL; ++k) {
          const int col = tid + 512 * k;
          v[q][p][k] = (okp && col < H) ? ld_dsmem(s_vec + q * H + col, p) : 0.f;
        }
      }
    float o[NCOL];
#pragma unroll
    for (int k = 0; k < NCOL; ++k) o[k] = 0.f;
#pragma unroll
    for (int q = 0; q < 3; ++q) {
      float mx = -INFINITY;
#pragma unroll
      for (int p = 0; p < 8; ++p) mx = fmaxf(mx, pm[q][p]);
      float w[8], den = 0.f;
#pragma unroll
      for (int p = 0; p < 8; ++p) {
        w[p] = (p < cs) ? expf(pm[q][p] - mx) : 0.f;
        den = fmaf(w[p], psum[q][p], den);
      }
      const float inv = 1.0f / den;
#pragma unroll
      for (int k = 0; k < NCOL; ++k) {
        float num = 0.f;
#pragma unroll
        for (int p = 0; p < 8; ++p) num = fmaf(w[p], v[q][p][k], num);
        o[k] = fmaf(num, inv, o[k]);
      }
    }
#pragma unroll
    for (int k = 0; k < NCOL; ++k) {
      const int col = tid + 512 * k;
      if (col < H) ctx[col] = beta * o[k];
    }
  }
  if (ktrace) ktrace[4] = clock64();
  cluster_sync_all();       // partners keep their shared memory alive until rank 0 has read it
  if (ktrace) ktrace[5] = clock64();
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

size_t stage_bytes(int R, int H) { return (static_cast<size_t>(3) * R * H + 4 * static_cast<size_t>(H)) * 4; }

template <int NCOL, int NV4, int RT, int HT>
int launch(const AttArgs &a, int cs, int nstages, cudaStream_t stream) {
  const size_t smem = nstages * stage_bytes(a.R, a.H) + (2 * RMAX * 2 + 2 * NWARPS * 4 + 2) * 4 + 2 * 8 * nstages + 64;
  static size_t smem_set = 0;
  if (smem > smem_set) {
    STAT_CUDA_CHECK(cudaFuncSetAttribute(att_stream_kernel<NCOL, NV4, RT, HT>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    smem_set = smem;
  }
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(static_cast<unsigned>(a.rows) * cs);
  cfg.blockDim = dim3(NTHREADS);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = cs;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  STAT_CUDA_CHECK(cudaLaunchKernelEx(&cfg, att_stream_kernel<NCOL, NV4, RT, HT>, a, nstages, cs));
  note_launch();
  return STAT_OK;
}

}  // namespace

// plan shared by the launcher and the workspace layout: cluster size and ring depth
bool att_stream_plan(int rows, int T, int R, int H, int *cluster, int *max_parts, int *nstages) {
  if ((H & 7) != 0 || H > 1024 || R > RMAX) return false;
  const size_t sb = stage_bytes(R, H);
  int ns = static_cast<int>((200 * 1024) / sb);
  if (ns < 2) return false;
  if (ns > 4) ns = 4;
  int cs = 1;
  while (cs < 8 && 2 * cs <= T && static_cast<long long>(rows) * 2 * cs <= sm_count()) cs *= 2;
  *cluster = cs;
  *max_parts = 1;          // partial states never leave the cluster
  *nstages = ns;
  return true;
}

static long long *g_att_trace = nullptr;
void att_set_trace(long long *p) { g_att_trace = p; }

int att_stream_launch(const AttArgs &a_in, cudaStream_t stream) {
  AttArgs a = a_in;
  a.trace = g_att_trace;
  int nchunks, max_parts, nstages;   // nchunks = cluster size
  STAT_REQUIRE(att_stream_plan(a.rows, a.T, a.R, a.H, &nchunks, &max_parts, &nstages), STAT_EINVAL,
               "att_stream: unsupported shape R=%d H=%d", a.R, a.H);
  const int H = a.H;
  if (H == 512 && a.R == 8) return launch<1, 2, 8, 512>(a, nchunks, nstages, stream);   // BASELINE shape
  if (H <= 256) return launch<1, 1, 0, 0>(a, nchunks, nstages, stream);
  if (H <= 512) return launch<1, 2, 0, 0>(a, nchunks, nstages, stream);
  if (H <= 768) return launch<2, 3, 0, 0>(a, nchunks, nstages, stream);
  return launch<2, 4, 0, 0>(a, nchunks, nstages, stream);
}

}  // namespace stat
