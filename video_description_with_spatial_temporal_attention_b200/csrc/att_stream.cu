// att_stream_kernel: the four soft-attentions of one decode step (model_attention.py:370-435,
// SURVEY App. A S1-S9) as a persistent, TMA-fed streaming kernel -- the HBM-bound kernel
// of the path.
//
// Work = the rows*T (row, frame) pairs of the step, cut into gridDim.x contiguous,
// equally sized chunks (one CTA per SM, 16 consumer warps + 1 producer warp), so a chunk
// may cover the tail of one decode row and the head of the next.  Per frame the producer warp issues seven
// cp.async.bulk copies (pctxl[t], ctxl0[t], qctxl[t]: R*H floats each; pctxg[t],
// pctxm[t], ctxg0[t], ctxm0[t]: H floats each) into one slot of a shared-memory ring
// guarded by full/empty mbarriers; every byte of the seven context blocks is read from
// HBM/L2 exactly once per step.  The consumer warps then
//   A. score the R regions (warp = region x column half): sum_h tanh(pctxl + h.Wdl) * Ul (S1)
//      and the g / m temporal scores of the frame (column-sliced partials)          (S4, S5)
//   C. softmax over R (S2), alpha-weighted sums cL = sum_r a_r ctxl0_r (S3) and
//      pLT = sum_r a_r qctxl_r + h.Wdlt + blt (S6, the :416 GEMM folded by linearity),
//      then the lt score partial                                                     (S7)
//   D. fold the frame into flash-style running (max, sum, weighted vector) states of
//      the g / m / lt temporal soft-maxes; the weighted vectors live in registers.
// When the chunk leaves a row, the partial state goes to global memory and the last
// CTA of that row (atomic ticket) merges the parts, applies the selector gate and
// writes ctx (S8, S9).
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
__device__ __forceinline__ void consumer_sync() { asm volatile("bar.sync 1, 512;" ::: "memory"); }

// e^x through the SFU (rel. error 2^-22): soft-max numerators
__device__ __forceinline__ float exp_fast(float x) { return ex2_approx(x * 1.4426950408889634f); }

// largest chunk index whose first frame is <= g, for n equal chunks of F frames:
// chunk i covers [i*F/n, (i+1)*F/n)
__host__ __device__ __forceinline__ int chunk_of(long long g, int n, long long F) {
  return static_cast<int>(((g + 1) * n - 1) / F);
}

// NCOL = ceil(H / 512) columns per thread in phases C/D (col = tid + 512 k);
// NV4  = ceil(H / 256) float4 chunks per lane in phase A (warp = (region, column half));
// RT   = compile-time R (8) or 0 for a runtime R <= 16.
template <int NCOL, int NV4, int RT>
__global__ void __launch_bounds__(NTHREADS, 1) att_stream_kernel(const AttArgs a, const int nstages) {
  extern __shared__ __align__(128) uint8_t smem_raw[];
  constexpr int RU = RT ? RT : RMAX;              // unroll bound of the region loops
  const int H = a.H, T = a.T;
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

  const long long F = static_cast<long long>(a.rows) * T;
  const int n = gridDim.x;
  const long long g0 = (static_cast<long long>(blockIdx.x) * F) / n;
  const long long g1 = (static_cast<long long>(blockIdx.x + 1) * F) / n;
  const int nframes = static_cast<int>(g1 - g0);
  const int row0 = static_cast<int>(g0 / T), t0 = static_cast<int>(g0 - static_cast<long long>(row0) * T);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

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
      int row = row0, t = t0, s = 0;
      uint32_t ph = 0;
      for (int i = 0; i < nframes; ++i) {
        mbar_wait(bar_empty + 8 * s, ph ^ 1);
        const int clip = a.row_clip ? a.row_clip[row] : row;
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
        if (++t == T) { t = 0; ++row; }
        if (++s == nstages) { s = 0; ph ^= 1; }
      }
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

  auto finalize = [&](int row) {
    // this chunk's share of `row` is complete: publish or merge
    const int first = chunk_of(static_cast<long long>(row) * T, n, F);
    const int last = chunk_of(static_cast<long long>(row) * T + T - 1, n, F);
    const int nparts = last - first + 1;
    float *ctx = a.ctx + static_cast<size_t>(row) * H;
    if (nparts == 1) {
#pragma unroll
      for (int k = 0; k < NCOL; ++k) {
        const int col = tid + 512 * k;
        if (col < H) ctx[col] = beta * (acc[0][k] / rs[0] + acc[1][k] / rs[1] + acc[2][k] / rs[2]);
      }
      return;
    }
    const int part = static_cast<int>(blockIdx.x) - first;
    float *rv = a.rec_vec + (static_cast<size_t>(row) * a.S + part) * 3 * H;
    float *rms = a.rec_ms + (static_cast<size_t>(row) * a.S + part) * 6;
#pragma unroll
    for (int q = 0; q < 3; ++q)
#pragma unroll
      for (int k = 0; k < NCOL; ++k) {
        const int col = tid + 512 * k;
        if (col < H) rv[q * H + col] = acc[q][k];
      }
    if (tid == 0) {
#pragma unroll
      for (int q = 0; q < 3; ++q) {
        rms[2 * q] = rm[q];
        rms[2 * q + 1] = rs[q];
      }
    }
    __threadfence();
    consumer_sync();
    if (tid == 0) {
      const unsigned int ticket = atomicAdd(a.counters + row, 1u);
      *s_flag = (ticket == static_cast<unsigned int>(nparts - 1)) ? 1 : 0;
    }
    consumer_sync();
    const int is_last = *s_flag;
    consumer_sync();                       // s_flag may be rewritten by the next finalize
    if (!is_last) return;
    __threadfence();
    const float *ms_all = a.rec_ms + static_cast<size_t>(row) * a.S * 6;
    const float *rv_all = a.rec_vec + static_cast<size_t>(row) * a.S * 3 * H;
    if (nparts <= 4) {
      // common case: all loads issued before any use (one memory round trip)
      float pm[4][3], psum[4][3], v[4][3][NCOL];
#pragma unroll
      for (int p = 0; p < 4; ++p)
#pragma unroll
        for (int q = 0; q < 3; ++q) {
          const bool ok = p < nparts;
          pm[p][q] = ok ? __ldcg(ms_all + p * 6 + 2 * q) : -INFINITY;
          psum[p][q] = ok ? __ldcg(ms_all + p * 6 + 2 * q + 1) : 0.f;
#pragma unroll
          for (int k = 0; k < NCOL; ++k) {
            const int col = tid + 512 * k;
            v[p][q][k] = (ok && col < H) ? __ldcg(rv_all + (static_cast<size_t>(p) * 3 + q) * H + col) : 0.f;
          }
        }
      float o[NCOL];
#pragma unroll
      for (int k = 0; k < NCOL; ++k) o[k] = 0.f;
#pragma unroll
      for (int q = 0; q < 3; ++q) {
        const float mx = fmaxf(fmaxf(pm[0][q], pm[1][q]), fmaxf(pm[2][q], pm[3][q]));
        float w[4], den = 0.f;
#pragma unroll
        for (int p = 0; p < 4; ++p) {
          w[p] = (p < nparts) ? expf(pm[p][q] - mx) : 0.f;
          den = fmaf(w[p], psum[p][q], den);
        }
        const float inv = 1.0f / den;
#pragma unroll
        for (int k = 0; k < NCOL; ++k) {
          float num = 0.f;
#pragma unroll
          for (int p = 0; p < 4; ++p) num = fmaf(w[p], v[p][q][k], num);
          o[k] = fmaf(num, inv, o[k]);
        }
      }
#pragma unroll
      for (int k = 0; k < NCOL; ++k) {
        const int col = tid + 512 * k;
        if (col < H) ctx[col] = beta * o[k];
      }
    } else {
      float M[3], inv[3];
#pragma unroll
      for (int q = 0; q < 3; ++q) {
        float mx = -INFINITY;
        for (int p = 0; p < nparts; ++p) mx = fmaxf(mx, __ldcg(ms_all + p * 6 + 2 * q));
        float den = 0.f;
        for (int p = 0; p < nparts; ++p)
          den += expf(__ldcg(ms_all + p * 6 + 2 * q) - mx) * __ldcg(ms_all + p * 6 + 2 * q + 1);
        M[q] = mx;
        inv[q] = 1.0f / den;
      }
#pragma unroll
      for (int k = 0; k < NCOL; ++k) {
        const int col = tid + 512 * k;
        if (col < H) {
          float o = 0.f;
#pragma unroll
          for (int q = 0; q < 3; ++q) {
            float num = 0.f;
            for (int p = 0; p < nparts; ++p)
              num = fmaf(expf(__ldcg(ms_all + p * 6 + 2 * q) - M[q]),
                         __ldcg(rv_all + (static_cast<size_t>(p) * 3 + q) * H + col), num);
            o += num * inv[q];
          }
          ctx[col] = beta * o;
        }
      }
    }
    if (tid == 0) a.counters[row] = 0u;
  };

  int row = row0, t = t0, s = 0, cur_row = -1;
  uint32_t ph = 0;
  for (int i = 0; i < nframes; ++i) {
    if (row != cur_row) {
      if (cur_row >= 0) finalize(cur_row);
      cur_row = row;
      const float *hp = a.hp + static_cast<size_t>(row) * a.ldhp;
      // the h-projections arrive as k-slice planes: summed here in plane order
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
      // (fixed trip count + predicate so that the loads of all planes are in flight together)
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        if (q >= a.hp_parts) break;
        const float *hq = hp + q * a.hp_plane;
#pragma unroll
        for (int j = 0; j < NV4; ++j) {
          const int c = 4 * lane + 128 * j;
          if (c < hw) {
            const float4 x = *reinterpret_cast<const float4 *>(hq + a.off_sl + half * hw + c);
            sl[j][0] += x.x; sl[j][1] += x.y; sl[j][2] += x.z; sl[j][3] += x.w;
          }
        }
#pragma unroll
        for (int k = 0; k < NCOL; ++k) {
          const int col = tid + 512 * k;
          if (col < H) {
            sg[k] += hq[a.off_sg + col];
            sm[k] += hq[a.off_sm + col];
            slt[k] += hq[a.off_slt + col];
          }
        }
        if (a.selector) bsel += hq[a.off_sel];
      }
      beta = a.selector ? sigmoid_acc(bsel) : 1.0f;
#pragma unroll
      for (int q = 0; q < 3; ++q) { rm[q] = -INFINITY; rs[q] = 0.f; }
    }
    const int buf = i & 1;
    const float *st = ring + static_cast<size_t>(s) * stage_floats;
    const float *pL = st, *cL0 = st + RH, *qL = st + 2 * RH;
    const float *pG = st + 3 * RH, *pM = pG + H, *G0 = pM + H, *M0 = G0 + H;
    long long *trace = (a.trace && blockIdx.x == 0 && tid == 0 && i < 12) ? a.trace + 5 * i : nullptr;
    if (trace) trace[0] = clock64();
    mbar_wait(bar_full + 8 * s, ph);
    if (trace) trace[1] = clock64();

    // ---- A: region scores (half rows of H) and the g / m score partials ---------------
    for (int r = warp & 7; r < R; r += 8) {
      float part = 0.f;
#pragma unroll
      for (int j = 0; j < NV4; ++j) {
        const int c = 4 * lane + 128 * j;
        if (c < hw) {
          const float4 x = *reinterpret_cast<const float4 *>(pL + r * H + half * hw + c);
          part = fmaf(tanh_fast(x.x + sl[j][0]), ul[j][0], part);
          part = fmaf(tanh_fast(x.y + sl[j][1]), ul[j][1], part);
          part = fmaf(tanh_fast(x.z + sl[j][2]), ul[j][2], part);
          part = fmaf(tanh_fast(x.w + sl[j][3]), ul[j][3], part);
        }
      }
      part = warp_sum(part);
      if (lane == 0) s_sc[(buf * RMAX + r) * 2 + half] = part;
    }
    float pg = 0.f, pm = 0.f;
#pragma unroll
    for (int k = 0; k < NCOL; ++k) {
      const int col = tid + 512 * k;
      if (col < H) {
        pg = fmaf(tanh_fast(pG[col] + sg[k]), ug[k], pg);
        pm = fmaf(tanh_fast(pM[col] + sm[k]), um[k], pm);
      }
    }
    pg = warp_sum(pg);
    pm = warp_sum(pm);
    consumer_sync();
    if (trace) trace[2] = clock64();

    // ---- C: softmax over regions, attended local context, its projection, lt partial ---
    float al[RU];
    {
      float mx = -INFINITY;
#pragma unroll
      for (int r = 0; r < RU; ++r) {
        const float2 hs = *reinterpret_cast<const float2 *>(s_sc + (buf * RMAX + r) * 2);
        al[r] = (r < R) ? (hs.x + hs.y) + cl : -INFINITY;
        mx = fmaxf(mx, al[r]);
      }
      float den = 0.f;
#pragma unroll
      for (int r = 0; r < RU; ++r) {
        al[r] = (r < R) ? exp_fast(al[r] - mx) : 0.f;
        den += al[r];
      }
      const float inv = 1.0f / den;
#pragma unroll
      for (int r = 0; r < RU; ++r) {
        al[r] *= inv;
        if (a.alpha_l && tid == r && r < R) a.alpha_l[(static_cast<size_t>(row) * T + t) * R + r] = al[r];
      }
    }
    float cLv[NCOL];
    float plt = 0.f;
#pragma unroll
    for (int k = 0; k < NCOL; ++k) {
      const int col = tid + 512 * k;
      float c0 = 0.f, p0 = 0.f;
      if (col < H) {
#pragma unroll
        for (int r = 0; r < RU; ++r) {
          if (r < R) {
            c0 = fmaf(al[r], cL0[r * H + col], c0);
            p0 = fmaf(al[r], qL[r * H + col], p0);
          }
        }
        plt = fmaf(tanh_fast(p0 + slt[k]), ult[k], plt);
      }
      cLv[k] = c0;
    }
    plt = warp_sum(plt);
    if (lane == 0) *reinterpret_cast<float4 *>(s_part + (buf * NWARPS + warp) * 4) = make_float4(pg, pm, plt, 0.f);
    consumer_sync();
    if (trace) trace[3] = clock64();

    // ---- D: fold the frame into the three running soft-max states ----------------------
    float sc3[3] = {cg, cm, clt};
#pragma unroll
    for (int w = 0; w < NWARPS; ++w) {
      const float4 p4 = *reinterpret_cast<const float4 *>(s_part + (buf * NWARPS + w) * 4);   // broadcast reads
      sc3[0] += p4.x;
      sc3[1] += p4.y;
      sc3[2] += p4.z;
    }
    if (a.att_scores && tid == 0) {
      const size_t plane = static_cast<size_t>(a.rows) * T, at = static_cast<size_t>(row) * T + t;
      a.att_scores[at] = sc3[0];
      a.att_scores[plane + at] = sc3[1];
      a.att_scores[2 * plane + at] = sc3[2];
    }
    float e[3], keep[3];
#pragma unroll
    for (int q = 0; q < 3; ++q) {
      const float mn = fmaxf(rm[q], sc3[q]);
      keep[q] = exp_fast(rm[q] - mn);
      e[q] = exp_fast(sc3[q] - mn);
      rs[q] = fmaf(rs[q], keep[q], e[q]);
      rm[q] = mn;
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
    if (trace) trace[4] = clock64();
    if (++t == T) { t = 0; ++row; }
    if (++s == nstages) { s = 0; ph ^= 1; }
  }
  if (cur_row >= 0) finalize(cur_row);
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

template <int NCOL, int NV4, int RT>
int launch(const AttArgs &a, int nchunks, int nstages, cudaStream_t stream) {
  const size_t smem = nstages * stage_bytes(a.R, a.H) + (2 * RMAX * 2 + 2 * NWARPS * 4 + 2) * 4 + 2 * 8 * nstages + 64;
  static size_t smem_set = 0;
  if (smem > smem_set) {
    STAT_CUDA_CHECK(cudaFuncSetAttribute(att_stream_kernel<NCOL, NV4, RT>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    smem_set = smem;
  }
  att_stream_kernel<NCOL, NV4, RT><<<nchunks, NTHREADS, smem, stream>>>(a, nstages);
  note_launch();
  STAT_CUDA_CHECK(cudaGetLastError());
  return STAT_OK;
}

}  // namespace

// plan shared by the launcher and the workspace layout
bool att_stream_plan(int rows, int T, int R, int H, int *nchunks, int *max_parts, int *nstages) {
  if ((H & 7) != 0 || H > 1024 || R > RMAX) return false;
  const size_t sb = stage_bytes(R, H);
  int ns = static_cast<int>((200 * 1024) / sb);
  if (ns < 2) return false;
  if (ns > 4) ns = 4;
  const long long F = static_cast<long long>(rows) * T;
  long long n = sm_count();
  if (n > F) n = F;
  const long long min_chunk = F / n;   // >= 1
  long long parts = (T + min_chunk - 1) / min_chunk + 1;
  if (parts > T) parts = T;
  if (parts < 1) parts = 1;
  *nchunks = static_cast<int>(n);
  *max_parts = static_cast<int>(parts);
  *nstages = ns;
  return true;
}

static long long *g_att_trace = nullptr;
void att_set_trace(long long *p) { g_att_trace = p; }

int att_stream_launch(const AttArgs &a_in, cudaStream_t stream) {
  AttArgs a = a_in;
  a.trace = g_att_trace;
  int nchunks, max_parts, nstages;
  STAT_REQUIRE(att_stream_plan(a.rows, a.T, a.R, a.H, &nchunks, &max_parts, &nstages), STAT_EINVAL,
               "att_stream: unsupported shape R=%d H=%d", a.R, a.H);
  STAT_REQUIRE(a.S >= max_parts, STAT_EINVAL, "att_stream: partial buffers hold %d parts per row, need %d", a.S,
               max_parts);
  const int H = a.H;
  if (H == 512 && a.R == 8) return launch<1, 2, 8>(a, nchunks, nstages, stream);   // BASELINE shape
  if (H <= 256) return launch<1, 1, 0>(a, nchunks, nstages, stream);
  if (H <= 512) return launch<1, 2, 0>(a, nchunks, nstages, stream);
  if (H <= 768) return launch<2, 3, 0>(a, nchunks, nstages, stream);
  return launch<2, 4, 0>(a, nchunks, nstages, stream);
}

}  // namespace stat
