// cell_kernel: everything of a decode step between two attentions in ONE launch (model_attention.py:437-457,
// :371-433 of the next step, :684-696) -- the dependent chain of a step is  attention -> cell  (2 launches)
// instead of  attention -> ctx_proj -> gates -> h_proj (-> readout activation).
//
//   phase 1   pre = ctx_t.Wc + h_{t-1}.U + EW[word]  ->  S10-S13 (gates, c_t, h_t)           (:437-457)
//             zadd = b + ctx_t.ff_logit_ctxglm_W (+ Wemb[word])                               (:689-693)
//   ---- grid-wide barrier: every CTA needs the whole new hidden state ----
//   phase 2   h_t.[Wdl | Wdg | Wdm | Wdlt | W_sel | U]  -> queries / selector logit / h.U of step t+1
//             z = 0.5 tanh(0.5 h_t.ff_logit_lstm_W + zadd)                                    (:684-696)
//
// Why plain fp32 FMAs and not the tensor-core GEMM: the products are 64 rows x (20 | 32) columns x 512 per SM; a
// tcgen05 tile of that size is bound by its fill / dependent-MMA / drain latency (measured 6-8 us per launch,
// three launches), the SIMT tile by 2 x 10^6 FMAs per SM (packed FFMA2: 128 FMA/clk/SM).  Results are exact fp32
// sums (no 3xTF32 split).
//
// One CTA per SM (grid = number of SMs, all co-resident: the barrier spins).  CTA i owns
//   phase 1: hidden units [i*upc, (i+1)*upc) -- all four gates of a unit, so the state update is thread-local --
//            and readout columns [i*epc, (i+1)*epc)
//   phase 2: columns [i*cpc, (i+1)*cpc) of [queries 4H | sel | 3 pad | h.U 4H (gate-interleaved) | readout E]
// Its weight slabs are packed once per parameter set as [k][column] (stat_prepare_params -> cell_pack) and copied
// with single bulk copies; the phase-2 slab is requested at kernel start (before the programmatic-dependency
// wait: weights are constants), so it flies during phase 1.  The activations arrive TRANSPOSED, [k][row] in
// 64-row chunks (the attention writes ctx that way, phase 1 writes h that way): a thread's four rows are one
// LDS.128, its four columns another, sixteen FMAs (eight FFMA2) per pair of loads.  K is split over warp groups
// (8 in phase 1, 4 in phase 2), each waiting only for its own slice of the activation copy; the partial tiles
// are summed through shared memory in a fixed order (bit-reproducible).
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "kernels.cuh"
#include "stat_common.cuh"

namespace stat {
namespace {

constexpr int CT = 512;        // threads per CTA
constexpr int RC = 64;         // decode rows per activation chunk
constexpr int W1C = 20;        // floats per k of a phase-1 slab: 4 unit slots x 4 gates | 4 readout columns
constexpr int W2C = 32;        // floats per k of a phase-2 slab
constexpr int NBAR = 11;       // 8 activation slices, phase-1 slab, phase-2 slab (3/4), phase-2 slab (last quarter)

__device__ __forceinline__ uint32_t smem_u32(const void *p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  uint32_t spins = 0;
  do {
    if (++spins > (1u << 26)) __trap();      // a lost copy becomes an error, never a hang
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
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}
// ... with an L2 eviction-priority policy (the weight slabs are re-read on every step)
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar, uint64_t policy) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::
          "r"(dst),
      "l"(src), "r"(bytes), "r"(bar), "l"(policy)
      : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ float4 ld4(const float *p) { return *reinterpret_cast<const float4 *>(p); }

// Grid-wide barrier on two words in global memory: [0] arrivals, [1] generation.  The last CTA to arrive resets
// the count and bumps the generation, so the pair needs no host reset between launches (zero once).
__device__ __forceinline__ void grid_barrier(unsigned int *bar, unsigned int nctas) {
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned int gen, old;
    // read the generation BEFORE arriving: it cannot change until this CTA has arrived
    asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(gen) : "l"(bar + 1) : "memory");
    // release at gpu scope: the CTA's phase-1 stores (ordered before this by the bar.sync above) are visible to
    // whoever acquires the generation word
    asm volatile("atom.add.acq_rel.gpu.global.u32 %0, [%1], 1;" : "=r"(old) : "l"(bar) : "memory");
    if (old == nctas - 1) {
      asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(bar), "r"(0u) : "memory");
      asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(bar + 1), "r"(gen + 1) : "memory");
    } else {
      const long long t0 = clock64();
      unsigned int now;
      do {
        asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(now) : "l"(bar + 1) : "memory");
        if (now == gen && clock64() - t0 > 4000000000ll) __trap();   // ~2 s: a CTA that never became resident
      } while (now == gen);
    }
    fence_proxy_async();                      // what follows reads the other CTAs' stores with bulk copies
  }
  __syncthreads();
}

struct CellArgs {
  int rows, H, E, V;
  int upc, epc, cpc;          // hidden units / readout columns / phase-2 columns per CTA
  int NQ1, NQ;                // 8H+4, 8H+4+E
  int do1, do2;               // which phases run
  int want_q, want_z;         // phase 2: store the queries / h.U block, the readout activation
  int prev2out;
  // phase 1
  const float *W1;            // [grid][H][20] slabs
  const float *ctxT;          // [chunk][H][64]
  const float *hu;            // (rows, ldq): h_{t-1}.U, gate-interleaved (column 4u+g)
  int ldq;
  const float *EW;            // (V+1, 4H) gate-interleaved token table
  const float *Wemb;          // (V, E)
  const float *bz;            // (E)
  const int64_t *tok_prev;    // (rows) or null
  const float *mask;          // (rows) or null
  const float *dp_gates;      // (rows, 3H) or null = 0.5
  const float *h_in, *c_in;   // (rows, H)
  float *h_out, *c_out;       // (rows, H), may alias the inputs
  float *h_all;               // (rows, H) or null
  float *hT;                  // [chunk][H][64] new hidden state, transposed (phase-2 operand)
  float *zadd;                // (rows, E)
  // phase 2
  const float *W2;            // [grid][H][32] slabs
  const float *bq;            // (8H+4)
  float *hq;                  // (rows, ldq): queries | sel | pad | h.U
  float *z;                   // (rows, E)
  const float *dp_z;          // (rows, E) or null = 0.5
  unsigned int *bar;          // grid barrier words
  long long *trace;           // debug: globaltimer stamps (ns) of thread 0 of every CTA, 16 per CTA, or null
  unsigned long long policy;  // L2 eviction priority of the weight-slab copies
};

__global__ void __launch_bounds__(CT, 1) cell_kernel(const CellArgs a) {
  extern __shared__ __align__(128) uint8_t smem_raw[];
  const int H = a.H;
  float *Xs = reinterpret_cast<float *>(smem_raw);          // [H][64] activations, later the partial tiles
  float *W1s = Xs + static_cast<size_t>(RC) * H;            // [H][20]; phase 2: last K quarter of the phase-2 slab
  float *W2s = W1s + static_cast<size_t>(W1C) * H;          // [3H/4][32]
  uint64_t *bars = reinterpret_cast<uint64_t *>(W2s + static_cast<size_t>(W2C) * (H / 4) * 3);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int cta = blockIdx.x;
  long long *tr = (a.trace && tid == 0) ? a.trace + static_cast<size_t>(cta) * 16 : nullptr;
  auto stamp = [&](int i) {
    if (tr) {
      unsigned long long g;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g));
      tr[i] = static_cast<long long>(g);
    }
  };
  stamp(0);
  const uint32_t bx = smem_u32(bars), bw1 = bx + 64, bw2 = bx + 72, bw3 = bx + 80;
  const int u0 = cta * a.upc, e0 = cta * a.epc, c0 = cta * a.cpc;
  const int nu = min(a.upc, H - u0), ne = min(a.epc, a.E - e0);          // may be <= 0
  const bool work1 = a.do1 && (nu > 0 || ne > 0);
  const bool work2 = a.do2 && c0 < a.NQ;
  const int nchunks = (a.rows + RC - 1) / RC;
  const uint32_t q_bytes = static_cast<uint32_t>(H / 4) * W2C * 4u;     // one K quarter of the phase-2 slab

  if (tid == 0) {
    for (int i = 0; i < NBAR; ++i) mbar_init(bx + 8 * i, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  // the weight slabs are constants: their copies start before the programmatic-dependency wait
  if (tid == 0) {
    if (work1) {
      mbar_expect_tx(bw1, static_cast<uint32_t>(H) * W1C * 4u);
      bulk_g2s(smem_u32(W1s), a.W1 + static_cast<size_t>(cta) * H * W1C, static_cast<uint32_t>(H) * W1C * 4u, bw1,
               a.policy);
    }
    if (work2) {
      const float *w2 = a.W2 + static_cast<size_t>(cta) * H * W2C;
      mbar_expect_tx(bw2, 3u * q_bytes);
      bulk_g2s(smem_u32(W2s), w2, 3u * q_bytes, bw2, a.policy);
      if (!work1) {      // the last quarter lives where the phase-1 slab was; without a phase 1 it can come now
        mbar_expect_tx(bw3, q_bytes);
        bulk_g2s(smem_u32(W1s), w2 + static_cast<size_t>(3) * (H / 4) * W2C, q_bytes, bw3, a.policy);
      }
    }
  }
  stamp(1);
  pdl_wait();
  pdl_trigger();
  stamp(2);

  // =============================== phase 1: gates =======================================
  // warps 0..7: one K slice each (H/8), a thread owns 8 rows x (4 gates of one unit slot + 1 readout column);
  // threads 0..255 then finish one (row, slot) each -- their operands are requested before the products start
  if (work1) {
    const int s = lane & 3, rg = lane >> 2;
    const int KQ = H / 8;
    for (int rc = 0; rc < nchunks; ++rc) {
      if (tid == 0) {
        const float *src = a.ctxT + static_cast<size_t>(rc) * H * RC;
        for (int i = 0; i < 8; ++i) {
          mbar_expect_tx(bx + 8 * i, static_cast<uint32_t>(KQ) * RC * 4u);
          bulk_g2s(smem_u32(Xs + static_cast<size_t>(i) * KQ * RC), src + static_cast<size_t>(i) * KQ * RC,
                   static_cast<uint32_t>(KQ) * RC * 4u, bx + 8 * i);
        }
      }
      // epilogue operands of (row, slot): in flight during the products
      const int frow = rc * RC + 8 * rg + (warp & 7);
      const bool fin = tid < 256 && frow < a.rows;
      float4 hu = make_float4(0.f, 0.f, 0.f, 0.f), ew = hu;
      float c_ = 0.f, h_ = 0.f, msk = 1.0f, di = 0.5f, df = 0.5f, dO = 0.5f, zb = 0.f;
      if (fin) {
        const long long tok = a.tok_prev ? a.tok_prev[frow] : -1;
        if (s < nu) {
          const int u = u0 + s;
          hu = __ldcg(reinterpret_cast<const float4 *>(a.hu + static_cast<size_t>(frow) * a.ldq) + u);
          ew = __ldg(reinterpret_cast<const float4 *>(a.EW + static_cast<size_t>(tok >= 0 ? tok : a.V) * 4 * H) + u);
          const size_t idx = static_cast<size_t>(frow) * H + u;
          c_ = a.c_in[idx];
          h_ = a.h_in[idx];
          if (a.dp_gates) {
            const float *dp = a.dp_gates + static_cast<size_t>(frow) * 3 * H + u;
            di = dp[0]; df = dp[H]; dO = dp[2 * H];
          }
        }
        if (a.mask) msk = a.mask[frow];
        if (s < ne) {
          zb = a.bz[e0 + s];
          if (a.prev2out && tok >= 0) zb += __ldg(a.Wemb + static_cast<size_t>(tok) * a.E + e0 + s);
        }
      }
      if (warp < 8) {
        const int kg = warp;
        if (rc == 0) mbar_wait(bw1, 0);
        if (rc == 0) stamp(3);
        mbar_wait(bx + 8 * kg, rc & 1);
        if (rc == 0) stamp(4);
        float2 acc[8][2], az[4];
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[i][0] = acc[i][1] = make_float2(0.f, 0.f);
#pragma unroll
        for (int i = 0; i < 4; ++i) az[i] = make_float2(0.f, 0.f);
        const float *xp = Xs + static_cast<size_t>(kg) * KQ * RC + 8 * rg;
        const float *wp = W1s + static_cast<size_t>(kg) * KQ * W1C;
#pragma unroll 2
        for (int k = 0; k < KQ; ++k) {
          const float4 xa = ld4(xp + k * RC), xb = ld4(xp + k * RC + 4);
          const float4 w = ld4(wp + k * W1C + 4 * s);
          const float wz = wp[k * W1C + 16 + s];
          const float2 w01 = make_float2(w.x, w.y), w23 = make_float2(w.z, w.w), wzz = make_float2(wz, wz);
          const float xr[8] = {xa.x, xa.y, xa.z, xa.w, xb.x, xb.y, xb.z, xb.w};
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const float2 xx = make_float2(xr[i], xr[i]);
            acc[i][0] = __ffma2_rn(xx, w01, acc[i][0]);
            acc[i][1] = __ffma2_rn(xx, w23, acc[i][1]);
          }
#pragma unroll
          for (int i = 0; i < 4; ++i) az[i] = __ffma2_rn(make_float2(xr[2 * i], xr[2 * i + 1]), wzz, az[i]);
        }
        __syncwarp();
        // partial tile of this K slice -> its own (drained) slice of the activation buffer:
        // [kg][lane][row i][4 gates | readout]   (40 floats per thread; KQ*64 >= 32*40 for H >= 160)
        float *pp = Xs + static_cast<size_t>(kg) * KQ * RC + lane * 40;
        const float zr[8] = {az[0].x, az[0].y, az[1].x, az[1].y, az[2].x, az[2].y, az[3].x, az[3].y};
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          pp[5 * i + 0] = acc[i][0].x; pp[5 * i + 1] = acc[i][0].y;
          pp[5 * i + 2] = acc[i][1].x; pp[5 * i + 3] = acc[i][1].y;
          pp[5 * i + 4] = zr[i];
        }
      }
      __syncthreads();
      if (rc == 0) stamp(5);
      if (fin) {
        // K slices summed in order
        float sum[5] = {0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const float *pq = Xs + static_cast<size_t>(q) * KQ * RC + lane * 40 + 5 * (warp & 7);
#pragma unroll
          for (int c = 0; c < 5; ++c) sum[c] += pq[c];
        }
        if (s < nu) {
          const int u = u0 + s;
          const float pi = (hu.x + ew.x) + sum[0], pf = (hu.y + ew.y) + sum[1];
          const float po = (hu.z + ew.z) + sum[2], pg = (hu.w + ew.w) + sum[3];
          const float ig = sigmoid_acc(pi * di), fg = sigmoid_acc(pf * df), og = sigmoid_acc(po * dO);
          const float gg = tanhf(pg);
          const size_t idx = static_cast<size_t>(frow) * H + u;
          float c = fg * c_ + ig * gg;
          c = msk * c + (1.0f - msk) * c_;
          float h = og * tanhf(c);
          h = msk * h + (1.0f - msk) * h_;
          a.c_out[idx] = c;
          a.h_out[idx] = h;
          if (a.h_all) a.h_all[idx] = h;
          a.hT[(static_cast<size_t>(rc) * H + u) * RC + (frow - rc * RC)] = h;
        }
        if (s < ne) a.zadd[static_cast<size_t>(frow) * a.E + e0 + s] = zb + sum[4];
      }
      fence_proxy_async_smem();   // the partial tiles (generic stores) are overwritten by the next bulk copy
      __syncthreads();
    }
    stamp(6);
    if (work2 && tid == 0) {      // last K quarter of the phase-2 slab -> where the phase-1 slab was
      mbar_expect_tx(bw3, q_bytes);
      bulk_g2s(smem_u32(W1s), a.W2 + static_cast<size_t>(cta) * H * W2C + static_cast<size_t>(3) * (H / 4) * W2C,
               q_bytes, bw3, a.policy);
    }
  }
  if (a.do1 && a.do2) grid_barrier(a.bar, gridDim.x);
  stamp(7);

  // =============================== phase 2: products of the new hidden state ==========================
  // 8 K slices (H/8) x 2 warps; a thread owns 8 rows x 4 columns; then thread (slice q, t64) finishes row
  // 8*rg + q of its 4 columns
  if (work2) {
    const int kg = warp >> 1, t64 = ((warp & 1) << 5) | lane, cg = t64 & 7, rg = t64 >> 3;
    const int KQ = H / 8;
    // K slices 0..5 sit in the phase-2 region, 6..7 (the last quarter) where the phase-1 slab was
    const float *wbase = (kg < 6) ? W2s + static_cast<size_t>(kg) * KQ * W2C : W1s + static_cast<size_t>(kg - 6) * KQ * W2C;
    const uint32_t par0 = work1 ? static_cast<uint32_t>(nchunks) : 0u;   // uses of the activation barriers so far
    float bqv[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int c = c0 + 4 * cg + i;
      bqv[i] = (c < a.NQ1) ? __ldg(a.bq + c) : 0.f;
    }
    for (int rc = 0; rc < nchunks; ++rc) {
      if (tid == 0) {
        const float *src = a.hT + static_cast<size_t>(rc) * H * RC;
        for (int i = 0; i < 8; ++i) {
          mbar_expect_tx(bx + 8 * i, static_cast<uint32_t>(KQ) * RC * 4u);
          bulk_g2s(smem_u32(Xs + static_cast<size_t>(i) * KQ * RC), src + static_cast<size_t>(i) * KQ * RC,
                   static_cast<uint32_t>(KQ) * RC * 4u, bx + 8 * i);
        }
      }
      // readout addend of this thread's (row, columns): in flight during the products
      const int frow = rc * RC + 8 * rg + kg;
      float za[4] = {0.f, 0.f, 0.f, 0.f}, zp[4] = {0.5f, 0.5f, 0.5f, 0.5f};
      if (a.want_z && frow < a.rows) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int cl = 4 * cg + i, c = c0 + cl;
          if (cl < a.cpc && c >= a.NQ1 && c < a.NQ) {
            const size_t zi = static_cast<size_t>(frow) * a.E + (c - a.NQ1);
            za[i] = __ldcg(a.zadd + zi);
            if (a.dp_z) zp[i] = a.dp_z[zi];
          }
        }
      }
      if (rc == 0) mbar_wait(kg < 6 ? bw2 : bw3, 0);
      if (rc == 0) stamp(8);
      mbar_wait(bx + 8 * kg, (par0 + rc) & 1);
      if (rc == 0) stamp(9);
      float2 acc[8][2];
#pragma unroll
      for (int i = 0; i < 8; ++i) acc[i][0] = acc[i][1] = make_float2(0.f, 0.f);
      {
        const float *xp = Xs + static_cast<size_t>(kg) * KQ * RC + 8 * rg;
        const float *wp = wbase + 4 * cg;
#pragma unroll 2
        for (int k = 0; k < KQ; ++k) {
          const float4 xa = ld4(xp + k * RC), xb = ld4(xp + k * RC + 4);
          const float4 w = ld4(wp + k * W2C);
          const float2 w01 = make_float2(w.x, w.y), w23 = make_float2(w.z, w.w);
          const float xr[8] = {xa.x, xa.y, xa.z, xa.w, xb.x, xb.y, xb.z, xb.w};
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const float2 xx = make_float2(xr[i], xr[i]);
            acc[i][0] = __ffma2_rn(xx, w01, acc[i][0]);
            acc[i][1] = __ffma2_rn(xx, w23, acc[i][1]);
          }
        }
      }
      // both warps of the slice are done with it: the partial tile goes where the slice was
      // [kg][t64][row i][4 columns]   (32 floats per thread; KQ*64 >= 64*32 for H >= 256)
      asm volatile("bar.sync %0, 64;" ::"r"(kg + 1) : "memory");
      {
        float *pp = Xs + static_cast<size_t>(kg) * KQ * RC + t64 * 32;
#pragma unroll
        for (int i = 0; i < 8; ++i)
          *reinterpret_cast<float4 *>(pp + 4 * i) = make_float4(acc[i][0].x, acc[i][0].y, acc[i][1].x, acc[i][1].y);
      }
      __syncthreads();
      if (rc == 0) stamp(10);
      {
        float4 sum = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const float4 v = ld4(Xs + static_cast<size_t>(q) * KQ * RC + t64 * 32 + 4 * kg);
          sum.x += v.x; sum.y += v.y; sum.z += v.z; sum.w += v.w;
        }
        if (frow < a.rows) {
          const float sv[4] = {sum.x, sum.y, sum.z, sum.w};
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int cl = 4 * cg + i, c = c0 + cl;
            if (cl >= a.cpc || c >= a.NQ) continue;
            if (c < a.NQ1) {
              if (a.want_q) a.hq[static_cast<size_t>(frow) * a.ldq + c] = sv[i] + bqv[i];
            } else if (a.want_z) {
              a.z[static_cast<size_t>(frow) * a.E + (c - a.NQ1)] = tanhf(fmaf(0.5f, sv[i], za[i])) * zp[i];
            }
          }
        }
      }
      fence_proxy_async_smem();
      __syncthreads();
    }
  }
  stamp(11);
}

// slabs of CTA `blockIdx.x` from the K-major packed weights: WcI (4H + E, H) rows 4u+g | readout rows;
// WqT (8H + 4 + E, H)
__global__ void __launch_bounds__(256) cell_pack_kernel(const float *WcI, const float *WqT, float *W1, float *W2, int H,
                                                        int E, int upc, int epc, int cpc, int NQ, int ctx2out) {
  const int cta = blockIdx.x;
  float *w1 = W1 + static_cast<size_t>(cta) * H * W1C;
  float *w2 = W2 + static_cast<size_t>(cta) * H * W2C;
  for (int i = threadIdx.x; i < H * W1C; i += blockDim.x) {
    const int col = i / H, k = i - col * H;        // k fastest: coalesced reads
    float v = 0.f;
    if (col < 16) {
      const int s = col >> 2, g = col & 3, u = cta * upc + s;
      if (s < upc && u < H) v = WcI[(static_cast<size_t>(4) * u + g) * H + k];
    } else {
      const int s = col - 16, e = cta * epc + s;
      if (ctx2out && s < epc && e < E) v = WcI[(static_cast<size_t>(4) * H + e) * H + k];
    }
    w1[static_cast<size_t>(k) * W1C + col] = v;
  }
  for (int i = threadIdx.x; i < H * W2C; i += blockDim.x) {
    const int col = i / H, k = i - col * H;
    const int c = cta * cpc + col;
    w2[static_cast<size_t>(k) * W2C + col] = (col < cpc && c < NQ) ? WqT[static_cast<size_t>(c) * H + k] : 0.f;
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

size_t cell_smem(int H) { return (static_cast<size_t>(RC + W1C) * H + static_cast<size_t>(W2C) * (H / 4) * 3) * 4 + NBAR * 8 + 16; }

}  // namespace

// Shapes the cell kernel takes: K split 8 ways into float4-aligned slices, the partial tiles fit in the activation
// buffer (H >= 256), everything in 227 KB of shared memory (H <= 512), at most 4 units / 4 readout columns /
// 32 phase-2 columns per CTA.
bool cell_plan(int H, int E, CellPlan *out) {
  CellPlan p;
  p.grid = sm_count();
  p.upc = (H + p.grid - 1) / p.grid;
  p.epc = (E + p.grid - 1) / p.grid;
  p.NQ1 = 8 * H + 4;
  p.NQ = p.NQ1 + E;
  p.cpc = ((p.NQ + p.grid - 1) / p.grid + 3) & ~3;
  p.w1_floats = static_cast<size_t>(p.grid) * H * W1C;
  p.w2_floats = static_cast<size_t>(p.grid) * H * W2C;
  if (out) *out = p;
  return H % 32 == 0 && H >= 256 && H <= 512 && p.upc <= 4 && p.epc <= 4 && p.cpc <= W2C && cell_smem(H) <= 227 * 1024;
}

int cell_pack_launch(const float *WcI, const float *WqT, float *W1, float *W2, int H, int E, int ctx2out,
                     cudaStream_t stream) {
  CellPlan p;
  if (!cell_plan(H, E, &p)) return STAT_OK;      // unsupported shape: the slabs are never read
  cell_pack_kernel<<<p.grid, 256, 0, stream>>>(WcI, WqT, W1, W2, H, E, p.upc, p.epc, p.cpc, p.NQ, ctx2out);
  STAT_CUDA_CHECK(cudaGetLastError());
  note_launch();
  return STAT_OK;
}

int cell_launch(const CellLaunch &c, cudaStream_t stream) {
  CellPlan p;
  STAT_REQUIRE(cell_plan(c.H, c.E, &p), STAT_EINVAL, "cell: unsupported shape H=%d E=%d", c.H, c.E);
  CellArgs a;
  memset(&a, 0, sizeof(a));
  a.rows = c.rows; a.H = c.H; a.E = c.E; a.V = c.V;
  a.upc = p.upc; a.epc = p.epc; a.cpc = p.cpc; a.NQ1 = p.NQ1; a.NQ = p.NQ;
  a.do1 = c.do1; a.do2 = c.do2; a.want_q = c.want_q; a.want_z = c.want_z; a.prev2out = c.prev2out;
  a.W1 = c.W1; a.ctxT = c.ctxT; a.hu = c.hq + 4 * c.H + 4; a.ldq = c.ldq; a.EW = c.EW; a.Wemb = c.Wemb; a.bz = c.bz;
  a.tok_prev = c.tok_prev; a.mask = c.mask; a.dp_gates = c.dp_gates;
  a.h_in = c.h_in; a.c_in = c.c_in; a.h_out = c.h_out; a.c_out = c.c_out; a.h_all = c.h_all;
  a.hT = c.hT; a.zadd = c.zadd;
  a.W2 = c.W2; a.bq = c.bq; a.hq = c.hq_out ? c.hq_out : c.hq; a.z = c.z; a.dp_z = c.dp_z; a.bar = c.bar;
  {
    static unsigned long long pol = 0;
    if (pol == 0) {
      const char *e = getenv("STAT_CELL_L2");       // last (default) | normal | first
      pol = 0x14F0000000000000ull;
      if (e && !strcmp(e, "normal")) pol = 0x1000000000000000ull;
      if (e && !strcmp(e, "first")) pol = 0x12F0000000000000ull;
    }
    a.policy = pol;
  }
  a.trace = gemm_get_trace() ? gemm_get_trace() + 4096 : nullptr;      // behind the GEMM's own stamps
  const size_t smem = cell_smem(c.H);
  static size_t smem_set[STAT_MAX_DEV] = {};
  STAT_TRY(ensure_dyn_smem(cell_kernel, smem, smem_set));
  return launch_pdl(cell_kernel, dim3(p.grid), dim3(CT), smem, stream, a);
}

}  // namespace stat
