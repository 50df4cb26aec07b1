// k_att: the four soft-attentions of one decode step (model_attention.py:370-435,
// SURVEY App. A S1-S9) for every decode row, in one launch.
//
// Work split: grid = (S segments of Tc frames) x rows; one warp owns one frame
// at a time.  For its frame the warp
//   A. streams pctxl[t, 0..R) once, scores a_r = sum_h tanh(pctxl + h.Wdl) * Ul   (S1)
//   B. softmax over the R regions in registers                                    (S2)
//   C. streams ctxl0[t, r] and qctxl[t, r] = ctxl0.Wclt once, accumulating
//      cL = sum_r alpha_r ctxl0_r  and  pLT = sum_r alpha_r qctxl_r  (+ h.Wdlt + blt)
//      -- the per-step (B*T,H)x(H,H) GEMM of :416 folded away by linearity      (S3, S6)
//   D. scores the three temporal attentions for this frame                        (S4, S5, S7)
// then the CTA turns its Tc frames into flash-style partials (max, sum, weighted
// vector) for the G / M / LT attentions.  The last CTA of a row to finish (atomic
// ticket) merges the S partials, applies the selector gate and writes ctx (S8, S9).
//
// Bytes per row and step: (3*R + 4) * T * H * 4 streamed exactly once (pctxl,
// ctxl0, qctxl, pctxg, ctxg0, pctxm, ctxm0); everything else stays on chip.
#include "stat_common.cuh"
#include "kernels.cuh"

namespace stat {
namespace {

template <int VEC>
__device__ __forceinline__ void ld_chunk(const float *__restrict__ p, float (&o)[VEC], bool ok) {
  if constexpr (VEC == 4) {
    float4 v = ok ? __ldg(reinterpret_cast<const float4 *>(p)) : make_float4(0.f, 0.f, 0.f, 0.f);
    o[0] = v.x; o[1] = v.y; o[2] = v.z; o[3] = v.w;
  } else {
    o[0] = ok ? __ldg(p) : 0.f;
  }
}

template <int VEC, int NV, int RMAX>
__global__ void __launch_bounds__(512, 1) att_step_kernel(const AttArgs a) {
  extern __shared__ float smem[];
  const int H = a.H, T = a.T, R = a.R;
  const int seg = blockIdx.x, row = blockIdx.y;
  const int clip = a.row_clip ? a.row_clip[row] : row;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  const int t0 = seg * a.Tc;
  const int nf = min(a.Tc, T - t0);

  // smem: vec[8][H] | cL[Tc][H] | sa[3][Tc] | se[3][Tc] | ms[3][2] | flag
  float *s_vec = smem;
  float *s_cL = s_vec + 8 * H;
  float *s_a = s_cL + a.Tc * H;
  float *s_e = s_a + 3 * a.Tc;
  float *s_ms = s_e + 3 * a.Tc;
  int *s_flag = reinterpret_cast<int *>(s_ms + 6);

  const float *hp = a.hp + static_cast<size_t>(row) * a.ldhp;
  for (int i = threadIdx.x; i < H; i += blockDim.x) {
    // the h-projections arrive as k-slice planes: summed here in plane order
    float v0 = 0.f, v1 = 0.f, v2 = 0.f, v3 = 0.f;
    for (int q = 0; q < a.hp_parts; ++q) {
      const float *hq = hp + q * a.hp_plane;
      v0 += hq[a.off_sl + i];
      v1 += hq[a.off_sg + i];
      v2 += hq[a.off_sm + i];
      v3 += hq[a.off_slt + i];
    }
    s_vec[0 * H + i] = v0;
    s_vec[1 * H + i] = a.Ul[i];
    s_vec[2 * H + i] = v1;
    s_vec[3 * H + i] = a.Ug[i];
    s_vec[4 * H + i] = v2;
    s_vec[5 * H + i] = a.Um[i];
    s_vec[6 * H + i] = v3;
    s_vec[7 * H + i] = a.Ult[i];
  }
  __syncthreads();

  // per-lane column ownership: chunk j covers columns (j*32+lane)*VEC .. +VEC
  float sl[NV][VEC], ul[NV][VEC];
  bool ok[NV];
#pragma unroll
  for (int j = 0; j < NV; ++j) {
    const int col = (j * 32 + lane) * VEC;
    ok[j] = col < H;
#pragma unroll
    for (int v = 0; v < VEC; ++v) {
      sl[j][v] = ok[j] ? s_vec[0 * H + col + v] : 0.f;
      ul[j][v] = ok[j] ? s_vec[1 * H + col + v] : 0.f;
    }
  }
  const float cl = a.cl[0], cg = a.cg[0], cm = a.cm[0], clt = a.clt[0];

  for (int f = warp; f < nf; f += nwarps) {
    const int t = t0 + f;
    const size_t frame = static_cast<size_t>(clip) * T + t;
    const float *pL = a.pctxl + frame * R * H;
    const float *cL0 = a.ctxl0 + frame * R * H;
    const float *qL = a.qctxl + frame * R * H;

    // ---- A: spatial scores -------------------------------------------------
    float sc[RMAX];
#pragma unroll
    for (int r = 0; r < RMAX; ++r) {
      sc[r] = -INFINITY;
      if (r < R) {
        float part = 0.f;
#pragma unroll
        for (int j = 0; j < NV; ++j) {
          float x[VEC];
          ld_chunk<VEC>(pL + static_cast<size_t>(r) * H + (j * 32 + lane) * VEC, x, ok[j]);
#pragma unroll
          for (int v = 0; v < VEC; ++v) part = fmaf(tanh_fast(x[v] + sl[j][v]), ul[j][v], part);
        }
        sc[r] = warp_sum(part) + cl;
      }
    }
    // ---- B: softmax over regions -------------------------------------------
    float mx = sc[0];
#pragma unroll
    for (int r = 1; r < RMAX; ++r) mx = fmaxf(mx, sc[r]);
    float den = 0.f;
#pragma unroll
    for (int r = 0; r < RMAX; ++r) {
      sc[r] = (r < R) ? expf(sc[r] - mx) : 0.f;
      den += sc[r];
    }
    const float inv = 1.0f / den;
#pragma unroll
    for (int r = 0; r < RMAX; ++r) {
      sc[r] *= inv;
      if (a.alpha_l && r < R && lane == r)
        a.alpha_l[(static_cast<size_t>(row) * T + t) * R + r] = sc[r];
    }
    // ---- C: attended local context and its projection ------------------------
    float cL[NV][VEC], pLT[NV][VEC];
#pragma unroll
    for (int j = 0; j < NV; ++j)
#pragma unroll
      for (int v = 0; v < VEC; ++v) { cL[j][v] = 0.f; pLT[j][v] = 0.f; }
#pragma unroll
    for (int r = 0; r < RMAX; ++r) {
      if (r < R) {
#pragma unroll
        for (int j = 0; j < NV; ++j) {
          float x[VEC], q[VEC];
          const size_t off = static_cast<size_t>(r) * H + (j * 32 + lane) * VEC;
          ld_chunk<VEC>(cL0 + off, x, ok[j]);
          ld_chunk<VEC>(qL + off, q, ok[j]);
#pragma unroll
          for (int v = 0; v < VEC; ++v) {
            cL[j][v] = fmaf(sc[r], x[v], cL[j][v]);
            pLT[j][v] = fmaf(sc[r], q[v], pLT[j][v]);
          }
        }
      }
    }
    // ---- D: the three temporal scores of this frame ---------------------------
    float pg = 0.f, pm = 0.f, plt = 0.f;
    const float *pG = a.pctxg + frame * H;
    const float *pM = a.pctxm + frame * H;
#pragma unroll
    for (int j = 0; j < NV; ++j) {
      const int col = (j * 32 + lane) * VEC;
      float xg[VEC], xm[VEC];
      ld_chunk<VEC>(pG + col, xg, ok[j]);
      ld_chunk<VEC>(pM + col, xm, ok[j]);
      if (ok[j]) {
#pragma unroll
        for (int v = 0; v < VEC; ++v) {
          pg = fmaf(tanh_fast(xg[v] + s_vec[2 * H + col + v]), s_vec[3 * H + col + v], pg);
          pm = fmaf(tanh_fast(xm[v] + s_vec[4 * H + col + v]), s_vec[5 * H + col + v], pm);
          plt = fmaf(tanh_fast(pLT[j][v] + s_vec[6 * H + col + v]), s_vec[7 * H + col + v], plt);
          s_cL[f * H + col + v] = cL[j][v];
        }
      }
    }
    pg = warp_sum(pg) + cg;
    pm = warp_sum(pm) + cm;
    plt = warp_sum(plt) + clt;
    if (lane == 0) {
      s_a[0 * a.Tc + f] = pg;
      s_a[1 * a.Tc + f] = pm;
      s_a[2 * a.Tc + f] = plt;
      if (a.att_scores) {
        const size_t st = static_cast<size_t>(a.rows) * T;
        a.att_scores[0 * st + static_cast<size_t>(row) * T + t] = pg;
        a.att_scores[1 * st + static_cast<size_t>(row) * T + t] = pm;
        a.att_scores[2 * st + static_cast<size_t>(row) * T + t] = plt;
      }
    }
  }
  __syncthreads();

  // ---- segment partials: m = max a, e_f = exp(a_f - m), s = sum e ---------------
  if (threadIdx.x < 3) {
    const int k = threadIdx.x;
    float m = -INFINITY;
    for (int f = 0; f < nf; ++f) m = fmaxf(m, s_a[k * a.Tc + f]);
    float s = 0.f;
    for (int f = 0; f < nf; ++f) {
      const float e = expf(s_a[k * a.Tc + f] - m);
      s_e[k * a.Tc + f] = e;
      s += e;
    }
    s_ms[2 * k] = m;
    s_ms[2 * k + 1] = s;
  }
  __syncthreads();

  const float *G0 = a.ctxg0 + (static_cast<size_t>(clip) * T + t0) * H;
  const float *M0 = a.ctxm0 + (static_cast<size_t>(clip) * T + t0) * H;
  const bool single = (a.S == 1);
  float beta = 1.0f;
  if (a.selector) {
    float bsel = 0.f;
    for (int q = 0; q < a.hp_parts; ++q) bsel += hp[q * a.hp_plane + a.off_sel];
    beta = sigmoid_acc(bsel);
  }
  float *rv = a.rec_vec + (static_cast<size_t>(row) * a.S + seg) * 3 * H;
  for (int col = threadIdx.x; col < H; col += blockDim.x) {
    float ag = 0.f, am = 0.f, alt = 0.f;
    for (int f = 0; f < nf; ++f) {
      ag = fmaf(s_e[0 * a.Tc + f], __ldg(G0 + static_cast<size_t>(f) * H + col), ag);
      am = fmaf(s_e[1 * a.Tc + f], __ldg(M0 + static_cast<size_t>(f) * H + col), am);
      alt = fmaf(s_e[2 * a.Tc + f], s_cL[f * H + col], alt);
    }
    if (single) {
      a.ctx[static_cast<size_t>(row) * (a.ldctx ? a.ldctx : H) + col] = beta * (ag / s_ms[1] + am / s_ms[3] + alt / s_ms[5]);
    } else {
      rv[0 * H + col] = ag;
      rv[1 * H + col] = am;
      rv[2 * H + col] = alt;
    }
  }
  if (single) return;

  float *rms = a.rec_ms + (static_cast<size_t>(row) * a.S + seg) * 6;
  if (threadIdx.x < 6) rms[threadIdx.x] = s_ms[threadIdx.x];
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned int ticket = atomicAdd(a.counters + row, 1u);
    *s_flag = (ticket == static_cast<unsigned int>(a.S - 1)) ? 1 : 0;
  }
  __syncthreads();
  if (*s_flag == 0) return;
  __threadfence();

  // ---- last CTA of this row: merge the S partials --------------------------------
  // weights w[k][s] = exp(m_s - M_k) / sum_s exp(m_s - M_k) * s_s  (kept in s_e)
  const float *ms_all = a.rec_ms + static_cast<size_t>(row) * a.S * 6;
  if (threadIdx.x < 3) {
    const int k = threadIdx.x;
    float M = -INFINITY;
    for (int s = 0; s < a.S; ++s) M = fmaxf(M, __ldcg(ms_all + s * 6 + 2 * k));
    float den = 0.f;
    for (int s = 0; s < a.S; ++s)
      den += expf(__ldcg(ms_all + s * 6 + 2 * k) - M) * __ldcg(ms_all + s * 6 + 2 * k + 1);
    s_ms[2 * k] = M;
    s_ms[2 * k + 1] = den;
  }
  __syncthreads();
  const float *rv_all = a.rec_vec + static_cast<size_t>(row) * a.S * 3 * H;
  for (int col = threadIdx.x; col < H; col += blockDim.x) {
    float out = 0.f;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      float num = 0.f;
      for (int s = 0; s < a.S; ++s)
        num = fmaf(expf(__ldcg(ms_all + s * 6 + 2 * k) - s_ms[2 * k]),
                   __ldcg(rv_all + (static_cast<size_t>(s) * 3 + k) * H + col), num);
      out += num / s_ms[2 * k + 1];
    }
    a.ctx[static_cast<size_t>(row) * (a.ldctx ? a.ldctx : H) + col] = beta * out;
  }
  if (threadIdx.x == 0) a.counters[row] = 0u;
}

template <int VEC, int NV, int RMAX>
int launch(const AttArgs &a, cudaStream_t stream) {
  const int nwarps = a.Tc < 16 ? a.Tc : 16;
  const size_t smem = sizeof(float) * (8 * a.H + static_cast<size_t>(a.Tc) * a.H + 6 * a.Tc + 6) + 16;
  static size_t smem_set[STAT_MAX_DEV] = {};
  if (smem > 48 * 1024) STAT_TRY(ensure_dyn_smem(att_step_kernel<VEC, NV, RMAX>, smem, smem_set));
  dim3 grid(a.S, a.rows);
  att_step_kernel<VEC, NV, RMAX><<<grid, nwarps * 32, smem, stream>>>(a);
  note_launch();
  STAT_CUDA_CHECK(cudaGetLastError());
  return STAT_OK;
}

template <int VEC, int NV>
int launch_r(const AttArgs &a, cudaStream_t stream) {
  if (a.R <= 8) return launch<VEC, NV, 8>(a, stream);
  return launch<VEC, NV, 16>(a, stream);
}

}  // namespace

int att_step_launch(const AttArgs &a, cudaStream_t stream) {
  STAT_REQUIRE(a.R >= 1 && a.R <= 16, STAT_EINVAL, "att: R=%d outside [1,16]", a.R);
  STAT_REQUIRE(a.H >= 1 && a.H <= 1024, STAT_EINVAL, "att: H=%d outside [1,1024]", a.H);
  STAT_REQUIRE(static_cast<size_t>(a.Tc) * a.H * 4 <= 160 * 1024, STAT_EINVAL,
               "att: segment of %d frames x H=%d does not fit shared memory", a.Tc, a.H);
  const int H = a.H;
  if (H % 128 == 0) {
    switch (H / 128) {
      case 1: return launch_r<4, 1>(a, stream);
      case 2: return launch_r<4, 2>(a, stream);
      case 3: return launch_r<4, 3>(a, stream);
      case 4: return launch_r<4, 4>(a, stream);
      case 8: return launch_r<4, 8>(a, stream);
      default: break;
    }
  }
  const int nv = (H + 31) / 32;
  if (nv <= 1) return launch_r<1, 1>(a, stream);
  if (nv <= 2) return launch_r<1, 2>(a, stream);
  if (nv <= 4) return launch_r<1, 4>(a, stream);
  if (nv <= 8) return launch_r<1, 8>(a, stream);
  if (nv <= 16) return launch_r<1, 16>(a, stream);
  return launch_r<1, 32>(a, stream);
}

}  // namespace stat
