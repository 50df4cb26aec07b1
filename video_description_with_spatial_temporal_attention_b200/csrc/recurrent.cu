// Small per-step kernels around the attention and the GEMMs: LSTM gate update,
// vocabulary reduction (log-softmax statistics / argmax / target gather), and the
// one-off helpers of the prologue (mean pool, weight transposes).
#include "stat_common.cuh"
#include "kernels.cuh"

namespace stat {
namespace {

// ---------------------------------------------------------------------------
// gates: S10-S13 of SURVEY App. A  (model_attention.py:437-457)
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(128) gates_kernel(const GateArgs a) {
  pdl_wait();
  pdl_trigger();
  const int row = blockIdx.y;
  const int j = blockIdx.x * 128 + threadIdx.x;
  const long long tok = a.tok_prev ? a.tok_prev[row] : -1;
  const int H = a.H;
  if (j < H) {
    const float *pc = a.pre_c + static_cast<size_t>(row) * a.ldpc;
    const float *u = a.hp + static_cast<size_t>(row) * a.ldhp + a.off_u;
    // k-slice planes of the two projections, summed in plane order (deterministic)
    float su[4] = {0.f, 0.f, 0.f, 0.f}, sc[4] = {0.f, 0.f, 0.f, 0.f};
    // (fixed trip count + predicate so that the loads of all planes are in flight together)
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      if (q < a.hp_parts) {
#pragma unroll
        for (int gi = 0; gi < 4; ++gi) su[gi] += u[q * a.hp_plane + gi * H + j];
      }
      if (q < a.pc_parts) {
#pragma unroll
        for (int gi = 0; gi < 4; ++gi) sc[gi] += pc[q * a.pc_plane + gi * H + j];
      }
    }
    const float *ew = a.EW + static_cast<size_t>(tok >= 0 ? tok : a.V) * 4 * H;
    float di = 0.5f, df = 0.5f, dO = 0.5f;
    if (a.dp_gates) {
      const float *dp = a.dp_gates + static_cast<size_t>(row) * 3 * H;
      di = dp[j];
      df = dp[H + j];
      dO = dp[2 * H + j];
    }
    // the token table is gate-interleaved: the four gate inputs of unit j are adjacent
    const float4 e4 = __ldg(reinterpret_cast<const float4 *>(ew) + j);
    const float pi = (su[0] + e4.x) + sc[0];
    const float pf = (su[1] + e4.y) + sc[1];
    const float po = (su[2] + e4.z) + sc[2];
    const float pg = (su[3] + e4.w) + sc[3];
    const float ig = sigmoid_acc(pi * di);
    const float fg = sigmoid_acc(pf * df);
    const float og = sigmoid_acc(po * dO);
    const float gg = tanhf(pg);
    const size_t idx = static_cast<size_t>(row) * H + j;
    const float c_ = a.c_in[idx], h_ = a.h_in[idx];
    const float m = a.mask ? a.mask[row] : 1.0f;
    float c = fg * c_ + ig * gg;
    c = m * c + (1.0f - m) * c_;
    float h = og * tanhf(c);
    h = m * h + (1.0f - m) * h_;
    a.c_out[idx] = c;
    a.h_out[idx] = h;
    if (a.h_all) a.h_all[idx] = h;
    if (a.dp_h) a.hd_out[idx] = h * a.dp_h[idx];
  }
  if (j < a.E) {
    // everything of the readout pre-activation that does not depend on the new h (:689-693)
    float z = a.bz[j];
    if (a.zc_off >= 0) {
#pragma unroll
      for (int q = 0; q < 8; ++q)
        if (q < a.pc_parts) z += a.pre_c[q * a.pc_plane + static_cast<size_t>(row) * a.ldpc + a.zc_off + j];
    }
    if (a.prev2out && tok >= 0) z += __ldg(a.Wemb + static_cast<size_t>(tok) * a.E + j);
    a.zadd[static_cast<size_t>(row) * a.E + j] = z;
  }
}

__global__ void __launch_bounds__(256) zact_kernel(const ZactArgs a) {
  pdl_wait();
  pdl_trigger();
  const int i = blockIdx.x * 256 + threadIdx.x;
  if (i >= a.rows * a.E) return;
  const int row = i / a.E, e = i - row * a.E;
  float s = 0.f;
#pragma unroll
  for (int q = 0; q < 8; ++q)
    if (q < a.parts) s += a.zpre[q * a.plane + static_cast<size_t>(row) * a.ldz + e];
  const float v = tanhf(fmaf(a.alpha, s, a.zadd[i]));
  a.z[i] = v * (a.dp_z ? a.dp_z[i] : 0.5f);
}

// gates for even H: two adjacent units per thread, every load of the thread (<= 4 + 4 k-slice
// planes x 4 gates, the token's table row, the old state) issued before the first use
__global__ void __launch_bounds__(128) gates2_kernel(const GateArgs a) {
  pdl_wait();
  pdl_trigger();
  const int row = blockIdx.y;
  const int j = (blockIdx.x * 128 + threadIdx.x) * 2;
  const long long tok = a.tok_prev ? a.tok_prev[row] : -1;
  const int H = a.H;
  if (j < H) {
    const float *pc = a.pre_c + static_cast<size_t>(row) * a.ldpc + j;
    const float *u = a.hp + static_cast<size_t>(row) * a.ldhp + a.off_u + j;
    const float *ew = a.EW + static_cast<size_t>(tok >= 0 ? tok : a.V) * 4 * H + 4 * j;
    float2 xu[4][4], xc[4][4], xe[4];
#pragma unroll
    for (int q = 0; q < 4; ++q)
#pragma unroll
      for (int gi = 0; gi < 4; ++gi) {
        xu[q][gi] = (q < a.hp_parts) ? *reinterpret_cast<const float2 *>(u + q * a.hp_plane + gi * H)
                                     : make_float2(0.f, 0.f);
        xc[q][gi] = (q < a.pc_parts) ? *reinterpret_cast<const float2 *>(pc + q * a.pc_plane + gi * H)
                                     : make_float2(0.f, 0.f);
      }
    {
      // gate-interleaved token table: units j, j+1 = 8 adjacent floats
      const float4 ea = __ldg(reinterpret_cast<const float4 *>(ew)), eb = __ldg(reinterpret_cast<const float4 *>(ew) + 1);
      xe[0] = make_float2(ea.x, eb.x); xe[1] = make_float2(ea.y, eb.y);
      xe[2] = make_float2(ea.z, eb.z); xe[3] = make_float2(ea.w, eb.w);
    }
    const size_t idx = static_cast<size_t>(row) * H + j;
    const float2 c_ = *reinterpret_cast<const float2 *>(a.c_in + idx);
    const float2 h_ = *reinterpret_cast<const float2 *>(a.h_in + idx);
    const float m = a.mask ? a.mask[row] : 1.0f;
    float2 dp[3] = {make_float2(0.5f, 0.5f), make_float2(0.5f, 0.5f), make_float2(0.5f, 0.5f)};
    if (a.dp_gates) {
      const float *d = a.dp_gates + static_cast<size_t>(row) * 3 * H + j;
#pragma unroll
      for (int gi = 0; gi < 3; ++gi) dp[gi] = *reinterpret_cast<const float2 *>(d + gi * H);
    }
    float2 pre[4];
#pragma unroll
    for (int gi = 0; gi < 4; ++gi) {
      float sux = 0.f, suy = 0.f, scx = 0.f, scy = 0.f;
#pragma unroll
      for (int q = 0; q < 4; ++q) {          // plane order: deterministic
        sux += xu[q][gi].x; suy += xu[q][gi].y;
        scx += xc[q][gi].x; scy += xc[q][gi].y;
      }
      pre[gi].x = (sux + xe[gi].x) + scx;
      pre[gi].y = (suy + xe[gi].y) + scy;
    }
    float2 c, h;
    {
      const float ig = sigmoid_acc(pre[0].x * dp[0].x), fg = sigmoid_acc(pre[1].x * dp[1].x);
      const float og = sigmoid_acc(pre[2].x * dp[2].x), gg = tanhf(pre[3].x);
      c.x = fg * c_.x + ig * gg;
      c.x = m * c.x + (1.0f - m) * c_.x;
      h.x = og * tanhf(c.x);
      h.x = m * h.x + (1.0f - m) * h_.x;
    }
    {
      const float ig = sigmoid_acc(pre[0].y * dp[0].y), fg = sigmoid_acc(pre[1].y * dp[1].y);
      const float og = sigmoid_acc(pre[2].y * dp[2].y), gg = tanhf(pre[3].y);
      c.y = fg * c_.y + ig * gg;
      c.y = m * c.y + (1.0f - m) * c_.y;
      h.y = og * tanhf(c.y);
      h.y = m * h.y + (1.0f - m) * h_.y;
    }
    *reinterpret_cast<float2 *>(a.c_out + idx) = c;
    *reinterpret_cast<float2 *>(a.h_out + idx) = h;
    if (a.h_all) *reinterpret_cast<float2 *>(a.h_all + idx) = h;
    if (a.dp_h) {
      const float2 dh = *reinterpret_cast<const float2 *>(a.dp_h + idx);
      *reinterpret_cast<float2 *>(a.hd_out + idx) = make_float2(h.x * dh.x, h.y * dh.y);
    }
  }
  // everything of the readout pre-activation that does not depend on the new h (:689-693)
  for (int e = j; e < min(j + 2, a.E); ++e) {
    float z = a.bz[e];
    if (a.zc_off >= 0) {
#pragma unroll
      for (int q = 0; q < 4; ++q)
        if (q < a.pc_parts) z += a.pre_c[q * a.pc_plane + static_cast<size_t>(row) * a.ldpc + a.zc_off + e];
    }
    if (a.prev2out && tok >= 0) z += __ldg(a.Wemb + static_cast<size_t>(tok) * a.E + e);
    a.zadd[static_cast<size_t>(row) * a.E + e] = z;
  }
}

// ---------------------------------------------------------------------------
// vocabulary reduction
// ---------------------------------------------------------------------------
// One CTA per row: two rolled passes over the V logits (maximum / arg-max, then the sum of exponentials), the CTA
// merges the per-thread results.
__global__ void __launch_bounds__(256) pick_kernel(const PickArgs a) {
  pdl_wait();
  pdl_trigger();
  __shared__ float s_val[8];
  __shared__ int s_idx[8];
  __shared__ float s_sum[8];
  __shared__ float s_bm;
  __shared__ int s_bi;
  __shared__ float s_bs;
  const int row = blockIdx.x;
  const float *l = a.logits + static_cast<size_t>(row) * a.ldl;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float m = -INFINITY, s = 0.f;
  int bi = 0x7fffffff;
  // Two short rolled passes over the row (50 KB: the second one hits L1 / L2): (1) maximum and first arg-max,
  // branch-free; (2) sum of exponentials against the thread's own maximum.  A first version kept 64 logits per
  // thread in registers with a data-dependent branch per element: 9-12 us per step, most of it divergence and cold
  // instruction fetches of the unrolled body.
  const int n4 = ((a.ldl & 3) == 0) ? (a.V >> 2) : 0;
  {
    const float4 *l4 = reinterpret_cast<const float4 *>(l);
#pragma unroll 4
    for (int k = threadIdx.x; k < n4; k += 256) {
      const float4 x = l4[k];
      const float mx = fmaxf(fmaxf(x.x, x.y), fmaxf(x.z, x.w));
      if (mx > m) {                // ascending k per thread: the first maximum wins ties
        m = mx;
        bi = 4 * k + (x.x == mx ? 0 : (x.y == mx ? 1 : (x.z == mx ? 2 : 3)));
      }
    }
    for (int v = 4 * n4 + threadIdx.x; v < a.V; v += 256) {
      const float x = l[v];
      if (x > m || (x == m && v < bi)) { m = x; bi = v; }
    }
    if (m > -INFINITY) {
#pragma unroll 4
      for (int k = threadIdx.x; k < n4; k += 256) {
        const float4 x = l4[k];
        s += (__expf(x.x - m) + __expf(x.y - m)) + (__expf(x.z - m) + __expf(x.w - m));
      }
      for (int v = 4 * n4 + threadIdx.x; v < a.V; v += 256) s += __expf(l[v] - m);
    }
  }
  // merge (max, arg-max, sum) over the warp, then over the 8 warps: larger value, then lower index
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float om = __shfl_xor_sync(0xffffffffu, m, o);
    const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
    const float os = __shfl_xor_sync(0xffffffffu, s, o);
    const float mn = fmaxf(m, om);
    // idle threads carry (-inf, 0): -inf - -inf would be NaN, so their weight is forced to 0
    const float w1 = (m == -INFINITY) ? 0.f : expf(m - mn);
    const float w2 = (om == -INFINITY) ? 0.f : expf(om - mn);
    s = s * w1 + os * w2;
    if (om > m || (om == m && oi < bi)) bi = oi;
    m = mn;
  }
  if (lane == 0) { s_val[warp] = m; s_idx[warp] = bi; s_sum[warp] = s; }
  __syncthreads();
  if (threadIdx.x == 0) {
    float b = s_val[0];
    int i = s_idx[0];
    for (int w = 1; w < 8; ++w)
      if (s_val[w] > b || (s_val[w] == b && s_idx[w] < i)) { b = s_val[w]; i = s_idx[w]; }
    float tot = 0.f;
    for (int w = 0; w < 8; ++w) tot += (s_val[w] == -INFINITY) ? 0.f : s_sum[w] * expf(s_val[w] - b);
    s_bm = b;
    s_bi = i;
    s_bs = tot;
    const int tok = i;
    if (a.tokens) {
      const bool live = a.alive[row] != 0;
      a.tokens[static_cast<size_t>(row) * a.maxlen + a.t] = live ? tok : -1;
      if (live) {
        a.scores[row] += logf(tot);              // -log p(argmax) = log sum exp(l - max)
        a.lengths[row] = a.t + 1;
        a.alive[row] = tok != 0;
        a.tok_prev[row] = tok;
      }
    }
    if (a.x_t) {
      const float p = expf(l[a.x_t[row]] - b) / tot;
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
  if (a.probs) {
    __syncthreads();
    const float mm = s_bm, inv = 1.0f / s_bs;
    float *p = a.probs + static_cast<size_t>(row) * a.V;
    for (int v = threadIdx.x; v < a.V; v += 256) p[v] = expf(l[v] - mm) * inv;
  }
  if (a.beam_k > 0) {
    // The cost score - log p[v] falls as the logit rises: the beam_k cheapest continuations of the row
    // are its beam_k largest logits (lower index first among equals), found by beam_k arg-max rounds.
    __shared__ int s_taken[BEAM_KMAX];
    __syncthreads();
    float *cc = a.cand_cost + static_cast<size_t>(row) * BEAM_KMAX;
    int32_t *cw = a.cand_word + static_cast<size_t>(row) * BEAM_KMAX;
    if (!a.row_alive[row]) {
      if (threadIdx.x < BEAM_KMAX) { cc[threadIdx.x] = INFINITY; cw[threadIdx.x] = 0; }
      return;
    }
    const float mm = s_bm, inv = 1.0f / s_bs, sc = a.row_score[row];
    // Every thread holds its own words (v = tid, tid + 256, ...) in registers -- one batch of independent
    // loads -- and keeps the best not-yet-taken one; after a round only the owner of the winner looks at
    // its words again.  Vocabularies beyond 256 * PICK_NB words re-read theirs from L2 instead.
    constexpr int PICK_NB = 64;
    const bool in_regs = a.V <= 256 * PICK_NB;
    float xv[PICK_NB];
    float bm = -INFINITY;
    int bv = 0x7fffffff;
    if (in_regs) {
#pragma unroll
      for (int s = 0; s < PICK_NB; ++s) {
        const int v = threadIdx.x + 256 * s;
        xv[s] = v < a.V ? l[v] : -INFINITY;
      }
#pragma unroll
      for (int s = 0; s < PICK_NB; ++s)
        if (xv[s] > bm) { bm = xv[s]; bv = threadIdx.x + 256 * s; }
    } else {
      for (int v = threadIdx.x; v < a.V; v += 256) {
        const float x = l[v];
        if (x > bm) { bm = x; bv = v; }
      }
    }
    for (int r = 0; r < a.beam_k; ++r) {
      float wm = bm;
      int wv = bv;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const float om = __shfl_xor_sync(0xffffffffu, wm, o);
        const int ov = __shfl_xor_sync(0xffffffffu, wv, o);
        if (om > wm || (om == wm && ov < wv)) { wm = om; wv = ov; }
      }
      __syncthreads();                       // previous round's readers of s_val / s_idx are done
      if (lane == 0) { s_val[warp] = wm; s_idx[warp] = wv; }
      __syncthreads();
      float b = s_val[0];
      int i = s_idx[0];
#pragma unroll
      for (int w = 1; w < 8; ++w)
        if (s_val[w] > b || (s_val[w] == b && s_idx[w] < i)) { b = s_val[w]; i = s_idx[w]; }
      const bool ok = i < a.V;               // V < beam_k: no candidate left
      if (threadIdx.x == 0) {
        s_taken[r] = i;
        cc[r] = ok ? sc - logf(expf(b - mm) * inv) : INFINITY;   // as the reference: -log of the fp32 probability
        cw[r] = ok ? i : 0;
      }
      if (ok && (i & 255) == static_cast<int>(threadIdx.x)) {
        // my best word was taken: next best of my words
        bm = -INFINITY;
        bv = 0x7fffffff;
        if (in_regs) {
          const int slot = i >> 8;
#pragma unroll
          for (int s = 0; s < PICK_NB; ++s) {
            if (s == slot) xv[s] = -INFINITY;
            if (xv[s] > bm) { bm = xv[s]; bv = threadIdx.x + 256 * s; }
          }
        } else {
          for (int v = threadIdx.x; v < a.V; v += 256) {
            bool taken = (v == i);
            for (int q = 0; q < r; ++q) taken |= (s_taken[q] == v);
            const float x = l[v];
            if (!taken && x > bm) { bm = x; bv = v; }
          }
        }
      }
    }
    if (threadIdx.x >= a.beam_k && threadIdx.x < BEAM_KMAX) { cc[threadIdx.x] = INFINITY; cw[threadIdx.x] = 0; }
  }
}

// The same reduction with the row held in registers: NT threads, NJ float4 of the row per thread (float4
// k = tid + NT j: coalesced), every load issued before the first value is used -- one memory round trip
// instead of two rolled passes -- and every later step (maximum, sum of exponentials, probabilities, the
// beam_k arg-max rounds of the beam search) works on registers.  Needs V <= 4 NT NJ and 16-byte aligned
// rows; other shapes take pick_kernel.  Ties: the lowest index wins (a thread walks its words in ascending
// order, merges compare (value, index)).
struct PickTriple {
  float m, s;
  int i;
};
__device__ __forceinline__ PickTriple pick_merge_warp(PickTriple t) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float om = __shfl_xor_sync(0xffffffffu, t.m, o);
    const int oi = __shfl_xor_sync(0xffffffffu, t.i, o);
    const float os = __shfl_xor_sync(0xffffffffu, t.s, o);
    const float mn = fmaxf(t.m, om);
    // empty partners carry (-inf, 0): -inf - -inf would be NaN, so their weight is forced to 0
    const float w1 = (t.m == -INFINITY) ? 0.f : expf(t.m - mn);
    const float w2 = (om == -INFINITY) ? 0.f : expf(om - mn);
    t.s = t.s * w1 + os * w2;
    if (om > t.m || (om == t.m && oi < t.i)) t.i = oi;
    t.m = mn;
  }
  return t;
}
__device__ __forceinline__ void argmax_merge_warp(float &m, int &i) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float om = __shfl_xor_sync(0xffffffffu, m, o);
    const int oi = __shfl_xor_sync(0xffffffffu, i, o);
    if (om > m || (om == m && oi < i)) { m = om; i = oi; }
  }
}

template <int NT, int NJ>
__global__ void __launch_bounds__(NT) pick_regs_kernel(const PickArgs a) {
  pdl_wait();
  pdl_trigger();
  constexpr int NW = NT / 32;
  __shared__ float s_val[2][NW];
  __shared__ int s_idx[2][NW];
  __shared__ float s_sum[NW];
  __shared__ float s_bm, s_bs;
  const int row = blockIdx.x, tid = threadIdx.x;
  const int lane = tid & 31, warp = tid >> 5;
  const float *l = a.logits + static_cast<size_t>(row) * a.ldl;
  const int nf4 = (a.V + 3) >> 2;       // ldl is a multiple of 4 and >= V: the last (partial) float4 is inside the row
  float x[NJ][4];
  {
    const float4 *l4 = reinterpret_cast<const float4 *>(l);
    float4 v[NJ];
#pragma unroll
    for (int j = 0; j < NJ; ++j) {
      const int k = tid + NT * j;
      v[j] = k < nf4 ? l4[k] : make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
    }
#pragma unroll
    for (int j = 0; j < NJ; ++j) {
      const int w0 = 4 * (tid + NT * j);
      x[j][0] = v[j].x;                                   // w0 < V whenever the float4 was loaded
      x[j][1] = w0 + 1 < a.V ? v[j].y : -INFINITY;
      x[j][2] = w0 + 2 < a.V ? v[j].z : -INFINITY;
      x[j][3] = w0 + 3 < a.V ? v[j].w : -INFINITY;
    }
  }
  PickTriple t;
  t.m = -INFINITY; t.s = 0.f; t.i = 0x7fffffff;
#pragma unroll
  for (int j = 0; j < NJ; ++j)
#pragma unroll
    for (int e = 0; e < 4; ++e)
      if (x[j][e] > t.m) { t.m = x[j][e]; t.i = 4 * (tid + NT * j) + e; }
  if (t.m > -INFINITY) {
#pragma unroll
    for (int j = 0; j < NJ; ++j)
      t.s += (__expf(x[j][0] - t.m) + __expf(x[j][1] - t.m)) + (__expf(x[j][2] - t.m) + __expf(x[j][3] - t.m));
  }
  const float own_m = t.m;
  const int own_i = t.i;
  t = pick_merge_warp(t);
  if (lane == 0) { s_val[0][warp] = t.m; s_idx[0][warp] = t.i; s_sum[warp] = t.s; }
  __syncthreads();
  if (warp == 0) {
    t.m = lane < NW ? s_val[0][lane] : -INFINITY;
    t.i = lane < NW ? s_idx[0][lane] : 0x7fffffff;
    t.s = lane < NW ? s_sum[lane] : 0.f;
    t = pick_merge_warp(t);
    if (lane == 0) {
      const float b = t.m, tot = t.s;
      const int tok = t.i;
      s_bm = b;
      s_bs = tot;
      if (a.tokens) {
        const bool live = a.alive[row] != 0;
        a.tokens[static_cast<size_t>(row) * a.maxlen + a.t] = live ? tok : -1;
        if (live) {
          a.scores[row] += logf(tot);              // -log p(argmax) = log sum exp(l - max)
          a.lengths[row] = a.t + 1;
          a.alive[row] = tok != 0;
          a.tok_prev[row] = tok;
        }
      }
      if (a.x_t) {
        const float p = expf(l[a.x_t[row]] - b) / tot;
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
  if (!a.probs && a.beam_k <= 0) return;
  __syncthreads();
  const float mm = s_bm, inv = 1.0f / s_bs;
  if (a.probs) {
    float *p = a.probs + static_cast<size_t>(row) * a.V;
#pragma unroll
    for (int j = 0; j < NJ; ++j)
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int v = 4 * (tid + NT * j) + e;
        if (v < a.V) p[v] = expf(x[j][e] - mm) * inv;
      }
  }
  if (a.beam_k > 0) {
    // The cost score - log p[v] falls as the logit rises: the beam_k cheapest continuations of the row are its
    // beam_k largest logits (lower index first among equals), found by beam_k arg-max rounds.  Every thread offers
    // the best of its not-yet-taken words; after a round only the owner of the winner looks at its words again.
    // One barrier per round: the per-warp winners alternate between two shared-memory buffers, every warp
    // reduces them for itself.
    float *cc = a.cand_cost + static_cast<size_t>(row) * BEAM_KMAX;
    int32_t *cw = a.cand_word + static_cast<size_t>(row) * BEAM_KMAX;
    if (!a.row_alive[row]) {
      if (tid < BEAM_KMAX) { cc[tid] = INFINITY; cw[tid] = 0; }
      return;
    }
    const float sc = a.row_score[row];
    float bm = own_m;
    int bv = own_i;
    for (int r = 0; r < a.beam_k; ++r) {
      const int par = (r + 1) & 1;          // buffer 0 was last read before the barrier above
      float wm = bm;
      int wv = bv;
      argmax_merge_warp(wm, wv);
      if (lane == 0) { s_val[par][warp] = wm; s_idx[par][warp] = wv; }
      __syncthreads();
      float b = lane < NW ? s_val[par][lane] : -INFINITY;
      int i = lane < NW ? s_idx[par][lane] : 0x7fffffff;
      argmax_merge_warp(b, i);
      const bool ok = i < a.V;               // V < beam_k: no candidate left
      if (tid == 0) {
        cc[r] = ok ? sc - logf(expf(b - mm) * inv) : INFINITY;   // as the reference: -log of the fp32 probability
        cw[r] = ok ? i : 0;
      }
      if (ok && ((i >> 2) & (NT - 1)) == tid) {
        // my best word was taken: next best of my words
        bm = -INFINITY;
        bv = 0x7fffffff;
#pragma unroll
        for (int j = 0; j < NJ; ++j)
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const int v = 4 * (tid + NT * j) + e;
            if (v == i) x[j][e] = -INFINITY;
            if (x[j][e] > bm) { bm = x[j][e]; bv = v; }
          }
      }
    }
    if (tid >= a.beam_k && tid < BEAM_KMAX) { cc[tid] = INFINITY; cw[tid] = 0; }
  }
}

// ---------------------------------------------------------------------------
// beam search bookkeeping (model_attention.py:852-994), one CTA per clip
// ---------------------------------------------------------------------------
// One CTA per clip: warp 0 does the bookkeeping (every global read it needs is issued up front: one memory
// round trip), then the whole CTA gathers the LSTM state (and, for the cell step, the h-products) of the clip's new
// row slots from the rows they continue -- the k slots of a clip only ever continue rows of the same clip.
constexpr int BEAM_SELECT_NT = 1024;
__global__ void __launch_bounds__(BEAM_SELECT_NT) beam_select_kernel(const BeamArgs a) {
  pdl_wait();
  pdl_trigger();
  __shared__ int s_hist[BEAM_KMAX][BEAM_LMAX];
  __shared__ int s_len[BEAM_KMAX];
  __shared__ int s_src[BEAM_KMAX], s_word[BEAM_KMAX];
  __shared__ float s_cost[BEAM_KMAX];
  __shared__ int s_nsel;
  __shared__ float s_cc[BEAM_KMAX][BEAM_KMAX];
  __shared__ int s_cw[BEAM_KMAX][BEAM_KMAX];
  __shared__ int s_alive[BEAM_KMAX];
  __shared__ int s_from[BEAM_KMAX];           // row whose state slot j continues, -1: nothing to gather
  const int b = blockIdx.x, tid = threadIdx.x, k = a.k;
  const int r0 = b * k;
  const int was_done = a.done[b];
  if (tid < 32) {
    const int lane = tid;
    if (was_done) {
      for (int j = lane; j < k; j += 32) { a.src_row[r0 + j] = r0 + j; a.tok_prev[r0 + j] = 0; s_from[j] = -1; }
    } else {
      // old histories (a hypothesis has at most t words before step t), candidates and flags of the clip's slots into
      // shared memory with all lanes: the selection below walks them serially, from global memory every probe
      // would be a dependent L2 round trip
      const int tmax = min(a.t, BEAM_LMAX);
      for (int e = lane; e < k * tmax; e += 32) {
        const int j = e / tmax, i = e - j * tmax;
        s_hist[j][i] = a.hist[static_cast<size_t>(r0 + j) * BEAM_LMAX + i];
      }
      for (int e = lane; e < k * k; e += 32) {
        const int j = e / k, c = e - j * k;
        s_cc[j][c] = a.cand_cost[static_cast<size_t>(r0 + j) * BEAM_KMAX + c];
        s_cw[j][c] = a.cand_word[static_cast<size_t>(r0 + j) * BEAM_KMAX + c];
      }
      for (int j = lane; j < k; j += 32) {
        s_alive[j] = a.alive[r0 + j];
        s_len[j] = a.hist_len[r0 + j];
      }
      const int dead0 = a.dead_k[b], nout0 = a.out_count[b];
      __syncwarp();
      if (lane == 0) {
        // the k - dead cheapest of the live slots' candidates: ascending cost, lower flat index (slot*V + word) first
        const int want = k - dead0;
        int head[BEAM_KMAX];                      // every slot's candidates are already in ascending order
        for (int j = 0; j < k; ++j) head[j] = 0;
        int n = 0;
        for (; n < want; ++n) {
          float bc = INFINITY;
          int bj = -1;
          long long bf = 0;
          for (int j = 0; j < k; ++j) {
            if (!s_alive[j] || head[j] >= k) continue;
            const float c = s_cc[j][head[j]];
            const long long f = static_cast<long long>(j) * a.V + s_cw[j][head[j]];
            if (c < bc || (c == bc && bj >= 0 && f < bf)) { bc = c; bj = j; bf = f; }
          }
          if (bj < 0 || bc == INFINITY) break;
          s_src[n] = bj;
          s_word[n] = s_cw[bj][head[bj]];
          s_cost[n] = bc;
          ++head[bj];
        }
        s_nsel = n;
      }
      __syncwarp();
      const int nsel = s_nsel;
      // new hypotheses in rank order: token 0 retires one, the others become the live slots 0, 1, ...
      int n_live = 0, n_out = nout0, dead = dead0;
      for (int n = 0; n < nsel; ++n) {
        const int src = s_src[n], word = s_word[n], len = s_len[src];
        if (word == 0) {
          int64_t *ot = a.out_tokens + (static_cast<size_t>(b) * k + n_out) * a.maxlen;
          for (int i = lane; i < a.maxlen; i += 32) ot[i] = i < len ? s_hist[src][i] : (i == len ? 0 : -1);
          if (lane == 0) { a.out_lengths[b * k + n_out] = len + 1; a.out_scores[b * k + n_out] = s_cost[n]; }
          ++n_out;
          ++dead;
        } else {
          const int row = r0 + n_live;
          int32_t *h = a.hist + static_cast<size_t>(row) * BEAM_LMAX;
          for (int i = lane; i < len; i += 32) h[i] = s_hist[src][i];
          if (lane == 0) {
            h[len] = word;
            a.hist_len[row] = len + 1;
            a.score[row] = s_cost[n];
            a.src_row[row] = r0 + src;
            a.tok_prev[row] = word;
            a.alive[row] = 1;
            s_from[n_live] = r0 + src;
          }
          ++n_live;
        }
      }
      __syncwarp();
      const bool finished = n_live < 1 || dead >= k || a.t + 1 >= a.maxlen;   // :969-972 and the end of the loop
      if (finished) {
        // the survivors follow the retired hypotheses (:975-979)
        for (int j = 0; j < n_live; ++j) {
          const int row = r0 + j;
          const int len = a.hist_len[row];
          int64_t *ot = a.out_tokens + (static_cast<size_t>(b) * k + n_out) * a.maxlen;
          for (int i = lane; i < a.maxlen; i += 32) ot[i] = i < len ? a.hist[static_cast<size_t>(row) * BEAM_LMAX + i] : -1;
          if (lane == 0) { a.out_lengths[b * k + n_out] = len; a.out_scores[b * k + n_out] = a.score[row]; }
          ++n_out;
        }
        n_live = 0;
      }
      for (int j = n_live + lane; j < k; j += 32) {
        a.alive[r0 + j] = 0;
        a.src_row[r0 + j] = r0 + j;
        a.tok_prev[r0 + j] = 0;
        s_from[j] = -1;
      }
      if (finished) for (int j = lane; j < k; j += 32) s_from[j] = -1;
      if (lane == 0) {
        a.out_count[b] = n_out;
        a.dead_k[b] = dead;
        a.done[b] = finished ? 1 : 0;
      }
    }
  }
  if (!a.dst_h) return;
  __syncthreads();
  // state (and h-products) of the live slots in their new order; the dead slots of a clip are never read again.
  // Four independent 16-byte loads per thread go out before the first store: the rows of a clip (k x 20 KB at H = 512)
  // move in two or three memory round trips.
  if ((a.H & 3) == 0) {
    const int H4 = a.H >> 2, Q4 = a.dst_q ? (a.ldq >> 2) : 0;
    const int per_row = 2 * H4 + Q4, total = k * per_row;
    for (int base = 0; base < total; base += 4 * BEAM_SELECT_NT) {
      float4 v[4];
      float4 *dst[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int e = base + u * BEAM_SELECT_NT + tid;
        dst[u] = nullptr;
        if (e >= total) continue;
        const int j = e / per_row, i = e - j * per_row;
        const int from = s_from[j];
        if (from < 0) continue;
        const size_t so = static_cast<size_t>(from), dd = static_cast<size_t>(r0 + j);
        const float *sp;
        float *dp;
        if (i < H4) { sp = a.src_h + so * a.H + 4 * i; dp = a.dst_h + dd * a.H + 4 * i; }
        else if (i < 2 * H4) { sp = a.src_c + so * a.H + 4 * (i - H4); dp = a.dst_c + dd * a.H + 4 * (i - H4); }
        else { sp = a.src_q + so * a.ldq + 4 * (i - 2 * H4); dp = a.dst_q + dd * a.ldq + 4 * (i - 2 * H4); }
        v[u] = *reinterpret_cast<const float4 *>(sp);
        dst[u] = reinterpret_cast<float4 *>(dp);
      }
#pragma unroll
      for (int u = 0; u < 4; ++u)
        if (dst[u]) *dst[u] = v[u];
    }
  } else {
    for (int j = 0; j < k; ++j) {
      const int from = s_from[j];
      if (from < 0) continue;
      const size_t so = static_cast<size_t>(from), dd = static_cast<size_t>(r0 + j);
      for (int i = tid; i < a.H; i += BEAM_SELECT_NT) {
        a.dst_h[dd * a.H + i] = a.src_h[so * a.H + i];
        a.dst_c[dd * a.H + i] = a.src_c[so * a.H + i];
      }
      if (a.dst_q)
        for (int i = tid; i < a.ldq; i += BEAM_SELECT_NT) a.dst_q[dd * a.ldq + i] = a.src_q[so * a.ldq + i];
    }
  }
}

// slot 0 of every clip starts from (h0, c0) with the empty hypothesis, "no previous word" (:881-893)
__global__ void beam_init_kernel(const BeamArgs a, const float *h0c0, float *h, float *c, int H, int32_t *row_clip) {
  const int row = blockIdx.x, b = row / a.k, j = row - b * a.k;
  for (int i = threadIdx.x; i < H; i += blockDim.x) {
    h[static_cast<size_t>(row) * H + i] = j == 0 ? h0c0[static_cast<size_t>(b) * 2 * H + i] : 0.f;
    c[static_cast<size_t>(row) * H + i] = j == 0 ? h0c0[static_cast<size_t>(b) * 2 * H + H + i] : 0.f;
  }
  for (int i = threadIdx.x; i < a.maxlen; i += blockDim.x)
    a.out_tokens[static_cast<size_t>(row) * a.maxlen + i] = -1;
  if (threadIdx.x == 0) {
    row_clip[row] = b;
    a.alive[row] = j == 0;
    a.score[row] = 0.f;
    a.hist_len[row] = 0;
    a.src_row[row] = row;
    a.tok_prev[row] = -1;
    a.out_lengths[row] = 0;
    a.out_scores[row] = 0.f;
    if (j == 0) { a.dead_k[b] = 0; a.done[b] = 0; a.out_count[b] = 0; }
  }
}

// ---------------------------------------------------------------------------
// prologue helpers
// ---------------------------------------------------------------------------
__global__ void meanpool_kernel(const float *__restrict__ ctxg, const float *__restrict__ mask,
                                float *__restrict__ gbar, int T, int D) {
  const int b = blockIdx.y;
  const int d = blockIdx.x * blockDim.x + threadIdx.x;
  if (d >= D) return;
  float cnt = 0.f;
  for (int t = 0; t < T; ++t) cnt += mask[b * T + t];
  float s = 0.f;
  const float *p = ctxg + static_cast<size_t>(b) * T * D + d;
  for (int t = 0; t < T; ++t) s += p[static_cast<size_t>(t) * D];
  gbar[static_cast<size_t>(b) * D + d] = s / cnt;        // :618, :649
}

__global__ void transpose_kernel(const float *__restrict__ src, int K, int N, float *__restrict__ dst,
                                 int ld_dst, int r0) {
  __shared__ float tile[32][33];
  const int k0 = blockIdx.y * 32, n0 = blockIdx.x * 32;
  for (int i = threadIdx.y; i < 32; i += 8) {
    const int k = k0 + i, n = n0 + threadIdx.x;
    tile[i][threadIdx.x] = (k < K && n < N) ? src[static_cast<size_t>(k) * N + n] : 0.f;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += 8) {
    const int n = n0 + i, k = k0 + threadIdx.x;
    if (n < N && k < K) dst[static_cast<size_t>(r0 + n) * ld_dst + k] = tile[threadIdx.x][i];
  }
}

// dst[(4*(n % H) + n / H) * ld_dst + c0 + k] = src[k * N + n], N = 4H: a (K, 4H) gate-major weight becomes K-major
// rows in gate-interleaved order (the four gates of a hidden unit adjacent), written at column offset c0
__global__ void transpose_il_kernel(const float *__restrict__ src, int K, int N, float *__restrict__ dst, int ld_dst,
                                    int c0, int H) {
  __shared__ float tile[32][33];
  const int k0 = blockIdx.y * 32, n0 = blockIdx.x * 32;
  for (int i = threadIdx.y; i < 32; i += 8) {
    const int k = k0 + i, n = n0 + threadIdx.x;
    tile[i][threadIdx.x] = (k < K && n < N) ? src[static_cast<size_t>(k) * N + n] : 0.f;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += 8) {
    const int n = n0 + i, k = k0 + threadIdx.x;
    if (n < N && k < K) dst[static_cast<size_t>(4 * (n % H) + n / H) * ld_dst + c0 + k] = tile[threadIdx.x][i];
  }
}

// dst[4*u + g] = src[g*H + u]
__global__ void interleave4_kernel(const float *__restrict__ src, float *__restrict__ dst, int H) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < 4 * H) dst[4 * (i % H) + i / H] = src[i];
}

// out_q[row][:] = softmax(sc[q][row][:]) for the planes q = blockIdx.y whose output pointer is not null
struct SoftmaxOuts {
  float *out[3];
};
__global__ void softmax_rows_kernel(const float *__restrict__ sc, SoftmaxOuts o, int nrows, int n) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  float *out = o.out[blockIdx.y];
  if (row >= nrows || !out) return;
  const float *p = sc + (static_cast<size_t>(blockIdx.y) * nrows + row) * n;
  float m = -INFINITY;
  for (int i = lane; i < n; i += 32) m = fmaxf(m, p[i]);
  m = warp_max(m);
  float s = 0.f;
  for (int i = lane; i < n; i += 32) s += expf(p[i] - m);
  s = warp_sum(s);
  for (int i = lane; i < n; i += 32) out[static_cast<size_t>(row) * n + i] = expf(p[i] - m) / s;
}

__global__ void scale_kernel(float *__restrict__ x, const float *__restrict__ f, size_t n) {
  const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < n) x[i] *= f[i];
}

__global__ void init_rows_kernel(int rows, int64_t *tok_prev, int32_t *alive, int32_t *lengths, float *scores,
                                 int64_t *tokens, int maxlen) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < rows) {
    tok_prev[i] = -1;
    alive[i] = 1;
    if (lengths) lengths[i] = 0;
    if (scores) scores[i] = 0.f;
  }
  if (tokens && i < rows * maxlen) tokens[i] = -1;
}

__global__ void add_vec_kernel(float *dst, const float *a, const float *b, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[i] = a[i] + (b ? b[i] : 0.f);
}

}  // namespace

int gates_launch(const GateArgs &a, cudaStream_t stream) {
  const int w = a.H > a.E ? a.H : a.E;
  const bool vec2 = (a.H % 2 == 0) && (a.ldhp % 2 == 0) && (a.ldpc % 2 == 0) && (a.off_u % 2 == 0) &&
                    a.hp_parts <= 4 && a.pc_parts <= 4 && (a.hp_plane % 2 == 0) && (a.pc_plane % 2 == 0);
  if (vec2) {
    dim3 grid((w / 2 + 127) / 128 + ((w % 2) ? 1 : 0), a.rows);
    set_launch_label("gates2_kernel");
  return launch_pdl(gates2_kernel, grid, dim3(128), 0, stream, a);
  }
  dim3 grid((w + 127) / 128, a.rows);
  set_launch_label("gates_kernel");
  return launch_pdl(gates_kernel, grid, dim3(128), 0, stream, a);
}

int zact_launch(const ZactArgs &a, cudaStream_t stream) {
  set_launch_label("zact_kernel");
  return launch_pdl(zact_kernel, dim3((a.rows * a.E + 255) / 256), dim3(256), 0, stream, a);
}

// STAT_PICK_NT = 256 | 512 | 1024: threads per row of the register-resident reduction (0: the two-pass kernel).
// Default: 1024 up to 128 rows (B = 64 greedy decode in the captured graph: 43.5 us per step against 44.6 at 512 and
// 45.4 with the two-pass kernel), 512 beyond (beam search over 160 rows: more than one CTA per SM).
int pick_launch(const PickArgs &a, cudaStream_t stream) {
  static int nt_env = -1;
  if (nt_env < 0) {
    const char *e = getenv("STAT_PICK_NT");
    nt_env = e ? atoi(e) : -2;
  }
  const int nt = nt_env >= 0 ? nt_env : (a.rows <= 128 ? 1024 : 512);
  const bool aligned = (a.ldl & 3) == 0 && (reinterpret_cast<uintptr_t>(a.logits) & 15) == 0;
  set_launch_label("pick_kernel");
  if (aligned && nt == 1024 && a.V <= 4 * 1024 * 4)
    return launch_pdl(pick_regs_kernel<1024, 4>, dim3(a.rows), dim3(1024), 0, stream, a);
  if (aligned && nt == 512 && a.V <= 4 * 512 * 7)
    return launch_pdl(pick_regs_kernel<512, 7>, dim3(a.rows), dim3(512), 0, stream, a);
  if (aligned && nt == 256 && a.V <= 4 * 256 * 13)
    return launch_pdl(pick_regs_kernel<256, 13>, dim3(a.rows), dim3(256), 0, stream, a);
  return launch_pdl(pick_kernel, dim3(a.rows), dim3(256), 0, stream, a);
}

int beam_select_launch(const BeamArgs &a, cudaStream_t stream) {
  set_launch_label("beam_select_kernel");
  return launch_pdl(beam_select_kernel, dim3(a.B), dim3(BEAM_SELECT_NT), 0, stream, a);
}

int beam_init_launch(const BeamArgs &a, const float *h0c0, float *h, float *c, int H, int32_t *row_clip,
                     cudaStream_t stream) {
  beam_init_kernel<<<a.B * a.k, 128, 0, stream>>>(a, h0c0, h, c, H, row_clip);
  note_launch();
  STAT_CUDA_CHECK(cudaGetLastError());
  return STAT_OK;
}

int meanpool_launch(const float *ctxg, const float *mask, float *gbar, int B, int T, int D,
                    cudaStream_t stream) {
  dim3 grid((D + 255) / 256, B);
  meanpool_kernel<<<grid, 256, 0, stream>>>(ctxg, mask, gbar, T, D);
  note_launch();
  STAT_CUDA_CHECK(cudaGetLastError());
  return STAT_OK;
}

int transpose_launch(const float *src, int K, int N, float *dst, int ld_dst, int r0, cudaStream_t stream) {
  dim3 grid((N + 31) / 32, (K + 31) / 32);
  transpose_kernel<<<grid, dim3(32, 8), 0, stream>>>(src, K, N, dst, ld_dst, r0);
  note_launch();
  STAT_CUDA_CHECK(cudaGetLastError());
  return STAT_OK;
}

int transpose_il_launch(const float *src, int K, int H, float *dst, int ld_dst, int c0, cudaStream_t stream) {
  dim3 grid((4 * H + 31) / 32, (K + 31) / 32);
  transpose_il_kernel<<<grid, dim3(32, 8), 0, stream>>>(src, K, 4 * H, dst, ld_dst, c0, H);
  note_launch();
  STAT_CUDA_CHECK(cudaGetLastError());
  return STAT_OK;
}

int interleave4_launch(const float *src, float *dst, int H, cudaStream_t stream) {
  interleave4_kernel<<<(4 * H + 255) / 256, 256, 0, stream>>>(src, dst, H);
  note_launch();
  STAT_CUDA_CHECK(cudaGetLastError());
  return STAT_OK;
}

int softmax_rows3_launch(const float *scores, float *out0, float *out1, float *out2, int nrows, int n,
                         cudaStream_t stream) {
  if (!out0 && !out1 && !out2) return STAT_OK;
  SoftmaxOuts o;
  o.out[0] = out0; o.out[1] = out1; o.out[2] = out2;
  softmax_rows_kernel<<<dim3((nrows + 3) / 4, 3), 128, 0, stream>>>(scores, o, nrows, n);
  note_launch();
  STAT_CUDA_CHECK(cudaGetLastError());
  return STAT_OK;
}

int scale_launch(float *x, const float *f, size_t n, cudaStream_t stream) {
  scale_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, stream>>>(x, f, n);
  note_launch();
  STAT_CUDA_CHECK(cudaGetLastError());
  return STAT_OK;
}

int init_rows_launch(int rows, int64_t *tok_prev, int32_t *alive, int32_t *lengths, float *scores,
                     int64_t *tokens, int maxlen, cudaStream_t stream) {
  const int n = tokens ? rows * maxlen : rows;
  init_rows_kernel<<<(n + 255) / 256, 256, 0, stream>>>(rows, tok_prev, alive, lengths, scores, tokens, maxlen);
  note_launch();
  STAT_CUDA_CHECK(cudaGetLastError());
  return STAT_OK;
}

int add_vec_launch(float *dst, const float *a, const float *b, int n, cudaStream_t stream) {
  add_vec_kernel<<<(n + 255) / 256, 256, 0, stream>>>(dst, a, b, n);
  note_launch();
  STAT_CUDA_CHECK(cudaGetLastError());
  return STAT_OK;
}

}  // namespace stat
