// Backward pass of the teacher-forced training step: what
// `tensor.grad(cost, wrt=itemlist(tparams))` yields in the reference (model_attention.py:1193)
// for the cost of model_attention.py:1129-1147, as explicit kernels + the dense primitive.
//
// Structure.  The forward (stat_precompute + stat_forward_teacher) leaves the
// projected context blocks, the four attention weights and the hidden states of every step.
//   A  recompute, batched over all N = L*B (step, clip) rows: hidden-state projections, fused
//      contexts, selector, gate pre-activations, the cell-state chain, readout activation, logits;
//   B  readout backward, batched: soft-max/NLL gradient in place of the logits, the ff_logit*
//      gradients, the readout's contribution to dh_t and dctx_t;
//   C  back-propagation through time, t = L-1 .. 0: cell, selector, the four soft-attentions
//      (per (clip, frame) blocks; every element of the step-invariant gradient blocks dpctx*,
//      dctx*0, dqctxl is owned by one thread: plain +=, no atomics, bit-reproducible), then ONE
//      product (B, 8H+1) x (8H+1, H) for everything that flows into h_{t-1};
//   D  every weight gradient as one tall-K product over the stacked steps / frames, the K0
//      backward (tanh feature projections), the init-state path, the embedding scatter, decay.
//
// First, correctness-oriented version: plain SIMT kernels around gemm_launch, no fusion, no
// tuning (per-step products run as single 128-row tiles) -- kept behind STAT_BW_FAST=0.  The default adds the
// first round of optimisations measured on the B200 in round 1 (k-split products, deferred accumulation of the
// step-invariant blocks, aligned weight copies: 24.2 -> 21.5 ms per B=128 step) plus the owner-block embedding
// scatter (the row-wise scatter of round 1 cost 5.3 ms and is gone); stat_grad_profile_* gives the per-phase
// device time of either mode.  The same translation unit compiles
// under g++ with -DSTAT_EMU against tests/emu/cuda_emu.h (threads-as-CUDA-threads emulation,
// test infrastructure) so that kernels and orchestration can be checked against the gradient
// oracle without a GPU; the product build never defines STAT_EMU.
#ifdef STAT_EMU
#include "cuda_emu.h"
#else
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <vector>

#include "kernels.cuh"
#include "stat_common.cuh"
#endif

namespace stat {
namespace bw {

constexpr int NT = 128;      // threads of the per-(clip, frame) / per-row blocks
constexpr int HMAX_PT = 8;   // columns per thread in att_main: H <= NT * HMAX_PT = 1024
constexpr int RMAX = 16;

inline size_t up(size_t x, size_t a) { return (x + a - 1) / a * a; }

#ifdef STAT_EMU
#define BW_LAUNCH(kernel, grid, block, stream, ...) (emu::launch(kernel, grid, block, __VA_ARGS__), STAT_OK)
#else
template <typename... KArgs, typename... Args>
int bw_launch(void (*kernel)(KArgs...), dim3 grid, dim3 block, cudaStream_t stream, Args... args) {
  kernel<<<grid, block, 0, stream>>>(KArgs(args)...);
  STAT_CUDA_CHECK(cudaGetLastError());
  note_launch();
  return STAT_OK;
}
#define BW_LAUNCH(kernel, grid, block, stream, ...) bw_launch(kernel, grid, block, stream, __VA_ARGS__)
#endif

// ---------------------------------------------------------------------------
// optional phase timing (stat_grad_profile_*): CUDA events on the caller's stream around each group of
// launches.  Off by default; process-wide, not for concurrent use, not under stream capture.
// ---------------------------------------------------------------------------
enum BwPhase { BP_LAYOUT = 0, BP_RECOMPUTE, BP_LOGITS, BP_READOUT_BW, BP_CELL_BW, BP_DCTX_GEMM, BP_SELECTOR_BW,
               BP_ATT_DOTS, BP_ATT_SOFT, BP_ATT_MAIN, BP_REDUCE_T, BP_DH_GEMM, BP_WGRAD_STEPS, BP_EMBEDDING,
               BP_CTX_BLOCKS, BP_INIT_STATE, BP_DECAY, BP_COUNT };
const char *const kBwPhaseNames[BP_COUNT] = {
    "operand_layouts", "recompute", "logits", "readout_backward", "loop_cell_backward", "loop_dctx_gemm",
    "loop_selector_backward", "loop_att_dots", "loop_att_soft", "loop_att_main", "loop_reduce_frames",
    "loop_dh_gemm", "weight_grads_steps", "embedding", "context_blocks", "init_state", "weight_decay"};
// bw_mark(phase, st): closes the phase that was open and opens `phase` (BP_COUNT = just close)
#ifdef STAT_EMU
inline void bw_mark(int, cudaStream_t) {}
#else
struct BwRec { int phase; cudaEvent_t a, b; };
bool g_bw_prof_on = false;
bool g_bw_open = false;
BwRec g_bw_cur;
std::vector<BwRec> g_bw_recs;
inline void bw_mark(int phase, cudaStream_t st) {
  if (!g_bw_prof_on) return;
  if (g_bw_open) {
    cudaEventRecord(g_bw_cur.b, st);
    g_bw_recs.push_back(g_bw_cur);
    g_bw_open = false;
  }
  if (phase >= BP_COUNT) return;
  g_bw_cur.phase = phase;
  cudaEventCreate(&g_bw_cur.a);
  cudaEventCreate(&g_bw_cur.b);
  cudaEventRecord(g_bw_cur.a, st);
  g_bw_open = true;
}
#endif

// ---------------------------------------------------------------------------
// device helpers
// ---------------------------------------------------------------------------
__device__ __forceinline__ float sigm(float x) { return 1.0f / (1.0f + expf(-x)); }
// tanh of the attention-score backward (11 per column and step in k_att_main, L*R per element in k_att_accum: the
// accurate tanhf made those kernels instruction-bound).  Same 2-MUFU form as the forward attention kernel, absolute
// error ~1.5e-7; the CPU emulation keeps tanhf.
#ifdef STAT_EMU
__device__ __forceinline__ float tanh_bw(float x) { return tanhf(x); }
#else
__device__ __forceinline__ float tanh_bw(float x) { return tanh_fast(x); }
#endif

// Sum over the block, result in every thread (same summation order everywhere).  All threads
// of the block must call; blockDim.x is a multiple of 32; sh holds 32 floats.
__device__ __forceinline__ float block_sum(float v, float *sh) {
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
  __syncthreads();
  if (lane == 0) sh[w] = v;
  __syncthreads();
  float r = 0.f;
  for (int i = 0; i < nw; ++i) r += sh[i];
  return r;
}

// N sums over the block at once (one barrier pair for all of them; same summation order as block_sum):
// v[i] <- block total of v[i], in every thread.  sh holds 32 * N floats.
template <int N>
__device__ __forceinline__ void block_sum_n(float (&v)[N], int n, float *sh) {
#pragma unroll
  for (int i = 0; i < N; ++i)
    if (i < n)
      for (int o = 16; o > 0; o >>= 1) v[i] += __shfl_xor_sync(0xffffffffu, v[i], o);
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
  __syncthreads();
  if (lane == 0) {
#pragma unroll
    for (int i = 0; i < N; ++i)
      if (i < n) sh[w * N + i] = v[i];
  }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < N; ++i) {
    if (i < n) {
      float r = 0.f;
      for (int k = 0; k < nw; ++k) r += sh[k * N + i];
      v[i] = r;
    }
  }
}

__device__ __forceinline__ float block_max(float v, float *sh) {
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
  __syncthreads();
  if (lane == 0) sh[w] = v;
  __syncthreads();
  float r = sh[0];
  for (int i = 1; i < nw; ++i) r = fmaxf(r, sh[i]);
  return r;
}

// ---------------------------------------------------------------------------
// generic kernels
// ---------------------------------------------------------------------------
// dst[c*ldd + r] = src[r*lds + c]   (rows x cols -> cols x rows); block 256 = 32 x 8
__global__ void k_transpose(const float *src, int rows, int cols, int lds, float *dst, int ldd) {
  __shared__ float tile[32][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
  for (int i = ty; i < 32; i += 8) {
    const int r = r0 + i, c = c0 + tx;
    tile[i][tx] = (r < rows && c < cols) ? src[static_cast<size_t>(r) * lds + c] : 0.f;
  }
  __syncthreads();
  for (int i = ty; i < 32; i += 8) {
    const int c = c0 + i, r = r0 + tx;
    if (c < cols && r < rows) dst[static_cast<size_t>(c) * ldd + r] = tile[tx][i];
  }
}

// dst[chunk][c] = sum over the chunk's rows of src[r][c]; chunk = blockIdx.y, `per` rows each
__global__ void k_colsum(const float *src, int rows, int cols, int ld, int per, float *dst, int ldd) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  const int r0 = blockIdx.y * per, r1 = min(rows, r0 + per);
  if (c >= cols) return;
  float s = 0.f;
  for (int r = r0; r < r1; ++r) s += src[static_cast<size_t>(r) * ld + c];
  dst[static_cast<size_t>(blockIdx.y) * ldd + c] = s;
}

__global__ void k_add2(float *out, const float *a, const float *b, size_t n) {
  const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < n) out[i] = a[i] + b[i];
}

__global__ void k_add_inplace(float *dst, const float *src, size_t n) {
  const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < n) dst[i] += src[i];
}

// dy *= 1 - y^2
__global__ void k_tanh_bw(float *dy, const float *y, size_t n) {
  const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < n) dy[i] *= 1.0f - y[i] * y[i];
}

// g += 2*c*p   (weight decay, model_attention.py:1130-1136)
__global__ void k_decay(float *g, const float *p, size_t n, float c2) {
  const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < n) g[i] = fmaf(c2, p[i], g[i]);
}

// out = x * (f ? f : 0.5)
__global__ void k_scale_dp(float *out, const float *x, const float *f, size_t n) {
  const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < n) out[i] = x[i] * (f ? f[i] : 0.5f);
}

// gbar[b][d] = sum_t ctxg[b][t][d] / sum_t mask[b][t]   (:618, :649)
__global__ void k_meanpool(const float *ctxg, const float *mask, float *gbar, int B, int T, int D) {
  const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= static_cast<size_t>(B) * D) return;
  const int b = static_cast<int>(i / D), d = static_cast<int>(i % D);
  float cnt = 0.f, s = 0.f;
  for (int t = 0; t < T; ++t) {
    cnt += mask[b * T + t];
    s += ctxg[(static_cast<size_t>(b) * T + t) * D + d];
  }
  gbar[i] = s / cnt;
}

// ---------------------------------------------------------------------------
// stage A: recompute
// ---------------------------------------------------------------------------
// Hprev[(t,b)] = t ? h_all[t-1][b] : h0[b]
__global__ void k_gather_prev(const float *h0c0, const float *h_all, float *Hprev, int L, int B, int H) {
  const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= static_cast<size_t>(L) * B * H) return;
  const int j = static_cast<int>(i % H);
  const size_t n = i / H;
  const int b = static_cast<int>(n % B);
  Hprev[i] = n < static_cast<size_t>(B) ? h0c0[static_cast<size_t>(b) * 2 * H + j] : h_all[i - static_cast<size_t>(B) * H];
}

// EMB[(t,b)] = t ? Wemb[x[t-1][b]] : 0   (:613-617)
__global__ void k_gather_emb(const float *Wemb, const int64_t *x, float *EMB, int L, int B, int E) {
  const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= static_cast<size_t>(L) * B * E) return;
  const int e = static_cast<int>(i % E);
  const size_t n = i / E;
  EMB[i] = n < static_cast<size_t>(B) ? 0.f : Wemb[static_cast<size_t>(x[n - B]) * E + e];
}

// csum[(t,b)][h] = cG + cM + cLT from the saved attention weights (:383,:399,:412,:426,:430)
__global__ void k_ctx_parts(const float *al, const float *ag, const float *am, const float *alt, const float *Lc,
                            const float *G, const float *M, float *csum, int L, int B, int T, int R, int H) {
  const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= static_cast<size_t>(L) * B * H) return;
  const int h = static_cast<int>(i % H);
  const size_t n = i / H;
  const int b = static_cast<int>(n % B);
  float s = 0.f;
  for (int t = 0; t < T; ++t) {
    const size_t bt = static_cast<size_t>(b) * T + t, nt = n * T + t;
    float cl = 0.f;
    for (int r = 0; r < R; ++r) cl = fmaf(al[nt * R + r], Lc[(bt * R + r) * H + h], cl);
    s += ag[nt] * G[bt * H + h] + am[nt] * M[bt * H + h] + alt[nt] * cl;
  }
  csum[i] = s;
}

// beta = sigmoid(selector logit), ctx = beta * csum   (:432-435)
__global__ void k_beta_ctx(const float *HQ, int ldq, int off_sel, const float *csum, float *beta, float *ctx, size_t N,
                           int H, int selector) {
  const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= N * H) return;
  const size_t n = i / H;
  const float be = selector ? sigm(HQ[n * ldq + off_sel]) : 1.0f;
  ctx[i] = be * csum[i];
  if (i % H == 0) beta[n] = be;
}

// the cell-state chain of one (clip, unit) over all steps (:437-454); gate activations kept
__global__ void k_cell_forward(const float *HQ, int ldq, int off_u, const float *XW, const float *CW,
                               const float *h0c0, const float *mask, const float *dp_gates, float *GATES, float *Call,
                               int L, int B, int H) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * H) return;
  const int b = i / H, j = i % H;
  float c = h0c0[static_cast<size_t>(b) * 2 * H + H + j];
  for (int t = 0; t < L; ++t) {
    const size_t n = static_cast<size_t>(t) * B + b;
    float pre[4];
    for (int q = 0; q < 4; ++q)
      pre[q] = HQ[n * ldq + off_u + q * H + j] + XW[n * 4 * H + q * H + j] + CW[n * 4 * H + q * H + j];
    const float di = dp_gates ? dp_gates[n * 3 * H + j] : 0.5f;
    const float df = dp_gates ? dp_gates[n * 3 * H + H + j] : 0.5f;
    const float dO = dp_gates ? dp_gates[n * 3 * H + 2 * H + j] : 0.5f;
    const float gi = sigm(pre[0] * di), gf = sigm(pre[1] * df), go = sigm(pre[2] * dO), gg = tanhf(pre[3]);
    const float m = mask[n];
    const float cn = gf * c + gi * gg;
    c = m * cn + (1.0f - m) * c;
    GATES[n * 4 * H + j] = gi;
    GATES[n * 4 * H + H + j] = gf;
    GATES[n * 4 * H + 2 * H + j] = go;
    GATES[n * 4 * H + 3 * H + j] = gg;
    Call[n * H + j] = c;
  }
}

// ZT = tanh(ZP + emb + ctx.Wctx), Z = ZT * dp_z   (:684-696); ZP = (dp_h*h).Wl + b
__global__ void k_zact(const float *ZP, const float *EMB, const float *ZC, const float *dp_z, float *ZT, float *Z,
                       size_t n) {
  const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float z = ZP[i];
  if (EMB) z += EMB[i];
  if (ZC) z += ZC[i];
  const float t = tanhf(z);
  ZT[i] = t;
  Z[i] = t * (dp_z ? dp_z[i] : 0.5f);
}

// ---------------------------------------------------------------------------
// stage B: readout backward
// ---------------------------------------------------------------------------
// In place: logits row -> d cost / d logits.  cost = inv_batch * sum_b -sum_t mask*log(p[x]+1e-8)
// (:711-715, :1129): dlogit_j = -mask*inv_batch * p_x/(p_x+1e-8) * (delta_xj - p_j).  One block per row.
__global__ void k_softmax_nll(float *LOG, int ldl, int V, const int64_t *x, const float *mask, float inv_batch) {
  __shared__ float sh[32];
  const size_t n = blockIdx.x;
  float *row = LOG + n * ldl;
  float mx = -3.0e38f;
  for (int j = threadIdx.x; j < V; j += blockDim.x) mx = fmaxf(mx, row[j]);
  mx = block_max(mx, sh);
  float s = 0.f;
  for (int j = threadIdx.x; j < V; j += blockDim.x) s += expf(row[j] - mx);
  s = block_sum(s, sh);
  const int xt = static_cast<int>(x[n]);
  const float px = expf(row[xt] - mx) / s;
  const float coef = -mask[n] * inv_batch * (px / (px + 1e-8f));
  __syncthreads();   // every thread has read row[xt] before anyone overwrites it
  for (int j = threadIdx.x; j < V; j += blockDim.x) {
    const float p = expf(row[j] - mx) / s;
    row[j] = coef * ((j == xt ? 1.0f : 0.f) - p);
  }
}

// DZP = DZ * dp_z * (1 - ZT^2)
__global__ void k_dzp(const float *DZ, const float *ZT, const float *dp_z, float *DZP, size_t n) {
  const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < n) DZP[i] = DZ[i] * (dp_z ? dp_z[i] : 0.5f) * (1.0f - ZT[i] * ZT[i]);
}

// coverage regulariser (:1138-1147): d/d alpha[step][i] of alpha_c * mean_n sum_b (1 - sum_steps alpha)^2
__global__ void k_cov(const float *alpha, int L, size_t cnt, float scale, float *out) {
  const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= cnt) return;
  float s = 0.f;
  for (int t = 0; t < L; ++t) s += alpha[static_cast<size_t>(t) * cnt + i];
  out[i] = -scale * (1.0f - s);
}

// ---------------------------------------------------------------------------
// stage C: back-propagation through time
// ---------------------------------------------------------------------------
// cell backward of step t (:441-457).  In: DHc/DCc = gradient w.r.t. (h_t, c_t) from the future,
// DHR = readout's dh_t.  Out: gate pre-activation gradients into DHQ[:, off_u ...), DCc <- dc_{t-1},
// DHm <- the part of dh_t that bypasses the cell through the mask.
__global__ void k_cell_backward(int t, const float *DHc, float *DCc, const float *DHR, const float *GATES,
                                const float *Call, const float *h0c0, const float *mask, const float *dp_gates,
                                float *DHQ, int ldq, int off_u, float *DHm, int B, int H) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * H) return;
  const int b = i / H, j = i % H;
  const size_t n = static_cast<size_t>(t) * B + b;
  const float m = mask[n];
  const float gi = GATES[n * 4 * H + j], gf = GATES[n * 4 * H + H + j], go = GATES[n * 4 * H + 2 * H + j],
              gg = GATES[n * 4 * H + 3 * H + j];
  const float c = Call[n * H + j];
  const float cprev = t > 0 ? Call[(n - B) * H + j] : h0c0[static_cast<size_t>(b) * 2 * H + H + j];
  const float di = dp_gates ? dp_gates[n * 3 * H + j] : 0.5f;
  const float df = dp_gates ? dp_gates[n * 3 * H + H + j] : 0.5f;
  const float dO = dp_gates ? dp_gates[n * 3 * H + 2 * H + j] : 0.5f;
  const float dh = DHc[i] + DHR[n * H + j];
  const float dht = m * dh;
  DHm[i] = (1.0f - m) * dh;
  const float tc = tanhf(c);
  const float d_o = dht * tc;
  const float dct = DCc[i] + dht * go * (1.0f - tc * tc);
  const float dcn = m * dct;
  DCc[i] = (1.0f - m) * dct + dcn * gf;
  const float d_f = dcn * cprev, d_i = dcn * gg, d_g = dcn * gi;
  float *o = DHQ + n * ldq + off_u;
  o[j] = d_i * gi * (1.0f - gi) * di;
  o[H + j] = d_f * gf * (1.0f - gf) * df;
  o[2 * H + j] = d_o * go * (1.0f - go) * dO;
  o[3 * H + j] = d_g * (1.0f - gg * gg);
}

// selector backward (:432-435): dctx = DCTX (gates) + DCR (readout); one block per clip
__global__ void k_selector_bw(int t, const float *DCTX, const float *DCR, const float *csum, const float *beta,
                              float *DHQ, int ldq, int off_sel, float *DC, int B, int H, int selector) {
  __shared__ float sh[32];
  const int b = blockIdx.x;
  const size_t n = static_cast<size_t>(t) * B + b;
  float part = 0.f;
  for (int h = threadIdx.x; h < H; h += blockDim.x) {
    const float d = DCTX[static_cast<size_t>(b) * H + h] + DCR[n * H + h];
    part = fmaf(d, csum[n * H + h], part);
  }
  const float dbeta = block_sum(part, sh);
  const float be = beta[n];
  for (int h = threadIdx.x; h < H; h += blockDim.x) {
    const float d = DCTX[static_cast<size_t>(b) * H + h] + DCR[n * H + h];
    DC[static_cast<size_t>(b) * H + h] = selector ? be * d : d;
  }
  if (threadIdx.x == 0 && selector) DHQ[n * ldq + off_sel] = dbeta * be * (1.0f - be);
}

// per (clip, frame): cL = sum_r alpha_l * Lc, and the three d alpha = dC . value   (:383,:399,:412,:426)
// (RT = compile-time R, 0 = runtime: with the region loops unrolled the loads of a column go out together instead of
// one DRAM round trip per region)
template <int RT>
__global__ void k_att_dots(int t_step, const float *al, const float *Lc, const float *G, const float *M,
                           const float *DC, float *CL, float *DA3, int B, int T, int R_, int H) {
  __shared__ float sh[32];
  const int R = RT ? RT : R_;
  constexpr int RU = RT ? RT : RMAX;
  const int bt = blockIdx.x, b = bt / T;
  const size_t n = static_cast<size_t>(t_step) * B + b;
  const size_t nt = n * T + (bt % T);
  float alr[RU];
#pragma unroll
  for (int r = 0; r < RU; ++r) alr[r] = (RT || r < R) ? al[nt * R + r] : 0.f;
  float pg = 0.f, pm = 0.f, pl = 0.f;
  for (int h = threadIdx.x; h < H; h += blockDim.x) {
    float lc[RU];
#pragma unroll
    for (int r = 0; r < RU; ++r) lc[r] = (RT || r < R) ? Lc[(static_cast<size_t>(bt) * R + r) * H + h] : 0.f;
    const float d = DC[static_cast<size_t>(b) * H + h];
    const float gv = G[static_cast<size_t>(bt) * H + h], mv = M[static_cast<size_t>(bt) * H + h];
    float cl = 0.f;
#pragma unroll
    for (int r = 0; r < RU; ++r)
      if (RT || r < R) cl = fmaf(alr[r], lc[r], cl);
    CL[static_cast<size_t>(bt) * H + h] = cl;
    pg = fmaf(d, gv, pg);
    pm = fmaf(d, mv, pm);
    pl = fmaf(d, cl, pl);
  }
  pg = block_sum(pg, sh);
  pm = block_sum(pm, sh);
  pl = block_sum(pl, sh);
  if (threadIdx.x == 0) {
    const size_t BT = static_cast<size_t>(B) * T;
    DA3[bt] = pg;
    DA3[BT + bt] = pm;
    DA3[2 * BT + bt] = pl;
  }
}

// temporal soft-max backward of (clip, which): da_t = alpha_t * (dalpha_t - sum_s alpha_s dalpha_s),
// dalpha including the coverage term; also the score-bias gradients (DCACC columns 1..3)
__global__ void k_att_soft(int t_step, const float *ag, const float *am, const float *alt, const float *DA3,
                           const float *COV3, float *DS3, float *DCACC, int B, int T) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= 3 * B) return;
  const int w = i / B, b = i % B;
  const float *alpha = (w == 0 ? ag : (w == 1 ? am : alt)) + (static_cast<size_t>(t_step) * B + b) * T;
  const size_t base = (static_cast<size_t>(w) * B + b) * T;
  float s = 0.f;
  for (int t = 0; t < T; ++t) s = fmaf(alpha[t], DA3[base + t] + COV3[base + t], s);
  for (int t = 0; t < T; ++t) {
    const float da = alpha[t] * (DA3[base + t] + COV3[base + t] - s);
    DS3[base + t] = da;
    DCACC[(static_cast<size_t>(b) * T + t) * 4 + 1 + w] += da;
  }
}

// per (clip, frame): everything behind the four scores (:371-426).  Accumulates the step-invariant
// gradient blocks and writes this frame's share of the query gradients (DSQP, summed over frames by
// k_reduce_t) -- see the derivation in DESIGN.md section 9.
struct AttBw {
  int t_step, B, T, R, H, ldq, global_proj;
  const float *al, *ag, *am, *alt;        // saved attention weights (L,B,T[,R])
  const float *HQ;                        // (N,ldq): [sl | sg | sm | slt(+blt)] at columns 0..4H
  const float *pL, *Lc, *Q;               // (B,T,R,H)
  const float *pG, *pM;                   // (B,T,H)
  const float *Ul, *Ug, *Um, *Ult;        // (H)
  const float *DC;                        // (B,H)
  const float *DS3;                       // (3,B,T) score gradients g, m, lt
  const float *COVL;                      // (B,T,R) coverage term of alpha_l
  float *DPL, *DLC, *DQ;                  // (B,T,R,H) +=
  float *DPG, *DPM, *DG, *DM;             // (B,T,H) +=
  float *DUACC;                           // (B,T,4H) += [l | g | m | lt]
  float *DCACC;                           // (B,T,4)  += column 0
  float *DSQP;                            // (B,T,4H) = [dsl | dsg | dsm | dslt]
  // STAT_BW_FAST: instead of read-modify-writing DPL / DLC / DQ every step, keep this step's small factors
  // (k_att_accum sums all steps in registers after the loop); null = accumulate in place
  float *DAL;                             // (L,B,T,R) spatial score gradients of every step
  float *DPLT;                            // (L,B,T,H) dpLT of every step
};

template <int RT>
__global__ void k_att_main(const AttBw a) {
  __shared__ float sh[32 * RMAX];
  constexpr int RU = RT ? RT : RMAX;
  const int H = a.H, R = RT ? RT : a.R, T = a.T;
  const int bt = blockIdx.x, b = bt / T;
  const size_t n = static_cast<size_t>(a.t_step) * a.B + b;
  const size_t nt = n * T + (bt % T);
  const size_t BT = static_cast<size_t>(a.B) * T;
  const float daG = a.DS3[bt], daM = a.DS3[BT + bt], daLT = a.DS3[2 * BT + bt];
  const float aG = a.ag[nt], aM = a.am[nt], aLT = a.alt[nt];
  float alr[RU], part[RU];
#pragma unroll
  for (int r = 0; r < RU; ++r) {
    alr[r] = (RT || r < R) ? a.al[nt * R + r] : 0.f;
    part[r] = 0.f;
  }
  const float *hq = a.HQ + n * a.ldq;
  float dcl[HMAX_PT], dpl[HMAX_PT];
  int k = 0;
  for (int h = threadIdx.x; h < H; h += blockDim.x, ++k) {
    const size_t o = static_cast<size_t>(bt) * H + h, o4 = static_cast<size_t>(bt) * 4 * H + h;
    // the loads of this column first (independent addresses: one memory round trip), then the arithmetic
    float q[RU], lc[RU];
#pragma unroll
    for (int r = 0; r < RU; ++r) {
      const size_t ol = (static_cast<size_t>(bt) * R + r) * H + h;
      q[r] = (RT || r < R) ? a.Q[ol] : 0.f;
      lc[r] = (RT || r < R) ? a.Lc[ol] : 0.f;
    }
    const float dc = a.DC[static_cast<size_t>(b) * H + h];
    const float pg = a.pG[o], pm = a.pM[o], sg = hq[H + h], sm = hq[2 * H + h], slt = hq[3 * H + h];
    const float ug = a.Ug[h], um = a.Um[h], ult = a.Ult[h];
    // global / motion attention (:389-412)
    // (deferred mode, a.DAL != null: the per-(clip, frame) accumulators DPG / DPM / DG / DM / DUACC are not touched
    // here -- k_att_accum / k_att_accum2 sum all steps after the loop)
    const bool rmw = a.DAL == nullptr;
    const float tg = tanh_bw(pg + sg);
    const float dqg = daG * ug * (1.0f - tg * tg);
    a.DSQP[o4 + H] = dqg;
    const float tm = tanh_bw(pm + sm);
    const float dqm = daM * um * (1.0f - tm * tm);
    a.DSQP[o4 + 2 * H] = dqm;
    if (rmw) {
      a.DPG[o] += dqg;
      a.DUACC[o4 + H] += daG * tg;
      if (a.global_proj) a.DG[o] += aG * dc;
      a.DPM[o] += dqm;
      a.DUACC[o4 + 2 * H] += daM * tm;
      a.DM[o] += aM * dc;
    }
    // local-temporal attention (:415-426), pLT = sum_r alpha_l Q_r + blt + slt
    float plt = slt;
#pragma unroll
    for (int r = 0; r < RU; ++r)
      if (RT || r < R) plt = fmaf(alr[r], q[r], plt);
    const float tl = tanh_bw(plt);
    const float dp = daLT * ult * (1.0f - tl * tl);
    a.DSQP[o4 + 3 * H] = dp;
    if (rmw) a.DUACC[o4 + 3 * H] += daLT * tl;
    const float dcL = aLT * dc;
    dcl[k] = dcL;
    dpl[k] = dp;
    if (a.DPLT) a.DPLT[nt * H + h] = dp;
#pragma unroll
    for (int r = 0; r < RU; ++r)
      if (RT || r < R) part[r] = fmaf(dcL, lc[r], fmaf(dp, q[r], part[r]));
  }
  // spatial soft-max backward (:380-383)
  block_sum_n<RU>(part, R, sh);
  float s = 0.f;
#pragma unroll
  for (int r = 0; r < RU; ++r) {
    if (RT || r < R) {
      part[r] += a.COVL[static_cast<size_t>(bt) * R + r];
      s = fmaf(alr[r], part[r], s);
    }
  }
  float dal[RU], dsum = 0.f;
#pragma unroll
  for (int r = 0; r < RU; ++r) {
    dal[r] = (RT || r < R) ? alr[r] * (part[r] - s) : 0.f;
    dsum += dal[r];
  }
  if (threadIdx.x == 0) a.DCACC[static_cast<size_t>(bt) * 4] += dsum;
  if (a.DAL) {
#pragma unroll
    for (int r = 0; r < RU; ++r)
      if ((RT || r < R) && threadIdx.x == r) a.DAL[nt * R + r] = dal[r];
  }
  k = 0;
  for (int h = threadIdx.x; h < H; h += blockDim.x, ++k) {
    const float sl = hq[h], ul = a.Ul[h];
    float pl[RU];
#pragma unroll
    for (int r = 0; r < RU; ++r) pl[r] = (RT || r < R) ? a.pL[(static_cast<size_t>(bt) * R + r) * H + h] : 0.f;
    float dsl = 0.f, dul = 0.f;
#pragma unroll
    for (int r = 0; r < RU; ++r) {
      if (RT || r < R) {
        const size_t ol = (static_cast<size_t>(bt) * R + r) * H + h;
        const float tl = tanh_bw(pl[r] + sl);
        const float dq = dal[r] * ul * (1.0f - tl * tl);
        dsl += dq;
        dul = fmaf(dal[r], tl, dul);
        if (!a.DAL) {
          a.DPL[ol] += dq;
          a.DLC[ol] = fmaf(alr[r], dcl[k], a.DLC[ol]);
          a.DQ[ol] = fmaf(alr[r], dpl[k], a.DQ[ol]);
        }
      }
    }
    const size_t o4 = static_cast<size_t>(bt) * 4 * H + h;
    a.DSQP[o4] = dsl;
    if (!a.DAL) a.DUACC[o4] += dul;
  }
}

// STAT_BW_FAST: the step-invariant local gradient blocks in one pass after the loop -- per (clip, frame, region,
// column) the sum over all steps, in registers, of what k_att_main otherwise read-modify-writes every step:
//   DPL = sum_s dal_r(s) Ul (1 - tanh^2(pL + sl_s)),  DLC = sum_s alpha_l,r(s) alpha_lt(s) dC_s,  DQ = sum_s alpha_l,r(s) dpLT_s
__global__ void k_att_accum(int L, int B, int T, int R, int H, int ldq, const float *al, const float *alt,
                            const float *DAL, const float *DPLT, const float *DCS, const float *HQ, const float *pL,
                            const float *Ul, float *DPL, float *DLC, float *DQ, float *DUACC) {
  const int bt = blockIdx.x, b = bt / T, t = bt % T;
  for (int h = threadIdx.x; h < H; h += blockDim.x) {
    const float ul = Ul[h];
    float dul = 0.f;                 // sum over steps and regions of dal * tanh(pL + sl): the Ul-gradient share
    for (int r = 0; r < R; ++r) {
      const size_t ol = (static_cast<size_t>(bt) * R + r) * H + h;
      const float pl = pL[ol];
      float apl = 0.f, alc = 0.f, aq = 0.f;
#pragma unroll 4
      for (int s = L - 1; s >= 0; --s) {
        const size_t n = static_cast<size_t>(s) * B + b, nt = n * T + t;
        const float ar = al[nt * R + r];
        const float da = DAL[nt * R + r];
        const float tl = tanh_bw(pl + HQ[n * ldq + h]);
        apl = fmaf(da * ul, 1.0f - tl * tl, apl);
        dul = fmaf(da, tl, dul);
        alc = fmaf(ar, alt[nt] * DCS[n * H + h], alc);
        aq = fmaf(ar, DPLT[nt * H + h], aq);
      }
      DPL[ol] = apl;
      DLC[ol] = alc;
      DQ[ol] = aq;
    }
    DUACC[static_cast<size_t>(bt) * 4 * H + h] = dul;
  }
}

// Deferred mode, the global / motion / local-temporal shares: per (clip, frame, column) the sum over all steps of what
// k_att_main otherwise read-modify-writes every step (218 MB per step at B = 128):
//   DPG = sum_s daG_s Ug (1 - tanh^2(pG + sg_s)),  DUACC_g = sum_s daG_s tanh(pG + sg_s),  DG = sum_s alpha_g(s) dC_s
//   (same for m),  DUACC_lt = sum_s daLT_s tanh(sum_r alpha_l,r(s) Q_r + slt_s)
template <int RT>
__global__ void k_att_accum2(int L, int B, int T, int R_, int H, int ldq, int global_proj, const float *al,
                             const float *ag, const float *am, const float *DS3S, const float *DCS, const float *HQ,
                             const float *pG, const float *pM, const float *Q, const float *Ug, const float *Um,
                             float *DPG, float *DPM, float *DG, float *DM, float *DUACC) {
  const int R = RT ? RT : R_;
  constexpr int RU = RT ? RT : RMAX;
  const int bt = blockIdx.x, b = bt / T, t = bt % T;
  const size_t BT = static_cast<size_t>(B) * T;
  for (int h = threadIdx.x; h < H; h += blockDim.x) {
    const size_t o = static_cast<size_t>(bt) * H + h, o4 = static_cast<size_t>(bt) * 4 * H + h;
    const float pg = pG[o], pm = pM[o], ug = Ug[h], um = Um[h];
    float q[RU];
#pragma unroll
    for (int r = 0; r < RU; ++r) q[r] = (RT || r < R) ? Q[(static_cast<size_t>(bt) * R + r) * H + h] : 0.f;
    float dpg = 0.f, dpm = 0.f, dg = 0.f, dm = 0.f, dug = 0.f, dum = 0.f, dult = 0.f;
#pragma unroll 2
    for (int s = L - 1; s >= 0; --s) {
      const size_t n = static_cast<size_t>(s) * B + b, nt = n * T + t;
      const float *ds = DS3S + static_cast<size_t>(s) * 3 * BT;
      const float daG = ds[bt], daM = ds[BT + bt], daLT = ds[2 * BT + bt];
      const float *hq = HQ + n * ldq;
      const float dc = DCS[n * H + h];
      const float tg = tanh_bw(pg + hq[H + h]);
      dpg = fmaf(daG * ug, 1.0f - tg * tg, dpg);
      dug = fmaf(daG, tg, dug);
      dg = fmaf(ag[nt], dc, dg);
      const float tm = tanh_bw(pm + hq[2 * H + h]);
      dpm = fmaf(daM * um, 1.0f - tm * tm, dpm);
      dum = fmaf(daM, tm, dum);
      dm = fmaf(am[nt], dc, dm);
      float plt = hq[3 * H + h];
#pragma unroll
      for (int r = 0; r < RU; ++r)
        if (RT || r < R) plt = fmaf(al[nt * R + r], q[r], plt);
      dult = fmaf(daLT, tanh_bw(plt), dult);
    }
    DPG[o] = dpg;
    DPM[o] = dpm;
    if (global_proj) DG[o] = dg;
    DM[o] = dm;
    DUACC[o4 + H] = dug;
    DUACC[o4 + 2 * H] = dum;
    DUACC[o4 + 3 * H] = dult;
  }
}

// DHQ[(t,b)][col] = sum_frames DSQP[b][frame][col], col < 4H
__global__ void k_reduce_t(int t_step, const float *DSQP, float *DHQ, int ldq, int B, int T, int H4) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * H4) return;
  const int b = i / H4, c = i % H4;
  float s = 0.f;
  for (int t = 0; t < T; ++t) s += DSQP[(static_cast<size_t>(b) * T + t) * H4 + c];
  DHQ[(static_cast<size_t>(t_step) * B + b) * ldq + c] = s;
}

// DP0[b] = [dh0 * (1 - h0^2) | dc0 * (1 - c0^2)]   (:657-660)
__global__ void k_init_bw(const float *DHc, const float *DCc, const float *h0c0, float *DP0, int B, int H) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * H) return;
  const int b = i / H, j = i % H;
  const float h0 = h0c0[static_cast<size_t>(b) * 2 * H + j], c0 = h0c0[static_cast<size_t>(b) * 2 * H + H + j];
  DP0[static_cast<size_t>(b) * 2 * H + j] = DHc[i] * (1.0f - h0 * h0);
  DP0[static_cast<size_t>(b) * 2 * H + H + j] = DCc[i] * (1.0f - c0 * c0);
}

// dWemb[x[t-1][b]] += DEMB[(t,b)] for t >= 1 (:613-617); one thread per column, rows in order:
// deterministic without atomics
__global__ void k_scatter_emb(const float *DEMB, const int64_t *x, float *dWemb, int L, int B, int E) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= E) return;
  for (size_t n = B; n < static_cast<size_t>(L) * B; ++n)
    dWemb[static_cast<size_t>(x[n - B]) * E + e] += DEMB[n * E + e];
}

// Default embedding scatter: one block per (step, clip) row i; the block of the FIRST row whose previous word is v
// owns dWemb[v] and sums the rows of that word in row order (deterministic), every other block leaves after the
// scan of the earlier rows.  dWemb is zero-filled beforehand.  O(M^2 / threads) token compares, M = (L-1)*B.
__global__ void k_scatter_emb_owner(const float *DEMB, const int64_t *x, float *dWemb, int L, int B, int E) {
  constexpr int CH = 256, EPT = 8;
  __shared__ int s_found, s_cnt;
  __shared__ int s_flag[CH];
  __shared__ int s_list[CH];
  const int M = (L - 1) * B;
  const int i = blockIdx.x;
  const int nt = blockDim.x;
  const int64_t tok = x[i];
  if (threadIdx.x == 0) s_found = 0;
  __syncthreads();
  bool f = false;
  for (int j = threadIdx.x; j < i; j += nt) f = f || (x[j] == tok);
  if (f) s_found = 1;
  __syncthreads();
  if (s_found) return;
  for (int e0 = 0; e0 < E; e0 += nt * EPT) {
    float acc[EPT];
#pragma unroll
    for (int q = 0; q < EPT; ++q) acc[q] = 0.f;
    for (int base = i; base < M; base += CH) {
      for (int k = threadIdx.x; k < CH; k += nt) s_flag[k] = (base + k < M && x[base + k] == tok) ? 1 : 0;
      __syncthreads();
      if (threadIdx.x == 0) {
        int n = 0;
        for (int k = 0; k < CH; ++k)
          if (s_flag[k]) s_list[n++] = base + k;
        s_cnt = n;
      }
      __syncthreads();
      const int n = s_cnt;
      for (int m = 0; m < n; ++m) {
        const float *src = DEMB + (static_cast<size_t>(s_list[m]) + B) * E;
#pragma unroll
        for (int q = 0; q < EPT; ++q) {
          const int e = e0 + threadIdx.x + q * nt;
          if (e < E) acc[q] += src[e];
        }
      }
      __syncthreads();
    }
#pragma unroll
    for (int q = 0; q < EPT; ++q) {
      const int e = e0 + threadIdx.x + q * nt;
      if (e < E) dWemb[static_cast<size_t>(tok) * E + e] = acc[q];
    }
  }
}

// C[m][n] = sum over the k-slice planes, in order (STAT_BW_FAST: k-split products)
__global__ void k_sum_planes(float *C, int ldc, const float *planes, size_t plane, int ks, int M, int N) {
  const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= static_cast<size_t>(M) * N) return;
  float s = 0.f;
  for (int z = 0; z < ks; ++z) s += planes[static_cast<size_t>(z) * plane + i];
  C[(i / N) * ldc + (i % N)] = s;
}

// ---------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------
// Default (STAT_BW_FAST=0 switches it off): products whose output has too few
// tiles to fill the SMs are k-split into planes and summed in order; the embedding scatter runs one block per
// vocabulary row; weights that sit 4-byte aligned in the caller's flat buffer are copied to 16-byte aligned
// scratch so that they take the tensor-core path.
struct MmCtx {
  bool fast = false;
  float *planes = nullptr;
};
thread_local MmCtx g_mm;
constexpr size_t PLANES_FLOATS = static_cast<size_t>(148) * 128 * 128 + 64 * 64;

inline dim3 g1(size_t n, int bs = 256) { return dim3(static_cast<unsigned>((n + bs - 1) / bs)); }

// C (M,N) = alpha * A (M,K) . Bt (N,K)^T + bias[N]
int mm(const float *A, int lda, const float *Bt, int ldb, float *C, int ldc, int M, int N, int K, const float *bias,
       cudaStream_t st) {
  GemmArgs g;
  memset(&g, 0, sizeof(g));
  g.P = A; g.ldp = lda; g.NP = M;
  g.Q = Bt; g.ldq = ldb; g.NQ = N;
  g.K = K; g.feat_on_p = 0; g.nseg = 1; g.ksplit = 1;
  g.seg[0] = GemmSeg{C, ldc, bias, nullptr, 0, 1.f, 1.f, 0, 0, N};
  if (g_mm.fast && g_mm.planes) {
    const int bq = N > 64 ? 128 : (N > 32 ? 64 : 32);
    const int tiles = ((M + 127) / 128) * ((N + bq - 1) / bq);
    const int nk = (K + 31) / 32;
    int ks = 148 / tiles;
    if (ks > 32) ks = 32;
    if (ks > nk) ks = nk;
    const size_t plane = up(static_cast<size_t>(M) * N, 64);
    if (ks >= 2 && plane * ks <= PLANES_FLOATS) {
      g.ksplit = ks;
      g.plane = plane;
      g.seg[0].C = g_mm.planes;
      g.seg[0].ldc = N;
      STAT_TRY(gemm_launch(g, st));
      return BW_LAUNCH(k_sum_planes, g1(static_cast<size_t>(M) * N), dim3(256), st, C, ldc, g_mm.planes, plane, ks, M, N);
    }
  }
  return gemm_launch(g, st);
}

int transpose(const float *src, int rows, int cols, int lds, float *dst, int ldd, cudaStream_t st) {
  return BW_LAUNCH(k_transpose, dim3((cols + 31) / 32, (rows + 31) / 32), dim3(256), st, src, rows, cols, lds, dst, ldd);
}

// dst[c] = sum_r src[r][c]; scratch holds 2 * ceil(rows/64) * cols floats
int colsum(const float *src, int rows, int cols, int ld, float *dst, float *scratch, cudaStream_t st) {
  const int per = 64;
  float *buf[2] = {scratch, scratch + static_cast<size_t>((rows + per - 1) / per) * cols};
  int which = 0;
  while (true) {
    const int chunks = (rows + per - 1) / per;
    float *out = chunks == 1 ? dst : buf[which];
    STAT_TRY(BW_LAUNCH(k_colsum, dim3((cols + 127) / 128, chunks), dim3(128), st, src, rows, cols, ld, per, out, cols));
    if (chunks == 1) return STAT_OK;
    src = out; rows = chunks; ld = cols;
    which ^= 1;
  }
}

struct GW {   // float offsets into the gradient workspace
  size_t Whcat, WhT, bq, WcT, WctxT, WlT, WvT, WvP, WdT;
  size_t Hprev, EMB, HQ, csum, beta, ctx, XW, CW, ZC, GATES, Call, HD, ZT, Z, LOG, DLT;
  size_t DZ, DZP, DHR, DCR, DHQ, DHc, DCc, DHm, DCTX, DC, TMPH, CL, DA3, DS3, COV3, COVL, DSQP;
  size_t DPG, DPM, DG, DM, DPL, DLC, DQ, DUACC, DCACC;
  size_t T1, T2, T3, DEMB, DWH, GBAR, DP0, SMALL, CS, PLANES, ALN, DAL, DPLT, DCS, DS3S;
  size_t total;
  int ldq, Vp;
};

GW gw_layout(const StatDims &d, int L) {
  GW w;
  const size_t B = d.B, T = d.T, R = d.R, H = d.H, E = d.E, V = d.V, N = static_cast<size_t>(L) * B;
  const size_t ldq = up(8 * H + 1, 4), Vp = up(V, 4);
  const size_t BT = B * T, BTR = B * T * R;
  size_t o = 0;
  auto take = [&](size_t n) {
    size_t at = o;
    o = up(o + n, 64);
    return at;
  };
  w.ldq = static_cast<int>(ldq);
  w.Vp = static_cast<int>(Vp);
  w.Whcat = take(H * ldq); w.WhT = take(ldq * H); w.bq = take(ldq);
  w.WcT = take(4 * H * H); w.WctxT = take(E * H); w.WlT = take(E * H);
  w.WvT = take(V * E); w.WvP = take(E * Vp); w.WdT = take(4 * H * E);
  w.Hprev = take(N * H); w.EMB = take(N * E); w.HQ = take(N * ldq); w.csum = take(N * H); w.beta = take(N);
  w.ctx = take(N * H); w.XW = take(N * 4 * H); w.CW = take(N * 4 * H); w.ZC = take(N * E);
  w.GATES = take(N * 4 * H); w.Call = take(N * H); w.HD = take(N * H); w.ZT = take(N * E); w.Z = take(N * E);
  w.LOG = take(N * Vp); w.DLT = take(V * N);
  w.DZ = take(N * E); w.DZP = take(N * E); w.DHR = take(N * H); w.DCR = take(N * H); w.DHQ = take(N * ldq);
  w.DHc = take(B * H); w.DCc = take(B * H); w.DHm = take(B * H); w.DCTX = take(B * H); w.DC = take(B * H);
  w.TMPH = take(B * H); w.CL = take(BT * H); w.DA3 = take(3 * BT); w.DS3 = take(3 * BT); w.COV3 = take(3 * BT);
  w.COVL = take(BTR); w.DSQP = take(BT * 4 * H);
  w.DPG = take(BT * H); w.DPM = take(BT * H); w.DG = take(BT * H); w.DM = take(BT * H);
  w.DPL = take(BTR * H); w.DLC = take(BTR * H); w.DQ = take(BTR * H);
  w.DUACC = take(BT * 4 * H); w.DCACC = take(BT * 4);
  // transposed operands of the tall-K products: T1 the wider one, T2 H- / ldq-wide, T3 a (rows,H) temporary
  size_t dmax = std::max(std::max(static_cast<size_t>(d.Dr), static_cast<size_t>(d.Dm)),
                         std::max(static_cast<size_t>(d.Dg), std::max(H, E)));
  const size_t kmax = std::max(BTR, N);
  w.T1 = take(dmax * kmax);
  w.T2 = take(std::max(ldq, std::max(E, 2 * H)) * kmax);
  w.T3 = take(kmax * std::max(H, E));
  w.DEMB = take(N * E); w.DWH = take(H * ldq); w.GBAR = take(B * d.Dg); w.DP0 = take(B * 2 * H);
  w.SMALL = take(up(4 * H + 4, 64) + 64);
  const size_t cmax = std::max(std::max(ldq, Vp), 4 * H);
  w.CS = take(2 * ((kmax + 63) / 64 + 1) * cmax);
  w.PLANES = take(PLANES_FLOATS);
  w.ALN = take(8 * H * H + 6 * H * E + 8 * 64);
  w.DAL = take(N * T * R);
  w.DPLT = take(N * T * H);
  w.DCS = take(N * H);
  w.DS3S = take(static_cast<size_t>(L) * 3 * BT);
  w.total = o;
  return w;
}

}  // namespace bw
}  // namespace stat

using namespace stat;
using namespace stat::bw;

extern "C" {

int stat_grad_profile_phases(void) { return BP_COUNT; }

const char *stat_grad_profile_phase_name(int phase) {
  return (phase >= 0 && phase < BP_COUNT) ? kBwPhaseNames[phase] : nullptr;
}

int stat_grad_profile_enable(int on) {
#ifndef STAT_EMU
  for (auto &r : g_bw_recs) {
    cudaEventDestroy(r.a);
    cudaEventDestroy(r.b);
  }
  g_bw_recs.clear();
  g_bw_open = false;
  g_bw_prof_on = on != 0;
#else
  (void)on;
#endif
  return STAT_OK;
}

int stat_grad_profile_collect(float *ms_by_phase, int *count_by_phase, int nphase) {
  STAT_REQUIRE(ms_by_phase && count_by_phase && nphase >= BP_COUNT, STAT_EINVAL,
               "grad_profile_collect: need arrays of %d entries", static_cast<int>(BP_COUNT));
  for (int i = 0; i < nphase; ++i) {
    ms_by_phase[i] = 0.f;
    count_by_phase[i] = 0;
  }
#ifndef STAT_EMU
  for (auto &r : g_bw_recs) {
    STAT_CUDA_CHECK(cudaEventSynchronize(r.b));
    float ms = 0.f;
    STAT_CUDA_CHECK(cudaEventElapsedTime(&ms, r.a, r.b));
    ms_by_phase[r.phase] += ms;
    count_by_phase[r.phase] += 1;
    cudaEventDestroy(r.a);
    cudaEventDestroy(r.b);
  }
  g_bw_recs.clear();
#endif
  return STAT_OK;
}

size_t stat_grad_workspace_bytes(const StatDims *d, int L) {
  if (!d || L < 1 || d->B < 1 || d->H < 1) return 0;
  return gw_layout(*d, L).total * sizeof(float);
}

int stat_grad_shared(const StatDims *d, const StatParams *p, const StatFwdBlocks *f, int L, const int64_t *x,
                     const float *mask, const float *ctxg, const float *mask_ctxg, const float *ctxl,
                     const float *ctxm, const float *dp_gates, const float *dp_h, const float *dp_z,
                     const float *alpha_l, const float *alpha_g, const float *alpha_m, const float *alpha_lt,
                     const float *h_all, float inv_batch, float alpha_c, float decay_c, const StatParams *grads,
                     void *gws, void *stream) {
  STAT_REQUIRE(d && p && f && grads && gws && x && mask && ctxg && mask_ctxg && ctxl && ctxm, STAT_EINVAL,
               "grad_shared: NULL argument");
  STAT_REQUIRE(alpha_l && alpha_g && alpha_m && alpha_lt && h_all && L >= 1, STAT_EINVAL,
               "grad_shared: the forward's attention weights and hidden states are required");
  STAT_REQUIRE(d->R >= 1 && d->R <= RMAX && d->H >= 1 && d->H <= NT * HMAX_PT, STAT_EINVAL,
               "grad_shared: need 1<=R<=%d and H<=%d (R=%d H=%d)", RMAX, NT * HMAX_PT, d->R, d->H);
  STAT_REQUIRE(f->ctxg0 && f->pctxg && f->ctxm0 && f->pctxm && f->ctxl0 && f->pctxl && f->qctxl && f->h0c0,
               STAT_EINVAL, "grad_shared: forward blocks missing");
  STAT_REQUIRE((reinterpret_cast<uintptr_t>(gws) & 15) == 0, STAT_EALIGN, "grad_shared: workspace must be 16-byte aligned");
  STAT_REQUIRE(d->B >= 1 && d->T >= 1 && d->E >= 1 && d->V >= 2 && d->Dg >= 1 && d->Dm >= 1 && d->Dr >= 1, STAT_EINVAL,
               "grad_shared: bad dims");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int B = d->B, T = d->T, R = d->R, H = d->H, E = d->E, V = d->V, Dg = d->Dg, Dm = d->Dm, Dr = d->Dr;
  const bool sel = d->flags & STAT_SELECTOR, c2o = d->flags & STAT_CTX2OUT, p2o = d->flags & STAT_PREV2OUT,
             gp = d->flags & STAT_GLOBAL_PROJ;
  const int N = L * B, BT = B * T, BTR = B * T * R, NH = 8 * H + (sel ? 1 : 0);
  const GW w = gw_layout(*d, L);
  const int ldq = w.ldq, Vp = w.Vp;
  float *W = static_cast<float *>(gws);
  const size_t F = sizeof(float);
  // gradient outputs (the struct's pointers name writable buffers)
#define GRAD(name) const_cast<float *>(grads->name)
#define NEEDG(name) STAT_REQUIRE(grads->name != nullptr && p->name != nullptr, STAT_EINVAL, "grad_shared: %s is NULL", #name)
  NEEDG(Wemb); NEEDG(ff_state_W); NEEDG(ff_state_b); NEEDG(ff_memory_W); NEEDG(ff_memory_b);
  NEEDG(ff_local_W); NEEDG(ff_local_b); NEEDG(ff_motion_W); NEEDG(ff_motion_b);
  NEEDG(decoder_W); NEEDG(decoder_U); NEEDG(decoder_b); NEEDG(decoder_Wc);
  NEEDG(decoder_Wcg_att); NEEDG(decoder_Wcm_att); NEEDG(decoder_Wclt_att);
  NEEDG(decoder_Wdg_att); NEEDG(decoder_Wdm_att); NEEDG(decoder_Wdlt_att);
  NEEDG(decoder_bg_att); NEEDG(decoder_bm_att); NEEDG(decoder_blt_att);
  NEEDG(decoder_Wcl_att); NEEDG(decoder_Wdl_att); NEEDG(decoder_bl_att);
  NEEDG(decoder_Ug_att); NEEDG(decoder_cg_att); NEEDG(decoder_Um_att); NEEDG(decoder_cm_att);
  NEEDG(decoder_Ult_att); NEEDG(decoder_clt_att); NEEDG(decoder_Ul_att); NEEDG(decoder_cl_att);
  NEEDG(ff_logit_lstm_W); NEEDG(ff_logit_lstm_b); NEEDG(ff_logit_W); NEEDG(ff_logit_b);
  if (sel) { NEEDG(decoder_W_sel); NEEDG(decoder_b_sel); }
  if (c2o) { NEEDG(ff_logit_ctxglm_W); NEEDG(ff_logit_ctxglm_b); }
  if (gp) { NEEDG(ff_global_W); NEEDG(ff_global_b); }
#undef NEEDG
  auto cp2d = [&](float *dst, int ldd, const float *src, int lds, int width, int rows) {
    return cudaMemcpy2DAsync(dst, F * ldd, src, F * lds, F * width, rows, cudaMemcpyDeviceToDevice, st);
  };
  auto zero = [&](size_t off, size_t n) { return cudaMemsetAsync(W + off, 0, n * F, st); };
  const char *fast_env = getenv("STAT_BW_FAST");
  const bool fast = !(fast_env && fast_env[0] == '0');     // default on; STAT_BW_FAST=0 = the plain first version
  g_mm.fast = fast;
  g_mm.planes = W + w.PLANES;
  // weight operands read in the caller's layout: 16-byte aligned copies when needed (fast mode only)
  size_t aln_used = 0;
  auto aligned = [&](const float *src, size_t n) -> const float * {
    if (!fast || (reinterpret_cast<uintptr_t>(src) & 15) == 0) return src;
    float *dst = W + w.ALN + aln_used;
    aln_used += up(n, 64);
    if (cudaMemcpyAsync(dst, src, n * F, cudaMemcpyDeviceToDevice, st) != cudaSuccess) return src;
    return dst;
  };
  const size_t HH = static_cast<size_t>(H) * H, HE = static_cast<size_t>(H) * E;
  const float *Wc_a = aligned(p->decoder_Wc, 4 * HH), *Wctx_a = c2o ? aligned(p->ff_logit_ctxglm_W, HE) : nullptr;
  const float *Wl_a = aligned(p->ff_logit_lstm_W, HE), *Wd_a = aligned(p->decoder_W, 4 * HE);
  const float *Wcg_a = aligned(p->decoder_Wcg_att, HH), *Wcm_a = aligned(p->decoder_Wcm_att, HH);
  const float *Wcl_a = aligned(p->decoder_Wcl_att, HH), *Wclt_a = aligned(p->decoder_Wclt_att, HH);

  bw_mark(BP_LAYOUT, st);
  // ---- operand layouts --------------------------------------------------------------------
  // Whcat (H, ldq) = [Wdl | Wdg | Wdm | Wdlt | U | W_sel | 0]: K-major operand of the dh_{t-1} product;
  // its transpose is the operand of the hidden-state projections
  STAT_CUDA_CHECK(zero(w.Whcat, static_cast<size_t>(H) * ldq));
  STAT_CUDA_CHECK(zero(w.bq, ldq));
  STAT_CUDA_CHECK(cp2d(W + w.Whcat, ldq, p->decoder_Wdl_att, H, H, H));
  STAT_CUDA_CHECK(cp2d(W + w.Whcat + H, ldq, p->decoder_Wdg_att, H, H, H));
  STAT_CUDA_CHECK(cp2d(W + w.Whcat + 2 * H, ldq, p->decoder_Wdm_att, H, H, H));
  STAT_CUDA_CHECK(cp2d(W + w.Whcat + 3 * H, ldq, p->decoder_Wdlt_att, H, H, H));
  STAT_CUDA_CHECK(cp2d(W + w.Whcat + 4 * H, ldq, p->decoder_U, 4 * H, 4 * H, H));
  STAT_CUDA_CHECK(cudaMemcpyAsync(W + w.bq + 3 * H, p->decoder_blt_att, F * H, cudaMemcpyDeviceToDevice, st));
  if (sel) {
    STAT_CUDA_CHECK(cp2d(W + w.Whcat + 8 * H, ldq, p->decoder_W_sel, 1, 1, H));
    STAT_CUDA_CHECK(cudaMemcpyAsync(W + w.bq + 8 * H, p->decoder_b_sel, F, cudaMemcpyDeviceToDevice, st));
  }
  STAT_TRY(transpose(W + w.Whcat, H, ldq, ldq, W + w.WhT, H, st));
  STAT_TRY(transpose(p->decoder_Wc, H, 4 * H, 4 * H, W + w.WcT, H, st));
  if (c2o) STAT_TRY(transpose(p->ff_logit_ctxglm_W, H, E, E, W + w.WctxT, H, st));
  STAT_TRY(transpose(p->ff_logit_lstm_W, H, E, E, W + w.WlT, H, st));
  STAT_TRY(transpose(p->ff_logit_W, E, V, V, W + w.WvT, E, st));
  STAT_CUDA_CHECK(zero(w.WvP, static_cast<size_t>(E) * Vp));
  STAT_CUDA_CHECK(cp2d(W + w.WvP, Vp, p->ff_logit_W, V, V, E));
  STAT_TRY(transpose(p->decoder_W, E, 4 * H, 4 * H, W + w.WdT, E, st));

  bw_mark(BP_RECOMPUTE, st);
  // ---- A: recompute over all (step, clip) rows -------------------------------------------------
  const size_t NHs = static_cast<size_t>(N) * H, NEs = static_cast<size_t>(N) * E;
  STAT_TRY(BW_LAUNCH(k_gather_prev, g1(NHs), dim3(256), st, f->h0c0, h_all, W + w.Hprev, L, B, H));
  STAT_TRY(BW_LAUNCH(k_gather_emb, g1(NEs), dim3(256), st, p->Wemb, x, W + w.EMB, L, B, E));
  STAT_CUDA_CHECK(zero(w.HQ, static_cast<size_t>(N) * ldq));
  STAT_TRY(mm(W + w.Hprev, H, W + w.WhT, H, W + w.HQ, ldq, N, NH, H, W + w.bq, st));                 // :371,389,402,415,433,437
  STAT_TRY(BW_LAUNCH(k_ctx_parts, g1(NHs), dim3(256), st, alpha_l, alpha_g, alpha_m, alpha_lt, f->ctxl0, f->ctxg0,
                     f->ctxm0, W + w.csum, L, B, T, R, H));
  STAT_TRY(BW_LAUNCH(k_beta_ctx, g1(NHs), dim3(256), st, W + w.HQ, ldq, 8 * H, W + w.csum, W + w.beta, W + w.ctx,
                     static_cast<size_t>(N), H, sel ? 1 : 0));
  STAT_TRY(mm(W + w.EMB, E, W + w.WdT, E, W + w.XW, 4 * H, N, 4 * H, E, p->decoder_b, st));          // :334-335
  STAT_TRY(mm(W + w.ctx, H, W + w.WcT, H, W + w.CW, 4 * H, N, 4 * H, H, nullptr, st));               // :439
  STAT_TRY(BW_LAUNCH(k_cell_forward, g1(static_cast<size_t>(B) * H), dim3(256), st, W + w.HQ, ldq, 4 * H, W + w.XW,
                     W + w.CW, f->h0c0, mask, dp_gates, W + w.GATES, W + w.Call, L, B, H));
  STAT_TRY(BW_LAUNCH(k_scale_dp, g1(NHs), dim3(256), st, W + w.HD, h_all, dp_h, NHs));               // :684-686
  STAT_TRY(mm(W + w.HD, H, W + w.WlT, H, W + w.DZ, E, N, E, H, p->ff_logit_lstm_b, st));             // ZP in DZ for now
  if (c2o) STAT_TRY(mm(W + w.ctx, H, W + w.WctxT, H, W + w.ZC, E, N, E, H, p->ff_logit_ctxglm_b, st));
  STAT_TRY(BW_LAUNCH(k_zact, g1(NEs), dim3(256), st, W + w.DZ, p2o ? W + w.EMB : nullptr, c2o ? W + w.ZC : nullptr,
                     dp_z, W + w.ZT, W + w.Z, NEs));
  bw_mark(BP_LOGITS, st);
  STAT_CUDA_CHECK(zero(w.LOG, static_cast<size_t>(N) * Vp));
  STAT_TRY(mm(W + w.Z, E, W + w.WvT, E, W + w.LOG, Vp, N, V, E, p->ff_logit_b, st));                 // :704-705

  bw_mark(BP_READOUT_BW, st);
  // ---- B: readout backward ---------------------------------------------------------------------
  STAT_TRY(BW_LAUNCH(k_softmax_nll, dim3(N), dim3(256), st, W + w.LOG, Vp, V, x, mask, inv_batch));
  STAT_TRY(colsum(W + w.LOG, N, V, Vp, GRAD(ff_logit_b), W + w.CS, st));
  STAT_TRY(transpose(W + w.LOG, N, V, Vp, W + w.DLT, N, st));
  STAT_TRY(transpose(W + w.Z, N, E, E, W + w.T1, N, st));
  STAT_TRY(mm(W + w.T1, N, W + w.DLT, N, GRAD(ff_logit_W), V, E, V, N, nullptr, st));
  STAT_TRY(mm(W + w.LOG, Vp, W + w.WvP, Vp, W + w.DZ, E, N, E, Vp, nullptr, st));
  STAT_TRY(BW_LAUNCH(k_dzp, g1(NEs), dim3(256), st, W + w.DZ, W + w.ZT, dp_z, W + w.DZP, NEs));
  STAT_TRY(colsum(W + w.DZP, N, E, E, GRAD(ff_logit_lstm_b), W + w.CS, st));
  STAT_TRY(transpose(W + w.DZP, N, E, E, W + w.T2, N, st));                                          // DZP^T (E,N)
  STAT_TRY(transpose(W + w.HD, N, H, H, W + w.T1, N, st));
  STAT_TRY(mm(W + w.T1, N, W + w.T2, N, GRAD(ff_logit_lstm_W), E, H, E, N, nullptr, st));
  if (c2o) {
    STAT_CUDA_CHECK(cudaMemcpyAsync(GRAD(ff_logit_ctxglm_b), GRAD(ff_logit_lstm_b), F * E, cudaMemcpyDeviceToDevice, st));
    STAT_TRY(transpose(W + w.ctx, N, H, H, W + w.T1, N, st));
    STAT_TRY(mm(W + w.T1, N, W + w.T2, N, GRAD(ff_logit_ctxglm_W), E, H, E, N, nullptr, st));
    STAT_TRY(mm(W + w.DZP, E, Wctx_a, E, W + w.DCR, H, N, H, E, nullptr, st));
  } else {
    STAT_CUDA_CHECK(zero(w.DCR, NHs));
  }
  STAT_TRY(mm(W + w.DZP, E, Wl_a, E, W + w.DHR, H, N, H, E, nullptr, st));
  STAT_TRY(BW_LAUNCH(k_scale_dp, g1(NHs), dim3(256), st, W + w.DHR, W + w.DHR, dp_h, NHs));
  // coverage terms of the four attentions (:1138-1147): mean over T (T*R for alpha_l), sum over clips
  {
    const float sc = 2.0f * alpha_c / static_cast<float>(T), scl = 2.0f * alpha_c / static_cast<float>(T * R);
    STAT_TRY(BW_LAUNCH(k_cov, g1(BT), dim3(256), st, alpha_g, L, static_cast<size_t>(BT), sc, W + w.COV3));
    STAT_TRY(BW_LAUNCH(k_cov, g1(BT), dim3(256), st, alpha_m, L, static_cast<size_t>(BT), sc, W + w.COV3 + BT));
    STAT_TRY(BW_LAUNCH(k_cov, g1(BT), dim3(256), st, alpha_lt, L, static_cast<size_t>(BT), sc, W + w.COV3 + 2 * BT));
    STAT_TRY(BW_LAUNCH(k_cov, g1(BTR), dim3(256), st, alpha_l, L, static_cast<size_t>(BTR), scl, W + w.COVL));
  }

  // ---- C: back-propagation through time ------------------------------------------------------------
  STAT_CUDA_CHECK(zero(w.DHQ, static_cast<size_t>(N) * ldq));
  STAT_CUDA_CHECK(zero(w.DHc, static_cast<size_t>(B) * H));
  STAT_CUDA_CHECK(zero(w.DCc, static_cast<size_t>(B) * H));
  STAT_CUDA_CHECK(zero(w.DPG, static_cast<size_t>(BT) * H));
  STAT_CUDA_CHECK(zero(w.DPM, static_cast<size_t>(BT) * H));
  STAT_CUDA_CHECK(zero(w.DG, static_cast<size_t>(BT) * H));
  STAT_CUDA_CHECK(zero(w.DM, static_cast<size_t>(BT) * H));
  STAT_CUDA_CHECK(zero(w.DPL, static_cast<size_t>(BTR) * H));
  STAT_CUDA_CHECK(zero(w.DLC, static_cast<size_t>(BTR) * H));
  STAT_CUDA_CHECK(zero(w.DQ, static_cast<size_t>(BTR) * H));
  STAT_CUDA_CHECK(zero(w.DUACC, static_cast<size_t>(BT) * 4 * H));
  STAT_CUDA_CHECK(zero(w.DCACC, static_cast<size_t>(BT) * 4));
  for (int t = L - 1; t >= 0; --t) {
    float *DHQt = W + w.DHQ + static_cast<size_t>(t) * B * ldq;
    // fast mode keeps dC of every step (k_att_accum reads them after the loop)
    float *DCt = fast ? W + w.DCS + static_cast<size_t>(t) * B * H : W + w.DC;
    bw_mark(BP_CELL_BW, st);
    STAT_TRY(BW_LAUNCH(k_cell_backward, g1(static_cast<size_t>(B) * H), dim3(256), st, t, W + w.DHc, W + w.DCc,
                       W + w.DHR, W + w.GATES, W + w.Call, f->h0c0, mask, dp_gates, W + w.DHQ, ldq, 4 * H, W + w.DHm,
                       B, H));
    bw_mark(BP_DCTX_GEMM, st);
    // dctx_t = dpre . Wc^T (:439); decoder_Wc (H,4H) is K-major for this product as it stands
    STAT_TRY(mm(DHQt + 4 * H, ldq, Wc_a, 4 * H, W + w.DCTX, H, B, H, 4 * H, nullptr, st));
    bw_mark(BP_SELECTOR_BW, st);
    STAT_TRY(BW_LAUNCH(k_selector_bw, dim3(B), dim3(NT), st, t, W + w.DCTX, W + w.DCR, W + w.csum, W + w.beta,
                       W + w.DHQ, ldq, 8 * H, DCt, B, H, sel ? 1 : 0));
    bw_mark(BP_ATT_DOTS, st);
    STAT_TRY(BW_LAUNCH((R == 8 ? k_att_dots<8> : k_att_dots<0>), dim3(BT), dim3(NT), st, t, alpha_l, f->ctxl0, f->ctxg0, f->ctxm0, DCt,
                       W + w.CL, W + w.DA3, B, T, R, H));
    bw_mark(BP_ATT_SOFT, st);
    // deferred mode keeps the score gradients of every step (k_att_accum2 reads them after the loop)
    float *DS3t = fast ? W + w.DS3S + static_cast<size_t>(t) * 3 * BT : W + w.DS3;
    STAT_TRY(BW_LAUNCH(k_att_soft, g1(3 * B, 128), dim3(128), st, t, alpha_g, alpha_m, alpha_lt, W + w.DA3,
                       W + w.COV3, DS3t, W + w.DCACC, B, T));
    bw_mark(BP_ATT_MAIN, st);
    AttBw a;
    memset(&a, 0, sizeof(a));
    a.t_step = t; a.B = B; a.T = T; a.R = R; a.H = H; a.ldq = ldq; a.global_proj = gp ? 1 : 0;
    a.al = alpha_l; a.ag = alpha_g; a.am = alpha_m; a.alt = alpha_lt;
    a.HQ = W + w.HQ;
    a.pL = f->pctxl; a.Lc = f->ctxl0; a.Q = f->qctxl; a.pG = f->pctxg; a.pM = f->pctxm;
    a.Ul = p->decoder_Ul_att; a.Ug = p->decoder_Ug_att; a.Um = p->decoder_Um_att; a.Ult = p->decoder_Ult_att;
    a.DC = DCt; a.DS3 = DS3t; a.COVL = W + w.COVL;
    a.DPL = W + w.DPL; a.DLC = W + w.DLC; a.DQ = W + w.DQ;
    a.DPG = W + w.DPG; a.DPM = W + w.DPM; a.DG = W + w.DG; a.DM = W + w.DM;
    a.DUACC = W + w.DUACC; a.DCACC = W + w.DCACC; a.DSQP = W + w.DSQP;
    if (fast) { a.DAL = W + w.DAL; a.DPLT = W + w.DPLT; }
    STAT_TRY(BW_LAUNCH((R == 8 ? k_att_main<8> : k_att_main<0>), dim3(BT), dim3(NT), st, a));
    bw_mark(BP_REDUCE_T, st);
    STAT_TRY(BW_LAUNCH(k_reduce_t, g1(static_cast<size_t>(B) * 4 * H), dim3(256), st, t, W + w.DSQP, W + w.DHQ, ldq,
                       B, T, 4 * H));
    bw_mark(BP_DH_GEMM, st);
    // dh_{t-1} = [dsl | dsg | dsm | dslt | dpre | dsel] . Whcat^T + the masked bypass
    STAT_TRY(mm(DHQt, ldq, W + w.Whcat, ldq, W + w.TMPH, H, B, H, ldq, nullptr, st));
    STAT_TRY(BW_LAUNCH(k_add2, g1(static_cast<size_t>(B) * H), dim3(256), st, W + w.DHc, W + w.TMPH, W + w.DHm,
                       static_cast<size_t>(B) * H));
  }

  if (fast) {
    bw_mark(BP_ATT_MAIN, st);
    STAT_TRY(BW_LAUNCH(k_att_accum, dim3(BT), dim3(NT), st, L, B, T, R, H, ldq, alpha_l, alpha_lt, W + w.DAL, W + w.DPLT,
                       W + w.DCS, W + w.HQ, f->pctxl, p->decoder_Ul_att, W + w.DPL, W + w.DLC, W + w.DQ, W + w.DUACC));
    STAT_TRY(BW_LAUNCH((R == 8 ? k_att_accum2<8> : k_att_accum2<0>), dim3(BT), dim3(NT), st, L, B, T, R, H, ldq,
                       gp ? 1 : 0, alpha_l, alpha_g, alpha_m, W + w.DS3S, W + w.DCS, W + w.HQ, f->pctxg, f->pctxm,
                       f->qctxl, p->decoder_Ug_att, p->decoder_Um_att, W + w.DPG, W + w.DPM, W + w.DG, W + w.DM,
                       W + w.DUACC));
  }
  bw_mark(BP_WGRAD_STEPS, st);
  // ---- D: weight gradients over the stacked steps ----------------------------------------------------
  float *T1 = W + w.T1, *T2 = W + w.T2, *T3 = W + w.T3, *CS = W + w.CS, *SM = W + w.SMALL;
  STAT_TRY(transpose(W + w.Hprev, N, H, H, T1, N, st));
  STAT_TRY(transpose(W + w.DHQ, N, ldq, ldq, T2, N, st));                                            // (ldq, N)
  STAT_TRY(mm(T1, N, T2, N, W + w.DWH, ldq, H, NH, N, nullptr, st));
  STAT_CUDA_CHECK(cp2d(GRAD(decoder_Wdl_att), H, W + w.DWH, ldq, H, H));
  STAT_CUDA_CHECK(cp2d(GRAD(decoder_Wdg_att), H, W + w.DWH + H, ldq, H, H));
  STAT_CUDA_CHECK(cp2d(GRAD(decoder_Wdm_att), H, W + w.DWH + 2 * H, ldq, H, H));
  STAT_CUDA_CHECK(cp2d(GRAD(decoder_Wdlt_att), H, W + w.DWH + 3 * H, ldq, H, H));
  STAT_CUDA_CHECK(cp2d(GRAD(decoder_U), 4 * H, W + w.DWH + 4 * H, ldq, 4 * H, H));
  if (sel) STAT_CUDA_CHECK(cp2d(GRAD(decoder_W_sel), 1, W + w.DWH + 8 * H, ldq, 1, H));
  STAT_TRY(colsum(W + w.DHQ + 3 * H, N, H, ldq, GRAD(decoder_blt_att), CS, st));
  STAT_TRY(colsum(W + w.DHQ + 4 * H, N, 4 * H, ldq, GRAD(decoder_b), CS, st));
  if (sel) STAT_TRY(colsum(W + w.DHQ + 8 * H, N, 1, ldq, GRAD(decoder_b_sel), CS, st));
  const float *DPREt = T2 + static_cast<size_t>(4) * H * N;                                          // (4H, N)
  STAT_TRY(transpose(W + w.EMB, N, E, E, T1, N, st));
  STAT_TRY(mm(T1, N, DPREt, N, GRAD(decoder_W), 4 * H, E, 4 * H, N, nullptr, st));
  STAT_TRY(transpose(W + w.ctx, N, H, H, T1, N, st));
  STAT_TRY(mm(T1, N, DPREt, N, GRAD(decoder_Wc), 4 * H, H, 4 * H, N, nullptr, st));
  bw_mark(BP_EMBEDDING, st);
  // embedding: dEMB = dpre . W^T (+ the readout's prev2out term), scattered to the rows of Wemb
  STAT_TRY(mm(W + w.DHQ + 4 * H, ldq, Wd_a, 4 * H, W + w.DEMB, E, N, E, 4 * H, nullptr, st));
  if (p2o) STAT_TRY(BW_LAUNCH(k_add_inplace, g1(NEs), dim3(256), st, W + w.DEMB, W + w.DZP, NEs));
  STAT_CUDA_CHECK(cudaMemsetAsync(GRAD(Wemb), 0, F * V * E, st));
  if (fast) {
    if (L > 1) STAT_TRY(BW_LAUNCH(k_scatter_emb_owner, dim3((L - 1) * B), dim3(128), st, W + w.DEMB, x, GRAD(Wemb), L, B, E));
  } else {
    STAT_TRY(BW_LAUNCH(k_scatter_emb, g1(E, 128), dim3(128), st, W + w.DEMB, x, GRAD(Wemb), L, B, E));
  }
  bw_mark(BP_CTX_BLOCKS, st);
  // score vectors and biases
  STAT_TRY(colsum(W + w.DUACC, BT, 4 * H, 4 * H, SM, CS, st));
  STAT_CUDA_CHECK(cudaMemcpyAsync(GRAD(decoder_Ul_att), SM, F * H, cudaMemcpyDeviceToDevice, st));
  STAT_CUDA_CHECK(cudaMemcpyAsync(GRAD(decoder_Ug_att), SM + H, F * H, cudaMemcpyDeviceToDevice, st));
  STAT_CUDA_CHECK(cudaMemcpyAsync(GRAD(decoder_Um_att), SM + 2 * H, F * H, cudaMemcpyDeviceToDevice, st));
  STAT_CUDA_CHECK(cudaMemcpyAsync(GRAD(decoder_Ult_att), SM + 3 * H, F * H, cudaMemcpyDeviceToDevice, st));
  float *SC = SM + up(4 * H, 64);
  STAT_TRY(colsum(W + w.DCACC, BT, 4, 4, SC, CS, st));
  STAT_CUDA_CHECK(cudaMemcpyAsync(GRAD(decoder_cl_att), SC, F, cudaMemcpyDeviceToDevice, st));
  STAT_CUDA_CHECK(cudaMemcpyAsync(GRAD(decoder_cg_att), SC + 1, F, cudaMemcpyDeviceToDevice, st));
  STAT_CUDA_CHECK(cudaMemcpyAsync(GRAD(decoder_cm_att), SC + 2, F, cudaMemcpyDeviceToDevice, st));
  STAT_CUDA_CHECK(cudaMemcpyAsync(GRAD(decoder_clt_att), SC + 3, F, cudaMemcpyDeviceToDevice, st));
  // context projections (:322-326) and the feature layers under them (:661-667)
  // (value block Y (rows,H), its projection gradient DP, its direct gradient DY; raw features X (rows,D))
  auto ctx_path = [&](const float *Y, float *DP, float *DY, int rows, const float *Wc_att, float *gWc, float *gbc,
                      bool through) -> int {
    STAT_TRY(transpose(Y, rows, H, H, T1, rows, st));
    STAT_TRY(transpose(DP, rows, H, H, T2, rows, st));
    STAT_TRY(mm(T1, rows, T2, rows, gWc, H, H, H, rows, nullptr, st));
    STAT_TRY(colsum(DP, rows, H, H, gbc, CS, st));
    if (!through) return STAT_OK;
    STAT_TRY(mm(DP, H, Wc_att, H, T3, H, rows, H, H, nullptr, st));                                   // dY += DP . Wc^T
    const size_t cnt = static_cast<size_t>(rows) * H;
    STAT_TRY(BW_LAUNCH(k_add_inplace, g1(cnt), dim3(256), st, DY, T3, cnt));
    return STAT_OK;
  };
  auto feat_path = [&](const float *Y, float *DY, int rows, const float *X, int D, float *gWf, float *gbf) -> int {
    const size_t cnt = static_cast<size_t>(rows) * H;
    STAT_TRY(BW_LAUNCH(k_tanh_bw, g1(cnt), dim3(256), st, DY, Y, cnt));
    STAT_TRY(transpose(X, rows, D, D, T1, rows, st));
    STAT_TRY(transpose(DY, rows, H, H, T2, rows, st));
    STAT_TRY(mm(T1, rows, T2, rows, gWf, H, D, H, rows, nullptr, st));
    return colsum(DY, rows, H, H, gbf, CS, st);
  };
  STAT_TRY(ctx_path(f->ctxg0, W + w.DPG, W + w.DG, BT, Wcg_a, GRAD(decoder_Wcg_att),
                    GRAD(decoder_bg_att), gp));
  if (gp) STAT_TRY(feat_path(f->ctxg0, W + w.DG, BT, ctxg, Dg, GRAD(ff_global_W), GRAD(ff_global_b)));
  STAT_TRY(ctx_path(f->ctxm0, W + w.DPM, W + w.DM, BT, Wcm_a, GRAD(decoder_Wcm_att),
                    GRAD(decoder_bm_att), true));
  STAT_TRY(feat_path(f->ctxm0, W + w.DM, BT, ctxm, Dm, GRAD(ff_motion_W), GRAD(ff_motion_b)));
  STAT_TRY(ctx_path(f->ctxl0, W + w.DPL, W + w.DLC, BTR, Wcl_a, GRAD(decoder_Wcl_att),
                    GRAD(decoder_bl_att), true));
  {   // Q = ctxl0 . Wclt_att (the :416 product made step-invariant): dWclt = ctxl0^T . dQ, dctxl0 += dQ . Wclt^T
    STAT_TRY(transpose(f->ctxl0, BTR, H, H, T1, BTR, st));
    STAT_TRY(transpose(W + w.DQ, BTR, H, H, T2, BTR, st));
    STAT_TRY(mm(T1, BTR, T2, BTR, GRAD(decoder_Wclt_att), H, H, H, BTR, nullptr, st));
    STAT_TRY(mm(W + w.DQ, H, Wclt_a, H, T3, H, BTR, H, H, nullptr, st));
    const size_t cnt = static_cast<size_t>(BTR) * H;
    STAT_TRY(BW_LAUNCH(k_add_inplace, g1(cnt), dim3(256), st, W + w.DLC, T3, cnt));
  }
  STAT_TRY(feat_path(f->ctxl0, W + w.DLC, BTR, ctxl, Dr, GRAD(ff_local_W), GRAD(ff_local_b)));
  bw_mark(BP_INIT_STATE, st);
  // init state (:618,:649,:657-660)
  STAT_TRY(BW_LAUNCH(k_meanpool, g1(static_cast<size_t>(B) * Dg), dim3(256), st, ctxg, mask_ctxg, W + w.GBAR, B, T, Dg));
  STAT_TRY(BW_LAUNCH(k_init_bw, g1(static_cast<size_t>(B) * H), dim3(256), st, W + w.DHc, W + w.DCc, f->h0c0,
                     W + w.DP0, B, H));
  STAT_TRY(transpose(W + w.GBAR, B, Dg, Dg, T1, B, st));
  STAT_TRY(transpose(W + w.DP0, B, 2 * H, 2 * H, T2, B, st));
  STAT_TRY(mm(T1, B, T2, B, GRAD(ff_state_W), H, Dg, H, B, nullptr, st));
  STAT_TRY(mm(T1, B, T2 + static_cast<size_t>(H) * B, B, GRAD(ff_memory_W), H, Dg, H, B, nullptr, st));
  STAT_TRY(colsum(W + w.DP0, B, H, 2 * H, GRAD(ff_state_b), CS, st));
  STAT_TRY(colsum(W + w.DP0 + H, B, H, 2 * H, GRAD(ff_memory_b), CS, st));
  bw_mark(BP_DECAY, st);
  // weight decay (:1130-1136): every parameter, biases included
  if (decay_c > 0.f) {
    const float c2 = 2.0f * decay_c;
    // Tensors that lie back to back in both the gradient and the parameter buffer (the optimizer's flat buffers:
    // all of them, in init_params order) are decayed by one launch: 43 launches -> 1 for the Trainer.
    struct Run { float *g; const float *p; size_t n; };
    Run runs[64];
    int nruns = 0;
    auto dec = [&](const float *g, const float *pp, size_t n) -> int {
      if (!g || !pp || n == 0) return STAT_OK;
      STAT_REQUIRE(nruns < 64, STAT_EINVAL, "grad_shared: too many decayed tensors");
      runs[nruns++] = Run{const_cast<float *>(g), pp, n};
      return STAT_OK;
    };
    const size_t h = H, e = E, v = V;
#define DEC(name, n) STAT_TRY(dec(grads->name, p->name, (n)))
    DEC(Wemb, v * e);
    DEC(ff_state_W, Dg * h); DEC(ff_state_b, h); DEC(ff_memory_W, Dg * h); DEC(ff_memory_b, h);
    if (gp) { DEC(ff_global_W, Dg * h); DEC(ff_global_b, h); }
    DEC(ff_local_W, Dr * h); DEC(ff_local_b, h); DEC(ff_motion_W, Dm * h); DEC(ff_motion_b, h);
    DEC(decoder_W, e * 4 * h); DEC(decoder_U, h * 4 * h); DEC(decoder_b, 4 * h); DEC(decoder_Wc, h * 4 * h);
    DEC(decoder_Wcg_att, h * h); DEC(decoder_Wcm_att, h * h); DEC(decoder_Wclt_att, h * h);
    DEC(decoder_Wdg_att, h * h); DEC(decoder_Wdm_att, h * h); DEC(decoder_Wdlt_att, h * h);
    DEC(decoder_bg_att, h); DEC(decoder_bm_att, h); DEC(decoder_blt_att, h);
    DEC(decoder_Wcl_att, h * h); DEC(decoder_Wdl_att, h * h); DEC(decoder_bl_att, h);
    DEC(decoder_Ug_att, h); DEC(decoder_cg_att, 1); DEC(decoder_Um_att, h); DEC(decoder_cm_att, 1);
    DEC(decoder_Ult_att, h); DEC(decoder_clt_att, 1); DEC(decoder_Ul_att, h); DEC(decoder_cl_att, 1);
    if (sel) { DEC(decoder_W_sel, h); DEC(decoder_b_sel, 1); }
    DEC(ff_logit_lstm_W, h * e); DEC(ff_logit_lstm_b, e);
    if (c2o) { DEC(ff_logit_ctxglm_W, h * e); DEC(ff_logit_ctxglm_b, e); }
    DEC(ff_logit_W, e * v); DEC(ff_logit_b, v);
#undef DEC
    for (int i = 1; i < nruns; ++i)            // by gradient address (insertion sort: <= 43 entries)
      for (int j = i; j > 0 && runs[j].g < runs[j - 1].g; --j) { const Run t = runs[j]; runs[j] = runs[j - 1]; runs[j - 1] = t; }
    for (int i = 0; i < nruns;) {
      Run r = runs[i++];
      while (i < nruns && runs[i].g == r.g + r.n && runs[i].p == r.p + r.n) r.n += runs[i++].n;
      STAT_TRY(BW_LAUNCH(k_decay, g1(r.n), dim3(256), st, r.g, r.p, r.n, c2));
    }
  }
#undef GRAD
  bw_mark(BP_COUNT, st);
  return STAT_OK;
}

}  // extern "C"
