// Launch-argument structs and launchers of the non-GEMM kernels.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

namespace stat {

// ---- att_step.cu -------------------------------------------------------------
struct AttArgs {
  const float *pctxl, *ctxl0, *qctxl;            // (clips,T,R,H)
  const float *pctxg, *ctxg0, *pctxm, *ctxm0;    // (clips,T,H)
  const float *hp;                               // (parts)(rows,ldhp) hidden-state projections: k-slice
  int ldhp;                                      // planes hp_plane floats apart, summed in order by the reader
  int hp_parts;
  size_t hp_plane;
  int off_sl, off_sg, off_sm, off_slt, off_sel;  // column offsets inside a hp row
  const float *Ul, *Ug, *Um, *Ult;               // (H) score vectors
  const float *cl, *cg, *cm, *clt;               // (1) score biases
  const int32_t *row_clip;                       // (rows) or null = identity
  int rows, T, R, H;
  int S, Tc;                                     // segments per row, frames per segment
  int selector;
  float *ctx;                                    // (rows,ldctx) out: fused, gated context
  int ldctx;                                     // row pitch of ctx (0 = H)
  float *rec_vec;                                // (rows,S,3,H) partial weighted sums
  float *rec_ms;                                 // (rows,S,3,2) partial (max,sum)
  unsigned int *counters;                        // (rows) tickets, zero between launches
  float *att_scores;                             // (3,rows,T) raw temporal scores or null
  float *alpha_l;                                // (rows,T,R) spatial weights or null
  long long *trace;                              // debug: clock stamps per CTA and group (att_group), or null
  int reverse;                                   // att_group: walk the frames of a slice backwards (odd decode steps)
  float *ctx_t;                                  // [row / 64][H][64] transposed copy of ctx (cell_kernel operand) or null
  int rows_per_clip;                             // beam search: row = clip * rows_per_clip + slot, the slots of a clip share
                                                 // its context blocks (att_clip_kernel); 0 = no such layout (row_clip decides)
};
int att_step_launch(const AttArgs &a, cudaStream_t stream);
// att_group.cu: four independent four-warp groups per CTA, each streaming whole frames (H % 4 == 0,
// one frame fits in shared memory).  a.S = capacity (parts per row) of rec_vec / rec_ms.
bool att_group_plan(int rows, int T, int R, int H, int *nctas, int *groups, int *max_parts);
int att_group_launch(const AttArgs &a, cudaStream_t stream);
void att_group_set_trace(long long *p);
void att_group_set_share(int on);   // att_clip_kernel for rows_per_clip >= 2: 1 on, 0 off, -1 default (STAT_ATT_SHARE, else on)

// ---- recurrent.cu --------------------------------------------------------------
// LSTM gates + state update (model_attention.py:437-457) and the emb / bias part of
// the readout pre-activation (:689-693).
struct GateArgs {
  int rows, H, E, V;
  const float *pre_c;      // (rows,ldpc): [0,4H) ctx.Wc ; [zc_off, zc_off+E) ctx.Wctx (zc_off < 0: absent)
  int ldpc, zc_off;
  int pc_parts;            // k-slice planes of pre_c, pc_plane floats apart
  size_t pc_plane;
  const float *hp;         // (parts)(rows,ldhp): h_.U at off_u
  int ldhp, off_u;
  int hp_parts;
  size_t hp_plane;
  const float *EW;         // (V+1,4H) token -> emb.W + b, gate-interleaved columns 4*unit + gate
  const float *Wemb;       // (V,E)
  const int64_t *tok_prev; // (rows) or null = no previous word
  const float *mask;       // (rows) or null = 1
  const float *dp_gates;   // (rows,3H) or null = 0.5
  const float *dp_h;       // (rows,H) or null
  const float *h_in, *c_in;
  float *h_out, *c_out;    // may alias the inputs
  float *hd_out;           // (rows,H) h * dp_h when dp_h != null
  const float *bz;         // (E) readout bias: ff_logit_lstm_b (+ ff_logit_ctxglm_b)
  float *zadd;             // (rows,E) out: bz + ctx.Wctx (ctx2out) + Wemb[tok] (prev2out)
  int prev2out;
  float *h_all;            // (rows,H) copy of h_out or null
};
int gates_launch(const GateArgs &a, cudaStream_t stream);

// readout activation: z = post * tanh(alpha * sum_p zpre[p] + zadd), post = dp_z or 0.5 (:684-696)
struct ZactArgs {
  int rows, E;
  const float *zpre;       // (parts)(rows,ldz) k-slice planes of h.ff_logit_lstm_W
  int ldz, parts;
  size_t plane;
  float alpha;             // 0.5 at eval (h*0.5), 1 when zpre was computed from h*mask
  const float *zadd;       // (rows,E)
  const float *dp_z;       // (rows,E) or null = 0.5
  float *z;                // (rows,E)
};
int zact_launch(const ZactArgs &a, cudaStream_t stream);

// vocabulary reduction of one step: log-softmax statistics, argmax, bookkeeping
struct PickArgs {
  int rows, V, ldl;
  const float *logits;       // (rows,ldl)
  // greedy bookkeeping (all null in the other modes)
  int64_t *tokens;           // (rows,maxlen)
  int maxlen, t;
  int32_t *lengths;          // (rows)
  float *scores;             // (rows)
  int32_t *alive;            // (rows)
  int64_t *tok_prev;         // (rows)
  // teacher forcing
  const int64_t *x_t;        // (rows) target tokens of this step or null
  const float *mask_t;       // (rows)
  float *logprob;            // (rows) accumulated
  float *logprob_comp;       // (rows) Kahan compensation of the accumulation (zero at the first step) or null
  // f_next
  float *probs;              // (rows,V) or null
  // beam search: the beam_k most probable words of every live row and their cumulative costs
  int beam_k;                // 0 = off
  const int32_t *row_alive;  // (rows)
  const float *row_score;    // (rows) cumulative -log p of the hypothesis in this row
  float *cand_cost;          // (rows,BEAM_KMAX): row_score - log p[word], ascending; +inf when absent
  int32_t *cand_word;        // (rows,BEAM_KMAX)
};
constexpr int BEAM_KMAX = 16;      // beam width limit
constexpr int BEAM_LMAX = 64;      // hypothesis length limit (maxlen)
int pick_launch(const PickArgs &a, cudaStream_t stream);

// Beam bookkeeping of one step for every clip (model_attention.py:905-973): the k - dead best of the
// live hypotheses' candidates, retirement on token 0, survivors compacted to the front of the
// clip's k row slots.  One warp per clip.
struct BeamArgs {
  int B, k, V, maxlen, t;
  const float *cand_cost;    // (B*k,BEAM_KMAX)
  const int32_t *cand_word;  // (B*k,BEAM_KMAX)
  int32_t *alive;            // (B*k) live flag of a row slot
  float *score;              // (B*k) cumulative cost of the live hypothesis
  int32_t *hist;             // (B*k,BEAM_LMAX) its words so far
  int32_t *hist_len;         // (B*k)
  int32_t *src_row;          // (B*k) out: row whose LSTM state the slot continues from
  int64_t *tok_prev;         // (B*k) out: last word of the slot (input of the next step)
  int32_t *dead_k;           // (B)
  int32_t *done;             // (B)
  int64_t *out_tokens;       // (B,k,maxlen), -1 padded
  int32_t *out_lengths;      // (B,k)
  float *out_scores;         // (B,k)
  int32_t *out_count;        // (B) finished hypotheses so far
  // state gather of the next step, done by the same launch (dst_h null: none): slot j of a clip continues row
  // src_row[j]; src_q / dst_q (ldq floats per row, a multiple of 4; null: none) are the h-products of the cell step
  const float *src_h, *src_c;
  float *dst_h, *dst_c;
  int H;
  const float *src_q;
  float *dst_q;
  int ldq;
};
int beam_select_launch(const BeamArgs &a, cudaStream_t stream);
int beam_init_launch(const BeamArgs &a, const float *h0c0, float *h, float *c, int H, int32_t *row_clip,
                     cudaStream_t stream);

// ---- step_fused.cu: the fused products of a decode step (fused_tile.cuh) ----------------------------------
enum { FE_STORE = 0, FE_GATES = 1, FE_ZC = 2, FE_Z = 3, FE_PICK = 4 };
// epilogue parameters (the union of all kinds; a launch reads what its kinds need)
struct FusedEpi {
  int rows;
  // FE_STORE: out[r*ldo + j] = acc + bias[j]
  float *out; int ldo; const float *bias;
  // FE_GATES (model_attention.py:437-457) on gate-interleaved features j = 4*unit + gate
  int H, V;
  const float *EWi;                  // (V+1, 4H) gate-interleaved token table (row V = bias alone)
  const float *hu; int ld_hu;        // (rows, ld_hu) h_{t-1}.U, gate-interleaved, or null
  const int64_t *tok_prev;           // (rows) or null = no previous word
  const float *mask;                 // (rows) or null
  const float *dp_gates;             // (rows, 3H) or null = 0.5
  const float *c_in; float *c_out;   // (rows, H)
  const float *h_in; int ld_hin;     // previous hidden state (masked bypass)
  float *h_out; int ld_hout;         // new hidden state -> activation buffer of the next products
  float *h_copy;                     // (rows, H) dense copy or null
  float *h_all;                      // (rows, H) or null
  float *hd_out; const float *dp_h;  // h * dp_h (explicit dropout) or null
  // FE_ZC (:689-693): zadd = acc + bz (+ Wemb[tok])
  float *zadd; int E; const float *bz; const float *Wemb; int prev2out;
  // FE_Z (:684-696): z = post * tanh(z_alpha * acc + zadd), post = dp_z or 0.5
  float *z; float z_alpha; const float *dp_z;
  // FE_PICK (:704-709): per (row, half tile) partial (max, sum exp, arg-max) of acc + bv; target logit
  const float *bv; float *part; int npart; int part0;
  const int64_t *x_t; float *tgt;
};
struct FusedSegment {
  int kind;      // FE_*
  int wrow0;     // first weight row of the segment in W
  int nfeat;     // features
  int K;         // reduction length (floats)
  int xsel;      // activation operand 0 / 1
};
struct FusedPhase {
  int swap;                          // 1: weights on the 128-lane axis, 32 decode rows per tile; 0: rows on lanes
  const float *W; int wrows, wK, ldw;          // packed K-major weights
  const float *X[2]; int xK[2], ldx[2];        // K-major activations (rows, xK)
  int rows;
  int nseg;
  FusedSegment seg[2];
  FusedEpi e;
};
bool fused_supported(int H, int E);
int fused_phase_launch(const FusedPhase &p, cudaStream_t stream);
// combine of the FE_PICK partials of one step + greedy / teacher-forcing bookkeeping (PickArgs as pick_launch)
int pick_combine_launch(const PickArgs &a, const float *part, int npart, const float *tgt, cudaStream_t stream);

// ---- cell_step.cu: everything of a decode step between two attentions in one launch --------------------------
struct CellPlan {
  int grid;                  // CTAs = SMs (all co-resident: the kernel has a grid-wide barrier)
  int upc, epc, cpc;         // hidden units / readout columns (phase 1), output columns (phase 2) per CTA
  int NQ1, NQ;               // 8H+4 (queries | sel | pad | h.U), 8H+4+E (+ readout)
  size_t w1_floats, w2_floats;   // sizes of the packed weight slabs
};
bool cell_plan(int H, int E, CellPlan *out);
// slabs from the K-major gate-interleaved packs: WcI (4H+E, H), WqT (8H+4+E, H)
int cell_pack_launch(const float *WcI, const float *WqT, float *W1, float *W2, int H, int E, int ctx2out,
                     cudaStream_t stream);
struct CellLaunch {
  int rows, H, E, V;
  int do1, do2;              // phase 1: ctx.Wc -> gates -> h, c, zadd;  phase 2: products of the (new) hidden state
  int want_q, want_z;        // phase 2 outputs: queries / sel / h.U rows (hq), readout activation (z)
  int prev2out;
  const float *W1, *W2;      // packed slabs (cell_pack_launch)
  const float *ctxT;         // [chunk][H][64] fused context, transposed (written by the attention kernel)
  float *hT;                 // [chunk][H][64] hidden state, transposed: written by phase 1, read by phase 2
  float *hq;                 // (rows, ldq) [queries 4H | sel | 3 pad | h.U 4H gate-interleaved]: phase 1 reads the
  int ldq;                   //   h.U block of the previous launch, phase 2 writes the row (to hq_out when not null)
  float *hq_out;
  const float *EW, *Wemb, *bz, *bq;
  const int64_t *tok_prev;   // (rows) or null = no previous word
  const float *mask;         // (rows) or null
  const float *dp_gates;     // (rows, 3H) or null = 0.5
  const float *dp_z;         // (rows, E) or null = 0.5
  const float *h_in, *c_in;
  float *h_out, *c_out;      // may alias the inputs
  float *h_all;              // (rows, H) or null
  float *zadd, *z;           // (rows, E)
  unsigned int *bar;         // 2 words, zero before the first launch
};
int cell_launch(const CellLaunch &c, cudaStream_t stream);

// ---- optim.cu: gradient clipping and the reference's optimizers over one flat parameter buffer ----
size_t clip_scratch_bytes();
int grad_clip_launch(float *grads, size_t n, float clip_c, void *scratch, cudaStream_t stream);
int adam_launch(float *p, const float *g, float *m, float *v, size_t n, int step, cudaStream_t stream);
int coverage_launch(const float *alphas, int L, int rows, int n, void *scratch, float *out, cudaStream_t stream);
int adadelta_launch(float *p, const float *g, float *rg2, float *ru2, size_t n, int phase, cudaStream_t stream);

int meanpool_launch(const float *ctxg, const float *mask, float *gbar, int B, int T, int D,
                    cudaStream_t stream);
// dst (rows_out, ld_dst) [r0 + n][k] = src (K, N) [k][n]   (weights -> K-major)
int transpose_launch(const float *src, int K, int N, float *dst, int ld_dst, int r0, cudaStream_t stream);
// gate-interleaving variants (rows 4*unit + gate): weights (K, 4H) -> K-major rows at column offset c0; vectors (4H)
int transpose_il_launch(const float *src, int K, int H, float *dst, int ld_dst, int c0, cudaStream_t stream);
int interleave4_launch(const float *src, float *dst, int H, cudaStream_t stream);
// out_q = row-wise softmax of plane q of scores (3, nrows, n) for the non-null outputs, one launch
int softmax_rows3_launch(const float *scores, float *out0, float *out1, float *out2, int nrows, int n,
                         cudaStream_t stream);
int scale_launch(float *x, const float *f, size_t n, cudaStream_t stream);
int init_rows_launch(int rows, int64_t *tok_prev, int32_t *alive, int32_t *lengths, float *scores,
                     int64_t *tokens, int maxlen, cudaStream_t stream);
int add_vec_launch(float *dst, const float *a, const float *b, int n, cudaStream_t stream);

}  // namespace stat
