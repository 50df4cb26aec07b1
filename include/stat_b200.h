/*
 * stat_b200.h -- C ABI of the B200-native spatial-temporal-attention caption
 * decoder (libstat_b200.so).
 *
 * The reference (tuyunbin/Video-Description-with-Spatial-Temporal-Attention) has
 * no native code and no FFI: its boundary for this path is the set of Theano
 * callables built in model_attention.py.  Each entry point below names the
 * reference callable / graph fragment it replaces (file:line into the
 * reference).  The Python host mirror that keeps the reference's own names
 * (init_params / build_model / f_log_probs / build_sampler / f_init / f_next /
 * gen_sample) sits on top of this ABI, see INTEGRATION.md.
 *
 * Conventions
 *   - plain C, no C++ or torch types; every pointer is a DEVICE pointer owned
 *     by the caller (inputs, outputs and workspace); the library never
 *     allocates, frees or retains device memory;
 *   - every function enqueues work on the caller's stream and returns without
 *     synchronising; all entry points may be stream-captured into a CUDA graph;
 *   - return 0 on success, a negative STAT_E* code otherwise;
 *     stat_last_error() gives a thread-local message;
 *   - all tensors fp32 row-major, tokens int64 (model_attention.py:587,798);
 *   - threads and devices: entry points may be called from several host
 *     threads and on several devices of one process as long as concurrent calls
 *     use different streams and different workspaces; the library keeps its
 *     internal side streams / events per host thread and device.  Process-wide
 *     state: the launch counter, the profiling and trace hooks, the GEMM
 *     implementation switch, the persisting-L2 size (per device);
 *   - there is no CPU fallback: without a CUDA device every compute entry
 *     point fails with STAT_ECUDA.
 */
#ifndef STAT_B200_H
#define STAT_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define STAT_ABI_VERSION 1

/* error codes */
#define STAT_OK       0
#define STAT_EINVAL  -1   /* bad argument / unsupported shape            */
#define STAT_ECUDA   -2   /* CUDA runtime or driver error                */
#define STAT_EALIGN  -3   /* pointer or leading dimension not 16B aligned */

/* option flags (config.py:17-49; global_proj is decision D1, the layer the
 * reference left commented out at model_attention.py:553-554,661-662) */
#define STAT_SELECTOR     1
#define STAT_PREV2OUT     2
#define STAT_CTX2OUT      4
#define STAT_GLOBAL_PROJ  8

typedef struct StatDims {
  int32_t B;      /* clips in the batch                                        */
  int32_t T;      /* frames per clip                                           */
  int32_t R;      /* regions per frame (<= 16)                                 */
  int32_t Dg;     /* ctxg_dim  (multiple of 4)                                 */
  int32_t Dm;     /* ctxm_dim  (multiple of 4)                                 */
  int32_t Dr;     /* ctxl_dim  (multiple of 4)                                 */
  int32_t H;      /* dim       (multiple of 4)                                 */
  int32_t E;      /* dim_word  (multiple of 4)                                 */
  int32_t V;      /* n_words                                                   */
  int32_t flags;  /* STAT_* option flags                                       */
} StatDims;

/* The reference's parameter dict (init_params, model_attention.py:518-581;
 * SURVEY App. B order).  Optional tensors may be NULL when their flag is off. */
typedef struct StatParams {
  const float *Wemb;                            /* (V,E)            :522     */
  const float *ff_state_W, *ff_state_b;         /* (Dg,H),(H)       :549     */
  const float *ff_memory_W, *ff_memory_b;       /* (Dg,H),(H)       :551     */
  const float *ff_global_W, *ff_global_b;       /* (Dg,H),(H)  D1   :553     */
  const float *ff_local_W, *ff_local_b;         /* (Dr,H),(H)       :556     */
  const float *ff_motion_W, *ff_motion_b;       /* (Dm,H),(H)       :558     */
  const float *decoder_W, *decoder_U, *decoder_b, *decoder_Wc;  /* :189-207  */
  const float *decoder_Wcg_att, *decoder_Wcm_att, *decoder_Wclt_att; /* :210 */
  const float *decoder_Wdg_att, *decoder_Wdm_att, *decoder_Wdlt_att; /* :222 */
  const float *decoder_bg_att, *decoder_bm_att, *decoder_blt_att;    /* :232 */
  const float *decoder_Wcl_att, *decoder_Wdl_att, *decoder_bl_att;   /* :243 */
  const float *decoder_Ug_att, *decoder_cg_att;                      /* :255 */
  const float *decoder_Um_att, *decoder_cm_att;
  const float *decoder_Ult_att, *decoder_clt_att;
  const float *decoder_Ul_att, *decoder_cl_att;
  const float *decoder_W_sel, *decoder_b_sel;   /* (H,1),()  selector :276   */
  const float *ff_logit_lstm_W, *ff_logit_lstm_b;      /* (H,E),(E)   :566   */
  const float *ff_logit_ctxglm_W, *ff_logit_ctxglm_b;  /* (H,E),(E)   :569   */
  const float *ff_logit_W, *ff_logit_b;                /* (E,V),(V)   :578   */
} StatParams;

int         stat_version(void);
const char *stat_last_error(void);

/* 0 = tcgen05 3xTF32 tensor-core GEMM with the 128-lane operand in tensor memory
 * (default), 2 = the same with both operands in shared memory, 1 = plain fp32 SIMT
 * GEMM (device-side cross-checks used by the GPU tests; never a CPU path). */
int stat_set_gemm_impl(int impl);

/* Implementation of the decode step behind stat_forward_teacher / stat_decode_greedy / stat_decode_beam (same
 * results to fp32 summation order; all are device paths):
 *   0 = separate kernels (the default): attention -> ctx.[Wc|Wctx] -> gates -> h.[Wd*|U|W_sel|Wl] as k-split
 *       tensor-core products + elementwise kernels, readout activation -> logits -> vocabulary reduction beside the
 *       next attention.  Products over 129..192 rows (beam search) go out as a 128-row and a narrow skinny launch.
 *   2 = cell step (cell_step.cu; where the shape allows it: H % 32 == 0, 160 <= H <= 512, no explicit dropout mask
 *       on h): attention -> ONE kernel for ctx.Wc + gates (S10-S13) | grid barrier | every product of the new hidden
 *       state + readout activation.  Two dependent launches per step; measured slower than 0 for greedy decoding,
 *       the default of beam searches over more than 192 rows.
 *   1 = fused tile kernels (step_fused.cu): gates / readout activation / vocabulary reduction in tcgen05 tile
 *       epilogues; measured slower end to end in round 2, kept selectable.
 *  -1 = back to the default (environment STAT_STEP=0|1|2, else as above). */
int stat_set_step_impl(int impl);

/* Beam search attention (stat_decode_beam): 1 = the k row slots of a clip share ONE pass over the clip's context
 * blocks (att_clip_kernel: a frame is brought on chip once for all k hypotheses, as the reference broadcasts one
 * clip's context to its k live rows, model_attention.py:786-788), 0 = one cluster per row (k passes over the same
 * blocks), -1 = back to the default (environment STAT_ATT_SHARE=0|1, else 1).  Same results to fp32 summation order. */
int stat_set_beam_share(int on);

/* Fault hunting: arms the spin-wait give-up sites of the attention and GEMM kernels to record where they trap
 * (4 ints: site = file tag + source line, blockIdx.x, threadIdx.x, blockIdx.y << 16 | blockIdx.z) in host-mapped
 * memory, which survives the device fault.  Returns the host pointer (NULL on failure). */
int *stat_debug_trap_log(void);

/* ---- L2 residency of the context blocks (no reference counterpart) -----------
 * The attention kernel re-reads the projected context blocks of the batch on
 * every decode step and copies them with the L2 evict_last priority (everything
 * streamed between two steps is evict_first).  This call sets the persisting-L2
 * carve-out of the device (cudaLimitPersistingL2CacheSize, a per-device setting
 * of the calling process): bytes < 0 = the device maximum (79 MB of the 126 MB on
 * B200), 0 = none.  Measured in round 2: bulk copies with an evict_last hint do
 * not use the set-aside, so a carve-out only shrinks the L2 they compete for --
 * the host mirror sets none.  Returns the size now in effect, or a negative
 * STAT_E* code. */
long long stat_set_l2_persist(long long bytes);

/* ---- parameter preparation (once per parameter set) ----------------------
 * Packs the reference tensors into the K-major concatenations the kernels
 * read and builds the token->gate-input table EW = Wemb.decoder_W + decoder_b
 * (model_attention.py:334-335 hoisted out of the step: row x of EW is the
 * `state_below` row of token x; row V is the bias alone = "no previous word",
 * :803-804).  `prepared` must hold stat_prepared_bytes() bytes. */
size_t stat_prepared_bytes(const StatDims *d);
int    stat_prepare_params(const StatDims *d, const StatParams *p, void *prepared,
                           void *stream);

/* ---- workspace -----------------------------------------------------------
 * `rows` = decode rows (= B for greedy / teacher forcing, = B*k for beams). */
size_t stat_workspace_bytes(const StatDims *d, int rows);
/* byte offset/size of a named workspace region (tests read intermediates):
 * "ctxg0","pctxg","ctxm0","pctxm","ctxl0","pctxl","qctxl","h0","c0","h","c",
 * "hp","ctx","logits","att_scores","alpha_l". Returns STAT_EINVAL if unknown. */
int    stat_workspace_region(const StatDims *d, int rows, const char *name,
                             size_t *offset, size_t *bytes);

/* ---- K0: per-batch prologue ----------------------------------------------
 * Replaces build_model's prologue (model_attention.py:618,649 mean-pool,
 * :657-660 init state, :664-667 ff_local/ff_motion) and lstm_cond_layer's
 * context projections (:322-326), plus Q = ctxl0.Wclt_att (the :416 GEMM made
 * step-invariant by linearity).  Results stay in `ws`. */
int stat_precompute(const StatDims *d, const void *prepared,
                    const float *ctxg, const float *mask_ctxg,
                    const float *ctxl, const float *ctxm,
                    void *ws, void *stream);

/* ---- f_init (model_attention.py:791-795; graph :739-777) -------------------
 * Initial LSTM state of the B clips: mean over all T frames of ctxg divided by
 * the number of non-zero frames (:618,:649), then tanh(ff_state) / tanh(ff_memory)
 * (:657-660).  out_h0 / out_c0 (B,H).  Uses the gbar / h0 regions of `ws`. */
int stat_init_state(const StatDims *d, const void *prepared, const float *ctxg,
                    const float *mask_ctxg, void *ws, float *out_h0, float *out_c0,
                    void *stream);

/* ---- f_log_probs (model_attention.py:1126; graph :583-717) ----------------
 * Teacher-forced forward over L steps for the B clips already precomputed in
 * `ws`.  x (L,B) int64, mask (L,B).  dp_gates (L,B,3H) / dp_h (L,B,H) / dp_z
 * (L,B,E) are the dropout factors; NULL = eval mode constants 0.5
 * (use_noise=0, common.py:94-99, model_attention.py:469-477).
 * out_logprob (B) = sum_t mask*log(p[x]+1e-8).  Optional outputs (may be NULL):
 * out_alpha_l (L,B,T,R), out_alpha_g/m/lt (L,B,T), out_h (L,B,H). */
int stat_forward_teacher(const StatDims *d, const void *prepared, void *ws,
                         int L, const int64_t *x, const float *mask,
                         const float *dp_gates, const float *dp_h, const float *dp_z,
                         float *out_logprob,
                         float *out_alpha_l, float *out_alpha_g,
                         float *out_alpha_m, float *out_alpha_lt, float *out_h,
                         void *stream);

/* ---- greedy decode = gen_sample(k=1) for B clips at once -------------------
 * (model_attention.py:852-994 with k=1, stochastic=False).  out_tokens
 * (B,maxlen) int64, -1 after the eos; out_lengths (B) int32 incl. the eos;
 * out_scores (B) = cumulative -log p (un-normalised, :921). */
int stat_decode_greedy(const StatDims *d, const void *prepared, void *ws,
                       int maxlen, int64_t *out_tokens, int32_t *out_lengths,
                       float *out_scores, void *stream);

/* ---- beam search = gen_sample(k > 1) for B clips at once ---------------------
 * (model_attention.py:852-994, stochastic=False).  `ws` must have been sized for
 * rows = B*k (stat_workspace_bytes) and hold the stat_precompute results of the
 * B clips.  Per clip and step: the k - dead cheapest continuations
 * (cumulative -log p, fp32) of the live hypotheses, hypotheses ending in token 0
 * retire, the search stops when none is live, k have retired or maxlen steps are
 * done; the survivors follow the retired ones, exactly the reference's order.
 * out_tokens (B,k,maxlen) int64, -1 padded; out_lengths (B,k) int32 (incl. the
 * eos when there is one); out_scores (B,k) cumulative -log p; out_count (B)
 * hypotheses returned per clip (<= k).  1 <= k <= 16, 1 <= maxlen <= 64. */
int stat_decode_beam(const StatDims *d, const void *prepared, void *ws, int k,
                     int maxlen, int64_t *out_tokens, int32_t *out_lengths,
                     float *out_scores, int32_t *out_count, void *stream);

/* ---- f_next (model_attention.py:845-848): one step for `rows` hypotheses ---
 * row_clip (rows) int32 maps a hypothesis to its clip in `ws` (NULL = row i
 * uses clip i).  x (rows) int64, -1 = no previous word.  h_in/c_in (rows,H).
 * out_probs (rows,V), out_h/out_c (rows,H). */
int stat_step(const StatDims *d, const void *prepared, void *ws, int rows,
              const int32_t *row_clip, const int64_t *x,
              const float *h_in, const float *c_in,
              float *out_probs, float *out_h, float *out_c, void *stream);

/* ---- parameter update of the training step (common.py:178-230, ---------------
 * model_attention.py:1194-1203).  All parameters / gradients / optimizer states
 * live in flat fp32 buffers of n elements in init_params order (the layout the
 * data-parallel all-reduce uses); buffers 16-byte aligned.
 * stat_grad_clip: g2 = sum g^2; if g2 > clip_c^2 every gradient is scaled by
 *   clip_c / sqrt(g2) (clip_c <= 0: no clipping).  `scratch` holds
 *   stat_clip_scratch_bytes() bytes; out_g2 (optional, 2 floats) receives g2 and
 *   the factor applied.  Deterministic (fixed-shape two-stage reduction).
 * stat_adam_step: the reference's adam, constants as in common.py:204-207
 *   (lr0 = 2e-4 -- the lr argument of f_update is ignored there -- b1 = 0.1 and
 *   b2 = 0.001 weight the NEW gradient, e = 1e-8), `step` = 1-based update count.
 * stat_adadelta_step: phase 0 = the running-gradient update f_grad_shared does
 *   (rg2 <- 0.95 rg2 + 0.05 g^2), phase 1 = f_update's step.
 * stat_alpha_coverage: the attention-coverage regulariser of the training cost
 *   (model_attention.py:1138-1147) without its alpha_c factor: for alphas
 *   (L, rows, n) out[0] = ((1 - alphas.sum(0))**2).sum(0).mean(), i.e. the sum over
 *   rows and mean over the n attended positions; `scratch` as for stat_grad_clip. */
size_t stat_clip_scratch_bytes(void);
int stat_alpha_coverage(const float *alphas, int L, int rows, int n, void *scratch, float *out, void *stream);
int stat_grad_clip(float *grads, size_t n, float clip_c, void *scratch, float *out_g2, void *stream);
int stat_adam_step(float *params, const float *grads, float *m, float *v, size_t n, int step, void *stream);
int stat_adadelta_step(float *params, const float *grads, float *rg2, float *ru2, size_t n, int phase,
                       void *stream);

/* ---- f_grad_shared: the gradients of the training cost -------------
 * Replaces `grads = tensor.grad(cost, wrt=itemlist(tparams))` (model_attention.py:1193) for
 *   cost = inv_batch * sum_b( -sum_t mask*log(p[x]+1e-8) )                    (:1129)
 *        + decay_c * sum_params sum(p^2)                                      (:1130-1136)
 *        + alpha_c * sum over the four attentions of
 *              ((1 - alphas.sum(0))**2).sum(0).mean()                         (:1138-1147)
 * (clipping, :1194-1203, is stat_grad_clip).  Call order of one training step:
 * stat_prepare_params, stat_precompute, stat_forward_teacher with every optional output
 * requested, then this.
 *   p          the reference tensors themselves (not the prepared block);
 *   f          the forward's step-invariant blocks inside its workspace (offsets from
 *              stat_workspace_region: "ctxg0","pctxg","ctxm0","pctxm","ctxl0","pctxl","qctxl";
 *              h0c0 = the "h0" region, (B,2H) rows of [h0 | c0]);
 *   x, mask, ctxg, mask_ctxg, ctxl, ctxm, dp_*   the inputs the forward saw (dp_* NULL = eval factors 0.5);
 *   alpha_l (L,B,T,R), alpha_g/m/lt (L,B,T), h_all (L,B,H)   stat_forward_teacher's outputs;
 *   inv_batch  1 / (clips in the GLOBAL batch): with data parallelism every rank passes the same
 *              value, adds decay_c on one rank only (or decay_c/world on each) and SUM-all-reduces;
 *   grads      a StatParams whose pointers name the OUTPUT buffers (written, not read), one per
 *              parameter in the reference's shapes -- e.g. views into the flat gradient buffer the
 *              optimizer entry points use;
 *   gws        stat_grad_workspace_bytes(d, L) bytes of scratch.
 * Deterministic: no atomics, fixed summation orders.  H <= 1024, R <= 16.
 * Environment: STAT_BW_FAST=1 selects the optimised variants (k-split products, deferred accumulation of the
 * step-invariant gradient blocks, row-wise embedding scatter); same results up to fp32 summation order. */
typedef struct StatFwdBlocks {
  const float *ctxg0, *pctxg, *ctxm0, *pctxm;   /* (B,T,H)    */
  const float *ctxl0, *pctxl, *qctxl;           /* (B,T,R,H)  */
  const float *h0c0;                            /* (B,2H)     */
} StatFwdBlocks;
size_t stat_grad_workspace_bytes(const StatDims *d, int L);
int stat_grad_shared(const StatDims *d, const StatParams *p, const StatFwdBlocks *f, int L,
                     const int64_t *x, const float *mask, const float *ctxg, const float *mask_ctxg,
                     const float *ctxl, const float *ctxm, const float *dp_gates, const float *dp_h,
                     const float *dp_z, const float *alpha_l, const float *alpha_g,
                     const float *alpha_m, const float *alpha_lt, const float *h_all,
                     float inv_batch, float alpha_c, float decay_c, const StatParams *grads,
                     void *gws, void *stream);

/* Phase timing of stat_grad_shared (measurement hook, no reference counterpart): when enabled, CUDA
 * events on the caller's stream bracket each launch group; collect synchronises on them and returns the
 * summed milliseconds / occurrences per phase.  Process-wide; not while capturing a graph. */
int         stat_grad_profile_enable(int on);
int         stat_grad_profile_phases(void);
const char *stat_grad_profile_phase_name(int phase);
int         stat_grad_profile_collect(float *ms_by_phase, int *count_by_phase, int nphase);

/* ---- the attention fragment of one step alone (model_attention.py:370-435) ---
 * S1-S9 for `rows` decode rows whose hidden-state projections (h.Wd*_att, selector
 * logit) already sit in the "hp" region of `ws` (left there by the previous
 * stat_step / decode step); writes the fused, gated context to the "ctx" region.
 * Exposed so that the HBM-bound kernel of the path can be timed in isolation. */
int stat_attention(const StatDims *d, const void *prepared, void *ws, int rows,
                   const int32_t *row_clip, void *stream);

/* ---- the dense primitive, exposed for the parity tests ---------------------
 * C[m][n] = post * act(alpha * sum_k A[m][k]*Bt[n][k] + bias[n]),  A (M,K) and
 * Bt (N,K) row-major (both "K-major"); act: 0 none, 1 tanh.  swap != 0 runs
 * the same problem with the operands exchanged on the tensor core (Bt rows on
 * the 128-lane axis), the form used for skinny activations. */
int stat_gemm(const float *A, int lda, const float *Bt, int ldb, float *C, int ldc,
              int M, int N, int K, const float *bias, float alpha, float post,
              int act, int swap, void *stream);

/* ---- measurement hooks (no reference counterpart) ---------------------------
 * stat_launch_count: kernels launched by this library so far in this process
 * (launches recorded into a CUDA graph count once, at capture).
 * stat_profile_*: when enabled, every launch group of stat_precompute / the
 * decode step is bracketed by CUDA events on the caller's stream;
 * stat_profile_collect synchronises on them and returns the summed milliseconds
 * and number of occurrences per phase (stat_profile_phase_name).  Do not enable
 * while capturing a graph. */
/* debugging aid: when set to a device buffer of >= 64 int64, the tensor-core GEMM's CTA
 * (0,0,0) stores clock64() stamps of its pipeline phases there; NULL switches it off. */
int         stat_debug_gemm_trace(void *dev_buffer_64_int64);
unsigned long long stat_launch_count(void);
int         stat_profile_enable(int on);
int         stat_profile_phases(void);
const char *stat_profile_phase_name(int phase);
int         stat_profile_collect(float *ms_by_phase, int *count_by_phase, int nphase);

#ifdef __cplusplus
}
#endif
#endif /* STAT_B200_H */
