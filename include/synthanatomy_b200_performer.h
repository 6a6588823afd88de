/* synthanatomy_b200 -- C ABI of the Performer prior hot path (sm_100a).
 *
 * Same conventions as synthanatomy_b200.h: plain DEVICE pointers + sizes, every call enqueues on `stream`,
 * never allocates, never synchronises, returns SA_OK or a negative sa_status; NO CPU fallback.
 *
 * The reference has no FFI for this path: it is Python on top of torch / cuBLAS plus the third-party packages
 * performer-pytorch==1.0.11, local-attention and pytorch-fast-transformers (absent from /root/reference; see
 * oracle/performer_oracle.py).  Each entry point names the reference call site it replaces (paths relative to
 * /root/reference) and, where the arithmetic lives in a third-party package, the function of that package.
 *
 * Layout: activations are row-major [rows = batch * seq][channels]; "act dtype" is SA_F32 (parity path, CUDA-core
 * fp32 FMA) or SA_BF16 (tcgen05 path, fp32 accumulation).  q | k | v of one layer live in ONE [rows][3 * inner]
 * buffer (inner = heads * dim_head; head h = columns [h * dim_head, (h + 1) * dim_head) of each third), the first
 * `global_heads` heads are FAVOR+ heads, the rest local-window heads (performer-pytorch SelfAttention.forward).
 */
#ifndef SYNTHANATOMY_B200_PERFORMER_H
#define SYNTHANATOMY_B200_PERFORMER_H

#include "synthanatomy_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

/* ------------------------------------------------------------------------------------------------
 * Dense layers.  Replaces torch.nn.Linear / F.linear (cuBLAS) at: SelfAttention.to_q/to_k/to_v/to_out
 * (performer.py:194-219 -> performer-pytorch), FeedForward.w1/w2, Performer.to_out (performer.py:221,286),
 * and their autograd data / weight gradients; fuses ReZero scale + residual (performer-pytorch ReZero,
 * SequentialSequence), GELU, bias.
 *
 * sa_gemm_nt:  v[i][j] = sum_k A[i][k] * B[j][k]                      (A: [m][lda], B: [n][ldb], act dtype)
 *   epilogue, in this order (every pointer optional):
 *     v += bias[j]
 *     dot_out[0] += sum_ij v * dot_with[i][j]          (act dtype, leading dimension ldo)
 *     v *= scale * scale_dev[0]
 *     act == SA_ACT_GELU_FWD:   pre[i][j] = v; v = gelu(v)        (exact erf GELU = nn.GELU())
 *     act == SA_ACT_GELU_BWD:   v *= gelu'(pre[i][j])
 *     act == SA_ACT_GELU_FWD_D: pre[i][j] = gelu'(v); v = gelu(v)  (the forward pass keeps the DERIVATIVE for the backward
 *     act == SA_ACT_MUL_PRE:    v *= pre[i][j]                      pass: one erf evaluation per element instead of two)
 *     v += resid[i][j]                                  (fp32, may alias out_f32)
 *     out_f32[i][j] = v;  out_act[i][j] = (act dtype) v
 *   pre / resid / dot_with / out_* all have leading dimension ldo.
 * ---------------------------------------------------------------------------------------------- */
typedef enum sa_act {
  SA_ACT_NONE = 0, SA_ACT_GELU_FWD = 1, SA_ACT_GELU_BWD = 2, SA_ACT_GELU_FWD_D = 3, SA_ACT_MUL_PRE = 4
} sa_act;

typedef struct sa_gemm_epilogue {
  const float* bias;
  const void* dot_with;
  float* dot_out;
  const float* scale_dev;
  float scale;
  int32_t act;
  void* pre;
  const float* resid;
  float* out_f32;
  void* out_act;
} sa_gemm_epilogue;

int sa_gemm_nt(int64_t m, int n, int k, int dtype, const void* a, int64_t lda, const void* b, int64_t ldb,
               const sa_gemm_epilogue* epi, int64_t ldo, void* stream);

/* sa_gemm_tn (weight gradients):  D[i][j] (+)= scale * scale_dev[0] * sum_r A[r][i] * B[r][j]
 *   A: [m][lda] (na columns used), B: [m][ldb] (nb columns used), act dtype; D fp32 [na][nb] dense.
 *   accumulate == 0: D is zeroed on `stream` first. */
int sa_gemm_tn(int64_t m, int na, int nb, int dtype, const void* a, int64_t lda, const void* b, int64_t ldb,
               const float* scale_dev, float scale, float* d, int accumulate, void* stream);

/* sa_gemm_tn that also returns colsum[i] = sum_r A[r][i] (fp32 [na], overwritten, never scaled): the bias gradient that
 * goes with a dense layer's weight gradient (A = dy, B = the layer's input) without a second pass over dy.  On the
 * tensor-core path it is one more 16-column product A^T 1 of the same staged A tiles. */
int sa_gemm_tn_colsum(int64_t m, int na, int nb, int dtype, const void* a, int64_t lda, const void* b, int64_t ldb,
                      const float* scale_dev, float scale, float* d, int accumulate, float* colsum, void* stream);

/* bf16x3 "parity" arithmetic of the two dense entry points (see sa_conv3d_fwd_x3 in synthanatomy_b200.h): fp32 operands
 * and fp32 epilogue tensors (every `void*` of the epilogue is fp32 here), products on the bf16 tensor cores as
 * hi.hi + lo.hi + hi.lo with fp32 accumulation.  The reference runs these layers in fp32 storage with TF32 matmuls
 * (run_transformer.py:165 amp=False); this path is closer to fp32 than TF32 is.  `workspace`: 256-byte aligned,
 * sa_gemm_nt_x3_workspace / sa_gemm_tn_x3_workspace bytes. */
size_t sa_gemm_nt_x3_workspace(int64_t m, int n, int k);
int sa_gemm_nt_x3(int64_t m, int n, int k, const float* a, int64_t lda, const float* b, int64_t ldb,
                  const sa_gemm_epilogue* epi, int64_t ldo, void* workspace, size_t ws_bytes, void* stream);
size_t sa_gemm_tn_x3_workspace(int64_t m, int na, int nb);
int sa_gemm_tn_x3(int64_t m, int na, int nb, const float* a, int64_t lda, const float* b, int64_t ldb,
                  const float* scale_dev, float scale, float* d, int accumulate, void* workspace, size_t ws_bytes,
                  void* stream);

/* ------------------------------------------------------------------------------------------------
 * Embedding front end, performer.py:241-268 (conditioning off, dropout 0):
 *   x[b][n] = tok_w[tokens[b][n]] + sum_a (sp_idx[a][n] >= 0 ? sp_w[a][sp_idx[a][n]] : 0) + pos_w[n]
 * sp_idx[a][n] is the coordinate along axis a of sequence position n - 1 (AbsoluteSpatialPositionalEmbedding,
 * performer.py:23-40: table row = coordinate value, BOS row zero-padded => -1 at n = 0).  n_axes in [0, 3].
 * x_f32 / x_act: either may be NULL.  Backward scatters dx into the four tables (fp32, accumulated).
 * ---------------------------------------------------------------------------------------------- */
int sa_embed_fwd(const int64_t* tokens, const int32_t* sp_idx, int n_axes, const float* tok_w, const float* const* sp_w,
                 const float* pos_w, int batch, int seq, int dim, int num_tokens, float* x_f32, void* x_act,
                 int act_dtype, void* stream);
/* num_tokens / sp_rows[a]: rows of the token table / of spatial table a (the deterministic mode gathers per table row,
 * in ascending position order, instead of scattering with atomics) */
int sa_embed_bwd(const float* dx, const int64_t* tokens, const int32_t* sp_idx, int n_axes, int batch, int seq, int dim,
                 int num_tokens, const int32_t* sp_rows, float* d_tok_w, float* const* d_sp_w, float* d_pos_w,
                 void* stream);

/* ------------------------------------------------------------------------------------------------
 * FAVOR+ (global heads).  Replaces performer-pytorch 1.0.11 softmax_kernel + causal_linear_attention and
 * fast_transformers CausalDotProduct (reached from performer.py:270).
 * ---------------------------------------------------------------------------------------------- */
typedef struct sa_favor_desc {
  int32_t batch, seq;
  int32_t heads;      /* number of global (FAVOR+) heads */
  int32_t dim_head;   /* d (64) */
  int32_t m;          /* random features (266) */
  int32_t mp;         /* row stride of the feature tensors, >= m; columns [m, mp) are written as zeros */
  int32_t ld;         /* leading dimension (elements) of the q/k/v/out row-major buffers the head columns live in */
  int32_t act_dtype;  /* dtype of q/k/v/out and of the feature tensors */
} sa_favor_desc;

/* kmax[0] = packed (ordered fp32 bits << 32 | ~flat index) maximum of c * k . P^T over (batch, head, seq, m),
 * c = d^-1/4 -- the GLOBAL key stabiliser of softmax_kernel(is_query=False).  kmax must be zeroed by the caller.
 * k points at column 0 of head 0 of the k block. */
int sa_favor_kmax(const sa_favor_desc* d, const void* k, const float* proj, unsigned long long* kmax, void* stream);

/* feat[b][h][n][j] = r * (exp(c x.P_j - c^2 |x|^2 / 2 - stab) + eps),  r = m^-1/2;
 * stab = row max over j (is_query, argmax[b][h][n] receives its j) or the global maximum in kmax (keys). */
int sa_favor_featmap_fwd(const sa_favor_desc* d, const void* x, const float* proj, int is_query,
                         const unsigned long long* kmax, float eps, void* feat, int32_t* argmax, void* stream);

/* Backward of the feature map (the stabilisers are NOT detached, as in 1.0.11):
 *   dx[b][n][h*d + :] = gradient w.r.t. the head columns (written, act dtype, leading dimension d->ld).
 *   keys: gsum[0] += sum of g_D over everything (zeroed by the caller); call sa_favor_kmax_fixup afterwards. */
int sa_favor_featmap_bwd(const sa_favor_desc* d, const void* x, const float* proj, int is_query, float eps,
                         const void* feat, const void* dfeat, const int32_t* argmax, void* dx, float* gsum,
                         void* stream);
/* dk[row*][head* cols] -= gsum[0] * c * P[j*]  at the arg-max position decoded from kmax. */
int sa_favor_kmax_fixup(const sa_favor_desc* d, const float* proj, const unsigned long long* kmax, const float* gsum,
                        void* dk, void* stream);

/* Causal linear attention.  out[b][n][h*d + e] = (q'[n] . S[n])[e] / den[n],  S[n] = sum_{j<=n} k'[j] (x) v[j],
 * den[n] = q'[n] . (sum_{j<=n} k'[j] + eps_cumsum).  den [batch][heads][seq] fp32 is saved for the backward.
 * Workspace: sa_favor_scan_workspace(d, backward) bytes. */
size_t sa_favor_scan_workspace(const sa_favor_desc* d, int backward);
int sa_favor_scan_fwd(const sa_favor_desc* d, const void* qf, const void* kf, const void* v, float eps_cumsum, void* out,
                      int out_ld, float* den, void* workspace, size_t ws_bytes, void* stream);
int sa_favor_scan_bwd(const sa_favor_desc* d, const void* qf, const void* kf, const void* v, float eps_cumsum,
                      const void* out, const void* dout, int out_ld, const float* den, void* dqf, void* dkf, void* dv,
                      void* workspace, size_t ws_bytes, void* stream);
/* Variants that keep the per-chunk prefix states S of the forward pass for the backward pass instead of recomputing
 * them.  sa_favor_scan_states_bytes() is the size of that buffer (0 when the selected path does not save states: pass
 * NULL then); the contents are opaque and only valid for the same descriptor and k', v. */
size_t sa_favor_scan_states_bytes(const sa_favor_desc* d);
int sa_favor_scan_fwd_save(const sa_favor_desc* d, const void* qf, const void* kf, const void* v, float eps_cumsum,
                           void* out, int out_ld, float* den, void* workspace, size_t ws_bytes, void* states,
                           size_t states_bytes, void* stream);
int sa_favor_scan_bwd_saved(const sa_favor_desc* d, const void* qf, const void* kf, const void* v, float eps_cumsum,
                            const void* out, const void* dout, int out_ld, const float* den, void* dqf, void* dkf,
                            void* dv, void* workspace, size_t ws_bytes, const void* states, size_t states_bytes,
                            void* stream);

/* sa_favor_scan_bwd_saved + sa_favor_featmap_bwd of the queries and of the keys in one call: the gradients of the
 * feature tensors (dq', dk') are consumed block by block inside the dq' / dk' kernels and never stored.
 *   x_q, x_k : the q / k head columns the features were computed from (leading dimension d->ld);  dx_q, dx_k likewise
 *   argq     : arg-max feature per query row, as saved by sa_favor_featmap_fwd;  *gsum += sum over all key rows of dD
 *              (what sa_favor_kmax_fixup needs afterwards);  states may be NULL (recomputed).
 * Only the tcgen05 path has this form: ask sa_favor_scan_bwd_fused_supported() first (else SA_ERR_UNSUPPORTED). */
int sa_favor_scan_bwd_fused_supported(const sa_favor_desc* d);
int sa_favor_scan_bwd_fused(const sa_favor_desc* d, const void* qf, const void* kf, const void* x_q, const void* x_k,
                            const void* v, const float* proj, float eps_cumsum, float eps_feature, const void* out,
                            const void* dout, int out_ld, const float* den, const int32_t* argq, void* dx_q, void* dx_k,
                            void* dv, float* gsum, void* workspace, size_t ws_bytes, const void* states, size_t states_bytes,
                            void* stream);

/* ------------------------------------------------------------------------------------------------
 * Local-window heads.  Replaces local_attention.LocalAttention.forward (window w, causal, look_backward = 1,
 * autopad, scale d^-1/2, optional rotary position term with frequencies inv_freq[d/2]):
 * query p attends keys j with (floor(p / w) - 1) * w <= j <= p.   lse [batch][heads][seq] fp32 is saved.
 * ---------------------------------------------------------------------------------------------- */
typedef struct sa_local_desc {
  int32_t batch, seq;
  int32_t heads;      /* number of local heads */
  int32_t dim_head;
  int32_t window;
  int32_t ld;         /* leading dimension of q/k/v (and dq/dk/dv) */
  int32_t out_ld;     /* leading dimension of out / dout */
  int32_t act_dtype;
} sa_local_desc;

/* In-place rotary position term of the local heads (local_attention.apply_rotary_pos_emb with SinusoidalEmbeddings):
 *   x <- x * cos(n * inv_freq) + rotate_half(x) * sin(n * inv_freq)    for the `heads` head blocks starting at `buf`
 * inverse != 0 applies the transpose (the backward of the map).  buf: [batch * seq][ld], act dtype. */
int sa_rotary(void* buf, int dtype, int64_t ld, int batch, int seq, int heads, int dim_head, const float* inv_freq,
              int inverse, void* stream);
/* The q and the k head blocks of one row-major buffer in ONE launch: the second block starts k_offset elements after the first. */
int sa_rotary_qk(void* buf, int dtype, int64_t ld, int64_t k_offset, int batch, int seq, int heads, int dim_head,
                 const float* inv_freq, int inverse, void* stream);

/* inv_freq == NULL: q and k already carry the position term (sa_rotary) or none is wanted; the tcgen05 kernels take
 * this form only. */
int sa_local_attn_fwd(const sa_local_desc* d, const void* q, const void* k, const void* v, const float* inv_freq,
                      void* out, float* lse, void* stream);
int sa_local_attn_bwd(const sa_local_desc* d, const void* q, const void* k, const void* v, const float* inv_freq,
                      const void* out, const void* dout, const float* lse, void* dq, void* dk, void* dv, void* stream);
/* Same, with an optional scratch buffer delta_ws [batch][heads][seq] fp32 (may be NULL): the dq kernel leaves
 * delta = dO . O per query there and the dk/dv kernel reads it back instead of recomputing it for every key tile. */
int sa_local_attn_bwd_ws(const sa_local_desc* d, const void* q, const void* k, const void* v, const float* inv_freq,
                         const void* out, const void* dout, const float* lse, void* dq, void* dk, void* dv,
                         float* delta_ws, void* stream);
/* q and k already carry the rotary term (sa_rotary / sa_rotary_qk applied in place before sa_local_attn_fwd, inv_freq ==
 * NULL there): dq and dk leave through the transpose of the rotation, applied to the fp32 gradient rows inside the
 * kernels' epilogues instead of a separate in-place pass over the stored gradients.  rot_table: [seq][dim_head / 2]
 * (cos, sin) pairs of n * inv_freq, fp32, 16-byte aligned, written by sa_rotary_table (same sincosf as sa_rotary). */
int sa_rotary_table(const float* inv_freq, int seq, int dim_head, float* table, void* stream);
int sa_local_attn_bwd_rot(const sa_local_desc* d, const void* q, const void* k, const void* v, const float* inv_freq,
                          const float* rot_table, const void* out, const void* dout, const float* lse, void* dq, void* dk,
                          void* dv, float* delta_ws, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Autoregressive sampling with recurrent state.  The reference re-runs the whole network
 * on the growing prefix for every sampled token (src/networks/transformers/transformer.py:58-101); these entry points
 * advance the attention state of one layer by ONE position and return what the prefix forward returns at its last
 * position.  q / k / v: the [batch][ld] rows of the new position (column 0 of head 0 of each block).
 *
 * sa_favor_decode_step (global heads): t = number of keys already absorbed (= position of the new token).
 *   mhist[t] = ordered-uint encoding of the key stabiliser after t keys (mhist[0] = encoding of -inf = 0x007FFFFF,
 *   later entries zero); Se [batch*heads][m][64], ze [batch*heads][m], S1 [batch*heads][64] fp32, zero before position 0;
 *   scratch: (2 * batch * heads * m + batch * heads) floats.
 * sa_local_decode_step (local heads): kcache / vcache [batch][nmax][heads * 64] (act dtype) receive the rotated key and
 *   the value of position p, which then attends the cached positions (floor(p / w) - 1) w .. p.
 * ---------------------------------------------------------------------------------------------- */
int sa_favor_decode_step(int batch, int heads, int m, int dtype, int t, const int* t_dev, const void* q, const void* k,
                         const void* v, int ld, const float* proj, float eps, float eps_cumsum, unsigned int* mhist,
                         float* scratch, float* Se, float* ze, float* S1, void* out, int out_ld, void* stream);
int sa_local_decode_step(int batch, int heads, int window, int dtype, int p, const int* p_dev, int nmax, const void* q,
                         const void* k, const void* v, int ld, const float* inv_freq, void* kcache, void* vcache, void* out,
                         int out_ld, void* stream);
/* Embedding of ONE position (performer.py:241-268): tokens [batch] int64, sp_idx [n_axes][sp_ld] as in sa_embed_fwd.
 * In all three entry points a non-NULL t_dev / p_dev (device int) overrides the host position, so that one captured
 * CUDA graph of a decoding step can be replayed for every position. */
int sa_embed_step(const int64_t* tokens, const int32_t* sp_idx, int n_axes, int sp_ld, const float* tok_w,
                  const float* const* sp_w, const float* pos_w, int batch, int dim, int t, const int* t_dev, float* x_f32,
                  void* x_act, int act_dtype, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Output head: nn.LayerNorm (performer.py:273) and cross-entropy over the logits
 * (inferer/transformer.py:29 + losses/transformer/transformer.py:24-33).
 * ---------------------------------------------------------------------------------------------- */
int sa_layernorm_fwd(const float* x, const float* w, const float* b, int64_t rows, int dim, float eps, float* y_f32,
                     void* y_act, int act_dtype, float* mean, float* rstd, void* stream);
/* dx = LN backward; dw/db accumulated (zeroed by the caller) */
int sa_layernorm_bwd(const float* dy, const float* x, const float* w, const float* mean, const float* rstd, int64_t rows,
                     int dim, float* dx, float* dw, float* db, void* stream);
/* loss_sum[0] += sum_rows (logsumexp(logits[row]) - logits[row][target[row]])                      (may be NULL)
 * dlogits[row][j] = grad_scale * grad_scale_dev[0] * (softmax(logits[row])[j] - [j == target[row]])
 * (dlogits may alias logits, may be NULL; grad_scale_dev is a DEVICE scalar -- the incoming loss gradient -- or NULL) */
int sa_ce_fwd_bwd(const float* logits, int64_t ld, const int64_t* target, int64_t rows, int vocab, float grad_scale,
                  const float* grad_scale_dev, float* loss_sum, float* dlogits, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Token plumbing between the two models (integer gathers; all pointers are device pointers)
 * ---------------------------------------------------------------------------------------------- */
/* element type of a latent index grid: uint16 is what the extraction mode stores per subject (run_vqvae.py:484-498),
 * int64 what index_quantize returns (baseline.py:343-346) */
typedef enum sa_tok_dtype { SA_TOK_U16 = 0, SA_TOK_I32 = 1, SA_TOK_I64 = 2 } sa_tok_dtype;
/* prepare_batch (src/utils/transformer.py:259-282) in one launch:
 *   y[b][i] = grid[b][order[i]]   x_in[b][0] = bos   x_in[b][i] = y[b][i - 1]      grid [batch][n_src], order [n] (n <= n_src) */
int sa_tokens_prepare(const void* grid, int tok_dtype, int batch, int64_t n_src, const int64_t* order, int64_t n, int64_t bos,
                      int64_t* x_in, int64_t* y, void* stream);
/* out[b][i] = src[b][index[i]]  -- e.g. sampled sequence -> grid with the reverted ordering (src/inferer/transformer.py:63-71) */
int sa_tokens_gather(const void* src, int tok_dtype, int batch, int64_t n_src, const int64_t* index, int64_t n, int64_t* out,
                     void* stream);
/* out[i] = (uint16) src[i]; out_of_range[0] is set to 1 if any value is outside [0, 65535] (zeroed by the caller) */
int sa_tokens_narrow(const int64_t* src, int64_t n, uint16_t* out, int* out_of_range, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Small helpers
 * ---------------------------------------------------------------------------------------------- */
/* dst[r][c] = (dst dtype) src[r][c] for c < cols; columns [cols, dst_ld) of dst are zero-filled (pads the logits
 * gradient to a TMA-friendly leading dimension). */
int sa_cast2d(const void* src, int src_dtype, int64_t src_ld, void* dst, int dst_dtype, int64_t dst_ld, int64_t rows,
              int cols, void* stream);
/* ReZero gate gradient read off the UNSCALED weight gradient t = dx^T a (sa_gemm_tn without scale) of a gated layer
 * y = x + g (a W^T) (performer-pytorch ReZero):  dot[0] += sum t . w  ( = sum (dx W) . a );  t *= g[0]  (t becomes dW).
 * w: the layer's fp32 weight, same [out][in] layout as t; n = out * in.  Saves reading the [rows][in] activation again
 * in the data-gradient GEMM's epilogue. */
int sa_gate_wgrad(float* t, const float* w, int64_t n, const float* g, float* dot, void* stream);

/* ReZero backward bookkeeping (performer-pytorch ReZero: y = g * f(x)), f(x) = core(x) + bias:
 *   dg[0] = dot[0] + sum_c bias[c] * colsum[c];   dbias[c] = g[0] * colsum[c]
 * colsum = column sums of the incoming gradient, dot = sum(grad * core(x)) from the sa_gemm_nt epilogue.
 * bias / dbias / dot may be NULL (attention sub-layer: no bias). */
int sa_rezero_finish(const float* colsum, const float* bias, const float* g, const float* dot, int n, float* dbias,
                     float* dg, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* SYNTHANATOMY_B200_PERFORMER_H */
