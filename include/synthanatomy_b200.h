/* synthanatomy_b200 -- C ABI of the B200 (sm_100a) hot path of AmigoLab/SynthAnatomy.
 *
 * Drop-in boundary: the reference has no FFI of its own (pure Python on top of torch / cuDNN /
 * cuBLAS); what a maintainer binds is the set of torch operator calls on the hot path.  Every entry
 * point below names the reference call site(s) it replaces (paths relative to /root/reference).
 *
 * Conventions
 *  - plain pointers + sizes, no torch types; every pointer is a DEVICE pointer unless stated.
 *  - activations are channels-last "NDHWC" (batch, depth, height, width, channel), fp32 or bf16.
 *  - every function enqueues on `stream` (a cudaStream_t passed as void*) and never allocates,
 *    never synchronises; workspace (if any) is supplied by the caller.
 *  - returns SA_OK (0) or a negative sa_status; sa_last_error() gives the message of the last
 *    failure on the calling thread.  Unsupported configurations return SA_ERR_UNSUPPORTED:
 *    there is NO CPU fallback anywhere in this library.
 */
#ifndef SYNTHANATOMY_B200_H
#define SYNTHANATOMY_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum sa_status {
  SA_OK = 0,
  SA_ERR_INVALID = -1,      /* bad argument (null pointer, negative size, ...) */
  SA_ERR_UNSUPPORTED = -2,  /* configuration outside what the kernels implement */
  SA_ERR_CUDA = -3,         /* a CUDA runtime / driver call failed */
  SA_ERR_WORKSPACE = -4     /* workspace too small */
} sa_status;

typedef enum sa_dtype { SA_F32 = 0, SA_BF16 = 1 } sa_dtype;

/* which implementation a dispatching entry point used last on this thread (for tests / gpu_launches) */
typedef enum sa_path { SA_PATH_NONE = 0, SA_PATH_SIMT = 1, SA_PATH_TCGEN05 = 2 } sa_path;

const char* sa_last_error(void);
int sa_version(void);
int sa_last_path(void);
/* number of kernels launched by this library (all threads of the process: autograd runs backward passes on its own
 * device thread) since the last reset */
int64_t sa_launch_count(void);
void sa_launch_count_reset(void);
/* 0: dispatch normally; 1: force the CUDA-core (SIMT) kernels even where a tcgen05 kernel exists */
void sa_set_force_simt(int on);
/* 1: every cross-CTA floating-point accumulation (split-K partials, bias / gate / loss sums, codebook statistics) is
 * added in one fixed order (a turnstile of device counters per launch), so that a training step is bit-reproducible run
 * to run -- the reference's `deterministic=True` (run_vqvae.py:550, run_transformer.py:417; src/utils/general.py:333).
 * 0 (default): partials are added in arrival order with atomics. */
void sa_set_deterministic(int on);
int sa_get_deterministic(void);

/* ------------------------------------------------------------------------------------------------
 * Gather-GEMM convolution primitive.
 *
 *   transposed == 0 (FORM_CONV):
 *     Y[b, o, n] = sum_{t, c}  X[b, o*stride - pad + t, c] * Wp[t][n][c]
 *   transposed == 1 (FORM_TCONV):
 *     Y[b, o, n] = sum_{t, c : (o + pad - t) % stride == 0}  X[b, (o + pad - t)/stride, c] * Wp[t][n][c]
 *
 * o, t are 3-vectors (depth, height, width), t ranges over ksize^3 taps (t = (td*ksize + th)*ksize + tw),
 * out-of-range X reads are zero.  Wp is the PACKED weight [ksize^3][c_out][c_in] in the activation dtype
 * (see sa_pack_weight).  Epilogue, in this order:  v = acc (+ bias[n]) (+ addend[b,o,n]);
 * if relu: v = max(v, 0);  if mask: v = mask[b,o,n] > 0 ? v : 0.
 *
 * Replaces: nn.Conv3d / nn.ConvTranspose3d forward and their autograd data-gradients --
 *   src/networks/vqvae/baseline.py:153,156 (ResidualLayer convs), :218-227 (strided down-convs),
 *   :242-244, :258 (pre/post-quant convs), :283-293 (ConvTranspose3d), and the fused elementwise
 *   nn.ReLU / F.relu(x + .) at :154,160,228,296-297.
 * ---------------------------------------------------------------------------------------------- */
typedef struct sa_conv_desc {
  int32_t batch;
  int32_t in_dhw[3];   /* spatial extent of X (the gathered tensor) */
  int32_t out_dhw[3];  /* spatial extent of Y */
  int32_t c_in;        /* channels of X */
  int32_t c_out;       /* channels of Y */
  int32_t ksize, stride, pad;
  int32_t transposed;
  int32_t act_dtype;   /* sa_dtype of X, Y, addend, mask and Wp */
} sa_conv_desc;

int sa_conv3d_fwd(const sa_conv_desc* d, const void* x, const void* wp, const float* bias,
                  const void* addend, const void* mask, int relu, void* y, void* stream);

/* Weight gradient of the primitive above, always in FORM_CONV indexing:
 *   dWp[t][n][c] (+)= sum_{b, o}  P[b, o, n] * Q[b, o*stride - pad + t, c]
 * P has extent out_dhw with c_out channels, Q has extent in_dhw with c_in channels; dWp is fp32
 * [ksize^3][c_out][c_in].  If accumulate == 0 dWp is overwritten (it is zeroed on `stream` first).
 * Replaces: cuDNN wgrad reached through autograd of the convs listed above. */
int sa_conv3d_wgrad(const sa_conv_desc* d, const void* p, const void* q, float* dwp, int accumulate,
                    void* stream);

/* ------------------------------------------------------------------------------------------------
 * bf16x3 "parity" arithmetic of the two entry points above: fp32 tensors in and out, products on the bf16 tensor cores
 * with each fp32 operand split into hi + lo bf16 parts (x ~= hi + lo to 16 significand bits) and the three cross terms
 * hi.hi + lo.hi + hi.lo evaluated as ONE contraction three times as long (input channels concatenated for the forward /
 * data-gradient form, the batch concatenated for the weight gradient), fp32 accumulation.  Error ~1e-5 of the operand
 * scale: this is the tensor-core path that meets the reference's fp32 results to 1e-4 (the reference itself runs these
 * convs in fp16 autocast or TF32).  x / y / addend / mask / p / q are fp32 NDHWC, wp is the PACKED fp32 weight
 * (sa_pack_weight with dst_dtype SA_F32), d->act_dtype must be SA_F32.  `workspace`: sa_conv3d_x3_workspace(d, wgrad)
 * bytes, 256-byte aligned.  sa_conv3d_x3_supported tells whether the tensor-core kernels take the shape (otherwise the
 * caller uses sa_conv3d_fwd / sa_conv3d_wgrad with SA_F32, the CUDA-core fp32 kernels).
 * ---------------------------------------------------------------------------------------------- */
int sa_conv3d_x3_supported(const sa_conv_desc* d, int wgrad);
size_t sa_conv3d_x3_workspace(const sa_conv_desc* d, int wgrad);
int sa_conv3d_fwd_x3(const sa_conv_desc* d, const float* x, const float* wp, const float* bias, const float* addend,
                     const float* mask, int relu, float* y, void* workspace, size_t ws_bytes, void* stream);
int sa_conv3d_wgrad_x3(const sa_conv_desc* d, const float* p, const float* q, float* dwp, int accumulate,
                       void* workspace, size_t ws_bytes, void* stream);

/* Fused backward of the pointwise (1x1x1, 128 -> 128 channels) convolution of a ResidualLayer
 * (/root/reference/src/networks/vqvae/baseline.py:153-160: y = relu(x + conv1x1(h)), h = relu(conv3x3x3(x))).
 * g [m][c_out] = gradient w.r.t. the pre-activation of y, h [m][c_in] the saved activation (bf16, NDHWC flattened to
 * m positions), wp_t = sa_pack_weight(w1, transpose = 1).  One pass produces
 *   dh[m][c_in] = (g W1) * (h > 0),   dwp[c_out][c_in] += g^T h,   dbias[c_out] += column sums of g
 * (dwp / dbias are accumulated into: zero them first).  Replaces three launches of the general entry points. */
int sa_conv1x1_bwd_fused(int64_t m, int c_out, int c_in, const void* g, const void* h, const void* wp_t, void* dh,
                         float* dwp, float* dbias, void* stream);
/* The same pass with one more output: dbias_h[c_in] += column sums of dh (may be NULL) -- the bias gradient of the 3x3x3
 * conv that produced h (baseline.py:153), so that no separate reduction has to stream dh again. */
int sa_conv1x1_bwd_fused_dbh(int64_t m, int c_out, int c_in, const void* g, const void* h, const void* wp_t, void* dh,
                             float* dwp, float* dbias, float* dbias_h, void* stream);
/* Streaming forward twin: y[m][c_out] = relu?(x W^T + bias + addend), wp = sa_pack_weight(w, transpose = 0)
 * (128 -> 128 channels; the residual `addend` is required). */
int sa_conv1x1_fwd_fused(int64_t m, int c_out, int c_in, const void* x, const void* wp, const float* bias,
                         const void* addend, int relu, void* y, void* stream);

/* dst[t'][a][b] = src[a][b][t] (transpose == 0) or dst[t'][b][a] = src[a][b][t] (transpose == 1),
 * t' = flip ? taps-1-t : t.  src is the torch layout (fp32, [A][B][taps]); dst is sa_dtype dst_dtype.
 * sa_unpack_wgrad is the exact inverse on fp32 data (dst[a][b][t] (+)= src[...]), used to scatter dWp
 * back into the torch-layout .grad. */
int sa_pack_weight(const float* src, int A, int B, int taps, int transpose, int flip, void* dst,
                   int dst_dtype, void* stream);

/* sa_pack_weight for many weights in a few launches (`items` is host memory; one destination dtype per call) */
typedef struct sa_wpack_item {
  const float* src;
  void* dst;
  int A, B, taps;
  int transpose, flip;
} sa_wpack_item;
int sa_pack_weight_multi(const sa_wpack_item* items, int n, int dst_dtype, void* stream);
int sa_unpack_wgrad(const float* src, int A, int B, int taps, int transpose, int flip, float* dst,
                    int accumulate, void* stream);

/* Single-channel helpers that turn the two 1-channel layers (first Conv3d 1->C, baseline.py:218-227, and the last
 * ConvTranspose3d C->1, baseline.py:283-293) into 1x1x1 gather-GEMMs over a k^3-"channel" tensor:
 *   sa_im2col_c1:  cols[b, o, t] = x[b, o*stride - pad + t]              (x: [B, in_dhw], cols: [B, out_dhw, k^3])
 *   sa_col2im_c1:  y[b, o] = bias[0] + sum_{t: (o+pad-t) % stride == 0} cols[b, (o+pad-t)/stride, t]
 *                                                                        (cols: [B, in_dhw, k^3], y: [B, out_dhw])
 * dtype is the sa_dtype of x / cols / y. */
int sa_im2col_c1(const void* x, int dtype, int batch, const int* in_dhw, const int* out_dhw, int ksize, int stride,
                 int pad, void* cols, void* stream);
int sa_col2im_c1(const void* cols, int dtype, int batch, const int* in_dhw, const int* out_dhw, int ksize, int stride,
                 int pad, const float* bias, void* y, void* stream);

/* db[c] (+)= sum_rows dy[row][c]   (bias gradient of every conv above) */
int sa_bias_grad(const void* dy, int64_t rows, int c, int dtype, float* db, int accumulate, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Vector quantiser (EMA codebook).  Replaces Quantizer_impl.forward, baseline.py:38-87, and
 * Quantizer.forward's histogram, baseline.py:110-120.
 *
 * sa_vq_forward: z is [rows][dim] fp32 (NDHWC-flattened latents), codebook [n_embed][dim] fp32.
 *   idx[row]   = argmin_k ( ||z||^2 - 2 z.w_k + ||w_k||^2 ), evaluated in fp32 in that association,
 *                first (lowest) index on ties                                   (baseline.py:49-56)
 *   q[row]     = codebook[idx[row]]                                             (baseline.py:63)
 *                or, if straight_through != 0, (codebook[idx[row]] - z[row]) + z[row], the value of
 *                (quantized - x).detach() + x with its two fp32 roundings          (baseline.py:85)
 *   counts[k] += #rows with idx == k          (fp32; baseline.py:68)  -- may be NULL
 *   dw[k][:]  += sum of z rows with idx == k  (fp32; baseline.py:69)  -- may be NULL
 *   sse[0]    += sum (q - z)^2                (fp32; numerator of mse_loss, baseline.py:82) -- may be NULL
 * counts, dw and sse must be zeroed by the caller (they are accumulated so that ranks/tiles can share them).
 * idx is int64 (the reference returns LongTensor, baseline.py:56).
 * ---------------------------------------------------------------------------------------------- */
int sa_vq_forward(const float* z, const float* codebook, int64_t rows, int dim, int n_embed,
                  int64_t* idx, float* q, int straight_through, float* counts, float* dw, float* sse,
                  void* stream);

/* Backward of the quantiser outputs w.r.t. the latents (any consistent layout, n elements):
 *   dz = g_q + g_loss[0] * coef * (z - q)
 * g_q: gradient of the straight-through output (identity, baseline.py:85); g_loss: DEVICE scalar gradient of
 * the latent loss; coef = 2 * commitment_cost / numel (baseline.py:82).  g_q or g_loss may be NULL (= 0). */
int sa_vq_backward(const float* g_q, const float* g_loss, const float* z, const float* q, float coef,
                   int64_t n, float* dz, void* stream);

/* EMA + Laplace smoothing + codebook refresh, baseline.py:75-80 (counts / dw already all-reduced):
 *   N <- decay N + (1-decay) counts;  embed_avg <- decay embed_avg + (1-decay) dw;
 *   n = sum N;  W = (N + eps)/(n + n_embed eps) n;  codebook <- embed_avg / W
 * decay / eps are doubles because the reference forms `1 - decay` and `n_embed * eps` in Python doubles.
 * workspace: >= 4 bytes (receives n), may be NULL. */
int sa_vq_ema_update(float* N, float* embed_avg, float* codebook, const float* counts, const float* dw,
                     int n_embed, int dim, double decay, double eps, float* workspace, void* stream);

/* out[0] = exp(-sum_k p_k log(p_k + 1e-10)), p_k = counts[k] / total   (Quantizer.forward, baseline.py:110-120;
 * counts are the per-rank histogram written by sa_vq_forward). */
int sa_vq_perplexity(const float* counts, int n_embed, float total, float* out, void* stream);

/* q[row] = codebook[idx[row]]   (Quantizer_impl.embed, baseline.py:89-91) */
int sa_vq_embed(const int64_t* idx, const float* codebook, int64_t rows, int dim, int n_embed, float* q,
                void* stream);

/* ------------------------------------------------------------------------------------------------
 * PatchGAN discriminator blocks (src/networks/discriminator/baseline.py:43-79): nn.BatchNorm3d + nn.LeakyReLU over
 * channels-last activations [rows = B*D*H*W][channels] (act dtype fp32 / bf16, statistics and parameters fp32).
 * `workspace`: sa_bn_workspace(channels) bytes of device memory.
 * ---------------------------------------------------------------------------------------------- */
size_t sa_bn_workspace(int channels);
/* training-mode statistics: mean[c], rstd[c] = 1/sqrt(biased var + eps); running_mean / running_var (may be NULL)
 * are updated as nn.BatchNorm3d does: r <- (1 - momentum) r + momentum * (mean | unbiased var) */
int sa_bn_stats(const void* x, int dtype, int64_t rows, int channels, void* workspace, float eps, float momentum, float* mean,
                float* rstd, float* running_mean, float* running_var, void* stream);
/* eval-mode statistics: mean = running_mean, rstd = 1/sqrt(running_var + eps) */
int sa_bn_eval_stats(const float* running_mean, const float* running_var, int channels, float eps, float* mean, float* rstd,
                     void* stream);
/* y = leaky_relu(gamma * (x - mean) * rstd + beta, slope)   (slope = 1: plain BatchNorm) */
int sa_bn_lrelu_fwd(const void* x, int dtype, int64_t rows, int channels, const float* mean, const float* rstd, const float* gamma,
                    const float* beta, float slope, void* y, void* stream);
/* training-mode backward of the pair: g = dL/dy, x = BatchNorm input, y = the forward's output.
 *   dgamma, dbeta [channels];  dx (may be NULL) = gamma rstd (g' - mean(g') - xhat mean(g' xhat)),  g' = g lrelu'(y) */
int sa_bn_lrelu_bwd(const void* g, const void* x, const void* y, int dtype, int64_t rows, int channels, const float* mean,
                    const float* rstd, const float* gamma, float slope, void* workspace, float* dgamma, float* dbeta, void* dx,
                    void* stream);
/* in place: x <- leaky_relu(x, slope);   g <- g * leaky_relu'(y) */
int sa_lrelu_fwd(void* x, int dtype, int64_t n, float slope, void* stream);
int sa_lrelu_bwd(void* g, const void* y, int dtype, int64_t n, float slope, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Elementwise / layout helpers
 * ---------------------------------------------------------------------------------------------- */
/* dst[b][s][c] = src[b][c][s]  (NCDHW -> NDHWC, `spatial` = D*H*W) and back; dtypes are sa_dtype */
int sa_nchw_to_nhwc(const void* src, int src_dtype, void* dst, int dst_dtype, int64_t batch, int c,
                    int64_t spatial, void* stream);
int sa_nhwc_to_nchw(const void* src, int src_dtype, void* dst, int dst_dtype, int64_t batch, int c,
                    int64_t spatial, void* stream);
/* dst = (dst_dtype) src, n elements */
int sa_cast(const void* src, int src_dtype, void* dst, int dst_dtype, int64_t n, void* stream);

/* Per-step preparation of dense-layer weights, many tensors per launch: for each item dst[r][c] = bf16(src[r][c])
 * (leading dimension dst_ld; NULL: skipped) and dst_t[c][r] = bf16(src[r][c]) (leading dimension dst_t_ld; NULL: skipped);
 * src fp32 dense [rows][cols].  `items` is host memory.  Pointer + leading-dimension pairs let several sources land in
 * one destination (q | k | v weights -> one [3 inner x dim] operand and its transpose). */
typedef struct sa_wprep_item {
  const float* src;
  void* dst;
  void* dst_t;
  int rows, cols;
  int dst_ld, dst_t_ld;
} sa_wprep_item;
int sa_weight_prep(const sa_wprep_item* items, int n, void* stream);
/* sse[0] += sum (a-b)^2 ; grad = scale * scale_dev[0] * (a - b)  (F.mse_loss fwd+bwd,
 * src/losses/vqvae/vqvae.py:56).  a: prediction (a_dtype), b: target fp32; sse, grad (a_dtype) and the DEVICE
 * scalar scale_dev (the incoming loss gradient) may each be NULL. */
int sa_mse_fwd_bwd(const void* a, int a_dtype, const float* b, int64_t n, float scale, const float* scale_dev,
                   float* sse, void* grad, void* stream);
/* torch.optim.Adam step (run_vqvae.py:82, run_transformer.py:109): fp32 p, g, m, v; no weight decay.
 * `step` is the 1-based step count AFTER increment. */
int sa_adam_step(float* p, const float* g, float* m, float* v, int64_t n, float lr, float beta1, float beta2,
                 float eps, int step, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Spectral (Jukebox) reconstruction loss, src/losses/vqvae/vqvae.py:522-640 (torch.fft.fftn, norm="ortho", amplitude,
 * F.mse_loss).  The axis transforms are dense DFT-matrix products through sa_gemm_nt_x3 (bf16x3 on the tensor cores);
 * a complex tensor is stored as [..][2 = re, im][axis].  These two entry points are the non-GEMM parts:
 *   sa_swap_outer_inner:  dst[b][c][m][a] = src[b][a][m][c]  (fp32) -- brings the next axis to the contraction position
 *   sa_spectral_amp_loss: spectra [rows][2][L]; sse[0] += sum (|P| - |T|)^2 (may be NULL);
 *                         grad (may be NULL) = coef * coef_dev[0] * (|P| - |T|) * P / |P|   (coef_dev may be NULL = 1)
 * ---------------------------------------------------------------------------------------------- */
int sa_swap_outer_inner(const float* src, float* dst, int64_t batch, int A, int M, int C, void* stream);
int sa_spectral_amp_loss(const float* pred_spec, const float* target_spec, int64_t rows, int L, float coef,
                         const float* coef_dev, float* sse, float* grad, void* stream);

/* The same update for `count` parameter tensors at once (host arrays of device pointers / element counts; all tensors
 * share lr / betas / eps / step): 64 tensors per launch instead of one launch per tensor. */
int sa_adam_multi(int count, float* const* p, const float* const* g, float* const* m, float* const* v,
                  const int64_t* n, float lr, float beta1, float beta2, float eps, int step, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* SYNTHANATOMY_B200_H */
