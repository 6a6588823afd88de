// HBM-bound ends of the Performer path: embedding front end, final LayerNorm, cross-entropy.
// Reference call sites: /root/reference/src/networks/transformers/performer.py:241-268 (embeddings), :273 (LayerNorm),
// src/inferer/transformer.py:29 + src/losses/transformer/transformer.py:24-33 (cross-entropy over [B, V, N]).
#include "sa_pf_common.cuh"

namespace {

struct EmbPtrs {
  const float* sp_w[3];
  float* d_sp_w[3];
};

template <typename T>
__global__ void embed_fwd_kernel(const long long* __restrict__ tokens, const int* __restrict__ sp_idx, int n_axes,
                                 const float* __restrict__ tok_w, EmbPtrs sp, const float* __restrict__ pos_w, int B,
                                 int N, int dim, float* __restrict__ x_f32, T* __restrict__ x_act) {
  const long long total = (long long)B * N * dim;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % dim);
    const long long row = i / dim;
    const int n = (int)(row % N);
    float v = tok_w[tokens[row] * dim + c];                 // performer.py:241
    for (int a = 0; a < n_axes; ++a) {                      // :243-244 (sequential += per axis)
      const int s = sp_idx[a * N + n];
      if (s >= 0) v += sp.sp_w[a][(long long)s * dim + c];
    }
    v += pos_w[(long long)n * dim + c];                     // :266
    if (x_f32) x_f32[i] = v;
    if (x_act) sa_st(x_act, i, v);
  }
}

__global__ void embed_bwd_kernel(const float* __restrict__ dx, const long long* __restrict__ tokens,
                                 const int* __restrict__ sp_idx, int n_axes, int B, int N, int dim,
                                 float* __restrict__ d_tok_w, EmbPtrs sp, float* __restrict__ d_pos_w) {
  const long long total = (long long)N * dim;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % dim);
    const int n = (int)(i / dim);
    float sum = 0.f;
    for (int b = 0; b < B; ++b) {
      const long long row = (long long)b * N + n;
      const float g = dx[row * dim + c];
      sum += g;
      if (d_tok_w) atomicAdd(d_tok_w + tokens[row] * dim + c, g);
    }
    d_pos_w[i] += sum;
    for (int a = 0; a < n_axes; ++a) {
      const int s = sp_idx[a * N + n];
      if (s >= 0) atomicAdd(sp.d_sp_w[a] + (long long)s * dim + c, sum);
    }
  }
}

// Deterministic mode: the table gradients as ordered gathers instead of scattered atomics.  One block per table row
// `key`; the positions that index it are found 256 at a time (ballot + prefix: an ascending list) and added in that order.
//   KIND 0: token table   -- domain = the B x N token rows, contribution dx[row]
//   KIND 1: spatial table -- domain = the N positions of one axis, contribution sum_b dx[b * N + n] (ascending b)
template <int KIND>
__global__ void __launch_bounds__(256)
embed_bwd_ordered_kernel(const float* __restrict__ dx, const long long* __restrict__ tokens, const int* __restrict__ sp_idx,
                         int B, int N, int dim, float* __restrict__ d_w) {
  __shared__ int s_list[256];
  __shared__ int s_wcnt[8];
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
  const long long key = blockIdx.x;
  const long long domain = KIND == 0 ? (long long)B * N : (long long)N;
  float acc[4] = {0.f, 0.f, 0.f, 0.f};                 // dim <= 1024: columns t, t + 256, ...
  for (long long base = 0; base < domain; base += 256) {
    const long long r = base + t;
    const bool hit = r < domain && (KIND == 0 ? tokens[r] == key : (long long)sp_idx[r] == key);
    const unsigned m = __ballot_sync(0xffffffffu, hit);
    if (lane == 0) s_wcnt[warp] = __popc(m);
    __syncthreads();
    int off = 0, total = 0;
#pragma unroll
    for (int w = 0; w < 8; ++w) { if (w < warp) off += s_wcnt[w]; total += s_wcnt[w]; }
    if (hit) s_list[off + __popc(m & ((1u << lane) - 1u))] = t;
    __syncthreads();
    for (int j = 0; j < total; ++j) {
      const long long rr = base + s_list[j];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int c = t + 256 * k;
        if (c >= dim) break;
        if (KIND == 0) {
          acc[k] += dx[rr * dim + c];
        } else {
          float sum = 0.f;
          for (int b = 0; b < B; ++b) sum += dx[((long long)b * N + rr) * dim + c];
          acc[k] += sum;
        }
      }
    }
    __syncthreads();
  }
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int c = t + 256 * k;
    if (c < dim) d_w[key * dim + c] += acc[k];
  }
}

// one warp per row, dim <= 32 * 32
template <typename T>
__global__ void __launch_bounds__(256)
layernorm_fwd_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ b,
                     long long rows, int dim, float eps, float* __restrict__ y_f32, T* __restrict__ y_act,
                     float* __restrict__ mean, float* __restrict__ rstd) {
  const int lane = threadIdx.x & 31;
  const long long row = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (row >= rows) return;
  const float* xr = x + row * dim;
  float v[32];
  float s = 0.f;
  const int per = (dim + 31) / 32;
#pragma unroll
  for (int i = 0; i < 32; ++i) {
    if (i >= per) break;
    const int c = lane + 32 * i;
    v[i] = c < dim ? xr[c] : 0.f;
    s += v[i];
  }
  const float mu = sa_warp_sum(s) / dim;
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < 32; ++i) {
    if (i >= per) break;
    const int c = lane + 32 * i;
    const float dlt = c < dim ? v[i] - mu : 0.f;
    q = fmaf(dlt, dlt, q);
  }
  const float rs = rsqrtf(sa_warp_sum(q) / dim + eps);
  if (lane == 0) { mean[row] = mu; rstd[row] = rs; }
#pragma unroll
  for (int i = 0; i < 32; ++i) {
    if (i >= per) break;
    const int c = lane + 32 * i;
    if (c < dim) {
      const float o = (v[i] - mu) * rs * w[c] + b[c];
      if (y_f32) y_f32[row * dim + c] = o;
      if (y_act) sa_st(y_act, row * dim + c, o);
    }
  }
}

__global__ void __launch_bounds__(256)
layernorm_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ x, const float* __restrict__ w,
                     const float* __restrict__ mean, const float* __restrict__ rstd, long long rows, int dim,
                     float* __restrict__ dx, float* __restrict__ dw, float* __restrict__ db, long long rows_per_block,
                     unsigned* turn) {
  extern __shared__ float sm[];     // per warp: dw partial [dim], db partial [dim]
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float* s_dw = sm + (size_t)warp * 2 * dim;
  float* s_db = s_dw + dim;
  const int per = (dim + 31) / 32;
  const long long r0 = (long long)blockIdx.x * rows_per_block;
  const long long r1 = min(rows, r0 + rows_per_block);
  float adw[32], adb[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) { adw[i] = 0.f; adb[i] = 0.f; }
  for (long long row = r0 + warp; row < r1; row += 8) {
    const float mu = mean[row], rs = rstd[row];
    float g[32], xh[32];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int i = 0; i < 32; ++i) {
      if (i >= per) break;
      const int c = lane + 32 * i;
      if (c < dim) {
        const float d = dy[row * dim + c];
        xh[i] = (x[row * dim + c] - mu) * rs;
        g[i] = d * w[c];
        adw[i] = fmaf(d, xh[i], adw[i]);
        adb[i] += d;
      } else { xh[i] = 0.f; g[i] = 0.f; }
      s1 += g[i];
      s2 = fmaf(g[i], xh[i], s2);
    }
    s1 = sa_warp_sum(s1) / dim;
    s2 = sa_warp_sum(s2) / dim;
#pragma unroll
    for (int i = 0; i < 32; ++i) {
      if (i >= per) break;
      const int c = lane + 32 * i;
      if (c < dim) dx[row * dim + c] = rs * (g[i] - s1 - xh[i] * s2);
    }
  }
#pragma unroll
  for (int i = 0; i < 32; ++i) {
    if (i >= per) break;
    const int c = lane + 32 * i;
    if (c < dim) { s_dw[c] = adw[i]; s_db[c] = adb[i]; }
  }
  __syncthreads();
  // the eight warps' partials in warp order, the blocks in block order when the deterministic mode is on
  sa_block_turn_begin(turn, blockIdx.x);
  for (int c = threadIdx.x; c < dim; c += blockDim.x) {
    float tw = 0.f, tb = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) { tw += sm[(size_t)w * 2 * dim + c]; tb += sm[(size_t)w * 2 * dim + dim + c]; }
    atomicAdd(dw + c, tw);
    atomicAdd(db + c, tb);
  }
  sa_block_turn_end(turn, blockIdx.x);
}

// one warp per row
__global__ void __launch_bounds__(256)
ce_kernel(const float* logits, long long ld, const long long* __restrict__ target, long long rows, int V,
          float grad_scale, const float* __restrict__ grad_scale_dev, float* __restrict__ loss_sum, float* dlogits,
          float* __restrict__ partials) {
  __shared__ float s_loss[8];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const long long row = (long long)blockIdx.x * 8 + warp;
  float loss = 0.f;
  if (row < rows) {
    const float* lr = logits + row * ld;
    float mx = -INFINITY;
    for (int j = lane; j < V; j += 32) mx = fmaxf(mx, lr[j]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    float se = 0.f;
    for (int j = lane; j < V; j += 32) se += expf(lr[j] - mx);
    se = sa_warp_sum(se);
    const float lse = mx + logf(se);
    const long long tg = target[row];
    loss = lse - lr[tg];
    __syncwarp();   // dlogits may alias logits: every lane has read lr[tg] before any lane overwrites it
    if (dlogits) {
      if (grad_scale_dev) grad_scale *= __ldg(grad_scale_dev);
      float* dr = dlogits + row * ld;
      for (int j = lane; j < V; j += 32) {
        const float p = expf(lr[j] - lse);
        dr[j] = grad_scale * (p - (j == tg ? 1.0f : 0.f));
      }
    }
  }
  if (lane == 0) s_loss[warp] = loss;
  __syncthreads();
  if (threadIdx.x == 0 && loss_sum) {
    float s = 0.f;
    for (int w = 0; w < 8; ++w) s += s_loss[w];
    if (partials) partials[blockIdx.x] = s;      // deterministic mode: summed in block order by sa_ordered_sum
    else atomicAdd(loss_sum, s);
  }
}

// dst[r][c] = (TO) src[r][c] for c < cols, 0 for cols <= c < dst_ld
template <typename TI, typename TO>
__global__ void cast2d_kernel(const TI* __restrict__ src, long long src_ld, TO* __restrict__ dst, long long dst_ld,
                              long long rows, int cols) {
  const long long total = rows * dst_ld;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / dst_ld;
    const int c = (int)(i % dst_ld);
    sa_st(dst, i, c < cols ? sa_ld(src, r * src_ld + c) : 0.f);
  }
}

// ReZero bookkeeping of a sub-layer f(x) = core(x) + bias:  dg = dot + sum_c bias[c] * colsum[c];  dbias = g * colsum
__global__ void rezero_finish_kernel(const float* __restrict__ colsum, const float* __restrict__ bias,
                                     const float* __restrict__ g, const float* __restrict__ dot, int n,
                                     float* __restrict__ dbias, float* __restrict__ dg) {
  __shared__ float s_red[8];
  float acc = 0.f;
  const float gv = g[0];
  for (int c = threadIdx.x; c < n; c += blockDim.x) {
    const float cs = colsum[c];
    if (bias) acc = fmaf(bias[c], cs, acc);
    if (dbias) dbias[c] = gv * cs;
  }
  acc = sa_warp_sum(acc);
  if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float s = dot ? dot[0] : 0.f;
    for (int w = 0; w < 8; ++w) s += s_red[w];
    dg[0] = s;
  }
}

// ReZero gate gradient from the UNSCALED weight gradient T = dx^T a of the gated layer y = x + g (a W^T):
//   dg += sum_ij T[i][j] W[i][j]   ( = sum (dx W) . a: the same number as a dot product over the [rows x n] activations,
//                                    read off a [n_out x n_in] matrix instead)
//   T  *= g                        (T becomes the weight gradient)
__global__ void __launch_bounds__(256)
gate_wgrad_kernel(float* __restrict__ T, const float* __restrict__ W, long long n, const float* __restrict__ g,
                  float* __restrict__ dot, float* __restrict__ partials) {
  __shared__ float s_red[8];
  const float gv = g[0];
  float acc = 0.f;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float t = T[i];
    acc = fmaf(t, W[i], acc);
    T[i] = t * gv;
  }
  acc = sa_warp_sum(acc);
  if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float s = 0.f;
    for (int w = 0; w < 8; ++w) s += s_red[w];
    if (partials) partials[blockIdx.x] = s;
    else atomicAdd(dot, s);
  }
}

inline unsigned ew_grid(long long n) {
  long long b = sa_cdiv(n, 256);
  if (b > 148 * 32) b = 148 * 32;
  if (b < 1) b = 1;
  return (unsigned)b;
}

}  // namespace

extern "C" int sa_embed_fwd(const int64_t* tokens, const int32_t* sp_idx, int n_axes, const float* tok_w,
                            const float* const* sp_w, const float* pos_w, int batch, int seq, int dim, int num_tokens,
                            float* x_f32, void* x_act, int act_dtype, void* stream) {
  SA_CHECK_ARG(tokens && tok_w && pos_w && (x_f32 || x_act), "null pointer");
  SA_CHECK_ARG(batch > 0 && seq > 0 && dim > 0 && num_tokens > 0, "bad sizes");
  SA_CHECK_ARG(n_axes >= 0 && n_axes <= 3 && (n_axes == 0 || (sp_idx && sp_w)), "bad spatial axes");
  EmbPtrs sp = {};
  for (int a = 0; a < n_axes; ++a) sp.sp_w[a] = sp_w[a];
  const long long total = (long long)batch * seq * dim;
  if (act_dtype == SA_BF16)
    embed_fwd_kernel<__nv_bfloat16><<<ew_grid(total), 256, 0, sa_stream(stream)>>>(
        (const long long*)tokens, sp_idx, n_axes, tok_w, sp, pos_w, batch, seq, dim, x_f32, (__nv_bfloat16*)x_act);
  else
    embed_fwd_kernel<float><<<ew_grid(total), 256, 0, sa_stream(stream)>>>((const long long*)tokens, sp_idx, n_axes, tok_w,
                                                                           sp, pos_w, batch, seq, dim, x_f32, (float*)x_act);
  SA_LAUNCH_CHECK();
  return SA_OK;
}

extern "C" int sa_embed_bwd(const float* dx, const int64_t* tokens, const int32_t* sp_idx, int n_axes, int batch, int seq,
                            int dim, int num_tokens, const int32_t* sp_rows, float* d_tok_w, float* const* d_sp_w,
                            float* d_pos_w, void* stream) {
  SA_CHECK_ARG(dx && tokens && d_tok_w && d_pos_w, "null pointer");
  SA_CHECK_ARG(n_axes >= 0 && n_axes <= 3 && (n_axes == 0 || (sp_idx && d_sp_w && sp_rows)), "bad spatial axes");
  SA_CHECK_ARG(num_tokens > 0, "bad table size");
  cudaStream_t st = sa_stream(stream);
  EmbPtrs sp = {};
  for (int a = 0; a < n_axes; ++a) sp.d_sp_w[a] = d_sp_w[a];
  if (sa_deterministic() && dim <= 1024) {
    // positional table: one thread per element already adds in batch order; the token / spatial tables as ordered gathers
    embed_bwd_kernel<<<ew_grid((long long)seq * dim), 256, 0, st>>>(dx, (const long long*)tokens, sp_idx, 0, batch, seq, dim,
                                                                    nullptr, sp, d_pos_w);
    SA_LAUNCH_CHECK();
    embed_bwd_ordered_kernel<0><<<(unsigned)num_tokens, 256, 0, st>>>(dx, (const long long*)tokens, nullptr, batch, seq, dim,
                                                                     d_tok_w);
    SA_LAUNCH_CHECK();
    for (int a = 0; a < n_axes; ++a) {
      embed_bwd_ordered_kernel<1><<<(unsigned)sp_rows[a], 256, 0, st>>>(dx, nullptr, sp_idx + (size_t)a * seq, batch, seq, dim,
                                                                       d_sp_w[a]);
      SA_LAUNCH_CHECK();
    }
    return SA_OK;
  }
  embed_bwd_kernel<<<ew_grid((long long)seq * dim), 256, 0, st>>>(dx, (const long long*)tokens, sp_idx, n_axes, batch, seq,
                                                                  dim, d_tok_w, sp, d_pos_w);
  SA_LAUNCH_CHECK();
  return SA_OK;
}

extern "C" int sa_layernorm_fwd(const float* x, const float* w, const float* b, int64_t rows, int dim, float eps,
                                float* y_f32, void* y_act, int act_dtype, float* mean, float* rstd, void* stream) {
  SA_CHECK_ARG(x && w && b && mean && rstd && (y_f32 || y_act), "null pointer");
  SA_UNSUPPORTED(dim > 1024, "LayerNorm over more than 1024 channels");
  const unsigned grid = (unsigned)sa_cdiv(rows, 8);
  if (act_dtype == SA_BF16)
    layernorm_fwd_kernel<__nv_bfloat16><<<grid, 256, 0, sa_stream(stream)>>>(x, w, b, rows, dim, eps, y_f32,
                                                                             (__nv_bfloat16*)y_act, mean, rstd);
  else
    layernorm_fwd_kernel<float><<<grid, 256, 0, sa_stream(stream)>>>(x, w, b, rows, dim, eps, y_f32, (float*)y_act, mean,
                                                                     rstd);
  SA_LAUNCH_CHECK();
  return SA_OK;
}

extern "C" int sa_layernorm_bwd(const float* dy, const float* x, const float* w, const float* mean, const float* rstd,
                                int64_t rows, int dim, float* dx, float* dw, float* db, void* stream) {
  SA_CHECK_ARG(dy && x && w && mean && rstd && dx && dw && db, "null pointer");
  SA_UNSUPPORTED(dim > 1024, "LayerNorm over more than 1024 channels");
  long long blocks = sa_cdiv(rows, 64);
  if (blocks > 148 * 4) blocks = 148 * 4;
  const long long rpb = sa_cdiv(rows, blocks);
  blocks = sa_cdiv(rows, rpb);
  const size_t smem = (size_t)8 * 2 * dim * sizeof(float);
  if (smem > 48 * 1024)
    SA_CUDA(cudaFuncSetAttribute(layernorm_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  layernorm_bwd_kernel<<<(unsigned)blocks, 256, smem, sa_stream(stream)>>>(dy, x, w, mean, rstd, rows, dim, dx, dw, db, rpb,
                                                                           sa_turn_slot(1, sa_stream(stream)));
  SA_LAUNCH_CHECK();
  return SA_OK;
}

extern "C" int sa_ce_fwd_bwd(const float* logits, int64_t ld, const int64_t* target, int64_t rows, int vocab,
                             float grad_scale, const float* grad_scale_dev, float* loss_sum, float* dlogits, void* stream) {
  SA_CHECK_ARG(logits && target && rows > 0 && vocab > 0 && ld >= vocab, "bad arguments");
  const int64_t blocks = sa_cdiv(rows, 8);
  float* partials = loss_sum ? sa_partial_slot((int)blocks, sa_stream(stream)) : nullptr;
  ce_kernel<<<(unsigned)blocks, 256, 0, sa_stream(stream)>>>(logits, ld, (const long long*)target, rows, vocab, grad_scale,
                                                             grad_scale_dev, loss_sum, dlogits, partials);
  SA_LAUNCH_CHECK();
  if (partials) return sa_ordered_sum(partials, (int)blocks, loss_sum, sa_stream(stream));
  return SA_OK;
}

extern "C" int sa_cast2d(const void* src, int src_dtype, int64_t src_ld, void* dst, int dst_dtype, int64_t dst_ld,
                         int64_t rows, int cols, void* stream) {
  SA_CHECK_ARG(src && dst && rows > 0 && cols > 0 && src_ld >= cols && dst_ld >= cols, "bad arguments");
  const long long total = rows * dst_ld;
  cudaStream_t st = sa_stream(stream);
  using B = __nv_bfloat16;
  if (src_dtype == SA_F32 && dst_dtype == SA_BF16)
    cast2d_kernel<float, B><<<ew_grid(total), 256, 0, st>>>((const float*)src, src_ld, (B*)dst, dst_ld, rows, cols);
  else if (src_dtype == SA_F32 && dst_dtype == SA_F32)
    cast2d_kernel<float, float><<<ew_grid(total), 256, 0, st>>>((const float*)src, src_ld, (float*)dst, dst_ld, rows, cols);
  else if (src_dtype == SA_BF16 && dst_dtype == SA_F32)
    cast2d_kernel<B, float><<<ew_grid(total), 256, 0, st>>>((const B*)src, src_ld, (float*)dst, dst_ld, rows, cols);
  else if (src_dtype == SA_BF16 && dst_dtype == SA_BF16)
    cast2d_kernel<B, B><<<ew_grid(total), 256, 0, st>>>((const B*)src, src_ld, (B*)dst, dst_ld, rows, cols);
  else { sa_set_error("sa_cast2d: bad dtype"); return SA_ERR_INVALID; }
  SA_LAUNCH_CHECK();
  return SA_OK;
}

extern "C" int sa_gate_wgrad(float* t, const float* w, int64_t n, const float* g, float* dot, void* stream) {
  SA_CHECK_ARG(t && w && g && dot && n >= 0, "bad arguments");
  if (n == 0) return SA_OK;
  long long blocks = sa_cdiv(n, 256 * 4);
  if (blocks > 148 * 4) blocks = 148 * 4;
  float* partials = sa_partial_slot((int)blocks, sa_stream(stream));
  gate_wgrad_kernel<<<(unsigned)blocks, 256, 0, sa_stream(stream)>>>(t, w, n, g, dot, partials);
  SA_LAUNCH_CHECK();
  if (partials) return sa_ordered_sum(partials, (int)blocks, dot, sa_stream(stream));
  return SA_OK;
}

extern "C" int sa_rezero_finish(const float* colsum, const float* bias, const float* g, const float* dot, int n,
                                float* dbias, float* dg, void* stream) {
  SA_CHECK_ARG(g && dg && n >= 0 && (n == 0 || colsum), "bad arguments");
  rezero_finish_kernel<<<1, 256, 0, sa_stream(stream)>>>(colsum, bias, g, dot, n, dbias, dg);
  SA_LAUNCH_CHECK();
  return SA_OK;
}
