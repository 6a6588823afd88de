// Fused EMA vector-quantiser kernels.
//
// Replaces Quantizer_impl.forward (src/networks/vqvae/baseline.py:38-87): the reference materialises the
// [rows, n_embed] distance matrix and a one-hot of the same size (2 x 92 MB at rows = 11200), then runs
// ~15 small ATen kernels.  Here one kernel computes distance + argmin + gather + cluster statistics +
// commitment-loss numerator, never materialising either matrix, and a second tiny kernel does the EMA.
#include "sa_common.cuh"

namespace {

// 16 lanes cooperate on one latent row (each scans every 16th code, in ascending order): 11 200 rows (config 2 / 3) are
// 700 blocks, ~5 per SM, and a thread's serial chain is 128 codes x 32 FMAs.  (With 4 lanes per row the same launch was
// 175 blocks -- barely one 8-warp block per SM -- and took 171 us; the arithmetic and its order are unchanged.)
constexpr int VQ_ROWS_PER_BLOCK = 16;
constexpr int VQ_LANES_PER_ROW = 16;
constexpr int VQ_THREADS = VQ_ROWS_PER_BLOCK * VQ_LANES_PER_ROW;
constexpr int VQ_CODE_CHUNK = 256;      // codes staged in shared memory per pass

// ||w_k||^2 with the same sequential fp32 (mul, then add) association for every code, so duplicated codebook
// rows get bit-identical distances and the lowest index wins (baseline.py:52,56).
template <int DIM>
__global__ void __launch_bounds__(VQ_THREADS)
vq_forward_kernel(const float* __restrict__ z, const float* __restrict__ cb, int64_t rows, int n_embed,
                  int64_t* __restrict__ idx_out, float* __restrict__ q_out, int straight_through,
                  float* __restrict__ counts, float* __restrict__ dw, float* __restrict__ sse) {
  extern __shared__ float smem[];
  float* s_cb = smem;                          // [VQ_CODE_CHUNK][DIM + 1]  (+1: conflict-free row reads)
  float* s_w2 = smem + VQ_CODE_CHUNK * (DIM + 1);  // [VQ_CODE_CHUNK]
  __shared__ float s_red[VQ_THREADS / 32];

  const int t = threadIdx.x;
  const int sub = t & (VQ_LANES_PER_ROW - 1);
  const int64_t row = (int64_t)blockIdx.x * VQ_ROWS_PER_BLOCK + (t / VQ_LANES_PER_ROW);
  const bool valid = row < rows;

  float x[DIM];
  float x2 = 0.f;
#pragma unroll
  for (int c = 0; c < DIM; ++c) {
    x[c] = valid ? z[row * DIM + c] : 0.f;
    x2 = __fadd_rn(x2, __fmul_rn(x[c], x[c]));            // (flat ** 2).sum(1)   baseline.py:50
  }

  // best_k starts at a VALID code: a row of NaNs (a diverged step) compares false everywhere and must still index
  // inside the codebook (torch.max over an all-NaN row also returns an in-range index); lanes start at their own
  // first code so that the lane merge below keeps "lowest index on ties" for ordinary rows.
  float best = INFINITY;
  int best_k = sub < n_embed ? sub : 0;
  for (int k0 = 0; k0 < n_embed; k0 += VQ_CODE_CHUNK) {
    const int nk = min(VQ_CODE_CHUNK, n_embed - k0);
    __syncthreads();
    for (int i = t; i < nk * DIM; i += VQ_THREADS) {
      const int k = i / DIM, c = i - k * DIM;
      s_cb[k * (DIM + 1) + c] = cb[(int64_t)(k0 + k) * DIM + c];
    }
    __syncthreads();
    for (int k = t; k < nk; k += VQ_THREADS) {
      float w2 = 0.f;
#pragma unroll
      for (int c = 0; c < DIM; ++c) {
        const float w = s_cb[k * (DIM + 1) + c];
        w2 = __fadd_rn(w2, __fmul_rn(w, w));               // (weight ** 2).sum(1)  baseline.py:52
      }
      s_w2[k] = w2;
    }
    __syncthreads();
    // each of the 4 lanes of a row scans codes k = sub, sub+4, ... ; ascending k inside a lane
    for (int k = sub; k < nk; k += VQ_LANES_PER_ROW) {
      const float* w = s_cb + k * (DIM + 1);
      float dot = 0.f;
#pragma unroll
      for (int c = 0; c < DIM; ++c) dot = fmaf(x[c], w[c], dot);   // torch.mm row . column
      // (x2 - 2*dot) + w2, baseline.py:49-53
      const float d = __fadd_rn(__fsub_rn(x2, __fmul_rn(2.f, dot)), s_w2[k]);
      if (d < best) { best = d; best_k = k0 + k; }   // strict '<' keeps the lowest index inside a lane
    }
  }
  // combine the 4 lanes of a row: smaller distance, then smaller index
#pragma unroll
  for (int o = 1; o < VQ_LANES_PER_ROW; o <<= 1) {
    const float ob = __shfl_xor_sync(0xffffffffu, best, o);
    const int ok = __shfl_xor_sync(0xffffffffu, best_k, o);
    if (ob < best || (ob == best && ok < best_k)) { best = ob; best_k = ok; }
  }

  float local_sse = 0.f;
  if (valid) {
    if (sub == 0) {
      idx_out[row] = (int64_t)best_k;
      if (counts) atomicAdd(counts + best_k, 1.0f);
    }
    // the 4 lanes split the DIM channels of the row for the gather / statistics
#pragma unroll
    for (int c = 0; c < DIM; ++c) {
      if ((c & (VQ_LANES_PER_ROW - 1)) != sub) continue;
      const float qv = cb[(int64_t)best_k * DIM + c];
      const float df = __fsub_rn(qv, x[c]);
      // straight-through estimator value (quantized - x).detach() + x, two fp32 roundings (baseline.py:85)
      if (q_out) q_out[row * DIM + c] = straight_through ? __fadd_rn(df, x[c]) : qv;
      if (dw) atomicAdd(dw + (int64_t)best_k * DIM + c, x[c]);
      local_sse = fmaf(df, df, local_sse);
    }
  }
  if (sse) {
    local_sse = sa_warp_sum(local_sse);
    if ((t & 31) == 0) s_red[t >> 5] = local_sse;
    __syncthreads();
    if (t == 0) {
      float s = 0.f;
      for (int i = 0; i < VQ_THREADS / 32; ++i) s += s_red[i];
      atomicAdd(sse, s);
    }
  }
}

// deterministic mode: dw[k][c] = sum of z[row][c] over the rows quantised to code k (encodings.T @ flat, baseline.py:72),
// one block per code, every partial sum in a fixed row order and a fixed tree across the row lanes
template <int DIM>
__global__ void __launch_bounds__(256)
vq_dw_ordered_kernel(const float* __restrict__ z, const int64_t* __restrict__ idx, int64_t rows, float* __restrict__ dw) {
  constexpr int LANES = 256 / DIM;
  __shared__ float s_part[256];
  const int k = blockIdx.x;
  const int c = threadIdx.x % DIM, rl = threadIdx.x / DIM;
  float acc = 0.f;
  for (int64_t r = rl; r < rows; r += LANES)
    if (idx[r] == (int64_t)k) acc += z[r * DIM + c];
  s_part[threadIdx.x] = acc;
  __syncthreads();
  if (rl == 0) {
    float s = 0.f;
    for (int j = 0; j < LANES; ++j) s += s_part[j * DIM + c];
    dw[(int64_t)k * DIM + c] += s;
  }
}

// deterministic mode: the loss numerator sum (codebook[idx] - z)^2 as one partial per block (fixed tree inside the block),
// added in block order by sa_ordered_sum.  A separate kernel on purpose: any ordered-sum code inside vq_forward_kernel --
// a turnstile, a second store target, even an index computation for the atomic -- moved it from 127 to 155-170 us per launch
// at config 3 with the mode switched OFF (the kernel is a latency-bound 32-FMA chain per code; its code layout matters).
template <int DIM>
__global__ void __launch_bounds__(256)
vq_sse_ordered_kernel(const float* __restrict__ z, const float* __restrict__ cb, const int64_t* __restrict__ idx,
                      int64_t rows, float* __restrict__ partials) {
  constexpr int RPB = 256 / DIM;
  __shared__ float s_w[8];
  const int c = threadIdx.x % DIM;
  const int64_t row = (int64_t)blockIdx.x * RPB + threadIdx.x / DIM;
  float v = 0.f;
  if (row < rows) {
    const float df = __fsub_rn(cb[idx[row] * DIM + c], z[row * DIM + c]);
    v = df * df;
  }
  v = sa_warp_sum(v);
  if ((threadIdx.x & 31) == 0) s_w[threadIdx.x >> 5] = v;
  __syncthreads();
  if (threadIdx.x == 0) {
    float s = 0.f;
    for (int i = 0; i < 8; ++i) s += s_w[i];
    partials[blockIdx.x] = s;
  }
}

// one block: N/embed_avg EMA, n = sum N, Laplace smoothing, codebook refresh (baseline.py:75-80)
__global__ void __launch_bounds__(1024)
vq_ema_kernel(float* __restrict__ N, float* __restrict__ embed_avg, float* __restrict__ cb,
              const float* __restrict__ counts, const float* __restrict__ dw, int n_embed, int dim, float decay,
              float one_minus_decay, float eps, float k_eps, float* __restrict__ ws) {
  __shared__ float s_part[32];
  __shared__ float s_n;
  const int t = threadIdx.x;
  float part = 0.f;
  for (int k = t; k < n_embed; k += blockDim.x) {
    // N.mul_(decay).add_(encodings_sum * (1 - decay))
    const float nk = __fadd_rn(__fmul_rn(N[k], decay), __fmul_rn(counts[k], one_minus_decay));
    N[k] = nk;
    part += nk;
  }
  part = sa_warp_sum(part);
  if ((t & 31) == 0) s_part[t >> 5] = part;
  __syncthreads();
  if (t == 0) {
    float n = 0.f;
    for (int i = 0; i < (blockDim.x + 31) / 32; ++i) n += s_part[i];
    s_n = n;
    if (ws) ws[0] = n;
  }
  __syncthreads();
  const float n = s_n;
  const float denom = n + k_eps;
  for (int i = t; i < n_embed * dim; i += blockDim.x) {
    const int k = i / dim;
    const float ea = __fadd_rn(__fmul_rn(embed_avg[i], decay), __fmul_rn(dw[i], one_minus_decay));
    embed_avg[i] = ea;
    const float W = (N[k] + eps) / denom * n;     // (N + eps) / (n + K eps) * n
    cb[i] = ea / W;
  }
}

__global__ void vq_embed_kernel(const int64_t* __restrict__ idx, const float* __restrict__ cb, int64_t rows, int dim,
                                int n_embed, float* __restrict__ q) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows * dim) return;
  const int64_t r = i / dim;
  const int c = (int)(i - r * dim);
  int64_t k = idx[r];
  k = k < 0 ? 0 : (k >= n_embed ? n_embed - 1 : k);
  q[i] = cb[k * dim + c];
}

// dz = g_q + g_loss * coef * (z - q):  straight-through gradient plus the commitment term
// d/dz [beta * mse(q.detach(), z)] = 2 beta (z - q) / numel   (baseline.py:82-85)
__global__ void vq_backward_kernel(const float* __restrict__ g_q, const float* __restrict__ g_loss,
                                   const float* __restrict__ z, const float* __restrict__ q, float coef, int64_t n,
                                   float* __restrict__ dz) {
  const float gl = g_loss ? g_loss[0] * coef : 0.f;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float gq = g_q ? g_q[i] : 0.f;
    dz[i] = fmaf(gl, z[i] - q[i], gq);
  }
}

// perplexity = exp(-sum p log(p + 1e-10)), p = counts / total   (Quantizer.forward, baseline.py:110-120)
__global__ void __launch_bounds__(1024)
vq_perplexity_kernel(const float* __restrict__ counts, int n_embed, float total, float* __restrict__ out) {
  __shared__ float s_part[32];
  float acc = 0.f;
  for (int k = threadIdx.x; k < n_embed; k += blockDim.x) {
    const float p = counts[k] / total;
    acc += p * logf(p + 1e-10f);
  }
  acc = sa_warp_sum(acc);
  if ((threadIdx.x & 31) == 0) s_part[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float s = 0.f;
    for (int i = 0; i < (blockDim.x + 31) / 32; ++i) s += s_part[i];
    out[0] = expf(-s);
  }
}

}  // namespace

extern "C" int sa_vq_forward(const float* z, const float* codebook, int64_t rows, int dim, int n_embed, int64_t* idx,
                             float* q, int straight_through, float* counts, float* dw, float* sse, void* stream) {
  SA_CHECK_ARG(z && codebook && idx, "null pointer");
  SA_CHECK_ARG(rows >= 0 && n_embed > 0, "bad sizes");
  SA_UNSUPPORTED(!(dim == 8 || dim == 16 || dim == 32 || dim == 64), "embedding dim must be 8, 16, 32 or 64");
  if (rows == 0) return SA_OK;
  cudaStream_t st = sa_stream(stream);
  const unsigned grid = (unsigned)sa_cdiv(rows, VQ_ROWS_PER_BLOCK);
  const size_t smem = (size_t)(VQ_CODE_CHUNK * (dim + 1) + VQ_CODE_CHUNK) * sizeof(float);
  // deterministic mode: the forward kernel only counts (sums of ones: exact in fp32 below 2^24 rows per code, hence
  // order-free); the per-code sums of z and the loss sum come from two ordered kernels after it
  const bool det = sa_deterministic();
  float* dw_ordered = det ? dw : nullptr;
  float* sse_ordered = det ? sse : nullptr;
  if (det) { dw = nullptr; sse = nullptr; }
  const int sse_blocks = (int)sa_cdiv(rows, 256 / dim);
  float* sse_parts = sse_ordered ? sa_partial_slot(sse_blocks, st) : nullptr;
#define SA_VQ_LAUNCH(D)                                                                                            \
  do {                                                                                                             \
    SA_CUDA(cudaFuncSetAttribute(vq_forward_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));   \
    vq_forward_kernel<D><<<grid, VQ_THREADS, smem, st>>>(z, codebook, rows, n_embed, idx, q, straight_through, counts, \
                                                         dw, sse);                                                     \
    if (dw_ordered) vq_dw_ordered_kernel<D><<<(unsigned)n_embed, 256, 0, st>>>(z, idx, rows, dw_ordered);          \
    if (sse_parts) vq_sse_ordered_kernel<D><<<(unsigned)sse_blocks, 256, 0, st>>>(z, codebook, idx, rows, sse_parts); \
  } while (0)
  switch (dim) {
    case 8: SA_VQ_LAUNCH(8); break;
    case 16: SA_VQ_LAUNCH(16); break;
    case 32: SA_VQ_LAUNCH(32); break;
    default: SA_VQ_LAUNCH(64); break;
  }
#undef SA_VQ_LAUNCH
  SA_LAUNCH_CHECK();
  if (sse_parts) return sa_ordered_sum(sse_parts, sse_blocks, sse_ordered, st);
  return SA_OK;
}

extern "C" int sa_vq_ema_update(float* N, float* embed_avg, float* codebook, const float* counts, const float* dw,
                                int n_embed, int dim, double decay, double eps, float* workspace, void* stream) {
  SA_CHECK_ARG(N && embed_avg && codebook && counts && dw, "null pointer");
  SA_CHECK_ARG(n_embed > 0 && dim > 0, "bad sizes");
  // python scalars are doubles in the reference: `1 - decay` and `n_embed * eps` are formed in double and only
  // then rounded to fp32 by the tensor op (baseline.py:75-79)
  vq_ema_kernel<<<1, 1024, 0, sa_stream(stream)>>>(N, embed_avg, codebook, counts, dw, n_embed, dim, (float)decay,
                                                   (float)(1.0 - decay), (float)eps, (float)((double)n_embed * eps),
                                                   workspace);
  SA_LAUNCH_CHECK();
  return SA_OK;
}

extern "C" int sa_vq_embed(const int64_t* idx, const float* codebook, int64_t rows, int dim, int n_embed, float* q,
                           void* stream) {
  SA_CHECK_ARG(idx && codebook && q, "null pointer");
  if (rows == 0) return SA_OK;
  const int64_t n = rows * dim;
  vq_embed_kernel<<<(unsigned)sa_cdiv(n, 256), 256, 0, sa_stream(stream)>>>(idx, codebook, rows, dim, n_embed, q);
  SA_LAUNCH_CHECK();
  return SA_OK;
}

extern "C" int sa_vq_backward(const float* g_q, const float* g_loss, const float* z, const float* q, float coef,
                              int64_t n, float* dz, void* stream) {
  SA_CHECK_ARG(z && q && dz && n >= 0, "bad arguments");
  if (n == 0) return SA_OK;
  int64_t blocks = sa_cdiv(n, 256);
  if (blocks > 148 * 8) blocks = 148 * 8;
  vq_backward_kernel<<<(unsigned)blocks, 256, 0, sa_stream(stream)>>>(g_q, g_loss, z, q, coef, n, dz);
  SA_LAUNCH_CHECK();
  return SA_OK;
}

extern "C" int sa_vq_perplexity(const float* counts, int n_embed, float total, float* out, void* stream) {
  SA_CHECK_ARG(counts && out && n_embed > 0 && total > 0.f, "bad arguments");
  vq_perplexity_kernel<<<1, 1024, 0, sa_stream(stream)>>>(counts, n_embed, total, out);
  SA_LAUNCH_CHECK();
  return SA_OK;
}
