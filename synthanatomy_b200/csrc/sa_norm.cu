// BatchNorm3d (+ LeakyReLU) over channels-last activations [rows = B * D * H * W][C] for the PatchGAN discriminator
// (/root/reference/src/networks/discriminator/baseline.py:43-79: Conv3d -> BatchNorm3d -> LeakyReLU(0.2) blocks).
//
//   training forward   per-channel batch mean and biased variance in two passes (sum, then centred squares; fp32
//                      partials per thread, fp64 atomics across CTAs), running statistics updated with momentum and the
//                      unbiased variance as nn.BatchNorm3d does;  y = lrelu(gamma * (x - mean) * rstd + beta)
//   backward           g' = g * lrelu'(y);  dbeta = sum g';  dgamma = sum g' xhat;
//                      dx = gamma * rstd * (g' - dbeta / rows - xhat * dgamma / rows)
//
// All of it is HBM-bound: forward reads x twice (statistics) + once and writes y; backward reads g, x, y twice and
// writes dx.  Thread x = channel (coalesced along C), thread y = row inside the CTA's row stripe.
#include "sa_common.cuh"

namespace {

template <typename T> __device__ __forceinline__ float ldf(const T* p, long long i);
template <> __device__ __forceinline__ float ldf<float>(const float* p, long long i) { return p[i]; }
template <> __device__ __forceinline__ float ldf<__nv_bfloat16>(const __nv_bfloat16* p, long long i) { return __bfloat162float(p[i]); }
template <typename T> __device__ __forceinline__ void stf(T* p, long long i, float v);
template <> __device__ __forceinline__ void stf<float>(float* p, long long i, float v) { p[i] = v; }
template <> __device__ __forceinline__ void stf<__nv_bfloat16>(__nv_bfloat16* p, long long i, float v) { p[i] = __float2bfloat16(v); }

constexpr int NB_X = 32, NB_Y = 8;

// MODE 0: acc0 = sum x                      MODE 1: acc0 = sum (x - mean)^2
// MODE 2: acc0 = sum g', acc1 = sum g' xhat (g' = g * lrelu'(y))
template <typename T, int MODE>
__global__ void colreduce_kernel(const T* __restrict__ x, const T* __restrict__ g, const T* __restrict__ y, long long rows, int C,
                                 const float* __restrict__ mean, const float* __restrict__ rstd, float slope,
                                 double* __restrict__ out0, double* __restrict__ out1, unsigned* turn) {
  __shared__ float s0[NB_Y][NB_X + 1], s1[NB_Y][NB_X + 1];
  const int c = blockIdx.x * NB_X + threadIdx.x;
  float a0 = 0.f, a1 = 0.f;
  if (c < C) {
    const float mu = MODE ? mean[c] : 0.f;
    const float rs = MODE == 2 ? rstd[c] : 0.f;
    for (long long r = (long long)blockIdx.y * NB_Y + threadIdx.y; r < rows; r += (long long)gridDim.y * NB_Y) {
      const long long i = r * C + c;
      const float v = ldf(x, i);
      if (MODE == 0) a0 += v;
      if (MODE == 1) { const float d = v - mu; a0 = fmaf(d, d, a0); }
      if (MODE == 2) {
        const float gp = ldf(g, i) * (ldf(y, i) > 0.f ? 1.f : slope);
        a0 += gp;
        a1 = fmaf(gp, (v - mu) * rs, a1);
      }
    }
  }
  s0[threadIdx.y][threadIdx.x] = a0;
  s1[threadIdx.y][threadIdx.x] = a1;
  __syncthreads();
  unsigned* my_turn = turn ? turn + blockIdx.x : nullptr;      // deterministic mode: row slabs add in slab order
  sa_block_turn_begin(my_turn, blockIdx.y);
  if (threadIdx.y == 0 && c < C) {
    double t0 = 0.0, t1 = 0.0;
#pragma unroll
    for (int j = 0; j < NB_Y; ++j) { t0 += (double)s0[j][threadIdx.x]; t1 += (double)s1[j][threadIdx.x]; }
    atomicAdd(out0 + c, t0);
    if (MODE == 2) atomicAdd(out1 + c, t1);
  }
  sa_block_turn_end(my_turn, blockIdx.y);
}

__global__ void bn_mean_kernel(const double* __restrict__ sum, long long rows, int C, float* __restrict__ mean) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c < C) mean[c] = (float)(sum[c] / (double)rows);
}

__global__ void bn_finalize_kernel(const double* __restrict__ ss, long long rows, int C, float eps, float momentum,
                                   const float* __restrict__ mean, float* __restrict__ rstd, float* __restrict__ running_mean,
                                   float* __restrict__ running_var) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const double var = ss[c] / (double)rows;                       // biased: what normalises the batch
  rstd[c] = (float)(1.0 / sqrt(var + (double)eps));
  if (running_mean) running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * mean[c];
  if (running_var) {
    const double unbiased = rows > 1 ? ss[c] / (double)(rows - 1) : var;
    running_var[c] = (1.f - momentum) * running_var[c] + momentum * (float)unbiased;
  }
}

__global__ void bn_eval_stats_kernel(const float* __restrict__ running_mean, const float* __restrict__ running_var, int C, float eps,
                                     float* __restrict__ mean, float* __restrict__ rstd) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c < C) { mean[c] = running_mean[c]; rstd[c] = rsqrtf(running_var[c] + eps); }
}

template <typename T>
__global__ void bn_lrelu_fwd_kernel(const T* __restrict__ x, long long total, int C, const float* __restrict__ mean,
                                    const float* __restrict__ rstd, const float* __restrict__ gamma, const float* __restrict__ beta,
                                    float slope, T* __restrict__ y) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    const float v = fmaf((ldf(x, i) - mean[c]) * rstd[c], gamma[c], beta[c]);
    stf(y, i, v > 0.f ? v : v * slope);
  }
}

template <typename T>
__global__ void bn_lrelu_bwd_dx_kernel(const T* __restrict__ g, const T* __restrict__ x, const T* __restrict__ y, long long total, int C,
                                       long long rows, const float* __restrict__ mean, const float* __restrict__ rstd,
                                       const float* __restrict__ gamma, float slope, const double* __restrict__ sums,
                                       T* __restrict__ dx) {
  const double inv_rows = 1.0 / (double)rows;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    const float gp = ldf(g, i) * (ldf(y, i) > 0.f ? 1.f : slope);
    const float xh = (ldf(x, i) - mean[c]) * rstd[c];
    const float mb = (float)(sums[c] * inv_rows), mg = (float)(sums[C + c] * inv_rows);
    stf(dx, i, gamma[c] * rstd[c] * (gp - mb - xh * mg));
  }
}

__global__ void bn_param_grads_kernel(const double* __restrict__ sums, int C, float* __restrict__ dgamma, float* __restrict__ dbeta) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c < C) { dbeta[c] = (float)sums[c]; dgamma[c] = (float)sums[C + c]; }
}

template <typename T>
__global__ void lrelu_fwd_kernel(T* __restrict__ x, long long n, float slope) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float v = ldf(x, i);
    if (v <= 0.f) stf(x, i, v * slope);
  }
}
template <typename T>
__global__ void lrelu_bwd_kernel(T* __restrict__ g, const T* __restrict__ y, long long n, float slope) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    if (!(ldf(y, i) > 0.f)) stf(g, i, ldf(g, i) * slope);
}

unsigned ew_blocks(long long total) {
  long long b = (total + 255) / 256;
  if (b > 148 * 32) b = 148 * 32;
  return (unsigned)(b < 1 ? 1 : b);
}
dim3 red_grid(long long rows, int C) {
  const unsigned gx = (unsigned)((C + NB_X - 1) / NB_X);
  long long gy = (rows + NB_Y * 16 - 1) / (NB_Y * 16);            // >= 16 rows per thread before another CTA is worth it
  const long long cap = (148 * 8 + gx - 1) / gx;
  if (gy > cap) gy = cap;
  if (gy < 1) gy = 1;
  return dim3(gx, (unsigned)gy);
}

template <typename T>
int bn_stats_t(const T* x, long long rows, int C, double* ws, float eps, float momentum, float* mean, float* rstd,
               float* running_mean, float* running_var, cudaStream_t st) {
  SA_CUDA(cudaMemsetAsync(ws, 0, sizeof(double) * 2 * C, st));
  const dim3 blk(NB_X, NB_Y), grd = red_grid(rows, C);
  colreduce_kernel<T, 0><<<grd, blk, 0, st>>>(x, nullptr, nullptr, rows, C, nullptr, nullptr, 0.f, ws, nullptr, sa_turn_slot((int)grd.x, st));
  SA_LAUNCH_CHECK();
  bn_mean_kernel<<<(C + 127) / 128, 128, 0, st>>>(ws, rows, C, mean);
  SA_LAUNCH_CHECK();
  colreduce_kernel<T, 1><<<grd, blk, 0, st>>>(x, nullptr, nullptr, rows, C, mean, nullptr, 0.f, ws + C, nullptr, sa_turn_slot((int)grd.x, st));
  SA_LAUNCH_CHECK();
  bn_finalize_kernel<<<(C + 127) / 128, 128, 0, st>>>(ws + C, rows, C, eps, momentum, mean, rstd, running_mean, running_var);
  SA_LAUNCH_CHECK();
  return SA_OK;
}

template <typename T>
int bn_bwd_t(const T* g, const T* x, const T* y, long long rows, int C, const float* mean, const float* rstd, const float* gamma,
             float slope, double* ws, float* dgamma, float* dbeta, T* dx, cudaStream_t st) {
  SA_CUDA(cudaMemsetAsync(ws, 0, sizeof(double) * 2 * C, st));
  const dim3 grd2 = red_grid(rows, C);
  colreduce_kernel<T, 2><<<grd2, dim3(NB_X, NB_Y), 0, st>>>(x, g, y, rows, C, mean, rstd, slope, ws, ws + C,
                                                            sa_turn_slot((int)grd2.x, st));
  SA_LAUNCH_CHECK();
  bn_param_grads_kernel<<<(C + 127) / 128, 128, 0, st>>>(ws, C, dgamma, dbeta);
  SA_LAUNCH_CHECK();
  if (dx) {
    bn_lrelu_bwd_dx_kernel<T><<<ew_blocks(rows * C), 256, 0, st>>>(g, x, y, rows * C, C, rows, mean, rstd, gamma, slope, ws, dx);
    SA_LAUNCH_CHECK();
  }
  return SA_OK;
}

}  // namespace

extern "C" {

size_t sa_bn_workspace(int channels) { return sizeof(double) * 2 * (size_t)(channels > 0 ? channels : 0); }

int sa_bn_stats(const void* x, int dtype, int64_t rows, int channels, void* workspace, float eps, float momentum, float* mean,
                float* rstd, float* running_mean, float* running_var, void* stream) {
  SA_CHECK_ARG(x && workspace && mean && rstd, "null pointer");
  SA_CHECK_ARG(rows > 0 && channels > 0, "empty tensor");
  SA_CHECK_ARG(dtype == SA_F32 || dtype == SA_BF16, "bad dtype");
  sa_note_path(SA_PATH_SIMT);
  if (dtype == SA_F32)
    return bn_stats_t((const float*)x, rows, channels, (double*)workspace, eps, momentum, mean, rstd, running_mean, running_var,
                      sa_stream(stream));
  return bn_stats_t((const __nv_bfloat16*)x, rows, channels, (double*)workspace, eps, momentum, mean, rstd, running_mean,
                    running_var, sa_stream(stream));
}

int sa_bn_eval_stats(const float* running_mean, const float* running_var, int channels, float eps, float* mean, float* rstd,
                     void* stream) {
  SA_CHECK_ARG(running_mean && running_var && mean && rstd, "null pointer");
  SA_CHECK_ARG(channels > 0, "no channels");
  sa_note_path(SA_PATH_SIMT);
  bn_eval_stats_kernel<<<(channels + 127) / 128, 128, 0, sa_stream(stream)>>>(running_mean, running_var, channels, eps, mean, rstd);
  SA_LAUNCH_CHECK();
  return SA_OK;
}

int sa_bn_lrelu_fwd(const void* x, int dtype, int64_t rows, int channels, const float* mean, const float* rstd, const float* gamma,
                    const float* beta, float slope, void* y, void* stream) {
  SA_CHECK_ARG(x && y && mean && rstd && gamma && beta, "null pointer");
  SA_CHECK_ARG(rows > 0 && channels > 0, "empty tensor");
  SA_CHECK_ARG(dtype == SA_F32 || dtype == SA_BF16, "bad dtype");
  sa_note_path(SA_PATH_SIMT);
  const long long total = (long long)rows * channels;
  if (dtype == SA_F32)
    bn_lrelu_fwd_kernel<float><<<ew_blocks(total), 256, 0, sa_stream(stream)>>>((const float*)x, total, channels, mean, rstd, gamma,
                                                                              beta, slope, (float*)y);
  else
    bn_lrelu_fwd_kernel<__nv_bfloat16><<<ew_blocks(total), 256, 0, sa_stream(stream)>>>(
        (const __nv_bfloat16*)x, total, channels, mean, rstd, gamma, beta, slope, (__nv_bfloat16*)y);
  SA_LAUNCH_CHECK();
  return SA_OK;
}

int sa_bn_lrelu_bwd(const void* g, const void* x, const void* y, int dtype, int64_t rows, int channels, const float* mean,
                    const float* rstd, const float* gamma, float slope, void* workspace, float* dgamma, float* dbeta, void* dx,
                    void* stream) {
  SA_CHECK_ARG(g && x && y && mean && rstd && gamma && workspace && dgamma && dbeta, "null pointer");
  SA_CHECK_ARG(rows > 0 && channels > 0, "empty tensor");
  SA_CHECK_ARG(dtype == SA_F32 || dtype == SA_BF16, "bad dtype");
  sa_note_path(SA_PATH_SIMT);
  if (dtype == SA_F32)
    return bn_bwd_t((const float*)g, (const float*)x, (const float*)y, rows, channels, mean, rstd, gamma, slope, (double*)workspace,
                    dgamma, dbeta, (float*)dx, sa_stream(stream));
  return bn_bwd_t((const __nv_bfloat16*)g, (const __nv_bfloat16*)x, (const __nv_bfloat16*)y, rows, channels, mean, rstd, gamma,
                  slope, (double*)workspace, dgamma, dbeta, (__nv_bfloat16*)dx, sa_stream(stream));
}

int sa_lrelu_fwd(void* x, int dtype, int64_t n, float slope, void* stream) {
  SA_CHECK_ARG(x, "null pointer");
  SA_CHECK_ARG(n > 0, "empty tensor");
  SA_CHECK_ARG(dtype == SA_F32 || dtype == SA_BF16, "bad dtype");
  sa_note_path(SA_PATH_SIMT);
  if (dtype == SA_F32) lrelu_fwd_kernel<float><<<ew_blocks(n), 256, 0, sa_stream(stream)>>>((float*)x, n, slope);
  else lrelu_fwd_kernel<__nv_bfloat16><<<ew_blocks(n), 256, 0, sa_stream(stream)>>>((__nv_bfloat16*)x, n, slope);
  SA_LAUNCH_CHECK();
  return SA_OK;
}

int sa_lrelu_bwd(void* g, const void* y, int dtype, int64_t n, float slope, void* stream) {
  SA_CHECK_ARG(g && y, "null pointer");
  SA_CHECK_ARG(n > 0, "empty tensor");
  SA_CHECK_ARG(dtype == SA_F32 || dtype == SA_BF16, "bad dtype");
  sa_note_path(SA_PATH_SIMT);
  if (dtype == SA_F32) lrelu_bwd_kernel<float><<<ew_blocks(n), 256, 0, sa_stream(stream)>>>((float*)g, (const float*)y, n, slope);
  else
    lrelu_bwd_kernel<__nv_bfloat16><<<ew_blocks(n), 256, 0, sa_stream(stream)>>>((__nv_bfloat16*)g, (const __nv_bfloat16*)y, n, slope);
  SA_LAUNCH_CHECK();
  return SA_OK;
}

}  // extern "C"
