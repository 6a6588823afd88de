// Spectral (Jukebox) reconstruction loss: helpers around the dense DFT-matrix products.
//
// Replaces JukeboxLoss._get_fft_amplitude + F.mse_loss, /root/reference/src/losses/vqvae/vqvae.py:598-599, 617-630
// (torch.fft.fftn over the channel + spatial axes with norm="ortho" -> cuFFT in the reference).  The 160 x 224 x 160
// volume factors as 2^5.5 x 2^5.7 x 2^5.5: instead of mixed-radix FFT passes the three axis transforms are dense
// DFT-matrix products on the tensor cores (bf16x3 arithmetic, csrc/sa_x3.cu: 16 significand bits per operand, fp32
// accumulation), 0.17 TFLOP per batch-of-8 transform.  A complex tensor is stored as [..][part = re, im][axis] so that a
// complex axis transform is ONE real product with contraction length 2n (rows (re', k') = [C | S], rows (im', k') =
// [-S | C] of the matrix).  This file holds what is not a GEMM: the axis swap between two transforms and the fused
// amplitude / squared-difference / gradient pass.
#include "sa_common.cuh"

namespace {

// dst[b][c][m][a] = src[b][a][m][c]   (fp32; 32 x 32 tiles of the (a, c) plane through shared memory, both sides coalesced)
__global__ void __launch_bounds__(256)
swap_outer_inner_kernel(const float* __restrict__ src, float* __restrict__ dst, int A, int M, int C, long long bm_total) {
  __shared__ float tile[32][33];
  const int tx = threadIdx.x, ty = threadIdx.y;      // 32 x 8
  for (long long bm = blockIdx.z; bm < bm_total; bm += gridDim.z) {
    const long long b = bm / M;
    const int m = (int)(bm - b * M);
    const int a0 = blockIdx.y * 32, c0 = blockIdx.x * 32;
    const float* s = src + b * (long long)A * M * C;
    float* d = dst + b * (long long)A * M * C;
    for (int j = ty; j < 32; j += 8) {
      const int a = a0 + j, c = c0 + tx;
      tile[j][tx] = (a < A && c < C) ? s[((long long)a * M + m) * C + c] : 0.f;
    }
    __syncthreads();
    for (int j = ty; j < 32; j += 8) {
      const int c = c0 + j, a = a0 + tx;
      if (a < A && c < C) d[((long long)c * M + m) * A + a] = tile[tx][j];
    }
    __syncthreads();
  }
}

// spectra stored as [rows][2][L] (re | im).  amp = sqrt(re^2 + im^2) (vqvae.py:626-628).
//   sse[0]  += sum (amp_p - amp_t)^2                                       (forward; numerator of F.mse_loss, :599)
//   grad     = coef * coef_dev[0] * (amp_p - amp_t) * (re_p, im_p) / amp_p  (backward w.r.t. the prediction's spectrum)
__global__ void __launch_bounds__(256)
spectral_amp_kernel(const float* __restrict__ p, const float* __restrict__ t, long long rows, int L, float coef,
                    const float* __restrict__ coef_dev, float* __restrict__ sse, float* __restrict__ grad,
                    float* __restrict__ partials) {
  const long long total = rows * L;
  const float cf = coef * (coef_dev ? __ldg(coef_dev) : 1.0f);
  float acc = 0.f;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / L;
    const int l = (int)(i - r * L);
    const long long o = r * 2 * L + l;
    const float pr = p[o], pi = p[o + L], tr = t[o], ti = t[o + L];
    const float ap = sqrtf(pr * pr + pi * pi), at = sqrtf(tr * tr + ti * ti);
    const float df = ap - at;
    acc = fmaf(df, df, acc);
    if (grad) {
      const float g = ap > 0.f ? cf * df / ap : 0.f;
      grad[o] = g * pr;
      grad[o + L] = g * pi;
    }
  }
  if (sse) {
    acc = sa_warp_sum(acc);
    __shared__ float part[8];
    if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
      float s = 0.f;
      for (int w = 0; w < 8; ++w) s += part[w];
      if (partials) partials[blockIdx.x] = s;       // deterministic mode: summed in block order afterwards
      else atomicAdd(sse, s);
    }
  }
}

}  // namespace

extern "C" int sa_swap_outer_inner(const float* src, float* dst, int64_t batch, int A, int M, int C, void* stream) {
  SA_CHECK_ARG(src && dst && batch >= 0 && A > 0 && M > 0 && C > 0, "bad arguments");
  if (batch == 0) return SA_OK;
  const long long bm = (long long)batch * M;
  dim3 grid((unsigned)sa_cdiv(C, 32), (unsigned)sa_cdiv(A, 32), (unsigned)(bm < 32768 ? bm : 32768));
  SA_UNSUPPORTED(grid.y > 65535, "axis too long");
  swap_outer_inner_kernel<<<grid, dim3(32, 8), 0, sa_stream(stream)>>>(src, dst, A, M, C, bm);
  SA_LAUNCH_CHECK();
  return SA_OK;
}

extern "C" int sa_spectral_amp_loss(const float* pred_spec, const float* target_spec, int64_t rows, int L, float coef,
                                    const float* coef_dev, float* sse, float* grad, void* stream) {
  SA_CHECK_ARG(pred_spec && target_spec && rows >= 0 && L > 0 && (sse || grad), "bad arguments");
  if (rows == 0) return SA_OK;
  long long blocks = sa_cdiv(rows * (long long)L, 256);
  if (blocks > 148 * 16) blocks = 148 * 16;
  float* partials = sse ? sa_partial_slot((int)blocks, sa_stream(stream)) : nullptr;
  spectral_amp_kernel<<<(unsigned)blocks, 256, 0, sa_stream(stream)>>>(pred_spec, target_spec, rows, L, coef, coef_dev, sse, grad,
                                                                       partials);
  SA_LAUNCH_CHECK();
  if (partials) return sa_ordered_sum(partials, (int)blocks, sse, sa_stream(stream));
  return SA_OK;
}
