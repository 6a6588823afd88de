// bf16x3: fp32-class products on the bf16 tensor cores ("parity" arithmetic of the tcgen05 path).
//
// An fp32 operand x is split into two bf16 numbers, hi = bf16(x) and lo = bf16(x - hi); hi + lo carries 16 significand
// bits of x (|x - hi - lo| <= 2^-17 |x|).  A product a.b summed over a contraction index is then
//     sum a b  ~=  sum ah bh + sum al bh + sum ah bl        (the dropped al.bl term is <= 2^-16 relative)
// and the three sums are ONE contraction that is three times as long:
//     A' = [ ah | al | ah ],   B' = [ bh | bh | bl ]   =>   A' . B'  =  ah.bh + al.bh + ah.bl
// so the bf16 kernels of this library -- GEMM (contraction = columns), weight-gradient GEMM (contraction = rows), conv
// (contraction = input channels), conv weight gradient (contraction = batch x positions) -- compute it unchanged, with
// fp32 accumulation in TMEM, from operands that the split kernels below write.  Cost: 3x the tensor-core work of the
// bf16 path plus one streaming pass per operand; error ~1e-5 of the operand scale (TF32 would give ~5e-4).
//
// Reference call sites: the same nn.Linear / nn.Conv3d / nn.ConvTranspose3d calls as sa_gemm_nt / sa_conv3d_fwd
// (/root/reference/src/networks/vqvae/baseline.py:150-160, 213-299; performer.py:194-221,286), run by the reference in
// fp32 / TF32 (run_transformer.py:165 amp=False).
#include <algorithm>

#include "sa_pf_common.cuh"

bool sa_tc_gemm_nt_supported(int64_t m, int n, int k, int dtype, const void* a, int64_t lda, const void* b, int64_t ldb);
int sa_tc_gemm_nt(int64_t, int, int, const void*, int64_t, const void*, int64_t, const SaEpi&, cudaStream_t);
bool sa_tc_gemm_tn_supported(int64_t m, int na, int nb, int dtype, const void* a, int64_t lda, const void* b, int64_t ldb);
int sa_tc_gemm_tn(int64_t, int, int, const void*, int64_t, const void*, int64_t, const float*, float, float*, cudaStream_t);
bool sa_tc_conv3_supported(const sa_conv_desc*);
int sa_tc_conv3_fwd_ex(const sa_conv_desc*, const void*, const void*, const float*, const void*, const void*, int, void*,
                       bool, cudaStream_t);
bool sa_tc_wgrad3_supported(const sa_conv_desc*);
int sa_tc_conv3d_wgrad3(const sa_conv_desc*, const void*, const void*, float*, cudaStream_t);
bool sa_tc_wgrad_supported(const sa_conv_desc*);
int sa_tc_conv3d_wgrad(const sa_conv_desc*, const void*, const void*, float*, cudaStream_t);

namespace {

constexpr int X3_THREADS = 256;

__device__ __forceinline__ void split2(float x, __nv_bfloat16& hi, __nv_bfloat16& lo) {
  hi = __float2bfloat16_rn(x);
  lo = __float2bfloat16_rn(x - __bfloat162float(hi));
}

// dst [rows][3 * kp]: thirds (hi | lo | hi) for ORDER 0, (hi | hi | lo) for ORDER 1; columns [k, kp) of each third are 0.
// One thread per column PAIR (kp is even): 8-byte read, three 4-byte writes.
template <int ORDER>
__global__ void __launch_bounds__(X3_THREADS)
split_cols_kernel(const float* __restrict__ src, long long ld, long long rows, int k, int kp,
                  __nv_bfloat16* __restrict__ dst) {
  const int half = kp >> 1;
  const long long total = rows * half;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / half;
    const int c = (int)(i - r * half) * 2;
    const float x0 = c < k ? src[r * ld + c] : 0.f;
    const float x1 = c + 1 < k ? src[r * ld + c + 1] : 0.f;
    __nv_bfloat162 hi, lo;
    split2(x0, hi.x, lo.x);
    split2(x1, hi.y, lo.y);
    __nv_bfloat162* d = reinterpret_cast<__nv_bfloat162*>(dst + r * 3LL * kp + c);
    d[0] = hi;
    d[half] = ORDER == 0 ? lo : hi;
    d[2 * half] = ORDER == 0 ? hi : lo;
  }
}

// dst [3 * rows][kp]: row blocks (hi ; lo ; hi) for ORDER 0, (hi ; hi ; lo) for ORDER 1
template <int ORDER>
__global__ void __launch_bounds__(X3_THREADS)
split_rows_kernel(const float* __restrict__ src, long long ld, long long rows, int k, int kp,
                  __nv_bfloat16* __restrict__ dst) {
  const int half = kp >> 1;
  const long long total = rows * half;
  const long long blk = rows * (long long)half;         // bfloat162 elements per row block
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / half;
    const int c = (int)(i - r * half) * 2;
    const float x0 = c < k ? src[r * ld + c] : 0.f;
    const float x1 = c + 1 < k ? src[r * ld + c + 1] : 0.f;
    __nv_bfloat162 hi, lo;
    split2(x0, hi.x, lo.x);
    split2(x1, hi.y, lo.y);
    __nv_bfloat162* d = reinterpret_cast<__nv_bfloat162*>(dst) + i;
    d[0] = hi;
    d[blk] = ORDER == 0 ? lo : hi;
    d[2 * blk] = ORDER == 0 ? hi : lo;
  }
}

// the fused epilogue of sa_gemm_nt on an fp32 product tile (fp32 epilogue tensors): used when the call names tensors in
// the activation dtype (pre / dot_with / out_act), which the bf16 tensor-core kernel would read and write as bf16
__global__ void __launch_bounds__(X3_THREADS)
x3_epilogue_kernel(const float* __restrict__ v, long long m, int n, SaEpi e) {
  const float st = e.scale * (e.scale_dev ? __ldg(e.scale_dev) : 1.0f);
  const long long total = m * n;
  float dot = 0.f;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / n;
    const int c = (int)(i - r * n);
    dot += sa_epi_elem<float>(e, r, c, v[i], st);
  }
  if (e.dot_out) {
    dot = sa_warp_sum(dot);
    __shared__ float part[X3_THREADS / 32];
    if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = dot;
    __syncthreads();
    if (threadIdx.x == 0) {
      float s = 0.f;
      for (int w = 0; w < X3_THREADS / 32; ++w) s += part[w];
      atomicAdd(e.dot_out, s);
    }
  }
}

inline unsigned x3_grid(long long work) {
  long long b = sa_cdiv(work, X3_THREADS);
  const long long cap = 148LL * 16;
  return (unsigned)std::max(1LL, std::min(b, cap));
}

inline size_t up256(size_t v) { return (v + 255) & ~(size_t)255; }
inline int up(int v, int q) { return (v + q - 1) / q * q; }

int split_cols(const float* src, long long ld, long long rows, int k, int kp, int order, void* dst, cudaStream_t st) {
  if (order == 0) split_cols_kernel<0><<<x3_grid(rows * (kp / 2)), X3_THREADS, 0, st>>>(src, ld, rows, k, kp, (__nv_bfloat16*)dst);
  else split_cols_kernel<1><<<x3_grid(rows * (kp / 2)), X3_THREADS, 0, st>>>(src, ld, rows, k, kp, (__nv_bfloat16*)dst);
  SA_LAUNCH_CHECK();
  return SA_OK;
}

int split_rows(const float* src, long long ld, long long rows, int k, int kp, int order, void* dst, cudaStream_t st) {
  if (order == 0) split_rows_kernel<0><<<x3_grid(rows * (kp / 2)), X3_THREADS, 0, st>>>(src, ld, rows, k, kp, (__nv_bfloat16*)dst);
  else split_rows_kernel<1><<<x3_grid(rows * (kp / 2)), X3_THREADS, 0, st>>>(src, ld, rows, k, kp, (__nv_bfloat16*)dst);
  SA_LAUNCH_CHECK();
  return SA_OK;
}

long long positions(const int32_t* dhw) { return (long long)dhw[0] * dhw[1] * dhw[2]; }

}  // namespace

// ------------------------------------------------------------------------------------------------ dense layers
extern "C" size_t sa_gemm_nt_x3_workspace(int64_t m, int n, int k) {
  const int kp = up(k, 8);
  return up256((size_t)m * 3 * kp * 2) + up256((size_t)n * 3 * kp * 2) + up256((size_t)m * n * 4) + 256;
}

extern "C" int sa_gemm_nt_x3(int64_t m, int n, int k, const float* a, int64_t lda, const float* b, int64_t ldb,
                             const sa_gemm_epilogue* epi, int64_t ldo, void* workspace, size_t ws_bytes, void* stream) {
  SA_CHECK_ARG(a && b && epi && workspace, "null pointer");
  SA_CHECK_ARG(m > 0 && n > 0 && k > 0 && lda >= k && ldb >= k && ldo >= n, "bad sizes");
  SA_CHECK_ARG(epi->out_f32 || epi->out_act || epi->dot_out, "no output");
  SA_CHECK_ARG(epi->act == SA_ACT_NONE || epi->pre, "GELU epilogue needs `pre`");
  SA_CHECK_ARG(!epi->dot_with == !epi->dot_out, "dot_with / dot_out must come together");
  if (ws_bytes < sa_gemm_nt_x3_workspace(m, n, k)) { sa_set_error("sa_gemm_nt_x3: workspace too small"); return SA_ERR_WORKSPACE; }
  SA_UNSUPPORTED((reinterpret_cast<uintptr_t>(workspace) & 255) != 0, "workspace must be 256-byte aligned");
  cudaStream_t st = sa_stream(stream);
  const int kp = up(k, 8);
  uint8_t* w = (uint8_t*)workspace;
  void* a3 = w; w += up256((size_t)m * 3 * kp * 2);
  void* b3 = w; w += up256((size_t)n * 3 * kp * 2);
  float* tmp = (float*)w;
  SA_UNSUPPORTED(!sa_tc_gemm_nt_supported(m, n, 3 * kp, SA_BF16, a3, 3LL * kp, b3, 3LL * kp), "shape outside the tcgen05 GEMM");
  int rc;
  if ((rc = split_cols(a, lda, m, k, kp, 0, a3, st)) != SA_OK) return rc;
  if ((rc = split_cols(b, ldb, n, k, kp, 1, b3, st)) != SA_OK) return rc;
  const SaEpi e = sa_make_epi(epi, ldo);
  if (!e.dot_with && !e.pre && !e.out_act) {
    // bias / scale / residual / fp32 output only: the tensor-core kernel's own epilogue is already the fp32 one
    return sa_tc_gemm_nt(m, n, 3 * kp, a3, 3LL * kp, b3, 3LL * kp, e, st);
  }
  SaEpi plain = {};
  plain.scale = 1.0f; plain.out_f32 = tmp; plain.ldo = n;
  if ((rc = sa_tc_gemm_nt(m, n, 3 * kp, a3, 3LL * kp, b3, 3LL * kp, plain, st)) != SA_OK) return rc;
  x3_epilogue_kernel<<<x3_grid(m * (long long)n), X3_THREADS, 0, st>>>(tmp, m, n, e);
  SA_LAUNCH_CHECK();
  return SA_OK;
}

extern "C" size_t sa_gemm_tn_x3_workspace(int64_t m, int na, int nb) {
  return up256((size_t)3 * m * up(na, 8) * 2) + up256((size_t)3 * m * up(nb, 8) * 2) + 256;
}

extern "C" int sa_gemm_tn_x3(int64_t m, int na, int nb, const float* a, int64_t lda, const float* b, int64_t ldb,
                             const float* scale_dev, float scale, float* d, int accumulate, void* workspace,
                             size_t ws_bytes, void* stream) {
  SA_CHECK_ARG(a && b && d && workspace, "null pointer");
  SA_CHECK_ARG(m > 0 && na > 0 && nb > 0 && lda >= na && ldb >= nb, "bad sizes");
  if (ws_bytes < sa_gemm_tn_x3_workspace(m, na, nb)) { sa_set_error("sa_gemm_tn_x3: workspace too small"); return SA_ERR_WORKSPACE; }
  SA_UNSUPPORTED((reinterpret_cast<uintptr_t>(workspace) & 255) != 0, "workspace must be 256-byte aligned");
  cudaStream_t st = sa_stream(stream);
  const int nap = up(na, 8), nbp = up(nb, 8);
  uint8_t* w = (uint8_t*)workspace;
  void* a3 = w; w += up256((size_t)3 * m * nap * 2);
  void* b3 = w;
  SA_UNSUPPORTED(!sa_tc_gemm_tn_supported(3 * m, na, nb, SA_BF16, a3, nap, b3, nbp), "shape outside the tcgen05 GEMM");
  int rc;
  if (!accumulate) SA_CUDA(cudaMemsetAsync(d, 0, (size_t)na * nb * sizeof(float), st));
  if ((rc = split_rows(a, lda, m, na, nap, 0, a3, st)) != SA_OK) return rc;
  if ((rc = split_rows(b, ldb, m, nb, nbp, 1, b3, st)) != SA_OK) return rc;
  return sa_tc_gemm_tn(3 * m, na, nb, a3, nap, b3, nbp, scale_dev, scale, d, st);
}

// ------------------------------------------------------------------------------------------------ convolutions
namespace {

// the bf16 problem the tensor-core kernels see for a forward / data-gradient launch
sa_conv_desc x3_fwd_desc(const sa_conv_desc* d, int* kp_out) {
  sa_conv_desc e = *d;
  const int kp = up(d->c_in, 64);
  e.c_in = 3 * kp;
  e.act_dtype = SA_BF16;
  *kp_out = kp;
  return e;
}

sa_conv_desc x3_wgrad_desc(const sa_conv_desc* d) {
  sa_conv_desc e = *d;
  e.batch = 3 * d->batch;
  e.act_dtype = SA_BF16;
  return e;
}

}  // namespace

extern "C" int sa_conv3d_x3_supported(const sa_conv_desc* d, int wgrad) {
  if (!d || d->act_dtype != SA_F32 || d->batch < 1) return 0;
  if (wgrad) {
    if (d->transposed || (d->c_in & 7) || (d->c_out & 7)) return 0;
    const sa_conv_desc e = x3_wgrad_desc(d);
    return (sa_tc_wgrad3_supported(&e) || sa_tc_wgrad_supported(&e)) ? 1 : 0;
  }
  int kp;
  const sa_conv_desc e = x3_fwd_desc(d, &kp);
  return sa_tc_conv3_supported(&e) ? 1 : 0;
}

extern "C" size_t sa_conv3d_x3_workspace(const sa_conv_desc* d, int wgrad) {
  if (!d) return 0;
  const long long pin = (long long)d->batch * positions(d->in_dhw), pout = (long long)d->batch * positions(d->out_dhw);
  if (wgrad) return up256((size_t)3 * pout * d->c_out * 2) + up256((size_t)3 * pin * d->c_in * 2) + 256;
  const int kp = up(d->c_in, 64);
  const long long taps = (long long)d->ksize * d->ksize * d->ksize;
  return up256((size_t)pin * 3 * kp * 2) + up256((size_t)taps * d->c_out * 3 * kp * 2) + 256;
}

// x, y, addend, mask: fp32 NDHWC; wp: PACKED fp32 weights [ksize^3][c_out][c_in] (sa_pack_weight with dst_dtype SA_F32)
extern "C" int sa_conv3d_fwd_x3(const sa_conv_desc* d, const float* x, const float* wp, const float* bias,
                                const float* addend, const float* mask, int relu, float* y, void* workspace,
                                size_t ws_bytes, void* stream) {
  SA_CHECK_ARG(d && x && wp && y && workspace, "null pointer");
  SA_UNSUPPORTED(!sa_conv3d_x3_supported(d, 0), "shape outside the table-driven tcgen05 conv kernel");
  if (ws_bytes < sa_conv3d_x3_workspace(d, 0)) { sa_set_error("sa_conv3d_fwd_x3: workspace too small"); return SA_ERR_WORKSPACE; }
  SA_UNSUPPORTED((reinterpret_cast<uintptr_t>(workspace) & 255) != 0, "workspace must be 256-byte aligned");
  cudaStream_t st = sa_stream(stream);
  int kp;
  const sa_conv_desc e = x3_fwd_desc(d, &kp);
  const long long pin = (long long)d->batch * positions(d->in_dhw);
  const long long wrows = (long long)d->ksize * d->ksize * d->ksize * d->c_out;
  uint8_t* w = (uint8_t*)workspace;
  void* x3 = w; w += up256((size_t)pin * 3 * kp * 2);
  void* w3 = w;
  int rc;
  if ((rc = split_cols(x, d->c_in, pin, d->c_in, kp, 0, x3, st)) != SA_OK) return rc;
  if ((rc = split_cols(wp, d->c_in, wrows, d->c_in, kp, 1, w3, st)) != SA_OK) return rc;
  return sa_tc_conv3_fwd_ex(&e, x3, w3, bias, addend, mask, relu, y, true, st);
}

// p [batch, out_dhw, c_out], q [batch, in_dhw, c_in]: fp32 NDHWC; dwp fp32 [ksize^3][c_out][c_in] as sa_conv3d_wgrad
extern "C" int sa_conv3d_wgrad_x3(const sa_conv_desc* d, const float* p, const float* q, float* dwp, int accumulate,
                                  void* workspace, size_t ws_bytes, void* stream) {
  SA_CHECK_ARG(d && p && q && dwp && workspace, "null pointer");
  SA_UNSUPPORTED(!sa_conv3d_x3_supported(d, 1), "shape outside the tcgen05 weight-gradient kernels");
  if (ws_bytes < sa_conv3d_x3_workspace(d, 1)) { sa_set_error("sa_conv3d_wgrad_x3: workspace too small"); return SA_ERR_WORKSPACE; }
  SA_UNSUPPORTED((reinterpret_cast<uintptr_t>(workspace) & 255) != 0, "workspace must be 256-byte aligned");
  cudaStream_t st = sa_stream(stream);
  const sa_conv_desc e = x3_wgrad_desc(d);
  const long long pin = (long long)d->batch * positions(d->in_dhw), pout = (long long)d->batch * positions(d->out_dhw);
  uint8_t* w = (uint8_t*)workspace;
  void* p3 = w; w += up256((size_t)3 * pout * d->c_out * 2);
  void* q3 = w;
  int rc;
  if ((rc = split_rows(p, d->c_out, pout, d->c_out, d->c_out, 0, p3, st)) != SA_OK) return rc;
  if ((rc = split_rows(q, d->c_in, pin, d->c_in, d->c_in, 1, q3, st)) != SA_OK) return rc;
  const long long taps = (long long)d->ksize * d->ksize * d->ksize;
  if (!accumulate) SA_CUDA(cudaMemsetAsync(dwp, 0, (size_t)taps * d->c_out * d->c_in * sizeof(float), st));
  if (sa_tc_wgrad3_supported(&e)) return sa_tc_conv3d_wgrad3(&e, p3, q3, dwp, st);
  return sa_tc_conv3d_wgrad(&e, p3, q3, dwp, st);
}
