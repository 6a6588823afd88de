// Layout / elementwise / reduction helpers around the conv kernels (all HBM-bound, coalesced, grid-stride).
#include "sa_common.cuh"

namespace {

constexpr int EW_THREADS = 256;
inline unsigned ew_grid(int64_t n, int per_thread = 1) {
  int64_t b = sa_cdiv(n, (int64_t)EW_THREADS * per_thread);
  const int64_t cap = 148 * 32;
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return (unsigned)b;
}

template <typename TO>
__global__ void pack_weight_kernel(const float* __restrict__ src, int A, int B, int taps, int transpose, int flip,
                                   TO* __restrict__ dst) {
  const int64_t n = (int64_t)A * B * taps;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    // i indexes dst [t'][r][c]
    const int R = transpose ? B : A, C = transpose ? A : B;
    const int c = (int)(i % C);
    const int r = (int)((i / C) % R);
    const int tp = (int)(i / ((int64_t)C * R));
    const int t = flip ? taps - 1 - tp : tp;
    const int a = transpose ? c : r, b = transpose ? r : c;
    sa_st(dst, i, src[((int64_t)a * B + b) * taps + t]);
  }
}

__global__ void unpack_wgrad_kernel(const float* __restrict__ src, int A, int B, int taps, int transpose, int flip,
                                    float* __restrict__ dst, int accumulate) {
  const int64_t n = (int64_t)A * B * taps;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    // i indexes dst [a][b][t]
    const int t = (int)(i % taps);
    const int b = (int)((i / taps) % B);
    const int a = (int)(i / ((int64_t)taps * B));
    const int tp = flip ? taps - 1 - t : t;
    const int R = transpose ? B : A, C = transpose ? A : B;
    const int r = transpose ? b : a, c = transpose ? a : b;
    const float v = src[((int64_t)tp * R + r) * C + c];
    dst[i] = accumulate ? dst[i] + v : v;
  }
}

// db[c] += sum over rows; block = (C-lane, row-lane) so that global reads stay coalesced along c
template <typename T>
__global__ void __launch_bounds__(256)
bias_grad_kernel(const T* __restrict__ dy, int64_t rows, int C, float* __restrict__ db, int64_t rows_per_block,
                 float* __restrict__ parts) {
  __shared__ float s_acc[256];
  const int cl = min(C, 256);          // threads along c
  const int rl = 256 / cl;             // row lanes (>= 1)
  const int tc = threadIdx.x % cl, tr = threadIdx.x / cl;
  const int64_t r0 = (int64_t)blockIdx.x * rows_per_block;
  const int64_t r1 = min(rows, r0 + rows_per_block);
  for (int c0 = 0; c0 < C; c0 += cl) {
    const int c = c0 + tc;
    float acc = 0.f;
    if (tr < rl && c < C)
      for (int64_t r = r0 + tr; r < r1; r += rl) acc += sa_ld(dy, r * C + c);
    s_acc[threadIdx.x] = acc;
    __syncthreads();
    if (tr == 0 && c < C) {
      float s = 0.f;
      for (int j = 0; j < rl; ++j) s += s_acc[j * cl + tc];
      if (parts) parts[(int64_t)blockIdx.x * C + c] = s;     // deterministic mode: summed in block order afterwards
      else atomicAdd(db + c, s);
    }
    __syncthreads();
  }
}

// vectorised variant (C a multiple of the 16-byte vector, 16-byte aligned rows): 16-byte loads, 4 rows in flight
template <typename T>
__global__ void __launch_bounds__(256)
bias_grad_vec_kernel(const T* __restrict__ dy, int64_t rows, int C, float* __restrict__ db, int64_t rows_per_block,
                     float* __restrict__ parts) {
  constexpr int VEC = 16 / sizeof(T);
  __shared__ float s_acc[256 * VEC];
  const int cv = C / VEC;
  const int cl = min(cv, 256);
  const int rl = 256 / cl;
  const int tc = threadIdx.x % cl, tr = threadIdx.x / cl;
  const int64_t r0 = (int64_t)blockIdx.x * rows_per_block;
  const int64_t r1 = min(rows, r0 + rows_per_block);
  const uint4* base = reinterpret_cast<const uint4*>(dy);
  for (int v0 = 0; v0 < cv; v0 += cl) {
    const int v = v0 + tc;
    float acc[VEC];
#pragma unroll
    for (int i = 0; i < VEC; ++i) acc[i] = 0.f;
    if (tr < rl && v < cv) {
      int64_t r = r0 + tr;
      for (; r + 3 * rl < r1; r += 4 * rl) {
        uint4 u[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) u[k] = __ldg(base + (r + (int64_t)k * rl) * cv + v);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          if (sizeof(T) == 2) {
            const uint32_t w[4] = {u[k].x, u[k].y, u[k].z, u[k].w};
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              acc[(2 * i) % VEC] += __uint_as_float(w[i] << 16);
              acc[(2 * i + 1) % VEC] += __uint_as_float(w[i] & 0xffff0000u);
            }
          } else {
            acc[0] += __uint_as_float(u[k].x); acc[1] += __uint_as_float(u[k].y);
            acc[2] += __uint_as_float(u[k].z); acc[3] += __uint_as_float(u[k].w);
          }
        }
      }
      for (; r < r1; r += rl) {
        const uint4 u = __ldg(base + r * cv + v);
        if (sizeof(T) == 2) {
          const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            acc[(2 * i) % VEC] += __uint_as_float(w[i] << 16);
            acc[(2 * i + 1) % VEC] += __uint_as_float(w[i] & 0xffff0000u);
          }
        } else {
          acc[0] += __uint_as_float(u.x); acc[1] += __uint_as_float(u.y);
          acc[2] += __uint_as_float(u.z); acc[3] += __uint_as_float(u.w);
        }
      }
    }
#pragma unroll
    for (int i = 0; i < VEC; ++i) s_acc[i * 256 + threadIdx.x] = acc[i];
    __syncthreads();
    if (tr == 0 && v < cv) {
#pragma unroll
      for (int i = 0; i < VEC; ++i) {
        float t = 0.f;
        for (int j = 0; j < rl; ++j) t += s_acc[i * 256 + j * cl + tc];
        if (parts) parts[(int64_t)blockIdx.x * C + v * VEC + i] = t;
        else atomicAdd(db + v * VEC + i, t);
      }
    }
    __syncthreads();
  }
}

// [b][c][s] -> [b][s][c] through a 32x32 shared tile (both sides coalesced)
template <typename TI, typename TO>
__global__ void transpose_cs_kernel(const TI* __restrict__ src, TO* __restrict__ dst, int C, int64_t S, bool to_nhwc) {
  __shared__ float tile[32][33];
  const int64_t b = blockIdx.z;
  const int64_t s0 = (int64_t)blockIdx.x * 32;
  const int c0 = blockIdx.y * 32;
  const int tx = threadIdx.x, ty = threadIdx.y;  // 32 x 8
  if (to_nhwc) {
    for (int j = ty; j < 32; j += 8) {
      const int c = c0 + j; const int64_t s = s0 + tx;
      tile[j][tx] = (c < C && s < S) ? sa_ld(src, (b * C + c) * S + s) : 0.f;
    }
    __syncthreads();
    for (int j = ty; j < 32; j += 8) {
      const int64_t s = s0 + j; const int c = c0 + tx;
      if (c < C && s < S) sa_st(dst, (b * S + s) * C + c, tile[tx][j]);
    }
  } else {
    for (int j = ty; j < 32; j += 8) {
      const int64_t s = s0 + j; const int c = c0 + tx;
      tile[j][tx] = (c < C && s < S) ? sa_ld(src, (b * S + s) * C + c) : 0.f;
    }
    __syncthreads();
    for (int j = ty; j < 32; j += 8) {
      const int c = c0 + j; const int64_t s = s0 + tx;
      if (c < C && s < S) sa_st(dst, (b * C + c) * S + s, tile[tx][j]);
    }
  }
}

template <typename TI, typename TO>
__global__ void cast_kernel(const TI* __restrict__ src, TO* __restrict__ dst, int64_t n) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    sa_st(dst, i, sa_ld(src, i));
}

template <typename T>
__global__ void __launch_bounds__(EW_THREADS)
mse_kernel(const T* __restrict__ a, const float* __restrict__ b, int64_t n, float scale,
           const float* __restrict__ scale_dev, float* __restrict__ sse, T* __restrict__ grad, unsigned* turn) {
  __shared__ float s_red[EW_THREADS / 32];
  float acc = 0.f;
  if (scale_dev) scale *= scale_dev[0];
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float d = sa_ld(a, i) - b[i];
    acc = fmaf(d, d, acc);
    if (grad) sa_st(grad, i, scale * d);
  }
  acc = sa_warp_sum(acc);
  if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = acc;
  __syncthreads();
  sa_block_turn_begin(turn, blockIdx.x);
  if (threadIdx.x == 0 && sse) {
    float s = 0.f;
    for (int i = 0; i < EW_THREADS / 32; ++i) s += s_red[i];
    atomicAdd(sse, s);
  }
  sa_block_turn_end(turn, blockIdx.x);
}

__global__ void adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                            float* __restrict__ v, int64_t n, float beta1, float beta2, float eps, float step_size,
                            float inv_sqrt_bc2) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float gi = g[i];
    const float mi = fmaf(beta1, m[i], (1.f - beta1) * gi);          // exp_avg.lerp_(grad, 1 - beta1)
    const float vi = fmaf(beta2, v[i], (1.f - beta2) * gi * gi);     // exp_avg_sq.mul_(b2).addcmul_(g, g, 1 - b2)
    m[i] = mi; v[i] = vi;
    const float denom = sqrtf(vi) * inv_sqrt_bc2 + eps;              // (sqrt(v) / sqrt(bc2)) + eps
    p[i] = p[i] - step_size * (mi / denom);                          // p.addcdiv_(m, denom, value=-lr / bc1)
  }
}

// multi-tensor form: one launch updates up to ADAM_BATCH parameter tensors (blockIdx.y = tensor, grid-stride over its
// elements), so that an optimiser step over the 155 (VQ-VAE) / 249 (Performer) parameter tensors is 3-4 launches
constexpr int ADAM_BATCH = 64;
struct AdamBatch {
  float* p[ADAM_BATCH];
  const float* g[ADAM_BATCH];
  float* m[ADAM_BATCH];
  float* v[ADAM_BATCH];
  long long n[ADAM_BATCH];
};
__global__ void __launch_bounds__(256)
adam_multi_kernel(const __grid_constant__ AdamBatch B, float beta1, float beta2, float eps, float step_size,
                  float inv_sqrt_bc2) {
  const int t = blockIdx.y;
  float* __restrict__ p = B.p[t];
  const float* __restrict__ g = B.g[t];
  float* __restrict__ m = B.m[t];
  float* __restrict__ v = B.v[t];
  const long long n = B.n[t];
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float gi = g[i];
    const float mi = fmaf(beta1, m[i], (1.f - beta1) * gi);
    const float vi = fmaf(beta2, v[i], (1.f - beta2) * gi * gi);
    m[i] = mi; v[i] = vi;
    const float denom = sqrtf(vi) * inv_sqrt_bc2 + eps;
    p[i] = p[i] - step_size * (mi / denom);
  }
}

// cols[b, o, t] = x[b, o*s - p + t]  (single-channel x), t = (td*k + th)*k + tw.  One thread per (o, td, th):
// the k taps along w are contiguous in x and in cols, a position's k^3 taps form one contiguous row.
template <typename T>
__global__ void __launch_bounds__(256)
im2col_c1_kernel(const T* __restrict__ x, int B, int iD, int iH, int iW, int oD, int oH, int oW, int k, int s, int p,
                 T* __restrict__ cols) {
  const int kk = k * k;
  const int64_t total = (int64_t)B * oD * oH * oW * kk;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int64_t r = i;
    const int j = (int)(r % kk); r /= kk;
    const int ow = (int)(r % oW); r /= oW;
    const int oh = (int)(r % oH); r /= oH;
    const int od = (int)(r % oD); r /= oD;
    const int b = (int)r;
    const int th = j % k, td = j / k;
    const int id = od * s - p + td, ih = oh * s - p + th;
    const bool row_ok = id >= 0 && id < iD && ih >= 0 && ih < iH;
    const T* xr = x + (((int64_t)b * iD + id) * iH + ih) * iW;
    T* cr = cols + i * k;
    const T zero = T(0.f);
    for (int tw = 0; tw < k; ++tw) {
      const int iw = ow * s - p + tw;
      cr[tw] = (row_ok && iw >= 0 && iw < iW) ? xr[iw] : zero;
    }
  }
}

// y[b, o] = bias + sum_{t : (o + p - t) % s == 0} cols[b, (o + p - t)/s, t]   (single-channel y)
template <typename T>
__global__ void __launch_bounds__(256)
col2im_c1_kernel(const T* __restrict__ cols, int B, int iD, int iH, int iW, int oD, int oH, int oW, int k, int s, int p,
                 const float* __restrict__ bias, T* __restrict__ y) {
  // block = 32 (w) x 4 (h) x 2 (d) output brick so that neighbouring voxels re-use the cols rows through L1
  const int ow = blockIdx.x * 32 + (threadIdx.x & 31);
  const int oh = blockIdx.y * 4 + ((threadIdx.x >> 5) & 3);
  const int odb = blockIdx.z * 2 + (threadIdx.x >> 7);
  const int nd = (oD + 1) / 2;
  const int b = odb / (nd * 2), od = odb % (nd * 2);
  if (b >= B || od >= oD || oh >= oH || ow >= oW) return;
  const int taps = k * k * k;
  float acc = bias ? bias[0] : 0.f;
  for (int td = (od + p) % s; td < k; td += s) {
    const int id = (od + p - td) / s;
    if (od + p - td < 0 || id >= iD) continue;
    for (int th = (oh + p) % s; th < k; th += s) {
      const int ih = (oh + p - th) / s;
      if (oh + p - th < 0 || ih >= iH) continue;
      for (int tw = (ow + p) % s; tw < k; tw += s) {
        const int iw = (ow + p - tw) / s;
        if (ow + p - tw < 0 || iw >= iW) continue;
        const int64_t ip = (((int64_t)b * iD + id) * iH + ih) * iW + iw;
        acc += sa_ld(cols, ip * taps + (td * k + th) * k + tw);
      }
    }
  }
  sa_st(y, (((int64_t)b * oD + od) * oH + oh) * oW + ow, acc);
}

}  // namespace

extern "C" int sa_pack_weight(const float* src, int A, int B, int taps, int transpose, int flip, void* dst,
                              int dst_dtype, void* stream) {
  SA_CHECK_ARG(src && dst, "null pointer");
  SA_CHECK_ARG(A > 0 && B > 0 && taps > 0, "bad sizes");
  const int64_t n = (int64_t)A * B * taps;
  if (dst_dtype == SA_BF16)
    pack_weight_kernel<__nv_bfloat16><<<ew_grid(n), EW_THREADS, 0, sa_stream(stream)>>>(src, A, B, taps, transpose, flip,
                                                                                       (__nv_bfloat16*)dst);
  else
    pack_weight_kernel<float><<<ew_grid(n), EW_THREADS, 0, sa_stream(stream)>>>(src, A, B, taps, transpose, flip,
                                                                               (float*)dst);
  SA_LAUNCH_CHECK();
  return SA_OK;
}

// many weights per launch (blockIdx.y = weight): the per-pass packing of a whole network's conv weights, forward and
// transposed (data-gradient) forms, in a few launches instead of two per conv and step
namespace {
constexpr int WPACK_BATCH = 48;
struct WPackBatch {
  const float* src[WPACK_BATCH];
  void* dst[WPACK_BATCH];
  int A[WPACK_BATCH], B[WPACK_BATCH], taps[WPACK_BATCH];
  unsigned char transpose[WPACK_BATCH], flip[WPACK_BATCH];
};

template <typename TO>
__global__ void pack_weight_multi_kernel(const __grid_constant__ WPackBatch b) {
  const int w = blockIdx.y;
  const int A = b.A[w], B = b.B[w], taps = b.taps[w];
  const int transpose = b.transpose[w], flip = b.flip[w];
  const float* __restrict__ src = b.src[w];
  TO* __restrict__ dst = reinterpret_cast<TO*>(b.dst[w]);
  const int64_t n = (int64_t)A * B * taps;
  const int R = transpose ? B : A, C = transpose ? A : B;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    const int r = (int)((i / C) % R);
    const int tp = (int)(i / ((int64_t)C * R));
    const int t = flip ? taps - 1 - tp : tp;
    const int a = transpose ? c : r, bb = transpose ? r : c;
    sa_st(dst, i, src[((int64_t)a * B + bb) * taps + t]);
  }
}
}  // namespace

extern "C" int sa_pack_weight_multi(const sa_wpack_item* items, int n, int dst_dtype, void* stream) {
  SA_CHECK_ARG(items != nullptr && n >= 0, "bad arguments");
  SA_CHECK_ARG(dst_dtype == SA_F32 || dst_dtype == SA_BF16, "bad dtype");
  cudaStream_t st = sa_stream(stream);
  for (int base = 0; base < n; base += WPACK_BATCH) {
    WPackBatch b = {};
    const int cnt = n - base < WPACK_BATCH ? n - base : WPACK_BATCH;
    int64_t nmax = 1;
    for (int i = 0; i < cnt; ++i) {
      const sa_wpack_item& it = items[base + i];
      SA_CHECK_ARG(it.src && it.dst && it.A > 0 && it.B > 0 && it.taps > 0, "bad item");
      b.src[i] = it.src; b.dst[i] = it.dst; b.A[i] = it.A; b.B[i] = it.B; b.taps[i] = it.taps;
      b.transpose[i] = (unsigned char)(it.transpose != 0); b.flip[i] = (unsigned char)(it.flip != 0);
      const int64_t ni = (int64_t)it.A * it.B * it.taps;
      if (ni > nmax) nmax = ni;
    }
    unsigned gx = ew_grid(nmax, 4);
    if (gx > 148 * 2) gx = 148 * 2;
    if (dst_dtype == SA_BF16) pack_weight_multi_kernel<__nv_bfloat16><<<dim3(gx, (unsigned)cnt), EW_THREADS, 0, st>>>(b);
    else pack_weight_multi_kernel<float><<<dim3(gx, (unsigned)cnt), EW_THREADS, 0, st>>>(b);
    SA_LAUNCH_CHECK();
  }
  return SA_OK;
}

extern "C" int sa_unpack_wgrad(const float* src, int A, int B, int taps, int transpose, int flip, float* dst,
                               int accumulate, void* stream) {
  SA_CHECK_ARG(src && dst, "null pointer");
  SA_CHECK_ARG(A > 0 && B > 0 && taps > 0, "bad sizes");
  const int64_t n = (int64_t)A * B * taps;
  unpack_wgrad_kernel<<<ew_grid(n), EW_THREADS, 0, sa_stream(stream)>>>(src, A, B, taps, transpose, flip, dst,
                                                                        accumulate);
  SA_LAUNCH_CHECK();
  return SA_OK;
}

// ------------------------------------------------------------------------------------------------
// Per-step preparation of the dense layers' weights, many tensors per launch (blockIdx.y = tensor):
//   dst  [rows x cols]  (leading dimension dst_ld)    = bf16(src)          the NT-GEMM operand of the forward pass
//   dstt [cols x rows]  (leading dimension dstt_ld)   = bf16(src)^T        the NT-GEMM operand of the data gradient
// 32 x 32 tiles through shared memory, both stores coalesced.  Replaces one cast + one transpose launch (+ a concat for
// q | k | v) per weight and step: 218 launches per Performer step.
namespace {
constexpr int WPREP_BATCH = 48;
struct WPrepBatch {
  const float* src[WPREP_BATCH];
  __nv_bfloat16* dst[WPREP_BATCH];
  __nv_bfloat16* dstt[WPREP_BATCH];
  int rows[WPREP_BATCH], cols[WPREP_BATCH], dst_ld[WPREP_BATCH], dstt_ld[WPREP_BATCH];
};

__global__ void __launch_bounds__(256) weight_prep_kernel(const __grid_constant__ WPrepBatch b) {
  __shared__ float tile[32][33];
  const int w = blockIdx.y;
  const int rows = b.rows[w], cols = b.cols[w];
  const int tc = (cols + 31) >> 5, tr = (rows + 31) >> 5;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;          // 32 x 8
  for (int t = blockIdx.x; t < tr * tc; t += gridDim.x) {
    const int r0 = (t / tc) * 32, c0 = (t % tc) * 32;
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int r = r0 + ty + 8 * i, c = c0 + tx;
      const float v = (r < rows && c < cols) ? b.src[w][(long long)r * cols + c] : 0.f;
      tile[ty + 8 * i][tx] = v;
      if (b.dst[w] && r < rows && c < cols) b.dst[w][(long long)r * b.dst_ld[w] + c] = __float2bfloat16_rn(v);
    }
    __syncthreads();
    if (b.dstt[w]) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int c = c0 + ty + 8 * i, r = r0 + tx;
        if (r < rows && c < cols) b.dstt[w][(long long)c * b.dstt_ld[w] + r] = __float2bfloat16_rn(tile[tx][ty + 8 * i]);
      }
    }
  }
}
}  // namespace

extern "C" int sa_weight_prep(const sa_wprep_item* items, int n, void* stream) {
  SA_CHECK_ARG(items != nullptr && n >= 0, "bad arguments");
  cudaStream_t st = sa_stream(stream);
  for (int base = 0; base < n; base += WPREP_BATCH) {
    WPrepBatch b = {};
    const int cnt = n - base < WPREP_BATCH ? n - base : WPREP_BATCH;
    long long max_tiles = 1;
    for (int i = 0; i < cnt; ++i) {
      const sa_wprep_item& it = items[base + i];
      SA_CHECK_ARG(it.src && it.rows > 0 && it.cols > 0 && (it.dst || it.dst_t), "bad item");
      b.src[i] = it.src; b.dst[i] = (__nv_bfloat16*)it.dst; b.dstt[i] = (__nv_bfloat16*)it.dst_t;
      b.rows[i] = it.rows; b.cols[i] = it.cols; b.dst_ld[i] = it.dst_ld; b.dstt_ld[i] = it.dst_t_ld;
      const long long tiles = sa_cdiv(it.rows, 32) * sa_cdiv(it.cols, 32);
      if (tiles > max_tiles) max_tiles = tiles;
    }
    if (max_tiles > 148 * 2) max_tiles = 148 * 2;
    weight_prep_kernel<<<dim3((unsigned)max_tiles, (unsigned)cnt), 256, 0, st>>>(b);
    SA_LAUNCH_CHECK();
  }
  return SA_OK;
}

extern "C" int sa_bias_grad(const void* dy, int64_t rows, int c, int dtype, float* db, int accumulate, void* stream) {
  SA_CHECK_ARG(dy && db, "null pointer");
  SA_CHECK_ARG(rows >= 0 && c > 0, "bad sizes");
  cudaStream_t st = sa_stream(stream);
  if (!accumulate) SA_CUDA(cudaMemsetAsync(db, 0, sizeof(float) * c, st));
  if (rows == 0) return SA_OK;
  int64_t blocks = sa_cdiv(rows, 256);
  if (blocks > 148 * 8) blocks = 148 * 8;
  const int64_t rpb = sa_cdiv(rows, blocks);
  blocks = sa_cdiv(rows, rpb);
  const int vec = dtype == SA_BF16 ? 8 : 4;
  const bool vec_ok = (c % vec == 0) && ((reinterpret_cast<uintptr_t>(dy) & 15) == 0);
  float* parts = sa_parts_alloc(blocks, c, st);      // deterministic mode: per-block partials, summed in block order
  if (vec_ok && dtype == SA_BF16)
    bias_grad_vec_kernel<__nv_bfloat16><<<(unsigned)blocks, 256, 0, st>>>((const __nv_bfloat16*)dy, rows, c, db, rpb, parts);
  else if (vec_ok)
    bias_grad_vec_kernel<float><<<(unsigned)blocks, 256, 0, st>>>((const float*)dy, rows, c, db, rpb, parts);
  else if (dtype == SA_BF16)
    bias_grad_kernel<__nv_bfloat16><<<(unsigned)blocks, 256, 0, st>>>((const __nv_bfloat16*)dy, rows, c, db, rpb, parts);
  else
    bias_grad_kernel<float><<<(unsigned)blocks, 256, 0, st>>>((const float*)dy, rows, c, db, rpb, parts);
  SA_LAUNCH_CHECK();
  if (parts) {
    const int rc = sa_parts_reduce(parts, blocks, c, c, db, st);
    if (rc != SA_OK) return rc;
    return sa_parts_free(parts, st);
  }
  return SA_OK;
}

template <typename TI, typename TO>
static int launch_transpose(const void* src, void* dst, int64_t batch, int c, int64_t spatial, bool to_nhwc,
                            cudaStream_t st) {
  if (c == 1) {  // the two layouts coincide
    cast_kernel<TI, TO><<<ew_grid(batch * spatial, 4), EW_THREADS, 0, st>>>((const TI*)src, (TO*)dst, batch * spatial);
  } else {
    dim3 grid((unsigned)sa_cdiv(spatial, 32), (unsigned)sa_cdiv(c, 32), (unsigned)batch);
    transpose_cs_kernel<TI, TO><<<grid, dim3(32, 8), 0, st>>>((const TI*)src, (TO*)dst, c, spatial, to_nhwc);
  }
  SA_LAUNCH_CHECK();
  return SA_OK;
}

static int dispatch_transpose(const void* src, int sdt, void* dst, int ddt, int64_t batch, int c, int64_t spatial,
                              bool to_nhwc, void* stream) {
  SA_CHECK_ARG(src && dst, "null pointer");
  SA_CHECK_ARG(batch >= 0 && c > 0 && spatial >= 0, "bad sizes");
  SA_UNSUPPORTED(batch > 65535, "batch > 65535");
  if (batch == 0 || spatial == 0) return SA_OK;
  cudaStream_t st = sa_stream(stream);
  if (sdt == SA_F32 && ddt == SA_F32) return launch_transpose<float, float>(src, dst, batch, c, spatial, to_nhwc, st);
  if (sdt == SA_F32 && ddt == SA_BF16)
    return launch_transpose<float, __nv_bfloat16>(src, dst, batch, c, spatial, to_nhwc, st);
  if (sdt == SA_BF16 && ddt == SA_F32)
    return launch_transpose<__nv_bfloat16, float>(src, dst, batch, c, spatial, to_nhwc, st);
  return launch_transpose<__nv_bfloat16, __nv_bfloat16>(src, dst, batch, c, spatial, to_nhwc, st);
}

extern "C" int sa_nchw_to_nhwc(const void* src, int sdt, void* dst, int ddt, int64_t batch, int c, int64_t spatial,
                               void* stream) {
  return dispatch_transpose(src, sdt, dst, ddt, batch, c, spatial, true, stream);
}
extern "C" int sa_nhwc_to_nchw(const void* src, int sdt, void* dst, int ddt, int64_t batch, int c, int64_t spatial,
                               void* stream) {
  return dispatch_transpose(src, sdt, dst, ddt, batch, c, spatial, false, stream);
}

extern "C" int sa_cast(const void* src, int sdt, void* dst, int ddt, int64_t n, void* stream) {
  SA_CHECK_ARG(src && dst && n >= 0, "bad arguments");
  if (n == 0) return SA_OK;
  cudaStream_t st = sa_stream(stream);
  const unsigned g = ew_grid(n, 4);
  if (sdt == SA_F32 && ddt == SA_F32) cast_kernel<float, float><<<g, EW_THREADS, 0, st>>>((const float*)src, (float*)dst, n);
  else if (sdt == SA_F32) cast_kernel<float, __nv_bfloat16><<<g, EW_THREADS, 0, st>>>((const float*)src, (__nv_bfloat16*)dst, n);
  else if (ddt == SA_F32) cast_kernel<__nv_bfloat16, float><<<g, EW_THREADS, 0, st>>>((const __nv_bfloat16*)src, (float*)dst, n);
  else cast_kernel<__nv_bfloat16, __nv_bfloat16><<<g, EW_THREADS, 0, st>>>((const __nv_bfloat16*)src, (__nv_bfloat16*)dst, n);
  SA_LAUNCH_CHECK();
  return SA_OK;
}

extern "C" int sa_mse_fwd_bwd(const void* a, int a_dtype, const float* b, int64_t n, float scale,
                              const float* scale_dev, float* sse, void* grad, void* stream) {
  SA_CHECK_ARG(a && b && n >= 0, "bad arguments");
  if (n == 0) return SA_OK;
  cudaStream_t st = sa_stream(stream);
  unsigned g = ew_grid(n, 8);
  if (sa_deterministic() && g > 148 * 4) g = 148 * 4;      // ordered adds serialise the blocks' tails
  unsigned* turn = sse ? sa_turn_slot(1, st) : nullptr;
  if (a_dtype == SA_BF16)
    mse_kernel<__nv_bfloat16><<<g, EW_THREADS, 0, st>>>((const __nv_bfloat16*)a, b, n, scale, scale_dev, sse,
                                                        (__nv_bfloat16*)grad, turn);
  else
    mse_kernel<float><<<g, EW_THREADS, 0, st>>>((const float*)a, b, n, scale, scale_dev, sse, (float*)grad, turn);
  SA_LAUNCH_CHECK();
  return SA_OK;
}

extern "C" int sa_adam_step(float* p, const float* g, float* m, float* v, int64_t n, float lr, float beta1,
                            float beta2, float eps, int step, void* stream) {
  SA_CHECK_ARG(p && g && m && v && n >= 0 && step >= 1, "bad arguments");
  if (n == 0) return SA_OK;
  const double bc1 = 1.0 - pow((double)beta1, (double)step);
  const double bc2 = 1.0 - pow((double)beta2, (double)step);
  adam_kernel<<<ew_grid(n, 4), EW_THREADS, 0, sa_stream(stream)>>>(p, g, m, v, n, beta1, beta2, eps,
                                                                   (float)((double)lr / bc1), (float)(1.0 / sqrt(bc2)));
  SA_LAUNCH_CHECK();
  return SA_OK;
}

extern "C" int sa_adam_multi(int count, float* const* p, const float* const* g, float* const* m, float* const* v,
                             const int64_t* n, float lr, float beta1, float beta2, float eps, int step, void* stream) {
  SA_CHECK_ARG(count >= 0 && (count == 0 || (p && g && m && v && n)) && step >= 1, "bad arguments");
  const double bc1 = 1.0 - pow((double)beta1, (double)step);
  const double bc2 = 1.0 - pow((double)beta2, (double)step);
  cudaStream_t st = sa_stream(stream);
  for (int base = 0; base < count; base += ADAM_BATCH) {
    AdamBatch B;
    const int nb = count - base < ADAM_BATCH ? count - base : ADAM_BATCH;
    long long nmax = 0;
    for (int i = 0; i < nb; ++i) {
      SA_CHECK_ARG(p[base + i] && g[base + i] && m[base + i] && v[base + i] && n[base + i] >= 0, "null tensor");
      B.p[i] = p[base + i]; B.g[i] = g[base + i]; B.m[i] = m[base + i]; B.v[i] = v[base + i]; B.n[i] = n[base + i];
      if (B.n[i] > nmax) nmax = B.n[i];
    }
    if (nmax == 0) continue;
    long long bx = sa_cdiv(nmax, 256 * 4);
    if (bx > 148) bx = 148;
    adam_multi_kernel<<<dim3((unsigned)bx, (unsigned)nb), 256, 0, st>>>(B, beta1, beta2, eps, (float)((double)lr / bc1),
                                                                       (float)(1.0 / sqrt(bc2)));
    SA_LAUNCH_CHECK();
  }
  return SA_OK;
}

extern "C" int sa_im2col_c1(const void* x, int dtype, int batch, const int* in_dhw, const int* out_dhw, int k, int s,
                            int p, void* cols, void* stream) {
  SA_CHECK_ARG(x && cols && in_dhw && out_dhw && batch > 0 && k > 0 && s > 0 && p >= 0, "bad arguments");
  const int64_t total = (int64_t)batch * out_dhw[0] * out_dhw[1] * out_dhw[2] * k * k;
  cudaStream_t st = sa_stream(stream);
  if (dtype == SA_BF16)
    im2col_c1_kernel<__nv_bfloat16><<<ew_grid(total), EW_THREADS, 0, st>>>(
        (const __nv_bfloat16*)x, batch, in_dhw[0], in_dhw[1], in_dhw[2], out_dhw[0], out_dhw[1], out_dhw[2], k, s, p,
        (__nv_bfloat16*)cols);
  else
    im2col_c1_kernel<float><<<ew_grid(total), EW_THREADS, 0, st>>>((const float*)x, batch, in_dhw[0], in_dhw[1],
                                                                   in_dhw[2], out_dhw[0], out_dhw[1], out_dhw[2], k, s,
                                                                   p, (float*)cols);
  SA_LAUNCH_CHECK();
  return SA_OK;
}

extern "C" int sa_col2im_c1(const void* cols, int dtype, int batch, const int* in_dhw, const int* out_dhw, int k, int s,
                            int p, const float* bias, void* y, void* stream) {
  SA_CHECK_ARG(cols && y && in_dhw && out_dhw && batch > 0 && k > 0 && s > 0 && p >= 0, "bad arguments");
  cudaStream_t st = sa_stream(stream);
  const int nd = (out_dhw[0] + 1) / 2;
  dim3 grid((unsigned)sa_cdiv(out_dhw[2], 32), (unsigned)sa_cdiv(out_dhw[1], 4), (unsigned)(nd * batch));
  SA_UNSUPPORTED(grid.y > 65535 || grid.z > 65535, "volume too large for the col2im grid");
  if (dtype == SA_BF16)
    col2im_c1_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>((const __nv_bfloat16*)cols, batch, in_dhw[0], in_dhw[1],
                                                          in_dhw[2], out_dhw[0], out_dhw[1], out_dhw[2], k, s, p, bias,
                                                          (__nv_bfloat16*)y);
  else
    col2im_c1_kernel<float><<<grid, 256, 0, st>>>((const float*)cols, batch, in_dhw[0], in_dhw[1], in_dhw[2], out_dhw[0],
                                                  out_dhw[1], out_dhw[2], k, s, p, bias, (float*)y);
  SA_LAUNCH_CHECK();
  return SA_OK;
}
