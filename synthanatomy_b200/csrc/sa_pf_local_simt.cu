// Local-window causal attention on CUDA cores (fp32 accumulation), flash-style: no score tensor is materialised.
// fp32 "parity" path and generic fallback of the tensor-core kernel.
//
// Replaces local_attention.LocalAttention.forward (window w, causal, look_backward = 1, autopad, rotary position
// term) called by performer-pytorch SelfAttention for the local heads, reached from
// /root/reference/src/networks/transformers/performer.py:270 (ctor args :199-200).
//
//   query p attends keys j with lo(p) <= j <= p,  lo(p) = max(0, floor(p / w) - 1) * w
//   scores = (rot(q) . rot(k)) * d^-1/2,  softmax over the allowed keys,  out = probs . v
#include "sa_pf_common.cuh"

namespace {

constexpr int LT = 64;     // query / key tile
constexpr int LD = 65;     // shared-memory leading dimension (odd)

struct LaArgs {
  int B, N, H, d, W, ld, out_ld;
  float scale;
};

__device__ __forceinline__ int la_lo(int p, int W) { const int w = p / W - 1; return (w > 0 ? w : 0) * W; }

// load rows n0..n0+63 of head (b, h) from src into tile[64][LD]; optional rotary (rot != 0) and scaling
template <typename T>
__device__ __forceinline__ void la_load(const LaArgs& a, float* tile, const T* __restrict__ src, long long ld, int b, int h,
                                        int n0, const float* __restrict__ inv_freq, bool rot, float scale, int t) {
  const int half = a.d / 2;
  for (int i = t; i < LT * half; i += 256) {
    const int row = i / half, dd = i % half;
    const int n = n0 + row;
    float x1 = 0.f, x2 = 0.f;
    if (n < a.N) {
      const long long base = ((long long)b * a.N + n) * ld + h * a.d;
      x1 = sa_ld(src, base + dd);
      x2 = sa_ld(src, base + dd + half);
      if (rot) {
        float sn, cs;
        sincosf((float)n * inv_freq[dd], &sn, &cs);
        const float r1 = x1 * cs - x2 * sn;      // q * cos + rotate_half(q) * sin, rotate_half = (-x2, x1)
        const float r2 = x2 * cs + x1 * sn;
        x1 = r1; x2 = r2;
      }
      x1 *= scale; x2 *= scale;
    }
    tile[row * LD + dd] = x1;
    tile[row * LD + dd + half] = x2;
  }
}

// write a [64][d] register-distributed tile (rows ty + 16 r, columns tx + 16 s) through shared memory with the
// inverse rotation (transpose of the rotary map) applied
template <typename T>
__device__ __forceinline__ void la_store_unrot(const LaArgs& a, float* tile, float (&acc)[4][4], T* __restrict__ dst,
                                               int b, int h, int n0, const float* __restrict__ inv_freq, bool rot,
                                               float scale, int t, int ty, int tx) {
  __syncthreads();
#pragma unroll
  for (int r = 0; r < 4; ++r)
#pragma unroll
    for (int s = 0; s < 4; ++s) tile[(ty + 16 * r) * LD + tx + 16 * s] = acc[r][s] * scale;
  __syncthreads();
  const int half = a.d / 2;
  for (int i = t; i < LT * half; i += 256) {
    const int row = i / half, dd = i % half;
    const int n = n0 + row;
    if (n >= a.N) continue;
    float g1 = tile[row * LD + dd], g2 = tile[row * LD + dd + half];
    if (rot) {
      float sn, cs;
      sincosf((float)n * inv_freq[dd], &sn, &cs);
      const float u1 = g1 * cs + g2 * sn;
      const float u2 = g2 * cs - g1 * sn;
      g1 = u1; g2 = u2;
    }
    const long long base = ((long long)b * a.N + n) * a.ld + h * a.d;
    sa_st(dst, base + dd, g1);
    sa_st(dst, base + dd + half, g2);
  }
}

template <typename T>
__global__ void __launch_bounds__(256)
local_fwd_kernel(LaArgs a, const T* __restrict__ q, const T* __restrict__ k, const T* __restrict__ v,
                 const float* __restrict__ inv_freq, T* __restrict__ out, float* __restrict__ lse) {
  extern __shared__ float sm[];
  float* Qs = sm;
  float* Ks = Qs + LT * LD;
  float* Vs = Ks + LT * LD;
  float* Ps = Vs + LT * LD;
  const int t = threadIdx.x, ty = t >> 4, tx = t & 15;
  const int bh = blockIdx.y, b = bh / a.H, h = bh % a.H;
  const int i0 = blockIdx.x * LT;
  const bool rot = inv_freq != nullptr;
  la_load<T>(a, Qs, q, a.ld, b, h, i0, inv_freq, rot, a.scale, t);
  float o[4][4], mrow[4], lrow[4];
  int lo[4];
  sa_tile_zero(o);
#pragma unroll
  for (int r = 0; r < 4; ++r) { mrow[r] = -INFINITY; lrow[r] = 0.f; lo[r] = la_lo(i0 + ty + 16 * r, a.W); }
  const int j_beg = (la_lo(i0, a.W) / LT) * LT;
  const int j_last = min(a.N - 1, i0 + LT - 1);
  for (int j0 = j_beg; j0 <= j_last; j0 += LT) {
    __syncthreads();
    la_load<T>(a, Ks, k, a.ld, b, h, j0, inv_freq, rot, 1.0f, t);
    la_load<T>(a, Vs, v, a.ld, b, h, j0, nullptr, false, 1.0f, t);
    __syncthreads();
    float s[4][4];
    sa_tile_zero(s);
    sa_tile_mma<4, 4>(s, Qs, LD, 1, Ks, 1, LD, a.d, ty, tx);
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const int i = i0 + ty + 16 * r;
      float mx = -INFINITY;
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const int j = j0 + tx + 16 * c;
        const bool ok = j <= i && j >= lo[r] && j < a.N;
        s[r][c] = ok ? s[r][c] : -INFINITY;
        mx = fmaxf(mx, s[r][c]);
      }
      mx = sa_half_max(mx);
      const float mnew = fmaxf(mrow[r], mx);
      const float alpha = (mnew == -INFINITY) ? 1.0f : __expf(mrow[r] - mnew);
      float ps = 0.f;
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const float p = (s[r][c] == -INFINITY) ? 0.f : expf(s[r][c] - mnew);
        Ps[(ty + 16 * r) * LD + tx + 16 * c] = p;
        ps += p;
      }
      ps = sa_half_sum(ps);
      lrow[r] = lrow[r] * alpha + ps;
      mrow[r] = mnew;
#pragma unroll
      for (int c = 0; c < 4; ++c) o[r][c] *= alpha;
    }
    __syncthreads();
    sa_tile_mma<4, 4>(o, Ps, LD, 1, Vs, LD, 1, LT, ty, tx);
  }
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const int n = i0 + ty + 16 * r;
    if (n >= a.N) continue;
    const float inv = 1.0f / lrow[r];
    if (tx == 0) lse[(long long)bh * a.N + n] = mrow[r] + logf(lrow[r]);
#pragma unroll
    for (int c = 0; c < 4; ++c) sa_st(out, ((long long)b * a.N + n) * a.out_ld + h * a.d + tx + 16 * c, o[r][c] * inv);
  }
}

// delta[row] = sum_e dO[row][e] * O[row][e] for the 64 rows of a tile -> shared array
template <typename T>
__device__ __forceinline__ void la_delta(const LaArgs& a, float* delta, const T* __restrict__ out,
                                         const T* __restrict__ dout, int b, int h, int n0, int t) {
  const int row = t >> 2, l = t & 3;
  const int n = n0 + row;
  float part = 0.f;
  if (n < a.N) {
    const long long base = ((long long)b * a.N + n) * a.out_ld + h * a.d;
    for (int e = l; e < a.d; e += 4) part = fmaf(sa_ld(dout, base + e), sa_ld(out, base + e), part);
  }
  part += __shfl_xor_sync(0xffffffffu, part, 1);
  part += __shfl_xor_sync(0xffffffffu, part, 2);
  if (l == 0) delta[row] = part;
}

// shared pieces of the two backward kernels: P and dS of a (query tile, key tile) pair in registers
template <typename T>
__device__ __forceinline__ void la_p_ds(const LaArgs& a, const float* Qs, const float* Ks, const float* Vs,
                                        const float* dOs, const float* lse_s, const float* delta, int i0, int j0,
                                        int ty, int tx, float (&p)[4][4], float (&ds)[4][4]) {
  float s[4][4], dp[4][4];
  sa_tile_zero(s);
  sa_tile_zero(dp);
  sa_tile_mma<4, 4>(s, Qs, LD, 1, Ks, 1, LD, a.d, ty, tx);
  sa_tile_mma<4, 4>(dp, dOs, LD, 1, Vs, 1, LD, a.d, ty, tx);
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const int il = ty + 16 * r, i = i0 + il;
    const int lo = la_lo(i, a.W);
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const int j = j0 + tx + 16 * c;
      const bool ok = i < a.N && j <= i && j >= lo && j < a.N;
      p[r][c] = ok ? expf(s[r][c] - lse_s[il]) : 0.f;
      ds[r][c] = p[r][c] * (dp[r][c] - delta[il]);
    }
  }
}

template <typename T>
__global__ void __launch_bounds__(256)
local_bwd_dq_kernel(LaArgs a, const T* __restrict__ q, const T* __restrict__ k, const T* __restrict__ v,
                    const float* __restrict__ inv_freq, const T* __restrict__ out, const T* __restrict__ dout,
                    const float* __restrict__ lse, T* __restrict__ dq) {
  extern __shared__ float sm[];
  float* Qs = sm;
  float* Ks = Qs + LT * LD;
  float* Vs = Ks + LT * LD;
  float* dOs = Vs + LT * LD;
  float* Ss = dOs + LT * LD;
  float* lse_s = Ss + LT * LD;
  float* delta = lse_s + LT;
  const int t = threadIdx.x, ty = t >> 4, tx = t & 15;
  const int bh = blockIdx.y, b = bh / a.H, h = bh % a.H;
  const int i0 = blockIdx.x * LT;
  const bool rot = inv_freq != nullptr;
  la_load<T>(a, Qs, q, a.ld, b, h, i0, inv_freq, rot, a.scale, t);
  la_load<T>(a, dOs, dout, a.out_ld, b, h, i0, nullptr, false, 1.0f, t);
  la_delta<T>(a, delta, out, dout, b, h, i0, t);
  if (t < LT) lse_s[t] = (i0 + t < a.N) ? lse[(long long)bh * a.N + i0 + t] : 0.f;
  float dqa[4][4];
  sa_tile_zero(dqa);
  const int j_beg = (la_lo(i0, a.W) / LT) * LT;
  const int j_last = min(a.N - 1, i0 + LT - 1);
  for (int j0 = j_beg; j0 <= j_last; j0 += LT) {
    __syncthreads();
    la_load<T>(a, Ks, k, a.ld, b, h, j0, inv_freq, rot, 1.0f, t);
    la_load<T>(a, Vs, v, a.ld, b, h, j0, nullptr, false, 1.0f, t);
    __syncthreads();
    float p[4][4], ds[4][4];
    la_p_ds<T>(a, Qs, Ks, Vs, dOs, lse_s, delta, i0, j0, ty, tx, p, ds);
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
      for (int c = 0; c < 4; ++c) Ss[(ty + 16 * r) * LD + tx + 16 * c] = ds[r][c];
    __syncthreads();
    sa_tile_mma<4, 4>(dqa, Ss, LD, 1, Ks, LD, 1, LT, ty, tx);     // dS . k_rot
  }
  la_store_unrot<T>(a, Ss, dqa, dq, b, h, i0, inv_freq, rot, a.scale, t, ty, tx);
}

template <typename T>
__global__ void __launch_bounds__(256)
local_bwd_dkv_kernel(LaArgs a, const T* __restrict__ q, const T* __restrict__ k, const T* __restrict__ v,
                     const float* __restrict__ inv_freq, const T* __restrict__ out, const T* __restrict__ dout,
                     const float* __restrict__ lse, T* __restrict__ dk, T* __restrict__ dv) {
  extern __shared__ float sm[];
  float* Qs = sm;
  float* Ks = Qs + LT * LD;
  float* Vs = Ks + LT * LD;
  float* dOs = Vs + LT * LD;
  float* Ss = dOs + LT * LD;     // dS
  float* Ps = Ss + LT * LD;      // P
  float* lse_s = Ps + LT * LD;
  float* delta = lse_s + LT;
  const int t = threadIdx.x, ty = t >> 4, tx = t & 15;
  const int bh = blockIdx.y, b = bh / a.H, h = bh % a.H;
  const int j0 = blockIdx.x * LT;
  const bool rot = inv_freq != nullptr;
  la_load<T>(a, Ks, k, a.ld, b, h, j0, inv_freq, rot, 1.0f, t);
  la_load<T>(a, Vs, v, a.ld, b, h, j0, nullptr, false, 1.0f, t);
  float dka[4][4], dva[4][4];
  sa_tile_zero(dka);
  sa_tile_zero(dva);
  // the last key of the tile is visible up to the end of the window after its own
  const int j_hi = min(a.N - 1, j0 + LT - 1);
  const int i_last = min(a.N - 1, (j_hi / a.W + 2) * a.W - 1);
  for (int i0 = j0; i0 <= i_last; i0 += LT) {
    __syncthreads();
    la_load<T>(a, Qs, q, a.ld, b, h, i0, inv_freq, rot, a.scale, t);
    la_load<T>(a, dOs, dout, a.out_ld, b, h, i0, nullptr, false, 1.0f, t);
    la_delta<T>(a, delta, out, dout, b, h, i0, t);
    if (t < LT) lse_s[t] = (i0 + t < a.N) ? lse[(long long)bh * a.N + i0 + t] : 0.f;
    __syncthreads();
    float p[4][4], ds[4][4];
    la_p_ds<T>(a, Qs, Ks, Vs, dOs, lse_s, delta, i0, j0, ty, tx, p, ds);
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        Ss[(ty + 16 * r) * LD + tx + 16 * c] = ds[r][c];
        Ps[(ty + 16 * r) * LD + tx + 16 * c] = p[r][c];
      }
    __syncthreads();
    sa_tile_mma<4, 4>(dva, Ps, 1, LD, dOs, LD, 1, LT, ty, tx);    // P^T . dO
    sa_tile_mma<4, 4>(dka, Ss, 1, LD, Qs, LD, 1, LT, ty, tx);     // dS^T . (scale * q_rot)
  }
  la_store_unrot<T>(a, Ss, dka, dk, b, h, j0, inv_freq, rot, 1.0f, t, ty, tx);
  la_store_unrot<T>(a, Ss, dva, dv, b, h, j0, nullptr, false, 1.0f, t, ty, tx);
}

LaArgs make_la(const sa_local_desc* d) {
  LaArgs a;
  a.B = d->batch; a.N = d->seq; a.H = d->heads; a.d = d->dim_head; a.W = d->window; a.ld = d->ld; a.out_ld = d->out_ld;
  a.scale = 1.0f / sqrtf((float)d->dim_head);
  return a;
}

int check_local(const sa_local_desc* d) {
  SA_CHECK_ARG(d != nullptr, "null descriptor");
  SA_CHECK_ARG(d->batch > 0 && d->seq > 0 && d->heads > 0 && d->window > 0, "bad sizes");
  SA_CHECK_ARG(d->act_dtype == SA_F32 || d->act_dtype == SA_BF16, "bad dtype");
  SA_UNSUPPORTED(d->dim_head != 64, "dim_head != 64");
  SA_UNSUPPORTED((long long)d->batch * d->heads > 65535, "batch * heads > 65535");
  return SA_OK;
}

template <typename K>
int set_smem(K kernel, size_t bytes) {
  if (bytes > 48 * 1024) SA_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
  return SA_OK;
}

}  // namespace

int sa_simt_local_attn_fwd(const sa_local_desc* d, const void* q, const void* k, const void* v, const float* inv_freq,
                           void* out, float* lse, cudaStream_t st) {
  int rc = check_local(d);
  if (rc != SA_OK) return rc;
  sa_note_path(SA_PATH_SIMT);
  const LaArgs a = make_la(d);
  dim3 grid((unsigned)sa_cdiv(d->seq, LT), (unsigned)(d->batch * d->heads));
  const size_t smem = sizeof(float) * 4 * LT * LD;
  if (d->act_dtype == SA_F32) {
    if ((rc = set_smem(local_fwd_kernel<float>, smem)) != SA_OK) return rc;
    local_fwd_kernel<float><<<grid, 256, smem, st>>>(a, (const float*)q, (const float*)k, (const float*)v, inv_freq,
                                                     (float*)out, lse);
  } else {
    using B = __nv_bfloat16;
    if ((rc = set_smem(local_fwd_kernel<B>, smem)) != SA_OK) return rc;
    local_fwd_kernel<B><<<grid, 256, smem, st>>>(a, (const B*)q, (const B*)k, (const B*)v, inv_freq, (B*)out, lse);
  }
  SA_LAUNCH_CHECK();
  return SA_OK;
}

int sa_simt_local_attn_bwd(const sa_local_desc* d, const void* q, const void* k, const void* v, const float* inv_freq,
                           const void* out, const void* dout, const float* lse, void* dq, void* dk, void* dv,
                           cudaStream_t st) {
  int rc = check_local(d);
  if (rc != SA_OK) return rc;
  sa_note_path(SA_PATH_SIMT);
  const LaArgs a = make_la(d);
  dim3 grid((unsigned)sa_cdiv(d->seq, LT), (unsigned)(d->batch * d->heads));
  const size_t smem_q = sizeof(float) * (5 * LT * LD + 2 * LT);
  const size_t smem_kv = sizeof(float) * (6 * LT * LD + 2 * LT);
#define SA_LA_BWD(T)                                                                                               \
  do {                                                                                                             \
    if ((rc = set_smem(local_bwd_dq_kernel<T>, smem_q)) != SA_OK) return rc;                                       \
    local_bwd_dq_kernel<T><<<grid, 256, smem_q, st>>>(a, (const T*)q, (const T*)k, (const T*)v, inv_freq,          \
                                                      (const T*)out, (const T*)dout, lse, (T*)dq);                 \
    SA_LAUNCH_CHECK();                                                                                             \
    if ((rc = set_smem(local_bwd_dkv_kernel<T>, smem_kv)) != SA_OK) return rc;                                     \
    local_bwd_dkv_kernel<T><<<grid, 256, smem_kv, st>>>(a, (const T*)q, (const T*)k, (const T*)v, inv_freq,        \
                                                        (const T*)out, (const T*)dout, lse, (T*)dk, (T*)dv);       \
    SA_LAUNCH_CHECK();                                                                                             \
  } while (0)
  if (d->act_dtype == SA_F32) SA_LA_BWD(float); else SA_LA_BWD(__nv_bfloat16);
#undef SA_LA_BWD
  return SA_OK;
}
