// CUDA-core fp32-accumulate GEMMs of the Performer path: the fp32 "parity" path and the fallback for shapes
// the tcgen05 kernels do not take.  64 x 64 x 16 tiles, 256 threads, 4 x 4 outputs per thread.
//
// Replaces cuBLAS behind torch.nn.Linear / autograd at performer-pytorch SelfAttention.to_{q,k,v,out},
// FeedForward.w1/w2 and /root/reference/src/networks/transformers/performer.py:221,286 (to_out).
#include "sa_pf_common.cuh"

namespace {

constexpr int GB = 64;    // tile edge
constexpr int GK = 16;    // k step
constexpr int GP = GB + 4;

template <typename T>
__global__ void __launch_bounds__(256)
gemm_nt_kernel(long long m, int n, int k, const T* __restrict__ A, long long lda, const T* __restrict__ B,
               long long ldb, SaEpi e) {
  __shared__ __align__(16) float As[GK][GP];
  __shared__ __align__(16) float Bs[GK][GP];
  __shared__ float s_dot[8];
  const int t = threadIdx.x;
  const int ty = t >> 4, tx = t & 15;
  const long long i0 = (long long)blockIdx.y * GB;
  const int j0 = blockIdx.x * GB;
  float acc[4][4];
  sa_tile_zero(acc);
  const int lk = t & 15, lr = t >> 4;   // load lanes: 16 consecutive k per row, 16 rows per pass
  for (int k0 = 0; k0 < k; k0 += GK) {
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const int row = lr + 16 * r;
      const long long gi = i0 + row;
      const int gj = j0 + row;
      const int gk = k0 + lk;
      As[lk][row] = (gi < m && gk < k) ? sa_ld(A, gi * lda + gk) : 0.f;
      Bs[lk][row] = (gj < n && gk < k) ? sa_ld(B, (long long)gj * ldb + gk) : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < GK; ++kk) {
      const float4 a = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
      const float4 b = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
      const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int s = 0; s < 4; ++s) acc[r][s] = fmaf(av[r], bv[s], acc[r][s]);
    }
    __syncthreads();
  }
  const float st = e.scale * (e.scale_dev ? __ldg(e.scale_dev) : 1.0f);
  float dot = 0.f;
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const long long gi = i0 + ty * 4 + r;
    if (gi >= m) continue;
#pragma unroll
    for (int s = 0; s < 4; ++s) {
      const int gj = j0 + tx * 4 + s;
      if (gj < n) dot += sa_epi_elem<T>(e, gi, gj, acc[r][s], st);
    }
  }
  if (e.dot_out) {
    dot = sa_warp_sum(dot);
    if ((t & 31) == 0) s_dot[t >> 5] = dot;
    __syncthreads();
    if (t == 0) {
      float s = 0.f;
      for (int w = 0; w < 8; ++w) s += s_dot[w];
      atomicAdd(e.dot_out, s);
    }
  }
}

// D[i][j] += scale * sum_r A[r][i] B[r][j]; blockIdx.z = split over r
template <typename T>
__global__ void __launch_bounds__(256)
gemm_tn_kernel(long long m, int na, int nb, const T* __restrict__ A, long long lda, const T* __restrict__ B,
               long long ldb, const float* __restrict__ scale_dev, float scale, float* __restrict__ D,
               long long rows_per_split, unsigned* turn) {
  __shared__ __align__(16) float As[GK][GP];
  __shared__ __align__(16) float Bs[GK][GP];
  const int t = threadIdx.x;
  const int ty = t >> 4, tx = t & 15;
  const int i0 = blockIdx.y * GB, j0 = blockIdx.x * GB;
  const long long r_beg = (long long)blockIdx.z * rows_per_split;
  const long long r_end = min(m, r_beg + rows_per_split);
  float acc[4][4];
  sa_tile_zero(acc);
  const int lc = (t & 15) * 4, lr = t >> 4;   // 4 consecutive columns, 16 rows per pass
  for (long long r0 = r_beg; r0 < r_end; r0 += GK) {
    const long long gr = r0 + lr;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const int gi = i0 + lc + c, gj = j0 + lc + c;
      As[lr][lc + c] = (gr < r_end && gi < na) ? sa_ld(A, gr * lda + gi) : 0.f;
      Bs[lr][lc + c] = (gr < r_end && gj < nb) ? sa_ld(B, gr * ldb + gj) : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < GK; ++kk) {
      const float4 a = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
      const float4 b = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
      const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int s = 0; s < 4; ++s) acc[r][s] = fmaf(av[r], bv[s], acc[r][s]);
    }
    __syncthreads();
  }
  const float st = scale * (scale_dev ? __ldg(scale_dev) : 1.0f);
  unsigned* my_turn = turn ? turn + blockIdx.y * gridDim.x + blockIdx.x : nullptr;     // deterministic mode: split order
  sa_block_turn_begin(my_turn, blockIdx.z);
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const int gi = i0 + ty * 4 + r;
    if (gi >= na) continue;
#pragma unroll
    for (int s = 0; s < 4; ++s) {
      const int gj = j0 + tx * 4 + s;
      if (gj < nb) atomicAdd(D + (long long)gi * nb + gj, st * acc[r][s]);
    }
  }
  sa_block_turn_end(my_turn, blockIdx.z);
}

}  // namespace

int sa_simt_gemm_nt(int64_t m, int n, int k, int dtype, const void* a, int64_t lda, const void* b, int64_t ldb,
                    const SaEpi& e, cudaStream_t st) {
  sa_note_path(SA_PATH_SIMT);
  dim3 grid((unsigned)sa_cdiv(n, GB), (unsigned)sa_cdiv(m, GB));
  SA_UNSUPPORTED(grid.y > 65535, "gemm_nt: more than 65535 row tiles");
  if (dtype == SA_F32)
    gemm_nt_kernel<float><<<grid, 256, 0, st>>>(m, n, k, (const float*)a, lda, (const float*)b, ldb, e);
  else
    gemm_nt_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>(m, n, k, (const __nv_bfloat16*)a, lda, (const __nv_bfloat16*)b,
                                                        ldb, e);
  SA_LAUNCH_CHECK();
  return SA_OK;
}

int sa_simt_gemm_tn(int64_t m, int na, int nb, int dtype, const void* a, int64_t lda, const void* b, int64_t ldb,
                    const float* scale_dev, float scale, float* d, cudaStream_t st) {
  sa_note_path(SA_PATH_SIMT);
  const int64_t tiles = sa_cdiv(na, GB) * sa_cdiv(nb, GB);
  int64_t splits = sa_cdiv(148 * 4, tiles);
  const int64_t max_splits = sa_cdiv(m, 256);
  if (splits > max_splits) splits = max_splits;
  if (splits < 1) splits = 1;
  int64_t rps = sa_cdiv(sa_cdiv(m, splits), GK) * GK;
  splits = sa_cdiv(m, rps);
  dim3 grid((unsigned)sa_cdiv(nb, GB), (unsigned)sa_cdiv(na, GB), (unsigned)splits);
  unsigned* turn = sa_turn_slot((int)(grid.x * grid.y), st);
  if (dtype == SA_F32)
    gemm_tn_kernel<float><<<grid, 256, 0, st>>>(m, na, nb, (const float*)a, lda, (const float*)b, ldb, scale_dev, scale,
                                                d, rps, turn);
  else
    gemm_tn_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>(m, na, nb, (const __nv_bfloat16*)a, lda,
                                                        (const __nv_bfloat16*)b, ldb, scale_dev, scale, d, rps, turn);
  SA_LAUNCH_CHECK();
  return SA_OK;
}

// ------------------------------------------------------------------------------------------------
// "Skinny" NT GEMM for a handful of rows (m <= 8): the decoding step of autoregressive sampling multiplies one token per
// batch element with every weight matrix.  Weight-bandwidth bound: one warp per output column streams that weight row
// with 16-byte loads (the few A rows stay in L1), fp32 accumulation, warp reduction, then lane i finishes row i with the
// shared epilogue.
// ------------------------------------------------------------------------------------------------
namespace {

constexpr int SK_MAXM = 8;

__device__ __forceinline__ void sk_load8(const float* p, float (&f)[8]) {
  const float4 a = __ldg(reinterpret_cast<const float4*>(p)), b = __ldg(reinterpret_cast<const float4*>(p) + 1);
  f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w; f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
}
__device__ __forceinline__ void sk_load8(const __nv_bfloat16* p, float (&f)[8]) {
  const uint4 u = __ldg(reinterpret_cast<const uint4*>(p));
  f[0] = __uint_as_float(u.x << 16); f[1] = __uint_as_float(u.x & 0xffff0000u);
  f[2] = __uint_as_float(u.y << 16); f[3] = __uint_as_float(u.y & 0xffff0000u);
  f[4] = __uint_as_float(u.z << 16); f[5] = __uint_as_float(u.z & 0xffff0000u);
  f[6] = __uint_as_float(u.w << 16); f[7] = __uint_as_float(u.w & 0xffff0000u);
}

// ksplit (1, 2, 4 or 8) warps share one output column (interleaved K slices, partial sums through shared memory), so that
// narrow outputs (n = 512) still fill the machine: 8 / ksplit columns per CTA.
template <typename T>
__global__ void __launch_bounds__(256)
skinny_gemm_nt_kernel(int m, int n, int k, int ksplit, const T* __restrict__ A, long long lda, const T* __restrict__ B,
                      long long ldb, SaEpi e) {
  __shared__ float s_part[8][SK_MAXM];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int cpc = 8 / ksplit;
  const int cl = warp / ksplit, kp = warp - cl * ksplit;
  const int j = blockIdx.x * cpc + cl;
  const bool col_ok = j < n;
  float acc[SK_MAXM];
#pragma unroll
  for (int i = 0; i < SK_MAXM; ++i) acc[i] = 0.f;
  if (col_ok) {
    const T* brow = B + (long long)j * ldb;
    const int kstep = ksplit * 256;
    int k0 = (kp * 32 + lane) * 8;
    for (; k0 + kstep < k; k0 += 2 * kstep) {           // two weight vectors in flight
      float w0[8], w1[8];
      sk_load8(brow + k0, w0);
      sk_load8(brow + k0 + kstep, w1);
#pragma unroll
      for (int i = 0; i < SK_MAXM; ++i) {
        if (i < m) {
          float a0[8], a1[8];
          sk_load8(A + (long long)i * lda + k0, a0);
          sk_load8(A + (long long)i * lda + k0 + kstep, a1);
#pragma unroll
          for (int q = 0; q < 8; ++q) acc[i] = fmaf(a0[q], w0[q], acc[i]);
#pragma unroll
          for (int q = 0; q < 8; ++q) acc[i] = fmaf(a1[q], w1[q], acc[i]);
        }
      }
    }
    for (; k0 < k; k0 += kstep) {
      float w[8];
      sk_load8(brow + k0, w);
#pragma unroll
      for (int i = 0; i < SK_MAXM; ++i) {
        if (i < m) {
          float a[8];
          sk_load8(A + (long long)i * lda + k0, a);
#pragma unroll
          for (int q = 0; q < 8; ++q) acc[i] = fmaf(a[q], w[q], acc[i]);
        }
      }
    }
  }
#pragma unroll
  for (int i = 0; i < SK_MAXM; ++i) acc[i] = sa_warp_sum(acc[i]);
  float mine = 0.f;
#pragma unroll
  for (int i = 0; i < SK_MAXM; ++i) mine = (lane == i) ? acc[i] : mine;
  if (ksplit > 1) {
    if (lane < SK_MAXM) s_part[warp][lane] = mine;
    __syncthreads();
    if (kp == 0 && lane < SK_MAXM) {
      mine = 0.f;
      for (int p = 0; p < ksplit; ++p) mine += s_part[warp + p][lane];
    }
  }
  const float st = e.scale * (e.scale_dev ? __ldg(e.scale_dev) : 1.0f);
  float dot = 0.f;
  if (kp == 0 && col_ok && lane < m) dot = sa_epi_elem<T>(e, lane, j, mine, st);
  if (e.dot_out && kp == 0) {
    dot = sa_warp_sum(dot);
    if (lane == 0 && col_ok) atomicAdd(e.dot_out, dot);
  }
}

}  // namespace

bool sa_skinny_gemm_nt_supported(int64_t m, int n, int k, int dtype, const void* a, int64_t lda, const void* b, int64_t ldb) {
  if (m < 1 || m > SK_MAXM || n < 1 || k < 8 || (k & 7)) return false;
  const int64_t al = dtype == SA_BF16 ? 8 : 4;      // elements per 16 bytes
  if ((lda % al) || (ldb % al)) return false;
  if ((reinterpret_cast<uintptr_t>(a) & 15) || (reinterpret_cast<uintptr_t>(b) & 15)) return false;
  return true;
}

int sa_skinny_gemm_nt(int64_t m, int n, int k, int dtype, const void* a, int64_t lda, const void* b, int64_t ldb,
                      const SaEpi& e, cudaStream_t st) {
  sa_note_path(SA_PATH_SIMT);
  // enough CTAs for ~2 per SM: narrow outputs split K over the warps of a CTA
  int ksplit = 1;
  while (ksplit < 8 && sa_cdiv(n, 8 / ksplit) < 296 && k >= 512 * ksplit) ksplit *= 2;
  const unsigned grid = (unsigned)sa_cdiv(n, 8 / ksplit);
  if (dtype == SA_BF16)
    skinny_gemm_nt_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>((int)m, n, k, ksplit, (const __nv_bfloat16*)a, lda,
                                                               (const __nv_bfloat16*)b, ldb, e);
  else
    skinny_gemm_nt_kernel<float><<<grid, 256, 0, st>>>((int)m, n, k, ksplit, (const float*)a, lda, (const float*)b, ldb, e);
  SA_LAUNCH_CHECK();
  return SA_OK;
}
