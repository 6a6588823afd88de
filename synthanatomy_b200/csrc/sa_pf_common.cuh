// Shared device helpers of the Performer kernels (sm_100a).
#pragma once
#include "sa_common.cuh"
#include "../../include/synthanatomy_b200_performer.h"

// ------------------------------------------------------------------------------------------------
// GEMM epilogue shared by the CUDA-core and the tcgen05 kernels (see sa_gemm_nt in the header)
// ------------------------------------------------------------------------------------------------
struct SaEpi {
  const float* bias;
  const void* dot_with;
  float* dot_out;
  const float* scale_dev;
  float scale;
  int act;
  void* pre;
  const float* resid;
  float* out_f32;
  void* out_act;
  long long ldo;
};

static inline SaEpi sa_make_epi(const sa_gemm_epilogue* e, int64_t ldo) {
  SaEpi p;
  p.bias = e->bias; p.dot_with = e->dot_with; p.dot_out = e->dot_out; p.scale_dev = e->scale_dev; p.scale = e->scale;
  p.act = e->act; p.pre = e->pre; p.resid = e->resid; p.out_f32 = e->out_f32; p.out_act = e->out_act; p.ldo = ldo;
  return p;
}

#ifdef __CUDACC__
__device__ __forceinline__ float sa_gelu(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752f)); }
__device__ __forceinline__ float sa_gelu_grad(float x) {
  return 0.5f * (1.0f + erff(x * 0.70710678118654752f)) + x * 0.39894228040143268f * __expf(-0.5f * x * x);
}

// one element of the epilogue; `st` = scale * scale_dev[0] (hoisted by the caller); returns the dot-product term
template <typename T>
__device__ __forceinline__ float sa_epi_elem(const SaEpi& e, long long i, int j, float v, float st) {
  const long long o = i * e.ldo + j;
  float dot = 0.f;
  if (e.bias) v += __ldg(e.bias + j);
  if (e.dot_with) dot = v * sa_ld(reinterpret_cast<const T*>(e.dot_with), o);
  v *= st;
  if (e.act == SA_ACT_GELU_FWD) {
    sa_st(reinterpret_cast<T*>(e.pre), o, v);
    v = sa_gelu(v);
  } else if (e.act == SA_ACT_GELU_FWD_D) {
    sa_st(reinterpret_cast<T*>(e.pre), o, sa_gelu_grad(v));
    v = sa_gelu(v);
  } else if (e.act == SA_ACT_GELU_BWD) {
    v *= sa_gelu_grad(sa_ld(reinterpret_cast<const T*>(e.pre), o));
  } else if (e.act == SA_ACT_MUL_PRE) {
    v *= sa_ld(reinterpret_cast<const T*>(e.pre), o);
  }
  if (e.resid) v += e.resid[o];
  if (e.out_f32) e.out_f32[o] = v;
  if (e.out_act) sa_st(reinterpret_cast<T*>(e.out_act), o, v);
  return dot;
}

// ------------------------------------------------------------------------------------------------
// 16 x 16 thread tile product on shared-memory operands (CUDA-core fp32 path of the attention kernels):
//   acc[r][s] += sum_{k < K} A(ty + 16 r, k) * B(k, tx + 16 s)
// A(i, k) = A[i * a_is + k * a_ks], B(k, j) = B[k * b_ks + j * b_js].  A reads are warp broadcasts (2 distinct
// rows per warp); B reads are conflict-free when b_js == 1 or b_js is odd.
// ------------------------------------------------------------------------------------------------
template <int R, int S>
__device__ __forceinline__ void sa_tile_mma(float (&acc)[R][S], const float* __restrict__ A, int a_is, int a_ks,
                                            const float* __restrict__ B, int b_ks, int b_js, int K, int ty, int tx) {
#pragma unroll 4
  for (int k = 0; k < K; ++k) {
    float a[R], b[S];
#pragma unroll
    for (int r = 0; r < R; ++r) a[r] = A[(ty + 16 * r) * a_is + k * a_ks];
#pragma unroll
    for (int s = 0; s < S; ++s) b[s] = B[k * b_ks + (tx + 16 * s) * b_js];
#pragma unroll
    for (int r = 0; r < R; ++r)
#pragma unroll
      for (int s = 0; s < S; ++s) acc[r][s] = fmaf(a[r], b[s], acc[r][s]);
  }
}

template <int R, int S>
__device__ __forceinline__ void sa_tile_zero(float (&acc)[R][S]) {
#pragma unroll
  for (int r = 0; r < R; ++r)
#pragma unroll
    for (int s = 0; s < S; ++s) acc[r][s] = 0.f;
}

// reductions across the 16 lanes that share `ty` (tx = lane & 15)
__device__ __forceinline__ float sa_half_max(float v) {
#pragma unroll
  for (int o = 8; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ float sa_half_sum(float v) {
#pragma unroll
  for (int o = 8; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
#endif
