// Shared helpers for the synthanatomy_b200 CUDA library (sm_100a only).
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/synthanatomy_b200.h"

// ------------------------------------------------------------------------------------------------
// host-side status plumbing
// ------------------------------------------------------------------------------------------------
void sa_set_error(const char* fmt, ...);
void sa_note_launch(int n = 1);
void sa_note_path(int path);
bool sa_force_simt();

#define SA_CHECK_ARG(cond, msg)                                   \
  do {                                                            \
    if (!(cond)) {                                                \
      sa_set_error("%s: invalid argument: %s", __func__, msg);    \
      return SA_ERR_INVALID;                                      \
    }                                                             \
  } while (0)

#define SA_UNSUPPORTED(cond, msg)                                 \
  do {                                                            \
    if (cond) {                                                   \
      sa_set_error("%s: unsupported: %s", __func__, msg);         \
      return SA_ERR_UNSUPPORTED;                                  \
    }                                                             \
  } while (0)

#define SA_CUDA(call)                                                                   \
  do {                                                                                  \
    cudaError_t e__ = (call);                                                           \
    if (e__ != cudaSuccess) {                                                           \
      sa_set_error("%s: CUDA error %s at %s:%d", __func__, cudaGetErrorString(e__),     \
                   __FILE__, __LINE__);                                                 \
      return SA_ERR_CUDA;                                                               \
    }                                                                                   \
  } while (0)

#define SA_LAUNCH_CHECK()                                                               \
  do {                                                                                  \
    cudaError_t e__ = cudaGetLastError();                                               \
    if (e__ != cudaSuccess) {                                                           \
      sa_set_error("%s: launch failed: %s at %s:%d", __func__, cudaGetErrorString(e__), \
                   __FILE__, __LINE__);                                                 \
      return SA_ERR_CUDA;                                                               \
    }                                                                                   \
    sa_note_launch();                                                                   \
  } while (0)

static inline cudaStream_t sa_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }
static inline size_t sa_dtype_size(int dt) { return dt == SA_BF16 ? 2 : 4; }
static inline int64_t sa_cdiv(int64_t a, int64_t b) { return (a + b - 1) / b; }

// ------------------------------------------------------------------------------------------------
// device-side dtype helpers
// ------------------------------------------------------------------------------------------------
#ifdef __CUDACC__
__device__ __forceinline__ float sa_ld(const float* p, int64_t i) { return p[i]; }
__device__ __forceinline__ float sa_ld(const __nv_bfloat16* p, int64_t i) { return __bfloat162float(p[i]); }
__device__ __forceinline__ void sa_st(float* p, int64_t i, float v) { p[i] = v; }
__device__ __forceinline__ void sa_st(__nv_bfloat16* p, int64_t i, float v) { p[i] = __float2bfloat16_rn(v); }

__device__ __forceinline__ float sa_warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
#endif
