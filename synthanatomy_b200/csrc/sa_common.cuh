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
// Deterministic mode (sa_set_deterministic): every cross-CTA floating-point accumulation (split-K partials, column sums,
// loss sums) runs through a "turnstile" -- a zeroed device counter per output region; the CTA that holds partial number s
// waits until the counter reads s, adds its partial, and passes the turn on -- so the additions happen in one fixed order
// and a step is bit-reproducible run to run.  sa_turn_slot returns nullptr when the mode is off (the kernels then add in
// arrival order with plain atomics), else `n` zeroed counters valid for ONE launch on `st`.
bool sa_deterministic();
unsigned* sa_turn_slot(int n, cudaStream_t st);
// true (once) when a deterministic-mode buffer could not be provided since the last call on this thread: the launch that
// asked for it would have fallen back to arrival-order atomics, so SA_LAUNCH_CHECK turns it into an error instead
bool sa_det_alloc_failed();
// scalar sums (losses, gate / stabiliser gradients) in deterministic mode: every contributor writes its partial into its
// own word of a zeroed slot (sa_partial_slot; nullptr when the mode is off) and sa_ordered_sum adds them to out[0] in
// index order (one CTA, fixed tree)
float* sa_partial_slot(int n, cudaStream_t st);
int sa_ordered_sum(const float* partials, int n, float* out, cudaStream_t st);
// tensor-valued sums (split-K weight gradients, column sums) in deterministic mode: contributor p stores its partial --
// same layout as the output -- at parts + p * stride (stream-ordered allocation, nullptr when the mode is off), and
// sa_parts_reduce adds out[j] += sum_p parts[p * stride + j] (ascending p) for j < width; sa_parts_free releases it.
float* sa_parts_alloc(int64_t nparts, int64_t stride, cudaStream_t st);
int sa_parts_reduce(const float* parts, int64_t nparts, int64_t stride, int64_t width, float* out, cudaStream_t st);
int sa_parts_free(float* parts, cudaStream_t st);

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
    if (sa_det_alloc_failed()) {                                                        \
      sa_set_error("%s: deterministic mode: could not allocate the ordered-sum buffers", __func__); \
      return SA_ERR_CUDA;                                                               \
    }                                                                                   \
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

// turnstile (deterministic mode): see sa_turn_slot above.  `ctr` may be nullptr (mode off: no-ops).
__device__ __forceinline__ void sa_turn_wait(const unsigned* ctr, unsigned my) {
  if (ctr == nullptr) return;
  unsigned v;
  for (unsigned spins = 0;; ++spins) {
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(ctr) : "memory");
    if (v == my) break;
    if (spins > (1u << 26)) __trap();      // seconds: a lost turn is a bug, not a wait
    __nanosleep(40);
  }
}
// the caller has made the partial's additions visible (__threadfence) and joined its threads (barrier) before this
__device__ __forceinline__ void sa_turn_pass(unsigned* ctr, unsigned my) {
  if (ctr == nullptr) return;
  asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(ctr), "r"(my + 1u) : "memory");
}
// whole-block forms for the CUDA-core kernels (every thread of the block must reach them)
__device__ __forceinline__ void sa_block_turn_begin(const unsigned* ctr, unsigned my) {
  if (ctr == nullptr) return;
  if (threadIdx.x == 0) sa_turn_wait(ctr, my);
  __syncthreads();
}
__device__ __forceinline__ void sa_block_turn_end(unsigned* ctr, unsigned my) {
  if (ctr == nullptr) return;
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) sa_turn_pass(ctr, my);
}

// fp32 reduction of four consecutive, 16-byte-aligned words in one instruction (REDG.ADD.F32x4): a quarter of the
// instructions and L2 transactions of four scalar atomicAdds in the split-K epilogues (thread = row, 32+ columns each)
__device__ __forceinline__ void sa_red_add_v4(float* p, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

__device__ __forceinline__ float sa_warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
#endif
