// C-ABI entry points that dispatch between the tcgen05 kernels and the CUDA-core kernels, plus the
// thread-local status plumbing.  See include/synthanatomy_b200.h for the contract.
#include <stdarg.h>

#include <atomic>
#include <mutex>

#include "sa_tc_common.cuh"

// implemented in the kernel translation units
int sa_simt_conv3d_fwd(const sa_conv_desc*, const void*, const void*, const float*, const void*, const void*, int,
                       void*, cudaStream_t);
int sa_simt_conv3d_wgrad(const sa_conv_desc*, const void*, const void*, float*, cudaStream_t);
bool sa_tc_conv3d_supported(const sa_conv_desc*);
bool sa_tc_wgrad3_supported(const sa_conv_desc*);
int sa_tc_conv3d_wgrad3(const sa_conv_desc*, const void*, const void*, float*, cudaStream_t);
bool sa_tc_conv3_supported(const sa_conv_desc*);
int sa_tc_conv3_fwd(const sa_conv_desc*, const void*, const void*, const float*, const void*, const void*, int, void*,
                    cudaStream_t);
int sa_tc_conv3d_fwd(const sa_conv_desc*, const void*, const void*, const float*, const void*, const void*, int, void*,
                     cudaStream_t);
bool sa_tc_wgrad_supported(const sa_conv_desc*);
int sa_tc_conv3d_wgrad(const sa_conv_desc*, const void*, const void*, float*, cudaStream_t);

namespace {
thread_local char g_err[512] = "";
thread_local int g_path = SA_PATH_NONE;
// process-wide (not per thread): autograd runs the backward pass of CUDA graphs on its own device thread, and a step's
// launch count must include it
std::atomic<int64_t> g_launches{0};
int g_force_simt = 0;
int g_deterministic = 0;

// turn counters of the deterministic mode: one zero-initialised ring per device, a fresh slot per launch
constexpr size_t kTurnRing = 1u << 20;          // counters (4 MB)
struct TurnPool { unsigned* base = nullptr; size_t next = 0; };
TurnPool g_turn[64];
std::mutex g_turn_mu;
thread_local bool g_det_failed = false;
}  // namespace

bool sa_det_alloc_failed() {
  const bool f = g_det_failed;
  g_det_failed = false;
  return f;
}

bool sa_deterministic() { return g_deterministic != 0; }
extern "C" void sa_set_deterministic(int on) { g_deterministic = on; }
extern "C" int sa_get_deterministic(void) { return g_deterministic; }

unsigned* sa_turn_slot(int n, cudaStream_t st) {
  if (!g_deterministic || n <= 0) return nullptr;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) { g_det_failed = true; return nullptr; }
  const size_t want = ((size_t)n + 31) & ~(size_t)31;
  if (want > kTurnRing) { g_det_failed = true; return nullptr; }
  std::lock_guard<std::mutex> lock(g_turn_mu);
  TurnPool& pool = g_turn[dev];
  if (!pool.base) {
    if (cudaMalloc(&pool.base, kTurnRing * sizeof(unsigned)) != cudaSuccess) {
      pool.base = nullptr;
      cudaGetLastError();
      g_det_failed = true;
      return nullptr;
    }
  }
  if (pool.next + want > kTurnRing) pool.next = 0;
  unsigned* slot = pool.base + pool.next;
  pool.next += want;
  // zeroed on the launch's own stream: a slot comes round again only after 2^20 counters' worth of later launches
  if (cudaMemsetAsync(slot, 0, want * sizeof(unsigned), st) != cudaSuccess) { g_det_failed = true; return nullptr; }
  return slot;
}

float* sa_partial_slot(int n, cudaStream_t st) { return reinterpret_cast<float*>(sa_turn_slot(n, st)); }

namespace {
__global__ void __launch_bounds__(1024) ordered_sum_kernel(const float* __restrict__ part, int n, float* __restrict__ out) {
  __shared__ float s[1024];
  float acc = 0.f;
  for (int i = threadIdx.x; i < n; i += 1024) acc += part[i];
  s[threadIdx.x] = acc;
  __syncthreads();
  for (int o = 512; o > 0; o >>= 1) {
    if ((int)threadIdx.x < o) s[threadIdx.x] += s[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) out[0] += s[0];
}
}  // namespace

namespace {
__global__ void __launch_bounds__(256) parts_reduce_kernel(const float* __restrict__ parts, long long nparts, long long stride,
                                                            long long width, float* __restrict__ out) {
  for (long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x; j < width; j += (long long)gridDim.x * blockDim.x) {
    float acc = 0.f;
    for (long long p = 0; p < nparts; ++p) acc += parts[p * stride + j];
    out[j] += acc;
  }
}
}  // namespace

float* sa_parts_alloc(int64_t nparts, int64_t stride, cudaStream_t st) {
  if (!g_deterministic || nparts <= 0 || stride <= 0) return nullptr;
  void* p = nullptr;
  if (cudaMallocAsync(&p, (size_t)nparts * (size_t)stride * sizeof(float), st) != cudaSuccess) {
    cudaGetLastError();
    g_det_failed = true;
    return nullptr;
  }
  return (float*)p;
}

int sa_parts_reduce(const float* parts, int64_t nparts, int64_t stride, int64_t width, float* out, cudaStream_t st) {
  int64_t blocks = sa_cdiv(width, 256);
  if (blocks > 148 * 16) blocks = 148 * 16;
  parts_reduce_kernel<<<(unsigned)blocks, 256, 0, st>>>(parts, nparts, stride, width, out);
  SA_LAUNCH_CHECK();
  return SA_OK;
}

int sa_parts_free(float* parts, cudaStream_t st) {
  if (parts) SA_CUDA(cudaFreeAsync(parts, st));
  return SA_OK;
}

int sa_ordered_sum(const float* partials, int n, float* out, cudaStream_t st) {
  ordered_sum_kernel<<<1, 1024, 0, st>>>(partials, n, out);
  SA_LAUNCH_CHECK();
  return SA_OK;
}

void sa_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
void sa_note_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }
void sa_note_path(int path) { g_path = path; }
bool sa_force_simt() { return g_force_simt != 0; }

extern "C" const char* sa_last_error(void) { return g_err; }
extern "C" int sa_version(void) { return 100; }
extern "C" int sa_last_path(void) { return g_path; }
extern "C" int64_t sa_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }
extern "C" void sa_launch_count_reset(void) { g_launches.store(0, std::memory_order_relaxed); }
extern "C" void sa_set_force_simt(int on) { g_force_simt = on; }

// ------------------------------------------------------------------------------------------------
sa_tmap_encode_fn sa_get_tmap_encode() {
  static sa_tmap_encode_fn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<sa_tmap_encode_fn>(p);
  });
  return fn;
}

int sa_make_tmap(CUtensorMap* out, int dtype, const void* base, int rank, const uint64_t* dims,
                 const uint64_t* strides_bytes, const uint32_t* box) {
  sa_tmap_encode_fn enc = sa_get_tmap_encode();
  if (!enc) { sa_set_error("cuTensorMapEncodeTiled entry point not available"); return SA_ERR_CUDA; }
  cuuint64_t gdims[5], gstr[4];
  cuuint32_t gbox[5], estr[5];
  for (int i = 0; i < rank; ++i) { gdims[i] = dims[i]; gbox[i] = box[i]; estr[i] = 1; }
  for (int i = 1; i < rank; ++i) gstr[i - 1] = strides_bytes[i];
  if (reinterpret_cast<uintptr_t>(base) & 15) { sa_set_error("tensor map base not 16-byte aligned"); return SA_ERR_INVALID; }
  for (int i = 1; i < rank; ++i)
    if (strides_bytes[i] & 15) { sa_set_error("tensor map stride %d not a multiple of 16 bytes", i); return SA_ERR_INVALID; }
  const CUtensorMapDataType dt = dtype == SA_F32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
  CUresult r = enc(out, dt, (cuuint32_t)rank, const_cast<void*>(base), gdims, gstr, gbox,
                   estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    sa_set_error("cuTensorMapEncodeTiled failed with CUresult %d (rank %d dims %llu,%llu,%llu box %u,%u,%u)", (int)r, rank,
                 (unsigned long long)dims[0], (unsigned long long)dims[1], (unsigned long long)(rank > 2 ? dims[2] : 0),
                 box[0], box[1], rank > 2 ? box[2] : 0);
    return SA_ERR_CUDA;
  }
  return SA_OK;
}

int sa_make_tmap_bf16(CUtensorMap* out, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                      const uint32_t* box) {
  return sa_make_tmap(out, SA_BF16, base, rank, dims, strides_bytes, box);
}

// ------------------------------------------------------------------------------------------------
static int check_conv_desc(const sa_conv_desc* d) {
  SA_CHECK_ARG(d != nullptr, "null descriptor");
  SA_CHECK_ARG(d->batch > 0 && d->c_in > 0 && d->c_out > 0, "bad batch / channels");
  SA_CHECK_ARG(d->ksize > 0 && d->stride > 0 && d->pad >= 0, "bad kernel geometry");
  SA_CHECK_ARG(d->act_dtype == SA_F32 || d->act_dtype == SA_BF16, "bad dtype");
  for (int i = 0; i < 3; ++i) SA_CHECK_ARG(d->in_dhw[i] > 0 && d->out_dhw[i] > 0, "bad spatial extent");
  SA_UNSUPPORTED(d->ksize > 7, "ksize > 7");
  const int64_t in_pos = (int64_t)d->batch * d->in_dhw[0] * d->in_dhw[1] * d->in_dhw[2];
  const int64_t out_pos = (int64_t)d->batch * d->out_dhw[0] * d->out_dhw[1] * d->out_dhw[2];
  SA_UNSUPPORTED(in_pos * d->c_in >= (1LL << 40) || out_pos * d->c_out >= (1LL << 40), "tensor too large");
  return SA_OK;
}

extern "C" int sa_conv3d_fwd(const sa_conv_desc* d, const void* x, const void* wp, const float* bias, const void* addend,
                             const void* mask, int relu, void* y, void* stream) {
  int rc = check_conv_desc(d);
  if (rc != SA_OK) return rc;
  SA_CHECK_ARG(x && wp && y, "null pointer");
  cudaStream_t st = sa_stream(stream);
  if (!sa_force_simt() && sa_tc_conv3_supported(d)) return sa_tc_conv3_fwd(d, x, wp, bias, addend, mask, relu, y, st);
  if (!sa_force_simt() && sa_tc_conv3d_supported(d)) return sa_tc_conv3d_fwd(d, x, wp, bias, addend, mask, relu, y, st);
  return sa_simt_conv3d_fwd(d, x, wp, bias, addend, mask, relu, y, st);
}

extern "C" int sa_conv3d_wgrad(const sa_conv_desc* d, const void* p, const void* q, float* dwp, int accumulate,
                               void* stream) {
  int rc = check_conv_desc(d);
  if (rc != SA_OK) return rc;
  SA_CHECK_ARG(p && q && dwp, "null pointer");
  SA_UNSUPPORTED(d->transposed != 0, "wgrad is defined in FORM_CONV indexing only (swap P/Q for transposed convs)");
  cudaStream_t st = sa_stream(stream);
  if (!accumulate) {
    const size_t n = (size_t)d->ksize * d->ksize * d->ksize * d->c_out * d->c_in;
    SA_CUDA(cudaMemsetAsync(dwp, 0, n * sizeof(float), st));
  }
  const bool dw_aligned = (reinterpret_cast<uintptr_t>(dwp) & 15) == 0;      // the tcgen05 epilogues add 16-byte vectors
  if (!sa_force_simt() && dw_aligned && sa_tc_wgrad3_supported(d)) return sa_tc_conv3d_wgrad3(d, p, q, dwp, st);
  if (!sa_force_simt() && dw_aligned && sa_tc_wgrad_supported(d)) return sa_tc_conv3d_wgrad(d, p, q, dwp, st);
  return sa_simt_conv3d_wgrad(d, p, q, dwp, st);
}
