// Token plumbing between the two models, on the device (integer gathers only):
//
//   sa_tokens_prepare   latent index grid -> Performer input / target sequences.  Replaces the host-side chain
//                       reshape -> encoded[:, index_sequence] -> F.pad(BOS) -> .long() -> [:, :-1] / [:, 1:] of
//                       /root/reference/src/utils/transformer.py:259-282 with one gather that reads the grid in its
//                       on-disk type (uint16, run_vqvae.py:484-498) and writes both int64 sequences:
//                           target[b][i] = grid[b][order[i]]        input[b][0] = BOS, input[b][i] = target[b][i - 1]
//   sa_tokens_gather    sequence -> grid (the inverse permutation used after sampling,
//                       /root/reference/src/inferer/transformer.py:63-71: sample[:, revert_ordering]) or any other
//                       index gather of token rows, int64 out
//   sa_tokens_narrow    int64 indices -> uint16 (what the extraction mode stores per subject) with a range check
//
// All three are HBM-bound byte movers: 2 B read + 16 B written per token (prepare), 8 + 8 (gather), 8 + 2 (narrow).
#include "sa_pf_common.cuh"

namespace {

template <typename T>
__device__ __forceinline__ long long tok_load(const void* p, long long i) {
  return (long long)reinterpret_cast<const T*>(p)[i];
}
__device__ __forceinline__ long long tok_load_any(const void* p, int dtype, long long i) {
  switch (dtype) {
    case SA_TOK_U16: return tok_load<unsigned short>(p, i);
    case SA_TOK_I32: return tok_load<int>(p, i);
    default: return tok_load<long long>(p, i);
  }
}

__global__ void tokens_prepare_kernel(const void* __restrict__ grid, int dtype, long long n_src, const long long* __restrict__ order,
                                      long long n, long long bos, int batch, long long* __restrict__ x_in,
                                      long long* __restrict__ y) {
  const long long total = (long long)batch * n;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    const long long b = e / n, i = e - b * n;
    const long long v = tok_load_any(grid, dtype, b * n_src + order[i]);
    y[e] = v;
    if (i + 1 < n) x_in[e + 1] = v;
    if (i == 0) x_in[e] = bos;
  }
}

__global__ void tokens_gather_kernel(const void* __restrict__ src, int dtype, long long n_src, const long long* __restrict__ index,
                                     long long n, int batch, long long* __restrict__ out) {
  const long long total = (long long)batch * n;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    const long long b = e / n, i = e - b * n;
    out[e] = tok_load_any(src, dtype, b * n_src + index[i]);
  }
}

__global__ void tokens_narrow_kernel(const long long* __restrict__ src, long long n, unsigned short* __restrict__ out,
                                     int* __restrict__ bad) {
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (long long)gridDim.x * blockDim.x) {
    const long long v = src[e];
    if (v < 0 || v > 65535) atomicExch(bad, 1);
    out[e] = (unsigned short)v;
  }
}

unsigned blocks_for(long long total) {
  long long b = (total + 255) / 256;
  if (b > 148 * 16) b = 148 * 16;
  return (unsigned)(b < 1 ? 1 : b);
}

bool tok_dtype_ok(int dtype) { return dtype == SA_TOK_U16 || dtype == SA_TOK_I32 || dtype == SA_TOK_I64; }

}  // namespace

extern "C" {

int sa_tokens_prepare(const void* grid, int tok_dtype, int batch, int64_t n_src, const int64_t* order, int64_t n, int64_t bos,
                      int64_t* x_in, int64_t* y, void* stream) {
  SA_CHECK_ARG(grid && order && x_in && y, "null pointer");
  SA_CHECK_ARG(tok_dtype_ok(tok_dtype), "token dtype must be SA_TOK_U16 / SA_TOK_I32 / SA_TOK_I64");
  SA_CHECK_ARG(batch >= 0 && n >= 0 && n_src >= n, "need n <= n_src");
  if (batch == 0 || n == 0) return SA_OK;
  sa_note_path(SA_PATH_SIMT);
  tokens_prepare_kernel<<<blocks_for((long long)batch * n), 256, 0, (cudaStream_t)stream>>>(
      grid, tok_dtype, (long long)n_src, (const long long*)order, (long long)n, (long long)bos, batch, (long long*)x_in,
      (long long*)y);
  SA_LAUNCH_CHECK();
  return SA_OK;
}

int sa_tokens_gather(const void* src, int tok_dtype, int batch, int64_t n_src, const int64_t* index, int64_t n, int64_t* out,
                     void* stream) {
  SA_CHECK_ARG(src && index && out, "null pointer");
  SA_CHECK_ARG(tok_dtype_ok(tok_dtype), "token dtype must be SA_TOK_U16 / SA_TOK_I32 / SA_TOK_I64");
  SA_CHECK_ARG(batch >= 0 && n >= 0 && n_src >= 0, "negative size");
  if (batch == 0 || n == 0) return SA_OK;
  sa_note_path(SA_PATH_SIMT);
  tokens_gather_kernel<<<blocks_for((long long)batch * n), 256, 0, (cudaStream_t)stream>>>(
      src, tok_dtype, (long long)n_src, (const long long*)index, (long long)n, batch, (long long*)out);
  SA_LAUNCH_CHECK();
  return SA_OK;
}

int sa_tokens_narrow(const int64_t* src, int64_t n, uint16_t* out, int* out_of_range, void* stream) {
  SA_CHECK_ARG(src && out && out_of_range, "null pointer");
  SA_CHECK_ARG(n >= 0, "negative size");
  if (n == 0) return SA_OK;
  sa_note_path(SA_PATH_SIMT);
  tokens_narrow_kernel<<<blocks_for((long long)n), 256, 0, (cudaStream_t)stream>>>((const long long*)src, (long long)n, out,
                                                                                   out_of_range);
  SA_LAUNCH_CHECK();
  return SA_OK;
}

}  // extern "C"
