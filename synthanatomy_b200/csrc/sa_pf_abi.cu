// C-ABI entry points of the Performer path (include/synthanatomy_b200_performer.h): argument checks and the
// dispatch between the tcgen05 kernels (bf16 operands) and the CUDA-core kernels (fp32 parity path / fallback).
#include "sa_pf_common.cuh"

int sa_simt_gemm_nt(int64_t, int, int, int, const void*, int64_t, const void*, int64_t, const SaEpi&, cudaStream_t);
int sa_simt_gemm_tn(int64_t, int, int, int, const void*, int64_t, const void*, int64_t, const float*, float, float*,
                    cudaStream_t);
bool sa_skinny_gemm_nt_supported(int64_t m, int n, int k, int dtype, const void* a, int64_t lda, const void* b, int64_t ldb);
int sa_skinny_gemm_nt(int64_t, int, int, int, const void*, int64_t, const void*, int64_t, const SaEpi&, cudaStream_t);
bool sa_tc_gemm_nt_supported(int64_t m, int n, int k, int dtype, const void* a, int64_t lda, const void* b, int64_t ldb);
int sa_tc_gemm_nt(int64_t, int, int, const void*, int64_t, const void*, int64_t, const SaEpi&, cudaStream_t);
bool sa_tc_gemm_tn_supported(int64_t m, int na, int nb, int dtype, const void* a, int64_t lda, const void* b, int64_t ldb);
int sa_tc_gemm_tn_colsum(int64_t, int, int, const void*, int64_t, const void*, int64_t, const float*, float, float*, float*,
                         cudaStream_t);
int sa_tc_gemm_tn(int64_t, int, int, const void*, int64_t, const void*, int64_t, const float*, float, float*,
                  cudaStream_t);

int sa_simt_favor_kmax(const sa_favor_desc*, const void*, const float*, unsigned long long*, cudaStream_t);
int sa_simt_favor_featmap_fwd(const sa_favor_desc*, const void*, const float*, int, const unsigned long long*, float,
                              void*, int32_t*, cudaStream_t);
int sa_simt_favor_featmap_bwd(const sa_favor_desc*, const void*, const float*, int, float, const void*, const void*,
                              const int32_t*, void*, float*, cudaStream_t);
int sa_simt_favor_kmax_fixup(const sa_favor_desc*, const float*, const unsigned long long*, const float*, void*,
                             cudaStream_t);
size_t sa_simt_favor_scan_workspace(const sa_favor_desc*, int);
int sa_simt_favor_scan_fwd(const sa_favor_desc*, const void*, const void*, const void*, float, void*, int, float*, void*,
                           size_t, cudaStream_t);
int sa_simt_favor_scan_bwd(const sa_favor_desc*, const void*, const void*, const void*, float, const void*, const void*,
                           int, const float*, void*, void*, void*, void*, size_t, cudaStream_t);
int sa_simt_local_attn_fwd(const sa_local_desc*, const void*, const void*, const void*, const float*, void*, float*,
                           cudaStream_t);
int sa_simt_local_attn_bwd(const sa_local_desc*, const void*, const void*, const void*, const float*, const void*,
                           const void*, const float*, void*, void*, void*, cudaStream_t);

bool sa_tc_favor_supported(const sa_favor_desc*);
size_t sa_tc_favor_scan_workspace(const sa_favor_desc*, int);
int sa_tc_favor_featmap_fwd(const sa_favor_desc*, int, const void*, const float*, const unsigned long long*,
                            unsigned long long*, float, void*, int32_t*, cudaStream_t);
int sa_tc_favor_featmap_bwd(const sa_favor_desc*, const void*, const float*, int, float, const void*, const void*,
                            const int32_t*, void*, float*, cudaStream_t);
size_t sa_tc_favor_states_bytes(const sa_favor_desc*);
int sa_tc_favor_scan_fwd(const sa_favor_desc*, const void*, const void*, const void*, float, void*, int, float*, void*,
                         size_t, void*, cudaStream_t);
int sa_tc_favor_scan_bwd(const sa_favor_desc*, const void*, const void*, const void*, float, const void*, const void*,
                         int, const float*, void*, void*, void*, void*, size_t, const void*, cudaStream_t);

int sa_tc_favor_scan_bwd_fused(const sa_favor_desc*, const void*, const void*, const void*, const void*, const void*,
                               const float*, float, float, const void*, const void*, int, const float*, const int32_t*, void*,
                               void*, void*, float*, void*, size_t, const void*, cudaStream_t);

bool sa_tc_local_supported(const sa_local_desc*, const void*, const void*, const void*, const float*);
int sa_tc_local_attn_fwd(const sa_local_desc*, const void*, const void*, const void*, void*, float*, cudaStream_t);
int sa_tc_local_attn_bwd(const sa_local_desc*, const void*, const void*, const void*, const void*, const void*, const float*,
                         void*, void*, void*, float*, const float*, cudaStream_t);
int sa_rotary_table_launch(const float*, int, int, float*, cudaStream_t);
int sa_rotary_launch(void*, int, int64_t, int, int, int, int, const float*, int, cudaStream_t);
int sa_rotary_qk_launch(void*, int, int64_t, int64_t, int, int, int, int, const float*, int, cudaStream_t);

extern "C" int sa_gemm_nt(int64_t m, int n, int k, int dtype, const void* a, int64_t lda, const void* b, int64_t ldb,
                          const sa_gemm_epilogue* epi, int64_t ldo, void* stream) {
  SA_CHECK_ARG(a && b && epi, "null pointer");
  SA_CHECK_ARG(m > 0 && n > 0 && k > 0 && lda >= k && ldb >= k && ldo >= n, "bad sizes");
  SA_CHECK_ARG(dtype == SA_F32 || dtype == SA_BF16, "bad dtype");
  SA_CHECK_ARG(epi->out_f32 || epi->out_act || epi->dot_out, "no output");
  SA_CHECK_ARG(epi->act == SA_ACT_NONE || epi->pre, "GELU epilogue needs `pre`");
  SA_CHECK_ARG(!epi->dot_with == !epi->dot_out, "dot_with / dot_out must come together");
  const SaEpi e = sa_make_epi(epi, ldo);
  cudaStream_t st = sa_stream(stream);
  // a handful of rows (one decoding position per batch element): weight-streaming kernel, either dtype
  if (sa_skinny_gemm_nt_supported(m, n, k, dtype, a, lda, b, ldb)) return sa_skinny_gemm_nt(m, n, k, dtype, a, lda, b, ldb, e, st);
  if (!sa_force_simt() && sa_tc_gemm_nt_supported(m, n, k, dtype, a, lda, b, ldb)) return sa_tc_gemm_nt(m, n, k, a, lda, b, ldb, e, st);
  return sa_simt_gemm_nt(m, n, k, dtype, a, lda, b, ldb, e, st);
}

extern "C" int sa_gemm_tn(int64_t m, int na, int nb, int dtype, const void* a, int64_t lda, const void* b, int64_t ldb,
                          const float* scale_dev, float scale, float* d, int accumulate, void* stream) {
  SA_CHECK_ARG(a && b && d, "null pointer");
  SA_CHECK_ARG(m > 0 && na > 0 && nb > 0 && lda >= na && ldb >= nb, "bad sizes");
  SA_CHECK_ARG(dtype == SA_F32 || dtype == SA_BF16, "bad dtype");
  cudaStream_t st = sa_stream(stream);
  if (!accumulate) SA_CUDA(cudaMemsetAsync(d, 0, (size_t)na * nb * sizeof(float), st));
  if (!sa_force_simt() && sa_tc_gemm_tn_supported(m, na, nb, dtype, a, lda, b, ldb))
    return sa_tc_gemm_tn(m, na, nb, a, lda, b, ldb, scale_dev, scale, d, st);
  return sa_simt_gemm_tn(m, na, nb, dtype, a, lda, b, ldb, scale_dev, scale, d, st);
}

extern "C" int sa_bias_grad(const void* dy, int64_t rows, int c, int dtype, float* db, int accumulate, void* stream);

extern "C" int sa_gemm_tn_colsum(int64_t m, int na, int nb, int dtype, const void* a, int64_t lda, const void* b, int64_t ldb,
                                 const float* scale_dev, float scale, float* d, int accumulate, float* colsum, void* stream) {
  SA_CHECK_ARG(a && b && d && colsum, "null pointer");
  SA_CHECK_ARG(m > 0 && na > 0 && nb > 0 && lda >= na && ldb >= nb, "bad sizes");
  SA_CHECK_ARG(dtype == SA_F32 || dtype == SA_BF16, "bad dtype");
  cudaStream_t st = sa_stream(stream);
  if (!sa_force_simt() && sa_tc_gemm_tn_supported(m, na, nb, dtype, a, lda, b, ldb)) {
    // the column sums are one more 16-column product (A^T 1) of the CTAs that own the first column tile
    if (!accumulate) SA_CUDA(cudaMemsetAsync(d, 0, (size_t)na * nb * sizeof(float), st));
    SA_CUDA(cudaMemsetAsync(colsum, 0, (size_t)na * sizeof(float), st));
    return sa_tc_gemm_tn_colsum(m, na, nb, a, lda, b, ldb, scale_dev, scale, d, colsum, st);
  }
  SA_CHECK_ARG(lda == na, "column sums of a column slice need the tensor-core path");
  const int rc = sa_gemm_tn(m, na, nb, dtype, a, lda, b, ldb, scale_dev, scale, d, accumulate, stream);
  if (rc != SA_OK) return rc;
  return sa_bias_grad(a, m, na, dtype, colsum, 0, stream);
}

extern "C" int sa_favor_kmax(const sa_favor_desc* d, const void* k, const float* proj, unsigned long long* kmax,
                             void* stream) {
  SA_CHECK_ARG(d && k && proj && kmax, "null pointer");
  if (!sa_force_simt() && sa_tc_favor_supported(d))
    return sa_tc_favor_featmap_fwd(d, 0, k, proj, nullptr, kmax, 0.f, nullptr, nullptr, sa_stream(stream));
  return sa_simt_favor_kmax(d, k, proj, kmax, sa_stream(stream));
}

extern "C" int sa_favor_featmap_fwd(const sa_favor_desc* d, const void* x, const float* proj, int is_query,
                                    const unsigned long long* kmax, float eps, void* feat, int32_t* argmax, void* stream) {
  SA_CHECK_ARG(d && x && proj && feat, "null pointer");
  SA_CHECK_ARG(is_query ? argmax != nullptr : kmax != nullptr, "queries need argmax, keys need kmax");
  if (!sa_force_simt() && sa_tc_favor_supported(d))
    return sa_tc_favor_featmap_fwd(d, is_query ? 1 : 2, x, proj, kmax, nullptr, eps, feat, argmax, sa_stream(stream));
  return sa_simt_favor_featmap_fwd(d, x, proj, is_query, kmax, eps, feat, argmax, sa_stream(stream));
}

extern "C" int sa_favor_featmap_bwd(const sa_favor_desc* d, const void* x, const float* proj, int is_query, float eps,
                                    const void* feat, const void* dfeat, const int32_t* argmax, void* dx, float* gsum,
                                    void* stream) {
  SA_CHECK_ARG(d && x && proj && feat && dfeat && dx, "null pointer");
  SA_CHECK_ARG(is_query ? argmax != nullptr : gsum != nullptr, "queries need argmax, keys need gsum");
  if (!sa_force_simt() && sa_tc_favor_supported(d))
    return sa_tc_favor_featmap_bwd(d, x, proj, is_query, eps, feat, dfeat, argmax, dx, gsum, sa_stream(stream));
  return sa_simt_favor_featmap_bwd(d, x, proj, is_query, eps, feat, dfeat, argmax, dx, gsum, sa_stream(stream));
}

extern "C" int sa_favor_kmax_fixup(const sa_favor_desc* d, const float* proj, const unsigned long long* kmax,
                                   const float* gsum, void* dk, void* stream) {
  SA_CHECK_ARG(d && proj && kmax && gsum && dk, "null pointer");
  return sa_simt_favor_kmax_fixup(d, proj, kmax, gsum, dk, sa_stream(stream));
}

extern "C" size_t sa_favor_scan_workspace(const sa_favor_desc* d, int backward) {
  if (!d) return 0;
  // large enough for either path (the dispatch can change with sa_set_force_simt between the query and the call)
  const size_t simt = sa_simt_favor_scan_workspace(d, backward);
  const size_t tc = sa_tc_favor_supported(d) ? sa_tc_favor_scan_workspace(d, backward) : 0;
  return simt > tc ? simt : tc;
}

extern "C" size_t sa_favor_scan_states_bytes(const sa_favor_desc* d) {
  if (!d || sa_force_simt() || !sa_tc_favor_supported(d)) return 0;
  return sa_tc_favor_states_bytes(d);
}

extern "C" int sa_favor_scan_fwd_save(const sa_favor_desc* d, const void* qf, const void* kf, const void* v,
                                      float eps_cumsum, void* out, int out_ld, float* den, void* workspace,
                                      size_t ws_bytes, void* states, size_t states_bytes, void* stream) {
  SA_CHECK_ARG(d && qf && kf && v && out && den && workspace, "null pointer");
  if (!sa_force_simt() && sa_tc_favor_supported(d)) {
    SA_CHECK_ARG(!states || states_bytes >= sa_tc_favor_states_bytes(d), "states buffer too small");
    return sa_tc_favor_scan_fwd(d, qf, kf, v, eps_cumsum, out, out_ld, den, workspace, ws_bytes, states, sa_stream(stream));
  }
  SA_CHECK_ARG(!states, "the CUDA-core path does not save states (sa_favor_scan_states_bytes() == 0)");
  return sa_simt_favor_scan_fwd(d, qf, kf, v, eps_cumsum, out, out_ld, den, workspace, ws_bytes, sa_stream(stream));
}

extern "C" int sa_favor_scan_fwd(const sa_favor_desc* d, const void* qf, const void* kf, const void* v, float eps_cumsum,
                                 void* out, int out_ld, float* den, void* workspace, size_t ws_bytes, void* stream) {
  return sa_favor_scan_fwd_save(d, qf, kf, v, eps_cumsum, out, out_ld, den, workspace, ws_bytes, nullptr, 0, stream);
}

extern "C" int sa_favor_scan_bwd_saved(const sa_favor_desc* d, const void* qf, const void* kf, const void* v,
                                       float eps_cumsum, const void* out, const void* dout, int out_ld, const float* den,
                                       void* dqf, void* dkf, void* dv, void* workspace, size_t ws_bytes,
                                       const void* states, size_t states_bytes, void* stream) {
  SA_CHECK_ARG(d && qf && kf && v && out && dout && den && dqf && dkf && dv && workspace, "null pointer");
  if (!sa_force_simt() && sa_tc_favor_supported(d)) {
    SA_CHECK_ARG(!states || states_bytes >= sa_tc_favor_states_bytes(d), "states buffer too small");
    return sa_tc_favor_scan_bwd(d, qf, kf, v, eps_cumsum, out, dout, out_ld, den, dqf, dkf, dv, workspace, ws_bytes, states,
                                sa_stream(stream));
  }
  return sa_simt_favor_scan_bwd(d, qf, kf, v, eps_cumsum, out, dout, out_ld, den, dqf, dkf, dv, workspace, ws_bytes,
                                sa_stream(stream));
}

extern "C" int sa_favor_scan_bwd_fused_supported(const sa_favor_desc* d) {
  return d && !sa_force_simt() && sa_tc_favor_supported(d) ? 1 : 0;
}

extern "C" int sa_favor_scan_bwd_fused(const sa_favor_desc* d, const void* qf, const void* kf, const void* x_q,
                                       const void* x_k, const void* v, const float* proj, float eps_cumsum,
                                       float eps_feature, const void* out, const void* dout, int out_ld, const float* den,
                                       const int32_t* argq, void* dx_q, void* dx_k, void* dv, float* gsum, void* workspace,
                                       size_t ws_bytes, const void* states, size_t states_bytes, void* stream) {
  SA_CHECK_ARG(d && qf && kf && x_q && x_k && v && proj && out && dout && den && argq && dx_q && dx_k && dv && gsum && workspace,
               "null pointer");
  if (!sa_favor_scan_bwd_fused_supported(d)) {
    sa_set_error("sa_favor_scan_bwd_fused: only the tcgen05 path has this form (sa_favor_scan_bwd_fused_supported)");
    return SA_ERR_UNSUPPORTED;
  }
  SA_CHECK_ARG(!states || states_bytes >= sa_tc_favor_states_bytes(d), "states buffer too small");
  return sa_tc_favor_scan_bwd_fused(d, qf, kf, v, x_q, x_k, proj, eps_cumsum, eps_feature, out, dout, out_ld, den, argq, dx_q,
                                    dx_k, dv, gsum, workspace, ws_bytes, states, sa_stream(stream));
}

extern "C" int sa_favor_scan_bwd(const sa_favor_desc* d, const void* qf, const void* kf, const void* v, float eps_cumsum,
                                 const void* out, const void* dout, int out_ld, const float* den, void* dqf, void* dkf,
                                 void* dv, void* workspace, size_t ws_bytes, void* stream) {
  return sa_favor_scan_bwd_saved(d, qf, kf, v, eps_cumsum, out, dout, out_ld, den, dqf, dkf, dv, workspace, ws_bytes,
                                 nullptr, 0, stream);
}

extern "C" int sa_local_attn_fwd(const sa_local_desc* d, const void* q, const void* k, const void* v,
                                 const float* inv_freq, void* out, float* lse, void* stream) {
  SA_CHECK_ARG(d && q && k && v && out && lse, "null pointer");
  if (!sa_force_simt() && sa_tc_local_supported(d, q, k, v, inv_freq))
    return sa_tc_local_attn_fwd(d, q, k, v, out, lse, sa_stream(stream));
  return sa_simt_local_attn_fwd(d, q, k, v, inv_freq, out, lse, sa_stream(stream));
}

extern "C" int sa_local_attn_bwd_ws(const sa_local_desc* d, const void* q, const void* k, const void* v,
                                    const float* inv_freq, const void* out, const void* dout, const float* lse, void* dq,
                                    void* dk, void* dv, float* delta_ws, void* stream) {
  SA_CHECK_ARG(d && q && k && v && out && dout && lse && dq && dk && dv, "null pointer");
  if (!sa_force_simt() && sa_tc_local_supported(d, q, k, v, inv_freq))
    return sa_tc_local_attn_bwd(d, q, k, v, out, dout, lse, dq, dk, dv, delta_ws, nullptr, sa_stream(stream));
  return sa_simt_local_attn_bwd(d, q, k, v, inv_freq, out, dout, lse, dq, dk, dv, sa_stream(stream));
}

extern "C" int sa_rotary_table(const float* inv_freq, int seq, int dim_head, float* table, void* stream) {
  SA_CHECK_ARG(inv_freq && table, "null pointer");
  SA_CHECK_ARG(seq > 0 && dim_head > 0 && (dim_head & 1) == 0, "bad sizes");
  return sa_rotary_table_launch(inv_freq, seq, dim_head / 2, table, sa_stream(stream));
}

extern "C" int sa_local_attn_bwd_rot(const sa_local_desc* d, const void* q, const void* k, const void* v,
                                     const float* inv_freq, const float* rot_table, const void* out, const void* dout,
                                     const float* lse, void* dq, void* dk, void* dv, float* delta_ws, void* stream) {
  SA_CHECK_ARG(d && q && k && v && inv_freq && rot_table && out && dout && lse && dq && dk && dv, "null pointer");
  SA_CHECK_ARG((reinterpret_cast<uintptr_t>(rot_table) & 15) == 0, "rot_table must be 16-byte aligned");
  cudaStream_t st = sa_stream(stream);
  if (!sa_force_simt() && sa_tc_local_supported(d, q, k, v, nullptr))
    return sa_tc_local_attn_bwd(d, q, k, v, out, dout, lse, dq, dk, dv, delta_ws, rot_table, st);
  // CUDA-core kernels: the gradients of the rotated q / k, then the transpose of the rotation as its own in-place pass
  int rc = sa_simt_local_attn_bwd(d, q, k, v, nullptr, out, dout, lse, dq, dk, dv, st);
  if (rc != SA_OK) return rc;
  if ((rc = sa_rotary_launch(dq, d->act_dtype, d->ld, d->batch, d->seq, d->heads, d->dim_head, inv_freq, 1, st)) != SA_OK) return rc;
  return sa_rotary_launch(dk, d->act_dtype, d->ld, d->batch, d->seq, d->heads, d->dim_head, inv_freq, 1, st);
}

extern "C" int sa_local_attn_bwd(const sa_local_desc* d, const void* q, const void* k, const void* v,
                                 const float* inv_freq, const void* out, const void* dout, const float* lse, void* dq,
                                 void* dk, void* dv, void* stream) {
  return sa_local_attn_bwd_ws(d, q, k, v, inv_freq, out, dout, lse, dq, dk, dv, nullptr, stream);
}

extern "C" int sa_rotary(void* buf, int dtype, int64_t ld, int batch, int seq, int heads, int dim_head,
                         const float* inv_freq, int inverse, void* stream) {
  SA_CHECK_ARG(buf && inv_freq, "null pointer");
  SA_CHECK_ARG(batch > 0 && seq > 0 && heads > 0 && dim_head > 0 && (dim_head & 1) == 0 && ld >= (int64_t)heads * dim_head,
               "bad sizes");
  SA_CHECK_ARG(dtype == SA_F32 || dtype == SA_BF16, "bad dtype");
  return sa_rotary_launch(buf, dtype, ld, batch, seq, heads, dim_head, inv_freq, inverse, sa_stream(stream));
}

extern "C" int sa_rotary_qk(void* buf, int dtype, int64_t ld, int64_t k_offset, int batch, int seq, int heads, int dim_head,
                            const float* inv_freq, int inverse, void* stream) {
  SA_CHECK_ARG(buf && inv_freq, "null pointer");
  SA_CHECK_ARG(batch > 0 && seq > 0 && heads > 0 && dim_head > 0 && (dim_head & 1) == 0 && k_offset >= (int64_t)heads * dim_head &&
                   ld >= k_offset + (int64_t)heads * dim_head, "bad sizes");
  SA_CHECK_ARG(dtype == SA_F32 || dtype == SA_BF16, "bad dtype");
  return sa_rotary_qk_launch(buf, dtype, ld, k_offset, batch, seq, heads, dim_head, inv_freq, inverse, sa_stream(stream));
}
