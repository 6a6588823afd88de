// placeholder until the tcgen05 GEMMs land (next commit): nothing is claimed supported
#include "sa_pf_common.cuh"
bool sa_tc_gemm_nt_supported(int64_t, int, int, int, const void*, int64_t, const void*, int64_t) { return false; }
int sa_tc_gemm_nt(int64_t, int, int, const void*, int64_t, const void*, int64_t, const SaEpi&, cudaStream_t) { return SA_ERR_UNSUPPORTED; }
bool sa_tc_gemm_tn_supported(int64_t, int, int, int, const void*, int64_t, const void*, int64_t) { return false; }
int sa_tc_gemm_tn(int64_t, int, int, const void*, int64_t, const void*, int64_t, const float*, float, float*, cudaStream_t) { return SA_ERR_UNSUPPORTED; }
