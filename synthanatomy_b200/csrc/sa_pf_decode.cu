// Incremental (one token per call) attention kernels for autoregressive sampling with recurrent state, sm_100a.
//
// The reference samples by re-running the whole network on the growing prefix for every token
// (/root/reference/src/networks/transformers/transformer.py:58-101 -> O(N^2) layer evaluations).  These kernels advance
// the attention state by ONE position and give exactly what the prefix forward gives at its last position:
//
//   global (FAVOR+) heads   S = sum_j k'_j (x) v_j and sum_j k'_j are carried as
//                              Se[f][e] = sum_j exp(c k_j.P_f - c^2|k_j|^2/2 - M) v_j[e],   ze[f] = sum_j exp(...),
//                              S1[e] = sum_j v_j[e],  cnt = number of keys,
//                            so that k'_j = r (exp(.) + eps) is reconstructed exactly: S = r (Se + eps S1),
//                            k_cumsum = r (ze + eps cnt) + 1e-6.  M is the reference's key stabiliser: the maximum of
//                            c k.P over (batch, heads, positions so far, features); when a new key raises it the carried
//                            sums are rescaled by exp(M_old - M_new).  Mhist[t] holds M after t keys (ordered-uint).
//   local heads              rotated k and v of every position are appended to a cache; position p attends the cached
//                            keys j with (floor(p / w) - 1) w <= j <= p.
//
// CUDA-core kernels (the work per token is a few hundred KB): one CTA per (batch, head).
#include "sa_pf_common.cuh"

namespace {

__device__ __forceinline__ unsigned int dc_f2ord(float f) {
  const unsigned int b = __float_as_uint(f);
  return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ float dc_ord2f(unsigned int o) {
  return __uint_as_float((o & 0x80000000u) ? (o ^ 0x80000000u) : ~o);
}

__device__ __forceinline__ float block_max(float v, float* red) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  const int w = threadIdx.x >> 5, nw = blockDim.x >> 5;
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[w] = v;
  __syncthreads();
  float r = red[0];
  for (int i = 1; i < nw; ++i) r = fmaxf(r, red[i]);
  return r;
}
__device__ __forceinline__ float block_sum(float v, float* red) {
  v = sa_warp_sum(v);
  const int w = threadIdx.x >> 5, nw = blockDim.x >> 5;
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[w] = v;
  __syncthreads();
  float r = 0.f;
  for (int i = 0; i < nw; ++i) r += red[i];
  return r;
}

struct FsArgs {
  int B, H, m, ld, out_ld, t;      // t = number of keys already in the state (position of the new token)
  const int* t_dev;                // when set, the position is read from device memory (CUDA-graph replay of a step)
  float c, r, eps, eps_cumsum;
};

// phase 1: projections of the new query / key, row max of the query projection, running global key maximum
template <typename T>
__global__ void __launch_bounds__(256)
favor_step_project_kernel(FsArgs a, const T* __restrict__ q, const T* __restrict__ k, const float* __restrict__ proj,
                          unsigned int* __restrict__ mhist, float* __restrict__ dq, float* __restrict__ dk,
                          float* __restrict__ qmax) {
  __shared__ float sq[64], sk[64], red[8];
  if (a.t_dev) a.t = *a.t_dev;
  const int bh = blockIdx.x, b = bh / a.H, h = bh % a.H;
  if (threadIdx.x < 64) {
    sq[threadIdx.x] = a.c * sa_ld(q, (long long)b * a.ld + h * 64 + threadIdx.x);
    sk[threadIdx.x] = a.c * sa_ld(k, (long long)b * a.ld + h * 64 + threadIdx.x);
  }
  __syncthreads();
  float mq = -INFINITY, mk = -INFINITY;
  for (int f = threadIdx.x; f < a.m; f += blockDim.x) {
    const float4* p = reinterpret_cast<const float4*>(proj + (long long)f * 64);
    float aq = 0.f, ak = 0.f;
#pragma unroll
    for (int e4 = 0; e4 < 16; ++e4) {
      const float4 pv = __ldg(p + e4);
      aq = fmaf(sq[4 * e4], pv.x, aq); aq = fmaf(sq[4 * e4 + 1], pv.y, aq); aq = fmaf(sq[4 * e4 + 2], pv.z, aq); aq = fmaf(sq[4 * e4 + 3], pv.w, aq);
      ak = fmaf(sk[4 * e4], pv.x, ak); ak = fmaf(sk[4 * e4 + 1], pv.y, ak); ak = fmaf(sk[4 * e4 + 2], pv.z, ak); ak = fmaf(sk[4 * e4 + 3], pv.w, ak);
    }
    dq[(long long)bh * a.m + f] = aq; dk[(long long)bh * a.m + f] = ak;
    mq = fmaxf(mq, aq); mk = fmaxf(mk, ak);
  }
  mq = block_max(mq, red);
  mk = block_max(mk, red);
  if (threadIdx.x == 0) {
    qmax[bh] = mq;
    const float prev = dc_ord2f(mhist[a.t]);
    atomicMax(mhist + a.t + 1, dc_f2ord(fmaxf(prev, mk)));
  }
}

// phase 2: state update and the attention output of the new position
template <typename T>
__global__ void __launch_bounds__(256)
favor_step_update_kernel(FsArgs a, const T* __restrict__ q, const T* __restrict__ k, const T* __restrict__ v,
                         const unsigned int* __restrict__ mhist, const float* __restrict__ dq, const float* __restrict__ dk,
                         const float* __restrict__ qmax, float* __restrict__ Se, float* __restrict__ ze,
                         float* __restrict__ S1, T* __restrict__ out) {
  extern __shared__ float sm[];
  float* ef = sm;                  // [m]  exp(dk - nk - Mnew)
  float* qf = ef + a.m;            // [m]  r (exp(dq - nq - rowmax) + eps)
  float* sv = qf + a.m;            // [64] v
  float* s1 = sv + 64;             // [64] updated S1
  float* part = s1 + 64;           // [4][64] partial numerators
  __shared__ float red[8];
  if (a.t_dev) a.t = *a.t_dev;
  const int bh = blockIdx.x, b = bh / a.H, h = bh % a.H;
  const int t = threadIdx.x;
  float nq = 0.f, nk = 0.f;
  if (t < 64) {
    const long long o = (long long)b * a.ld + h * 64 + t;
    const float qv = sa_ld(q, o), kv = sa_ld(k, o), vv = sa_ld(v, o);
    nq = qv * qv; nk = kv * kv;
    sv[t] = vv;
    const float ns = S1[(long long)bh * 64 + t] + vv;
    S1[(long long)bh * 64 + t] = ns;
    s1[t] = ns;
  }
  nq = block_sum(nq, red) * (0.5f * a.c * a.c);
  nk = block_sum(nk, red) * (0.5f * a.c * a.c);
  const float m_old = dc_ord2f(mhist[a.t]), m_new = dc_ord2f(mhist[a.t + 1]);
  const float scale = (a.t == 0) ? 0.f : expf(m_old - m_new);
  const float qm = qmax[bh];
  const float cnt = (float)(a.t + 1);
  float den = 0.f;
  for (int f = t; f < a.m; f += blockDim.x) {
    const float e = expf(dk[(long long)bh * a.m + f] - nk - m_new);
    const float qq = a.r * (expf(dq[(long long)bh * a.m + f] - nq - qm) + a.eps);
    ef[f] = e; qf[f] = qq;
    const float z = ze[(long long)bh * a.m + f] * scale + e;
    ze[(long long)bh * a.m + f] = z;
    den = fmaf(qq, a.r * (z + a.eps * cnt) + a.eps_cumsum, den);
  }
  den = block_sum(den, red);         // (also orders the shared-memory writes above before the reads below)
  const int e = t & 63, g = t >> 6;
  float num = 0.f;
  float* srow = Se + (long long)bh * a.m * 64;
  const float reps1 = a.eps * s1[e];
  for (int f0 = g; f0 < a.m; f0 += 32) {          // 8 state rows per pass: the loads are issued before the stores
    float sold[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int f = f0 + 4 * u;
      sold[u] = f < a.m ? srow[(long long)f * 64 + e] : 0.f;
    }
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int f = f0 + 4 * u;
      if (f < a.m) {
        const float s = sold[u] * scale + ef[f] * sv[e];
        srow[(long long)f * 64 + e] = s;
        num = fmaf(qf[f], a.r * (s + reps1), num);
      }
    }
  }
  part[g * 64 + e] = num;
  __syncthreads();
  if (t < 64) {
    const float nsum = part[t] + part[64 + t] + part[128 + t] + part[192 + t];
    sa_st(out, (long long)b * a.out_ld + h * 64 + t, nsum / den);
  }
}

struct LsArgs {
  int B, H, W, ld, out_ld, p, nmax;   // p = position of the new token, nmax = rows of the caches per batch element
  const int* p_dev;                   // when set, the position is read from device memory
  float scale;
  int rotary;
};

// local heads: append the (rotated) key and the value of position p to the caches, attend [lo(p), p]
template <typename T>
__global__ void __launch_bounds__(128)
local_step_kernel(LsArgs a, const T* __restrict__ q, const T* __restrict__ k, const T* __restrict__ v,
                  const float* __restrict__ inv_freq, T* __restrict__ kcache, T* __restrict__ vcache, T* __restrict__ out) {
  extern __shared__ float sm[];
  float* sq = sm;                  // [64] rotated query
  float* sc = sq + 64;             // [2 W] scores
  float* part = sc + 2 * a.W;      // [2][64]
  __shared__ float red[4];
  if (a.p_dev) a.p = *a.p_dev;
  const int bh = blockIdx.x, b = bh / a.H, h = bh % a.H;
  const int t = threadIdx.x;
  const long long crow = ((long long)b * a.nmax + a.p) * (a.H * 64) + h * 64;
  if (t < 32) {
    const long long o = (long long)b * a.ld + h * 64 + t;
    float q1 = sa_ld(q, o), q2 = sa_ld(q, o + 32), k1 = sa_ld(k, o), k2 = sa_ld(k, o + 32);
    if (a.rotary) {
      float sn, cs;
      sincosf((float)a.p * inv_freq[t], &sn, &cs);
      const float a1 = q1 * cs - q2 * sn, a2 = q2 * cs + q1 * sn;
      const float b1 = k1 * cs - k2 * sn, b2 = k2 * cs + k1 * sn;
      q1 = a1; q2 = a2; k1 = b1; k2 = b2;
    }
    sq[t] = q1; sq[t + 32] = q2;
    sa_st(kcache, crow + t, k1); sa_st(kcache, crow + t + 32, k2);
  } else if (t < 96) {
    const int e = t - 32;
    sa_st(vcache, crow + e, sa_ld(v, (long long)b * a.ld + h * 64 + e));
  }
  __syncthreads();                 // cache rows of position p are visible to this CTA (the only reader in this launch)
  const int w = a.p / a.W - 1;
  const int lo = (w > 0 ? w : 0) * a.W;
  const int nk = a.p - lo + 1;
  float mx = -INFINITY;
  for (int j = t; j < nk; j += blockDim.x) {
    const T* kr = kcache + ((long long)b * a.nmax + lo + j) * (a.H * 64) + h * 64;
    float acc = 0.f;
#pragma unroll 8
    for (int e = 0; e < 64; ++e) acc = fmaf(sq[e], sa_ld(kr, e), acc);
    acc *= a.scale;
    sc[j] = acc;
    mx = fmaxf(mx, acc);
  }
  {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    if ((t & 31) == 0) red[t >> 5] = mx;
    __syncthreads();
    mx = fmaxf(fmaxf(red[0], red[1]), fmaxf(red[2], red[3]));
  }
  float sum = 0.f;
  for (int j = t; j < nk; j += blockDim.x) {
    const float pr = expf(sc[j] - mx);
    sc[j] = pr;
    sum += pr;
  }
  sum = sa_warp_sum(sum);
  __syncthreads();
  if ((t & 31) == 0) red[t >> 5] = sum;
  __syncthreads();
  sum = red[0] + red[1] + red[2] + red[3];
  const int e = t & 63, g = t >> 6;
  float acc = 0.f;
  for (int j = g; j < nk; j += 2)
    acc = fmaf(sc[j], sa_ld(vcache, ((long long)b * a.nmax + lo + j) * (a.H * 64) + h * 64 + e), acc);
  part[g * 64 + e] = acc;
  __syncthreads();
  if (t < 64) sa_st(out, (long long)b * a.out_ld + h * 64 + t, (part[t] + part[64 + t]) / sum);
}

}  // namespace

// q / k / v: [batch][ld] rows of the new position (column 0 of head 0 of each block); proj [m][64] fp32;
// mhist [>= t + 2] ordered-uint key maxima (mhist[0] must hold the encoding of -inf, later entries zero);
// scratch: 2 * batch * heads * m + batch * heads floats;  Se [batch*heads][m][64], ze [batch*heads][m], S1 [batch*heads][64].
extern "C" int sa_favor_decode_step(int batch, int heads, int m, int dtype, int t, const int* t_dev, const void* q,
                                    const void* k, const void* v, int ld, const float* proj, float eps, float eps_cumsum,
                                    unsigned int* mhist, float* scratch, float* Se, float* ze, float* S1, void* out,
                                    int out_ld, void* stream) {
  SA_CHECK_ARG(q && k && v && proj && mhist && scratch && Se && ze && S1 && out, "null pointer");
  SA_CHECK_ARG(batch > 0 && heads > 0 && m > 0 && m <= 1024 && t >= 0, "bad sizes");
  SA_CHECK_ARG(dtype == SA_F32 || dtype == SA_BF16, "bad dtype");
  cudaStream_t st = sa_stream(stream);
  sa_note_path(SA_PATH_SIMT);
  FsArgs a;
  a.B = batch; a.H = heads; a.m = m; a.ld = ld; a.out_ld = out_ld; a.t = t; a.t_dev = t_dev;
  a.c = powf(64.0f, -0.25f); a.r = powf((float)m, -0.5f); a.eps = eps; a.eps_cumsum = eps_cumsum;
  float* dq = scratch;
  float* dk = dq + (size_t)batch * heads * m;
  float* qmax = dk + (size_t)batch * heads * m;
  const size_t smem = sizeof(float) * ((size_t)2 * m + 64 + 64 + 256);
  const unsigned grid = (unsigned)(batch * heads);
  if (dtype == SA_F32) {
    favor_step_project_kernel<float><<<grid, 256, 0, st>>>(a, (const float*)q, (const float*)k, proj, mhist, dq, dk, qmax);
    SA_LAUNCH_CHECK();
    favor_step_update_kernel<float><<<grid, 256, smem, st>>>(a, (const float*)q, (const float*)k, (const float*)v, mhist, dq,
                                                             dk, qmax, Se, ze, S1, (float*)out);
  } else {
    favor_step_project_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>(a, (const __nv_bfloat16*)q, (const __nv_bfloat16*)k, proj,
                                                                   mhist, dq, dk, qmax);
    SA_LAUNCH_CHECK();
    favor_step_update_kernel<__nv_bfloat16><<<grid, 256, smem, st>>>(a, (const __nv_bfloat16*)q, (const __nv_bfloat16*)k,
                                                                     (const __nv_bfloat16*)v, mhist, dq, dk, qmax, Se, ze, S1,
                                                                     (__nv_bfloat16*)out);
  }
  SA_LAUNCH_CHECK();
  return SA_OK;
}

// kcache / vcache: [batch][nmax][heads * 64] (act dtype); position p must be < nmax.
extern "C" int sa_local_decode_step(int batch, int heads, int window, int dtype, int p, const int* p_dev, int nmax,
                                    const void* q, const void* k, const void* v, int ld, const float* inv_freq, void* kcache,
                                    void* vcache, void* out, int out_ld, void* stream) {
  SA_CHECK_ARG(q && k && v && kcache && vcache && out, "null pointer");
  SA_CHECK_ARG(batch > 0 && heads > 0 && window > 0 && p >= 0 && p < nmax, "bad sizes");
  SA_CHECK_ARG(dtype == SA_F32 || dtype == SA_BF16, "bad dtype");
  SA_UNSUPPORTED(window > 4096, "window > 4096");
  cudaStream_t st = sa_stream(stream);
  sa_note_path(SA_PATH_SIMT);
  LsArgs a;
  a.B = batch; a.H = heads; a.W = window; a.ld = ld; a.out_ld = out_ld; a.p = p; a.nmax = nmax; a.p_dev = p_dev;
  a.scale = 0.125f; a.rotary = inv_freq != nullptr;
  const size_t smem = sizeof(float) * ((size_t)64 + 2 * window + 128);
  const unsigned grid = (unsigned)(batch * heads);
  if (dtype == SA_F32) {
    if (smem > 48 * 1024) SA_CUDA(cudaFuncSetAttribute(local_step_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    local_step_kernel<float><<<grid, 128, smem, st>>>(a, (const float*)q, (const float*)k, (const float*)v, inv_freq,
                                                      (float*)kcache, (float*)vcache, (float*)out);
  } else {
    if (smem > 48 * 1024)
      SA_CUDA(cudaFuncSetAttribute(local_step_kernel<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    local_step_kernel<__nv_bfloat16><<<grid, 128, smem, st>>>(a, (const __nv_bfloat16*)q, (const __nv_bfloat16*)k,
                                                              (const __nv_bfloat16*)v, inv_freq, (__nv_bfloat16*)kcache,
                                                              (__nv_bfloat16*)vcache, (__nv_bfloat16*)out);
  }
  SA_LAUNCH_CHECK();
  return SA_OK;
}

namespace {
struct EsPtrs { const float* sp_w[3]; };

template <typename T>
__global__ void embed_step_kernel(const long long* __restrict__ tokens, const int* __restrict__ sp_idx, int n_axes, int sp_ld,
                                  const float* __restrict__ tok_w, EsPtrs sp, const float* __restrict__ pos_w, int B, int dim,
                                  int t, const int* __restrict__ t_dev, float* __restrict__ x_f32, T* __restrict__ x_act) {
  if (t_dev) t = *t_dev;
  const int total = B * dim;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int c = i % dim, b = i / dim;
    float v = tok_w[tokens[b] * dim + c];
    for (int a = 0; a < n_axes; ++a) {
      const int s = sp_idx[a * sp_ld + t];
      if (s >= 0) v += sp.sp_w[a][(long long)s * dim + c];
    }
    v += pos_w[(long long)t * dim + c];
    if (x_f32) x_f32[i] = v;
    if (x_act) sa_st(x_act, i, v);
  }
}
}  // namespace

// embedding of ONE position t (read from t_dev when given): tokens [batch] int64, sp_idx [n_axes][sp_ld] as in sa_embed_fwd
extern "C" int sa_embed_step(const int64_t* tokens, const int32_t* sp_idx, int n_axes, int sp_ld, const float* tok_w,
                             const float* const* sp_w, const float* pos_w, int batch, int dim, int t, const int* t_dev,
                             float* x_f32, void* x_act, int act_dtype, void* stream) {
  SA_CHECK_ARG(tokens && tok_w && pos_w && (x_f32 || x_act), "null pointer");
  SA_CHECK_ARG(n_axes >= 0 && n_axes <= 3 && (n_axes == 0 || (sp_idx && sp_w)), "bad spatial axes");
  EsPtrs sp = {};
  for (int a = 0; a < n_axes; ++a) sp.sp_w[a] = sp_w[a];
  const unsigned grid = (unsigned)sa_cdiv((int64_t)batch * dim, 256);
  if (act_dtype == SA_BF16)
    embed_step_kernel<__nv_bfloat16><<<grid, 256, 0, sa_stream(stream)>>>((const long long*)tokens, sp_idx, n_axes, sp_ld, tok_w,
                                                                          sp, pos_w, batch, dim, t, t_dev, x_f32,
                                                                          (__nv_bfloat16*)x_act);
  else
    embed_step_kernel<float><<<grid, 256, 0, sa_stream(stream)>>>((const long long*)tokens, sp_idx, n_axes, sp_ld, tok_w, sp,
                                                                  pos_w, batch, dim, t, t_dev, x_f32, (float*)x_act);
  SA_LAUNCH_CHECK();
  return SA_OK;
}
