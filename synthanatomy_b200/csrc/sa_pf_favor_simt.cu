// FAVOR+ (global heads) on CUDA cores, fp32 accumulation: random-feature map (forward / backward, with the
// non-detached stabilisers of performer-pytorch 1.0.11) and chunk-parallel causal linear attention
// (forward / backward).  This is the fp32 "parity" path and the generic fallback of the tensor-core kernels.
//
// Replaces performer_pytorch.softmax_kernel, causal_linear_attention and fast_transformers' CausalDotProduct
// CUDA extension, reached from /root/reference/src/networks/transformers/performer.py:270.
//
// Causal scan layout.  The sequence is cut into chunks of 64 tokens.  With v_aug = [v (d) | 1 | 0] (66 columns)
// and the state S[m][66] = sum_j k'[j] (x) v_aug[j] started at S0 = [0 ... 0 | 0 | eps], one product q' . S yields
// the numerator (columns < d) and, as column d + column d+1, the denominator q' . (cumsum(k') + eps).
//   pass 1  per-chunk outer-product sums                     (parallel over chunks)
//   pass 2  exclusive prefix (suffix for the backward) over chunks, in place
//   pass 3  per chunk: inter-chunk term q' . S plus the masked intra-chunk term (q' k'^T)_{j<=i} v_aug
#include "sa_pf_common.cuh"

namespace {

constexpr int FT_TOK = 16;       // tokens per feature-map tile
constexpr int FT_CHUNK = 256;    // tokens per feature-map CTA
constexpr int SC = 64;           // scan chunk (tokens)
constexpr int SE = 66;           // augmented value columns
constexpr int SL = 68;           // feature slab
constexpr int LDP = 81;          // padded shared-memory leading dimension (odd, >= 80)
constexpr int LDA = 65;

__device__ __forceinline__ unsigned int f2ord(float f) {
  const unsigned int b = __float_as_uint(f);
  return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ float ord2f(unsigned int o) {
  return __uint_as_float((o & 0x80000000u) ? (o ^ 0x80000000u) : ~o);
}

// ------------------------------------------------------------------------------------------------
// feature map
// ------------------------------------------------------------------------------------------------
struct FmArgs {
  int B, N, H, d, m, mp, ld;
  float c, r, eps;
};

// MODE 0: key max pre-pass; 1: forward query; 2: forward key
template <typename T, int MODE>
__global__ void __launch_bounds__(256)
featmap_fwd_kernel(FmArgs a, const T* __restrict__ x, const float* __restrict__ proj,
                   unsigned long long* __restrict__ kmax_out, const unsigned long long* __restrict__ kmax_in,
                   T* __restrict__ feat, int* __restrict__ argmax) {
  extern __shared__ float sm[];
  float* Ps = sm;                                 // [m][65]
  float* Xs = Ps + a.m * LDA;                     // [16][65]
  float* Ds = Xs + FT_TOK * LDA;                  // [16][mp + 1]
  float* diag = Ds + FT_TOK * (a.mp + 1);         // [16]
  float* rmax = diag + FT_TOK;                    // [16]
  int* ramax = reinterpret_cast<int*>(rmax + FT_TOK);
  const int t = threadIdx.x;
  const int bh = blockIdx.y, b = bh / a.H, h = bh % a.H;
  const int ldd = a.mp + 1;
  for (int i = t; i < a.m * a.d; i += 256) Ps[(i / a.d) * LDA + (i % a.d)] = proj[i];
  float stab = 0.f;
  if (MODE == 2) stab = ord2f((unsigned int)(kmax_in[0] >> 32));
  float best = -INFINITY;
  unsigned int best_idx = 0;
  const int n_beg = blockIdx.x * FT_CHUNK;
  const int n_end = min(a.N, n_beg + FT_CHUNK);
  for (int n0 = n_beg; n0 < n_end; n0 += FT_TOK) {
    __syncthreads();
    for (int i = t; i < FT_TOK * a.d; i += 256) {
      const int tok = i / a.d, dd = i % a.d;
      const int n = n0 + tok;
      Xs[tok * LDA + dd] = n < n_end ? sa_ld(x, ((long long)b * a.N + n) * a.ld + h * a.d + dd) : 0.f;
    }
    __syncthreads();
    if (t < FT_TOK * 16) {   // |x|^2 of each token: 16 lanes per token
      const int tok = t >> 4, l = t & 15;
      float s = 0.f;
      for (int dd = l; dd < a.d; dd += 16) s = fmaf(Xs[tok * LDA + dd], Xs[tok * LDA + dd], s);
      s = sa_half_sum(s);
      if (l == 0) diag[tok] = (s / 2.0f) * (a.c * a.c);
    }
    for (int j = t; j < a.m; j += 256) {
      float acc[FT_TOK];
#pragma unroll
      for (int k = 0; k < FT_TOK; ++k) acc[k] = 0.f;
      for (int dd = 0; dd < a.d; ++dd) {
        const float p = Ps[j * LDA + dd];
#pragma unroll
        for (int k = 0; k < FT_TOK; ++k) acc[k] = fmaf(a.c * Xs[k * LDA + dd], p, acc[k]);
      }
      if (MODE == 0) {
#pragma unroll
        for (int k = 0; k < FT_TOK; ++k) {
          if (n0 + k < n_end && acc[k] > best) {
            best = acc[k];
            best_idx = (unsigned int)((((long long)bh * a.N) + n0 + k) * a.m + j);
          }
        }
      } else {
#pragma unroll
        for (int k = 0; k < FT_TOK; ++k) Ds[k * ldd + j] = acc[k];
      }
    }
    if (MODE == 0) continue;
    __syncthreads();
    if (MODE == 1) {   // row max + arg max: 16 lanes per token
      const int tok = t >> 4, l = t & 15;
      float mx = -INFINITY; int am = 0;
      for (int j = l; j < a.m; j += 16) {
        const float v = Ds[tok * ldd + j];
        if (v > mx) { mx = v; am = j; }
      }
#pragma unroll
      for (int o = 8; o > 0; o >>= 1) {
        const float ov = __shfl_xor_sync(0xffffffffu, mx, o);
        const int oa = __shfl_xor_sync(0xffffffffu, am, o);
        if (ov > mx || (ov == mx && oa < am)) { mx = ov; am = oa; }
      }
      if (l == 0) { rmax[tok] = mx; ramax[tok] = am; }
      __syncthreads();
    }
    for (int i = t; i < FT_TOK * a.mp; i += 256) {
      const int tok = i / a.mp, j = i % a.mp;
      const int n = n0 + tok;
      if (n >= n_end) continue;
      float v = 0.f;
      if (j < a.m) {
        const float s = (MODE == 1) ? rmax[tok] : stab;
        v = a.r * (expf(Ds[tok * ldd + j] - diag[tok] - s) + a.eps);
      }
      sa_st(feat, (((long long)bh * a.N) + n) * a.mp + j, v);
    }
    if (MODE == 1 && t < FT_TOK && n0 + t < n_end) argmax[(long long)bh * a.N + n0 + t] = ramax[t];
  }
  if (MODE == 0) {
    unsigned long long packed = ((unsigned long long)f2ord(best) << 32) | (unsigned long long)(0xFFFFFFFFu - best_idx);
    if (best == -INFINITY) packed = 0ull;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const unsigned long long other = __shfl_xor_sync(0xffffffffu, packed, o);
      packed = other > packed ? other : packed;
    }
    if ((t & 31) == 0 && packed) atomicMax(kmax_out, packed);
  }
}

template <typename T>
__global__ void __launch_bounds__(256)
featmap_bwd_kernel(FmArgs a, const T* __restrict__ x, const float* __restrict__ proj, int is_query,
                   const T* __restrict__ feat, const T* __restrict__ dfeat, const int* __restrict__ argmax,
                   T* __restrict__ dx, float* __restrict__ gsum, float* __restrict__ gpart) {
  extern __shared__ float sm[];
  float* Ps = sm;                                 // [m][65]
  float* Xs = Ps + a.m * LDA;                     // [16][65]
  float* Gs = Xs + FT_TOK * LDA;                  // [16][mp + 1]  g_D then dD
  float* ssum = Gs + FT_TOK * (a.mp + 1);         // [16]
  const int t = threadIdx.x;
  const int bh = blockIdx.y, b = bh / a.H, h = bh % a.H;
  const int ldd = a.mp + 1;
  for (int i = t; i < a.m * a.d; i += 256) Ps[(i / a.d) * LDA + (i % a.d)] = proj[i];
  const int n_beg = blockIdx.x * FT_CHUNK;
  const int n_end = min(a.N, n_beg + FT_CHUNK);
  float cta_sum = 0.f;
  for (int n0 = n_beg; n0 < n_end; n0 += FT_TOK) {
    __syncthreads();
    for (int i = t; i < FT_TOK * a.d; i += 256) {
      const int tok = i / a.d, dd = i % a.d;
      const int n = n0 + tok;
      Xs[tok * LDA + dd] = n < n_end ? sa_ld(x, ((long long)b * a.N + n) * a.ld + h * a.d + dd) : 0.f;
    }
    for (int i = t; i < FT_TOK * a.mp; i += 256) {
      const int tok = i / a.mp, j = i % a.mp;
      const int n = n0 + tok;
      float g = 0.f;
      if (n < n_end && j < a.m) {
        const long long o = (((long long)bh * a.N) + n) * a.mp + j;
        g = sa_ld(dfeat, o) * (sa_ld(feat, o) - a.r * a.eps);     // dfeat * r * exp(.)
      }
      Gs[tok * ldd + j] = g;
    }
    __syncthreads();
    {
      const int tok = t >> 4, l = t & 15;
      float s = 0.f;
      for (int j = l; j < a.m; j += 16) s += Gs[tok * ldd + j];
      s = sa_half_sum(s);
      if (l == 0) {
        ssum[tok] = s;
        if (is_query && n0 + tok < n_end) Gs[tok * ldd + argmax[(long long)bh * a.N + n0 + tok]] -= s;
      }
    }
    __syncthreads();
    if (!is_query && t < FT_TOK) cta_sum += ssum[t];
    // dx[tok][dd] = c * sum_j dD[tok][j] P[j][dd] - s[tok] * c^2 * x[tok][dd]
    for (int o = t; o < 4 * a.d; o += 256) {
      const int dd = o % a.d, tg = o / a.d;
      float acc[4] = {0.f, 0.f, 0.f, 0.f};
      for (int j = 0; j < a.m; ++j) {
        const float p = Ps[j * LDA + dd];
#pragma unroll
        for (int k = 0; k < 4; ++k) acc[k] = fmaf(Gs[(tg * 4 + k) * ldd + j], p, acc[k]);
      }
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int tok = tg * 4 + k, n = n0 + tok;
        if (n < n_end)
          sa_st(dx, ((long long)b * a.N + n) * a.ld + h * a.d + dd,
                a.c * acc[k] - ssum[tok] * (a.c * a.c) * Xs[tok * LDA + dd]);
      }
    }
  }
  if (!is_query) {
    if (t < 32) {
      float s = (t < FT_TOK) ? cta_sum : 0.f;
      s = sa_warp_sum(s);
      if (t == 0) {
        if (gpart) gpart[blockIdx.y * gridDim.x + blockIdx.x] = s;      // deterministic mode: summed in block order afterwards
        else atomicAdd(gsum, s);
      }
    }
  }
}

template <typename T>
__global__ void kmax_fixup_kernel(FmArgs a, const float* __restrict__ proj, const unsigned long long* __restrict__ kmax,
                                  const float* __restrict__ gsum, T* __restrict__ dk) {
  const unsigned long long packed = kmax[0];
  if (packed == 0ull) return;
  const unsigned int flat = 0xFFFFFFFFu - (unsigned int)(packed & 0xFFFFFFFFull);
  const int j = flat % a.m;
  const long long row = flat / a.m;          // (b * H + h) * N + n
  const int n = (int)(row % a.N);
  const int bh = (int)(row / a.N);
  const int b = bh / a.H, h = bh % a.H;
  const int dd = threadIdx.x;
  if (dd < a.d) {
    const long long o = ((long long)b * a.N + n) * a.ld + h * a.d + dd;
    sa_st(dk, o, sa_ld(dk, o) - gsum[0] * a.c * proj[j * a.d + dd]);
  }
}

// ------------------------------------------------------------------------------------------------
// causal scan
// ------------------------------------------------------------------------------------------------
struct ScArgs {
  int B, N, H, d, m, mp, ld, out_ld, nchunks;
  float eps;
};

// fill a [64][LDP] tile with the augmented values of chunk tokens n0..n0+63
//   AUG 0: v_aug = [v | 1 | 0]      AUG 1: d_aug = [dout / den | dden | dden],  dden = -sum(dout * out) / den
template <typename T, int AUG>
__device__ __forceinline__ void load_aug(const ScArgs& a, float* tile, int b, int h, int n0, const T* __restrict__ v,
                                         const T* __restrict__ out, const T* __restrict__ dout,
                                         const float* __restrict__ den, int t) {
  if (AUG == 0) {
    for (int i = t; i < SC * 80; i += 256) {
      const int tok = i / 80, e = i % 80;
      const int n = n0 + tok;
      float val = 0.f;
      if (n < a.N) {
        if (e < a.d) val = sa_ld(v, ((long long)b * a.N + n) * a.ld + h * a.d + e);
        else if (e == a.d) val = 1.0f;
      }
      tile[tok * LDP + e] = val;
    }
  } else {
    // 4 lanes per token
    const int tok = t >> 2, l = t & 3;
    const int n = n0 + tok;
    float part = 0.f;
    float dn = 1.0f;
    if (n < a.N) {
      dn = den[((long long)(b * a.H + h)) * a.N + n];
      const long long base = ((long long)b * a.N + n) * a.out_ld + h * a.d;
      for (int e = l; e < a.d; e += 4) {
        const float g = sa_ld(dout, base + e);
        part = fmaf(g, sa_ld(out, base + e), part);
        tile[tok * LDP + e] = g / dn;
      }
    } else {
      for (int e = l; e < a.d; e += 4) tile[tok * LDP + e] = 0.f;
    }
    part += __shfl_xor_sync(0xffffffffu, part, 1);
    part += __shfl_xor_sync(0xffffffffu, part, 2);
    const float dden = n < a.N ? -part / dn : 0.f;
    for (int e = a.d + l; e < 80; e += 4) tile[tok * LDP + e] = (e < a.d + 2) ? dden : 0.f;
  }
}

// load feature slab [64 tokens][68 features] (+ zero pad to 80) of F[bh][n][mp]
template <typename T>
__device__ __forceinline__ void load_slab(const ScArgs& a, float* tile, const T* __restrict__ F, int bh, int n0, int m0,
                                          int t) {
  for (int i = t; i < SC * 80; i += 256) {
    const int tok = i / 80, c = i % 80;
    const int n = n0 + tok, mi = m0 + c;
    tile[tok * LDP + c] = (c < SL && n < a.N && mi < a.mp) ? sa_ld(F, ((long long)bh * a.N + n) * a.mp + mi) : 0.f;
  }
}

// pass 1: state[bh][chunk][mi][e] = sum_{tok in chunk} F[tok][mi] * aug[tok][e]
template <typename T, int AUG>
__global__ void __launch_bounds__(256)
scan_chunk_sum_kernel(ScArgs a, const T* __restrict__ F, const T* __restrict__ v, const T* __restrict__ out,
                      const T* __restrict__ dout, const float* __restrict__ den, float* __restrict__ state) {
  __shared__ float Fs[SC * LDP];
  __shared__ float Vs[SC * LDP];
  const int t = threadIdx.x, ty = t >> 4, tx = t & 15;
  const int chunk = blockIdx.x, bh = blockIdx.y, b = bh / a.H, h = bh % a.H;
  const int n0 = chunk * SC;
  load_aug<T, AUG>(a, Vs, b, h, n0, v, out, dout, den, t);
  float* dst = state + ((long long)bh * a.nchunks + chunk) * a.mp * SE;
  for (int m0 = 0; m0 < a.mp; m0 += SL) {
    __syncthreads();
    load_slab<T>(a, Fs, F, bh, n0, m0, t);
    __syncthreads();
    float acc[5][5];
    sa_tile_zero(acc);
    sa_tile_mma<5, 5>(acc, Fs, 1, LDP, Vs, LDP, 1, SC, ty, tx);   // A(mi, tok) = Fs[tok][mi]
#pragma unroll
    for (int r = 0; r < 5; ++r) {
      const int c = ty + 16 * r, mi = m0 + c;
      if (c >= SL || mi >= a.mp) continue;
#pragma unroll
      for (int s = 0; s < 5; ++s) {
        const int e = tx + 16 * s;
        if (e < SE) dst[mi * SE + e] = acc[r][s];
      }
    }
  }
}

// pass 2: exclusive prefix (reverse == 0, seeded with S0 = eps in column d+1) or exclusive suffix (reverse == 1, seed 0)
__global__ void scan_prefix_kernel(ScArgs a, float* __restrict__ state, int reverse) {
  const long long per = (long long)a.mp * SE;
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= per) return;
  const int bh = blockIdx.y;
  float* p = state + (long long)bh * a.nchunks * per + i;
  const int e = (int)(i % SE);
  float run = (!reverse && e == a.d + 1) ? a.eps : 0.f;
  if (!reverse) {
    for (int c = 0; c < a.nchunks; ++c) { const float v = p[c * per]; p[c * per] = run; run += v; }
  } else {
    for (int c = a.nchunks - 1; c >= 0; --c) { const float v = p[c * per]; p[c * per] = run; run += v; }
  }
}

__device__ __forceinline__ void load_state_slab(const ScArgs& a, float* tile, const float* __restrict__ st, int m0, int t) {
  for (int i = t; i < 80 * 80; i += 256) {     // rows [68, 80) and columns [66, 80) are zero padding
    const int c = i / 80, e = i % 80;
    const int mi = m0 + c;
    tile[c * LDP + e] = (c < SL && mi < a.mp && e < SE) ? st[mi * SE + e] : 0.f;
  }
}

// pass 3 (forward)
template <typename T>
__global__ void __launch_bounds__(256)
scan_fwd_kernel(ScArgs a, const T* __restrict__ qf, const T* __restrict__ kf, const T* __restrict__ v,
                const float* __restrict__ state, T* __restrict__ out, float* __restrict__ den) {
  extern __shared__ float sm[];
  float* Qs = sm;                    // [64][LDP]
  float* Ks = Qs + SC * LDP;         // [64][LDP]
  float* Ss = Ks + SC * LDP;         // [80][LDP]
  float* Vs = Ss + 80 * LDP;         // [64][LDP]
  float* As = Vs + SC * LDP;         // [64][LDA]
  float* dn = As + SC * LDA;         // [64][2]
  const int t = threadIdx.x, ty = t >> 4, tx = t & 15;
  const int chunk = blockIdx.x, bh = blockIdx.y, b = bh / a.H, h = bh % a.H;
  const int n0 = chunk * SC;
  const float* st = state + ((long long)bh * a.nchunks + chunk) * a.mp * SE;
  load_aug<T, 0>(a, Vs, b, h, n0, v, nullptr, nullptr, nullptr, t);
  float num[4][5], aij[4][4];
  sa_tile_zero(num);
  sa_tile_zero(aij);
  for (int m0 = 0; m0 < a.mp; m0 += SL) {
    __syncthreads();
    load_slab<T>(a, Qs, qf, bh, n0, m0, t);
    load_slab<T>(a, Ks, kf, bh, n0, m0, t);
    load_state_slab(a, Ss, st, m0, t);
    __syncthreads();
    sa_tile_mma<4, 5>(num, Qs, LDP, 1, Ss, LDP, 1, SL, ty, tx);     // q' . S
    sa_tile_mma<4, 4>(aij, Qs, LDP, 1, Ks, 1, LDP, SL, ty, tx);     // q' . k'^T
  }
#pragma unroll
  for (int r = 0; r < 4; ++r)
#pragma unroll
    for (int s = 0; s < 4; ++s) {
      const int i = ty + 16 * r, j = tx + 16 * s;
      As[i * LDA + j] = (j <= i) ? aij[r][s] : 0.f;
    }
  __syncthreads();
  sa_tile_mma<4, 5>(num, As, LDA, 1, Vs, LDP, 1, SC, ty, tx);
  if (tx < 2) {
#pragma unroll
    for (int r = 0; r < 4; ++r) dn[(ty + 16 * r) * 2 + tx] = num[r][4];   // columns d, d+1 (d == 64)
  }
  __syncthreads();
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const int i = ty + 16 * r, n = n0 + i;
    if (n >= a.N) continue;
    const float dd = dn[i * 2] + dn[i * 2 + 1];
    if (tx == 0) den[(long long)bh * a.N + n] = dd;
#pragma unroll
    for (int s = 0; s < 4; ++s)
      sa_st(out, ((long long)b * a.N + n) * a.out_ld + h * a.d + tx + 16 * s, num[r][s] / dd);
  }
}

// pass 3 (backward): dq', dk', dv of one chunk
template <typename T>
__global__ void __launch_bounds__(256)
scan_bwd_kernel(ScArgs a, const T* __restrict__ qf, const T* __restrict__ kf, const T* __restrict__ v,
                const T* __restrict__ out, const T* __restrict__ dout, const float* __restrict__ den,
                const float* __restrict__ stateS, const float* __restrict__ stateR, T* __restrict__ dqf,
                T* __restrict__ dkf, T* __restrict__ dv) {
  extern __shared__ float sm[];
  float* Qs = sm;                    // [64][LDP]
  float* Ks = Qs + SC * LDP;
  float* Ss = Ks + SC * LDP;         // [80][LDP]
  float* Rs = Ss + 80 * LDP;         // [80][LDP]
  float* Vs = Rs + 80 * LDP;         // [64][LDP]
  float* Ds = Vs + SC * LDP;         // [64][LDP]
  float* Bs = Ds + SC * LDP;         // [64][LDA]  masked d_aug . v_aug^T
  float* As = Bs + SC * LDA;         // [64][LDA]  masked q' . k'^T
  const int t = threadIdx.x, ty = t >> 4, tx = t & 15;
  const int chunk = blockIdx.x, bh = blockIdx.y, b = bh / a.H, h = bh % a.H;
  const int n0 = chunk * SC;
  const long long soff = ((long long)bh * a.nchunks + chunk) * a.mp * SE;
  load_aug<T, 0>(a, Vs, b, h, n0, v, nullptr, nullptr, nullptr, t);
  load_aug<T, 1>(a, Ds, b, h, n0, nullptr, out, dout, den, t);
  __syncthreads();
  {
    float bij[4][4];
    sa_tile_zero(bij);
    sa_tile_mma<4, 4>(bij, Ds, LDP, 1, Vs, 1, LDP, SE, ty, tx);     // d_aug[i] . v_aug[j]
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
      for (int s = 0; s < 4; ++s) {
        const int i = ty + 16 * r, j = tx + 16 * s;
        Bs[i * LDA + j] = (j <= i) ? bij[r][s] : 0.f;
      }
  }
  float aij[4][4], dvv[4][5];
  sa_tile_zero(aij);
  sa_tile_zero(dvv);
  for (int m0 = 0; m0 < a.mp; m0 += SL) {
    __syncthreads();
    load_slab<T>(a, Qs, qf, bh, n0, m0, t);
    load_slab<T>(a, Ks, kf, bh, n0, m0, t);
    load_state_slab(a, Ss, stateS + soff, m0, t);
    load_state_slab(a, Rs, stateR + soff, m0, t);
    __syncthreads();
    sa_tile_mma<4, 4>(aij, Qs, LDP, 1, Ks, 1, LDP, SL, ty, tx);     // q' . k'^T
    sa_tile_mma<4, 5>(dvv, Ks, LDP, 1, Rs, LDP, 1, SL, ty, tx);     // k' . R
    float acc[4][5];
    // dq'[i][mi] = d_aug[i] . S[mi] + sum_j B[i][j] k'[j][mi]
    sa_tile_zero(acc);
    sa_tile_mma<4, 5>(acc, Ds, LDP, 1, Ss, 1, LDP, SE, ty, tx);
    sa_tile_mma<4, 5>(acc, Bs, LDA, 1, Ks, LDP, 1, SC, ty, tx);
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const int n = n0 + ty + 16 * r;
      if (n >= a.N) continue;
#pragma unroll
      for (int s = 0; s < 5; ++s) {
        const int c = tx + 16 * s, mi = m0 + c;
        if (c < SL && mi < a.mp) sa_st(dqf, ((long long)bh * a.N + n) * a.mp + mi, acc[r][s]);
      }
    }
    // dk'[j][mi] = v_aug[j] . R[mi] + sum_i B[i][j] q'[i][mi]
    sa_tile_zero(acc);
    sa_tile_mma<4, 5>(acc, Vs, LDP, 1, Rs, 1, LDP, SE, ty, tx);
    sa_tile_mma<4, 5>(acc, Bs, 1, LDA, Qs, LDP, 1, SC, ty, tx);
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const int n = n0 + ty + 16 * r;
      if (n >= a.N) continue;
#pragma unroll
      for (int s = 0; s < 5; ++s) {
        const int c = tx + 16 * s, mi = m0 + c;
        if (c < SL && mi < a.mp) sa_st(dkf, ((long long)bh * a.N + n) * a.mp + mi, acc[r][s]);
      }
    }
  }
#pragma unroll
  for (int r = 0; r < 4; ++r)
#pragma unroll
    for (int s = 0; s < 4; ++s) {
      const int i = ty + 16 * r, j = tx + 16 * s;
      As[i * LDA + j] = (j <= i) ? aij[r][s] : 0.f;
    }
  __syncthreads();
  sa_tile_mma<4, 5>(dvv, As, 1, LDA, Ds, LDP, 1, SC, ty, tx);        // sum_i A[i][j] d_aug[i][e]
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const int n = n0 + ty + 16 * r;
    if (n >= a.N) continue;
#pragma unroll
    for (int s = 0; s < 4; ++s)
      sa_st(dv, ((long long)b * a.N + n) * a.ld + h * a.d + tx + 16 * s, dvv[r][s]);
  }
}

FmArgs make_fm(const sa_favor_desc* d, float eps) {
  FmArgs a;
  a.B = d->batch; a.N = d->seq; a.H = d->heads; a.d = d->dim_head; a.m = d->m; a.mp = d->mp; a.ld = d->ld;
  a.c = powf((float)d->dim_head, -0.25f);
  a.r = powf((float)d->m, -0.5f);
  a.eps = eps;
  return a;
}

ScArgs make_sc(const sa_favor_desc* d, int out_ld, float eps) {
  ScArgs a;
  a.B = d->batch; a.N = d->seq; a.H = d->heads; a.d = d->dim_head; a.m = d->m; a.mp = d->mp; a.ld = d->ld;
  a.out_ld = out_ld; a.nchunks = (int)sa_cdiv(d->seq, SC); a.eps = eps;
  return a;
}

size_t fm_smem(const FmArgs& a) { return sizeof(float) * ((size_t)a.m * LDA + FT_TOK * LDA + FT_TOK * (a.mp + 1) + 64); }

template <typename K>
int set_smem(K kernel, size_t bytes) {
  if (bytes > 48 * 1024) SA_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
  return SA_OK;
}

int check_favor(const sa_favor_desc* d) {
  SA_CHECK_ARG(d != nullptr, "null descriptor");
  SA_CHECK_ARG(d->batch > 0 && d->seq > 0 && d->heads > 0 && d->m > 0 && d->mp >= d->m, "bad sizes");
  SA_CHECK_ARG(d->act_dtype == SA_F32 || d->act_dtype == SA_BF16, "bad dtype");
  SA_UNSUPPORTED(d->dim_head != 64, "dim_head != 64");
  SA_UNSUPPORTED(d->m > 512, "more than 512 random features");
  SA_UNSUPPORTED((long long)d->batch * d->heads * d->seq * d->m >= (1LL << 32), "feature tensor has >= 2^32 elements");
  SA_UNSUPPORTED((long long)d->batch * d->heads > 65535, "batch * heads > 65535");
  return SA_OK;
}

}  // namespace

// ------------------------------------------------------------------------------------------------ host entry points
int sa_simt_favor_kmax(const sa_favor_desc* d, const void* k, const float* proj, unsigned long long* kmax,
                       cudaStream_t st) {
  int rc = check_favor(d);
  if (rc != SA_OK) return rc;
  sa_note_path(SA_PATH_SIMT);
  const FmArgs a = make_fm(d, 0.f);
  const size_t smem = fm_smem(a);
  dim3 grid((unsigned)sa_cdiv(d->seq, FT_CHUNK), (unsigned)(d->batch * d->heads));
  if (d->act_dtype == SA_F32) {
    if ((rc = set_smem(featmap_fwd_kernel<float, 0>, smem)) != SA_OK) return rc;
    featmap_fwd_kernel<float, 0><<<grid, 256, smem, st>>>(a, (const float*)k, proj, kmax, nullptr, nullptr, nullptr);
  } else {
    if ((rc = set_smem(featmap_fwd_kernel<__nv_bfloat16, 0>, smem)) != SA_OK) return rc;
    featmap_fwd_kernel<__nv_bfloat16, 0><<<grid, 256, smem, st>>>(a, (const __nv_bfloat16*)k, proj, kmax, nullptr,
                                                                  nullptr, nullptr);
  }
  SA_LAUNCH_CHECK();
  return SA_OK;
}

int sa_simt_favor_featmap_fwd(const sa_favor_desc* d, const void* x, const float* proj, int is_query,
                              const unsigned long long* kmax, float eps, void* feat, int32_t* argmax, cudaStream_t st) {
  int rc = check_favor(d);
  if (rc != SA_OK) return rc;
  sa_note_path(SA_PATH_SIMT);
  const FmArgs a = make_fm(d, eps);
  const size_t smem = fm_smem(a);
  dim3 grid((unsigned)sa_cdiv(d->seq, FT_CHUNK), (unsigned)(d->batch * d->heads));
#define SA_FM_LAUNCH(T, MODE)                                                                                     \
  do {                                                                                                            \
    if ((rc = set_smem(featmap_fwd_kernel<T, MODE>, smem)) != SA_OK) return rc;                                   \
    featmap_fwd_kernel<T, MODE><<<grid, 256, smem, st>>>(a, (const T*)x, proj, nullptr, kmax, (T*)feat, argmax);  \
  } while (0)
  if (d->act_dtype == SA_F32) {
    if (is_query) SA_FM_LAUNCH(float, 1); else SA_FM_LAUNCH(float, 2);
  } else {
    if (is_query) SA_FM_LAUNCH(__nv_bfloat16, 1); else SA_FM_LAUNCH(__nv_bfloat16, 2);
  }
#undef SA_FM_LAUNCH
  SA_LAUNCH_CHECK();
  return SA_OK;
}

int sa_simt_favor_featmap_bwd(const sa_favor_desc* d, const void* x, const float* proj, int is_query, float eps,
                              const void* feat, const void* dfeat, const int32_t* argmax, void* dx, float* gsum,
                              cudaStream_t st) {
  int rc = check_favor(d);
  if (rc != SA_OK) return rc;
  sa_note_path(SA_PATH_SIMT);
  const FmArgs a = make_fm(d, eps);
  const size_t smem = fm_smem(a);
  dim3 grid((unsigned)sa_cdiv(d->seq, FT_CHUNK), (unsigned)(d->batch * d->heads));
  const int nblocks = (int)(grid.x * grid.y);
  float* gpart = (gsum && !is_query) ? sa_partial_slot(nblocks, st) : nullptr;
  if (d->act_dtype == SA_F32) {
    if ((rc = set_smem(featmap_bwd_kernel<float>, smem)) != SA_OK) return rc;
    featmap_bwd_kernel<float><<<grid, 256, smem, st>>>(a, (const float*)x, proj, is_query, (const float*)feat,
                                                       (const float*)dfeat, argmax, (float*)dx, gsum, gpart);
  } else {
    if ((rc = set_smem(featmap_bwd_kernel<__nv_bfloat16>, smem)) != SA_OK) return rc;
    featmap_bwd_kernel<__nv_bfloat16><<<grid, 256, smem, st>>>(a, (const __nv_bfloat16*)x, proj, is_query,
                                                               (const __nv_bfloat16*)feat, (const __nv_bfloat16*)dfeat,
                                                               argmax, (__nv_bfloat16*)dx, gsum, gpart);
  }
  SA_LAUNCH_CHECK();
  if (gpart) return sa_ordered_sum(gpart, nblocks, gsum, st);
  return SA_OK;
}

int sa_simt_favor_kmax_fixup(const sa_favor_desc* d, const float* proj, const unsigned long long* kmax,
                             const float* gsum, void* dk, cudaStream_t st) {
  int rc = check_favor(d);
  if (rc != SA_OK) return rc;
  sa_note_path(SA_PATH_SIMT);
  const FmArgs a = make_fm(d, 0.f);
  if (d->act_dtype == SA_F32) kmax_fixup_kernel<float><<<1, 64, 0, st>>>(a, proj, kmax, gsum, (float*)dk);
  else kmax_fixup_kernel<__nv_bfloat16><<<1, 64, 0, st>>>(a, proj, kmax, gsum, (__nv_bfloat16*)dk);
  SA_LAUNCH_CHECK();
  return SA_OK;
}

size_t sa_simt_favor_scan_workspace(const sa_favor_desc* d, int backward) {
  const size_t one = (size_t)d->batch * d->heads * sa_cdiv(d->seq, SC) * d->mp * SE * sizeof(float);
  return backward ? 2 * one : one;
}

int sa_simt_favor_scan_fwd(const sa_favor_desc* d, const void* qf, const void* kf, const void* v, float eps, void* out,
                           int out_ld, float* den, void* ws, size_t ws_bytes, cudaStream_t st) {
  int rc = check_favor(d);
  if (rc != SA_OK) return rc;
  sa_note_path(SA_PATH_SIMT);
  if (ws_bytes < sa_simt_favor_scan_workspace(d, 0)) { sa_set_error("favor_scan_fwd: workspace too small"); return SA_ERR_WORKSPACE; }
  const ScArgs a = make_sc(d, out_ld, eps);
  float* state = (float*)ws;
  dim3 grid((unsigned)a.nchunks, (unsigned)(d->batch * d->heads));
  const size_t smem = sizeof(float) * ((size_t)3 * SC * LDP + 80 * LDP + SC * LDA + SC * 2);
  const long long per = (long long)a.mp * SE;
  dim3 pgrid((unsigned)sa_cdiv(per, 256), (unsigned)(d->batch * d->heads));
#define SA_SCAN_FWD(T)                                                                                             \
  do {                                                                                                             \
    scan_chunk_sum_kernel<T, 0><<<grid, 256, 0, st>>>(a, (const T*)kf, (const T*)v, nullptr, nullptr, nullptr, state); \
    SA_LAUNCH_CHECK();                                                                                             \
    scan_prefix_kernel<<<pgrid, 256, 0, st>>>(a, state, 0);                                                        \
    SA_LAUNCH_CHECK();                                                                                             \
    if ((rc = set_smem(scan_fwd_kernel<T>, smem)) != SA_OK) return rc;                                             \
    scan_fwd_kernel<T><<<grid, 256, smem, st>>>(a, (const T*)qf, (const T*)kf, (const T*)v, state, (T*)out, den);  \
    SA_LAUNCH_CHECK();                                                                                             \
  } while (0)
  if (d->act_dtype == SA_F32) SA_SCAN_FWD(float); else SA_SCAN_FWD(__nv_bfloat16);
#undef SA_SCAN_FWD
  return SA_OK;
}

int sa_simt_favor_scan_bwd(const sa_favor_desc* d, const void* qf, const void* kf, const void* v, float eps,
                           const void* out, const void* dout, int out_ld, const float* den, void* dqf, void* dkf,
                           void* dv, void* ws, size_t ws_bytes, cudaStream_t st) {
  int rc = check_favor(d);
  if (rc != SA_OK) return rc;
  sa_note_path(SA_PATH_SIMT);
  if (ws_bytes < sa_simt_favor_scan_workspace(d, 1)) { sa_set_error("favor_scan_bwd: workspace too small"); return SA_ERR_WORKSPACE; }
  const ScArgs a = make_sc(d, out_ld, eps);
  float* stateS = (float*)ws;
  float* stateR = stateS + (size_t)d->batch * d->heads * a.nchunks * a.mp * SE;
  dim3 grid((unsigned)a.nchunks, (unsigned)(d->batch * d->heads));
  const size_t smem = sizeof(float) * ((size_t)4 * SC * LDP + 2 * 80 * LDP + 2 * SC * LDA);
  const long long per = (long long)a.mp * SE;
  dim3 pgrid((unsigned)sa_cdiv(per, 256), (unsigned)(d->batch * d->heads));
#define SA_SCAN_BWD(T)                                                                                              \
  do {                                                                                                              \
    scan_chunk_sum_kernel<T, 0><<<grid, 256, 0, st>>>(a, (const T*)kf, (const T*)v, nullptr, nullptr, nullptr, stateS); \
    SA_LAUNCH_CHECK();                                                                                              \
    scan_chunk_sum_kernel<T, 1><<<grid, 256, 0, st>>>(a, (const T*)qf, nullptr, (const T*)out, (const T*)dout, den, stateR); \
    SA_LAUNCH_CHECK();                                                                                              \
    scan_prefix_kernel<<<pgrid, 256, 0, st>>>(a, stateS, 0);                                                        \
    SA_LAUNCH_CHECK();                                                                                              \
    scan_prefix_kernel<<<pgrid, 256, 0, st>>>(a, stateR, 1);                                                        \
    SA_LAUNCH_CHECK();                                                                                              \
    if ((rc = set_smem(scan_bwd_kernel<T>, smem)) != SA_OK) return rc;                                              \
    scan_bwd_kernel<T><<<grid, 256, smem, st>>>(a, (const T*)qf, (const T*)kf, (const T*)v, (const T*)out,          \
                                                (const T*)dout, den, stateS, stateR, (T*)dqf, (T*)dkf, (T*)dv);     \
    SA_LAUNCH_CHECK();                                                                                              \
  } while (0)
  if (d->act_dtype == SA_F32) SA_SCAN_BWD(float); else SA_SCAN_BWD(__nv_bfloat16);
#undef SA_SCAN_BWD
  return SA_OK;
}
