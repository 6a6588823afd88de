// CUDA-core (SIMT) gather-GEMM convolution kernels: the fp32 "parity" path and the catch-all for
// shapes the tcgen05 kernels do not take (c_in == 1, c_out == 1, c_in == 32 ...).
//
// Reference call sites replaced: nn.Conv3d / nn.ConvTranspose3d forward + autograd
// (src/networks/vqvae/baseline.py:153-156, 218-227, 242-244, 258, 283-293).
#include "sa_common.cuh"

namespace {

struct ConvGeom {
  int B;
  int iD, iH, iW;   // X extent
  int oD, oH, oW;   // Y extent
  int Cin, Cout;
  int k, s, p;
  int taps;         // k^3
};

__host__ __device__ inline ConvGeom make_geom(const sa_conv_desc& d) {
  ConvGeom g;
  g.B = d.batch;
  g.iD = d.in_dhw[0]; g.iH = d.in_dhw[1]; g.iW = d.in_dhw[2];
  g.oD = d.out_dhw[0]; g.oH = d.out_dhw[1]; g.oW = d.out_dhw[2];
  g.Cin = d.c_in; g.Cout = d.c_out;
  g.k = d.ksize; g.s = d.stride; g.p = d.pad;
  g.taps = d.ksize * d.ksize * d.ksize;
  return g;
}

// Input position gathered by output position (od,oh,ow) and tap (td,th,tw); returns -1 if it is
// padding / a hole of the transposed form.
template <bool TRANSPOSED>
__device__ __forceinline__ int64_t gather_pos(const ConvGeom& g, int b, int od, int oh, int ow, int tap) {
  const int tw = tap % g.k;
  const int th = (tap / g.k) % g.k;
  const int td = tap / (g.k * g.k);
  int id, ih, iw;
  if (!TRANSPOSED) {
    id = od * g.s - g.p + td;
    ih = oh * g.s - g.p + th;
    iw = ow * g.s - g.p + tw;
  } else {
    id = od + g.p - td;
    ih = oh + g.p - th;
    iw = ow + g.p - tw;
    if (id < 0 || ih < 0 || iw < 0) return -1;
    if ((id % g.s) | (ih % g.s) | (iw % g.s)) return -1;
    id /= g.s; ih /= g.s; iw /= g.s;
  }
  if (id < 0 || id >= g.iD || ih < 0 || ih >= g.iH || iw < 0 || iw >= g.iW) return -1;
  return (((int64_t)b * g.iD + id) * g.iH + ih) * g.iW + iw;
}

constexpr int TM = 64, TN = 64, TK = 16;

// Y[pos][n] = epi( sum_{kidx} X[gather(pos, kidx / Cin)][kidx % Cin] * Wp[kidx / Cin][n][kidx % Cin] )
template <typename T, bool TRANSPOSED>
__global__ void __launch_bounds__(256)
gg_fwd_kernel(ConvGeom g, const T* __restrict__ x, const T* __restrict__ wp, const float* __restrict__ bias,
              const T* __restrict__ addend, const T* __restrict__ mask, int relu, T* __restrict__ y) {
  __shared__ float As[TK][TM + 4];
  __shared__ float Bs[TK][TN + 4];
  const int t = threadIdx.x;
  const int64_t M = (int64_t)g.B * g.oD * g.oH * g.oW;
  const int64_t m0 = (int64_t)blockIdx.x * TM;
  const int n0 = blockIdx.y * TN;
  const int Ktot = g.taps * g.Cin;

  // loader role: row lm = t / 4, k-offset lk = (t % 4) * 4
  const int lm = t >> 2, lk = (t & 3) << 2;
  const int64_t lpos = m0 + lm;
  int lb = 0, lod = 0, loh = 0, low = 0;
  const bool lvalid = lpos < M;
  if (lvalid) {
    int64_t r = lpos;
    low = (int)(r % g.oW); r /= g.oW;
    loh = (int)(r % g.oH); r /= g.oH;
    lod = (int)(r % g.oD); r /= g.oD;
    lb = (int)r;
  }
  const int ln = n0 + lm;  // weight row loaded by this thread

  const int ty = t >> 4, tx = t & 15;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  for (int k0 = 0; k0 < Ktot; k0 += TK) {
    // ---- stage A (gathered activations) and B (weights) ----
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int kidx = k0 + lk + e;
      float av = 0.f, bv = 0.f;
      if (kidx < Ktot) {
        const int tap = kidx / g.Cin, ci = kidx - tap * g.Cin;
        if (lvalid) {
          const int64_t ip = gather_pos<TRANSPOSED>(g, lb, lod, loh, low, tap);
          if (ip >= 0) av = sa_ld(x, ip * g.Cin + ci);
        }
        if (ln < g.Cout) bv = sa_ld(wp, ((int64_t)tap * g.Cout + ln) * g.Cin + ci);
      }
      As[lk + e][lm] = av;
      Bs[lk + e][lm] = bv;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < TK; ++kk) {
      float a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = As[kk][ty * 4 + i];
#pragma unroll
      for (int j = 0; j < 4; ++j) b[j] = Bs[kk][tx * 4 + j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }

#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int64_t pos = m0 + ty * 4 + i;
    if (pos >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tx * 4 + j;
      if (n >= g.Cout) continue;
      float v = acc[i][j];
      const int64_t o = pos * g.Cout + n;
      if (bias) v += bias[n];
      if (addend) v += sa_ld(addend, o);
      if (relu) v = fmaxf(v, 0.f);
      if (mask) v = sa_ld(mask, o) > 0.f ? v : 0.f;
      sa_st(y, o, v);
    }
  }
}

// One warp per output position, for tiny c_out (the final ConvTranspose3d 128 -> 1, baseline.py:283-293).
template <typename T, bool TRANSPOSED>
__global__ void __launch_bounds__(256)
gg_fwd_narrow_kernel(ConvGeom g, const T* __restrict__ x, const T* __restrict__ wp, const float* __restrict__ bias,
                     const T* __restrict__ addend, const T* __restrict__ mask, int relu, T* __restrict__ y) {
  const int lane = threadIdx.x & 31;
  const int64_t M = (int64_t)g.B * g.oD * g.oH * g.oW;
  const int64_t pos = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (pos >= M) return;
  int64_t r = pos;
  const int ow = (int)(r % g.oW); r /= g.oW;
  const int oh = (int)(r % g.oH); r /= g.oH;
  const int od = (int)(r % g.oD); r /= g.oD;
  const int b = (int)r;
  for (int n = 0; n < g.Cout; ++n) {
    float acc = 0.f;
    for (int tap = 0; tap < g.taps; ++tap) {
      const int64_t ip = gather_pos<TRANSPOSED>(g, b, od, oh, ow, tap);
      if (ip < 0) continue;  // warp-uniform
      const T* xr = x + ip * g.Cin;
      const T* wr = wp + ((int64_t)tap * g.Cout + n) * g.Cin;
      for (int c = lane; c < g.Cin; c += 32) acc = fmaf(sa_ld(xr, c), sa_ld(wr, c), acc);
    }
    acc = sa_warp_sum(acc);
    if (lane == 0) {
      float v = acc;
      const int64_t o = pos * g.Cout + n;
      if (bias) v += bias[n];
      if (addend) v += sa_ld(addend, o);
      if (relu) v = fmaxf(v, 0.f);
      if (mask) v = sa_ld(mask, o) > 0.f ? v : 0.f;
      sa_st(y, o, v);
    }
  }
}

// dWp[tap][n][c] += sum_pos P[pos][n] * Q[gather(pos, tap)][c];  GEMM with M = n, N = (tap, c), K = pos.
template <typename T>
__global__ void __launch_bounds__(256)
gg_wgrad_kernel(ConvGeom g, const T* __restrict__ p, const T* __restrict__ q, float* __restrict__ dwp,
                int64_t pos_per_split, unsigned* turn) {
  __shared__ float As[TK][TM + 4];  // [pos][n]
  __shared__ float Bs[TK][TN + 4];  // [pos][j]
  const int t = threadIdx.x;
  const int64_t M = (int64_t)g.B * g.oD * g.oH * g.oW;
  const int n0 = blockIdx.x * TM;
  const int j0 = blockIdx.y * TN;
  const int J = g.taps * g.Cin;
  const int64_t pbeg = (int64_t)blockIdx.z * pos_per_split;
  const int64_t pend = min(M, pbeg + pos_per_split);

  const int lp = t >> 4;          // position within the chunk loaded by this thread
  const int lc = (t & 15) << 2;   // 4 consecutive n / j
  const int ty = t >> 4, tx = t & 15;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  for (int64_t pc = pbeg; pc < pend; pc += TK) {
    const int64_t pos = pc + lp;
    const bool valid = pos < pend;
    int b = 0, od = 0, oh = 0, ow = 0;
    if (valid) {
      int64_t r = pos;
      ow = (int)(r % g.oW); r /= g.oW;
      oh = (int)(r % g.oH); r /= g.oH;
      od = (int)(r % g.oD); r /= g.oD;
      b = (int)r;
    }
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int n = n0 + lc + e;
      As[lp][lc + e] = (valid && n < g.Cout) ? sa_ld(p, pos * g.Cout + n) : 0.f;
      const int j = j0 + lc + e;
      float bv = 0.f;
      if (valid && j < J) {
        const int tap = j / g.Cin, c = j - tap * g.Cin;
        const int64_t ip = gather_pos<false>(g, b, od, oh, ow, tap);
        if (ip >= 0) bv = sa_ld(q, ip * g.Cin + c);
      }
      Bs[lp][lc + e] = bv;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < TK; ++kk) {
      float a[4], bb[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = As[kk][ty * 4 + i];
#pragma unroll
      for (int j = 0; j < 4; ++j) bb[j] = Bs[kk][tx * 4 + j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], bb[j], acc[i][j]);
    }
    __syncthreads();
  }
  // deterministic mode: the position splits of one output tile add in split order
  unsigned* my_turn = turn ? turn + blockIdx.y * gridDim.x + blockIdx.x : nullptr;
  sa_block_turn_begin(my_turn, blockIdx.z);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int n = n0 + ty * 4 + i;
    if (n >= g.Cout) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int jj = j0 + tx * 4 + j;
      if (jj >= J) continue;
      const int tap = jj / g.Cin, c = jj - tap * g.Cin;
      atomicAdd(dwp + ((int64_t)tap * g.Cout + n) * g.Cin + c, acc[i][j]);
    }
  }
  sa_block_turn_end(my_turn, blockIdx.z);
}

template <typename T>
int launch_fwd(const sa_conv_desc* d, const void* x, const void* wp, const float* bias, const void* addend,
               const void* mask, int relu, void* y, cudaStream_t st) {
  const ConvGeom g = make_geom(*d);
  const int64_t M = (int64_t)g.B * g.oD * g.oH * g.oW;
  if (g.Cout <= 4 && g.Cin >= 32) {
    const int wpb = 8;
    dim3 grid((unsigned)sa_cdiv(M, wpb));
    if (d->transposed)
      gg_fwd_narrow_kernel<T, true><<<grid, wpb * 32, 0, st>>>(g, (const T*)x, (const T*)wp, bias, (const T*)addend,
                                                               (const T*)mask, relu, (T*)y);
    else
      gg_fwd_narrow_kernel<T, false><<<grid, wpb * 32, 0, st>>>(g, (const T*)x, (const T*)wp, bias,
                                                                (const T*)addend, (const T*)mask, relu, (T*)y);
  } else {
    dim3 grid((unsigned)sa_cdiv(M, TM), (unsigned)sa_cdiv(g.Cout, TN));
    if (d->transposed)
      gg_fwd_kernel<T, true><<<grid, 256, 0, st>>>(g, (const T*)x, (const T*)wp, bias, (const T*)addend,
                                                   (const T*)mask, relu, (T*)y);
    else
      gg_fwd_kernel<T, false><<<grid, 256, 0, st>>>(g, (const T*)x, (const T*)wp, bias, (const T*)addend,
                                                    (const T*)mask, relu, (T*)y);
  }
  SA_LAUNCH_CHECK();
  return SA_OK;
}

template <typename T>
int launch_wgrad(const sa_conv_desc* d, const void* p, const void* q, float* dwp, cudaStream_t st) {
  const ConvGeom g = make_geom(*d);
  const int64_t M = (int64_t)g.B * g.oD * g.oH * g.oW;
  const int J = g.taps * g.Cin;
  const int64_t tiles = sa_cdiv(g.Cout, TM) * sa_cdiv(J, TN);
  int64_t splits = sa_cdiv(148 * 8, tiles);
  const int64_t max_splits = sa_cdiv(M, 4 * TK);
  if (splits > max_splits) splits = max_splits;
  if (splits < 1) splits = 1;
  if (splits > 65535) splits = 65535;
  int64_t pps = sa_cdiv(sa_cdiv(M, splits), TK) * TK;
  splits = sa_cdiv(M, pps);
  dim3 grid((unsigned)sa_cdiv(g.Cout, TM), (unsigned)sa_cdiv(J, TN), (unsigned)splits);
  gg_wgrad_kernel<T><<<grid, 256, 0, st>>>(g, (const T*)p, (const T*)q, dwp, pps, sa_turn_slot((int)tiles, st));
  SA_LAUNCH_CHECK();
  return SA_OK;
}

}  // namespace

int sa_simt_conv3d_fwd(const sa_conv_desc* d, const void* x, const void* wp, const float* bias, const void* addend,
                       const void* mask, int relu, void* y, cudaStream_t st) {
  sa_note_path(SA_PATH_SIMT);
  if (d->act_dtype == SA_F32) return launch_fwd<float>(d, x, wp, bias, addend, mask, relu, y, st);
  return launch_fwd<__nv_bfloat16>(d, x, wp, bias, addend, mask, relu, y, st);
}

int sa_simt_conv3d_wgrad(const sa_conv_desc* d, const void* p, const void* q, float* dwp, cudaStream_t st) {
  sa_note_path(SA_PATH_SIMT);
  if (d->act_dtype == SA_F32) return launch_wgrad<float>(d, p, q, dwp, st);
  return launch_wgrad<__nv_bfloat16>(d, p, q, dwp, st);
}
