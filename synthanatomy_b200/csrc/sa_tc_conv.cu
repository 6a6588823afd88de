// tcgen05 implicit-GEMM 3-D convolution ("shift-GEMM") for sm_100a.
//
//   Y[b, g*os + oq, n] = epi( sum_{taps t} sum_{c}  Xmap[t][b, g + off[t], c] * Wp[wrow[t] + n][c] )
//
// One CTA computes a 128-position x N-channel output tile: the 128 positions are a (td x th x tw) box of the
// tile grid G, so for every tap the A operand is ONE 5-D TMA box load of the NDHWC activation tensor at a
// shifted coordinate (halo / padding = TMA out-of-bounds zero fill) and lands in shared memory already in the
// K-major SWIZZLE_128B layout tcgen05.mma wants (row = position, 64 channels = 128 B).  The B operand is the
// packed weight slice [n][c-chunk].  Accumulation over taps x channel chunks happens in TMEM (fp32); the
// epilogue (bias, residual addend, ReLU, ReLU-mask) runs out of TMEM and writes bf16 NDHWC.
//
//   stride-1 conv (3x3x3, 1x1x1)         : one activation map, offsets t - pad
//   4/2/1 strided conv                   : 8 parity views of X (doubled strides), 64 taps, offsets {-1,0,+1}
//   4/2/1 transposed conv (and dgrad of the strided conv): 8 launches (output parity phases) of 8 taps each,
//                                          output written with stride 2
//
// Warp roles (192 threads): warp 0 = TMA producer, warp 1 = TMEM owner + MMA issuer, warps 2-5 = epilogue.
//
// Reference call sites replaced: cuDNN conv fprop / dgrad behind nn.Conv3d / nn.ConvTranspose3d,
// src/networks/vqvae/baseline.py:153-160, 218-228, 242-244, 258, 283-297.
#include <mutex>
#include <stdlib.h>

#include "sa_tc_common.cuh"

using namespace satc;

namespace {

constexpr int TC_THREADS = 192;
constexpr int TC_M = 128;        // positions per tile == UMMA M
constexpr int TC_KC = 64;        // channels per pipeline stage (128 B rows)
constexpr int TC_MAX_TAPS = 64;
constexpr int TC_MAX_STAGES = 8;

struct TcTap {
  int8_t map, dd, dh, dw;  // activation map id and tile-coordinate offset
  int32_t wrow;            // first row of this tap's [N][Cin] slice in the packed weight matrix
};

struct TcConvParams {
  CUtensorMap amap[8];
  CUtensorMap wmap;
  TcTap taps[TC_MAX_TAPS];
  int ntaps, cchunks, N, stages, tmem_cols;
  int gD, gH, gW;           // tile grid extent (per batch element)
  int td, th, tw;           // tile box, td*th*tw == 128
  int ntd, nth, ntw;        // tiles per dim
  int oD, oH, oW;           // output tensor extent
  int os, oqd, oqh, oqw;    // output position = g*os + oq
  int relu;
  const float* bias;
  const __nv_bfloat16* addend;
  const __nv_bfloat16* mask;
  __nv_bfloat16* y;
};

__global__ void __launch_bounds__(TC_THREADS)
tc_conv_kernel(const __grid_constant__ TcConvParams P) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ uint64_t full_bar[TC_MAX_STAGES];
  __shared__ uint64_t empty_bar[TC_MAX_STAGES];
  __shared__ uint64_t tmem_full_bar;
  __shared__ uint32_t tmem_base_slot;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t a_bytes = TC_M * 128;
  const uint32_t b_bytes = (uint32_t)P.N * 128;
  const uint32_t stage_bytes = a_bytes + b_bytes;
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);

  // tile coordinates
  int tile = blockIdx.x;
  const int tw_i = tile % P.ntw; tile /= P.ntw;
  const int th_i = tile % P.nth; tile /= P.nth;
  const int td_i = tile % P.ntd; tile /= P.ntd;
  const int b = tile;
  const int g0d = td_i * P.td, g0h = th_i * P.th, g0w = tw_i * P.tw;

  if (warp == 0 && lane == 0) {
    for (int i = 0; i < P.stages; ++i) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], 1); }
    mbar_init(&tmem_full_bar, 1);
    fence_mbar_init();
    fence_proxy_async();
  }
  if (warp == 1) {
    tmem_alloc(&tmem_base_slot, (uint32_t)P.tmem_cols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_slot;

  const int iters = P.ntaps * P.cchunks;

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer
    if (lane == 0) {
      prefetch_tmap(&P.wmap);
      prefetch_tmap(&P.amap[0]);
      int stage = 0; uint32_t phase = 0;
      for (int it = 0; it < iters; ++it) {
        const int tap = it / P.cchunks, cc = it - tap * P.cchunks;
        const TcTap tp = P.taps[tap];
        mbar_wait(&empty_bar[stage], phase ^ 1);
        mbar_expect_tx(&full_bar[stage], stage_bytes);
        uint8_t* sa = smem + (size_t)stage * stage_bytes;
        tma_load_5d(sa, &P.amap[tp.map], &full_bar[stage], cc * TC_KC, g0w + tp.dw, g0h + tp.dh, g0d + tp.dd, b);
        tma_load_2d(sa + a_bytes, &P.wmap, &full_bar[stage], cc * TC_KC, tp.wrow);
        if (++stage == P.stages) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer (one elected lane)
    if (lane == 0) {
      const uint32_t idesc = make_idesc_bf16(TC_M, P.N, 0, 0);
      int stage = 0; uint32_t phase = 0;
      for (int it = 0; it < iters; ++it) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        const uint32_t sa = smem_u32(smem + (size_t)stage * stage_bytes);
        const uint32_t sb = sa + a_bytes;
#pragma unroll
        for (int k = 0; k < TC_KC / 16; ++k) {
          // K-major SWIZZLE_128B: 8-row groups 1024 B apart; a K step of 16 bf16 = 32 B inside the swizzle atom
          const uint64_t da = make_smem_desc(sa + k * 32, 16, 1024, 2);
          const uint64_t db = make_smem_desc(sb + k * 32, 16, 1024, 2);
          umma_bf16(tmem_base, da, db, idesc, (it | k) != 0);
        }
        umma_commit(&empty_bar[stage]);   // frees the smem slot once these MMAs retire
        if (++stage == P.stages) { stage = 0; phase ^= 1; }
      }
      umma_commit(&tmem_full_bar);        // accumulator complete
    }
  } else {
    // ------------------------------------------------------------ epilogue: TMEM -> regs -> bf16 NDHWC
    const int quad = warp & 3;                 // TMEM lane quadrant this warp may access
    const int r = quad * 32 + lane;            // tile row == TMEM lane
    const int wx = r % P.tw, hy = (r / P.tw) % P.th, dz = r / (P.tw * P.th);
    const int gd = g0d + dz, gh = g0h + hy, gw = g0w + wx;
    const bool valid = gd < P.gD && gh < P.gH && gw < P.gW;
    const int64_t opos = (((int64_t)b * P.oD + (gd * P.os + P.oqd)) * P.oH + (gh * P.os + P.oqh)) * P.oW +
                         (gw * P.os + P.oqw);
    const int64_t obase = opos * P.N;
    mbar_wait(&tmem_full_bar, 0);
    tc_fence_after();
    for (int c0 = 0; c0 < P.N; c0 += 32) {
      uint32_t v[32];
      tmem_ld_32x32(tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)c0, v);
      tmem_ld_wait();
      if (valid) {
        const int nc = min(32, P.N - c0);      // N is a multiple of 16
        float f[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[j]);
        if (P.bias) {
#pragma unroll
          for (int j = 0; j < 32; ++j) if (j < nc) f[j] += __ldg(P.bias + c0 + j);
        }
        if (P.addend) {
          const uint4* ap = reinterpret_cast<const uint4*>(P.addend + obase + c0);
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            if (q * 8 < nc) {
              const uint4 u = __ldg(ap + q);
              f[q * 8 + 0] += bf16lo(u.x); f[q * 8 + 1] += bf16hi(u.x);
              f[q * 8 + 2] += bf16lo(u.y); f[q * 8 + 3] += bf16hi(u.y);
              f[q * 8 + 4] += bf16lo(u.z); f[q * 8 + 5] += bf16hi(u.z);
              f[q * 8 + 6] += bf16lo(u.w); f[q * 8 + 7] += bf16hi(u.w);
            }
          }
        }
        if (P.relu) {
#pragma unroll
          for (int j = 0; j < 32; ++j) f[j] = fmaxf(f[j], 0.f);
        }
        if (P.mask) {
          const uint4* mp = reinterpret_cast<const uint4*>(P.mask + obase + c0);
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            if (q * 8 < nc) {
              const uint4 u = __ldg(mp + q);
              const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                if (!(bf16lo(w[e]) > 0.f)) f[q * 8 + 2 * e] = 0.f;
                if (!(bf16hi(w[e]) > 0.f)) f[q * 8 + 2 * e + 1] = 0.f;
              }
            }
          }
        }
        uint4* yp = reinterpret_cast<uint4*>(P.y + obase + c0);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          if (q * 8 < nc) {
            uint4 u;
            u.x = pack_bf16x2(f[q * 8 + 0], f[q * 8 + 1]);
            u.y = pack_bf16x2(f[q * 8 + 2], f[q * 8 + 3]);
            u.z = pack_bf16x2(f[q * 8 + 4], f[q * 8 + 5]);
            u.w = pack_bf16x2(f[q * 8 + 6], f[q * 8 + 7]);
            yp[q] = u;
          }
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, (uint32_t)P.tmem_cols);
}

// ---------------------------------------------------------------------------------------------- host
int pow2_ceil(int v) { int p = 1; while (p < v) p <<= 1; return p; }

bool choose_tile(int gD, int gH, int gW, int* td, int* th, int* tw) {
  // powers of two with td*th*tw == 128 and each t_i <= pow2_ceil(G_i) (a box may overhang the tensor: TMA zero-fills
  // and the epilogue masks); minimise padded volume, then prefer wide tw
  int64_t best = -1; int bd = 0, bh = 0, bw = 0;
  for (int w = 1; w <= 128; w <<= 1) {
    if (w > pow2_ceil(gW)) break;
    for (int h = 1; h * w <= 128; h <<= 1) {
      if (h > pow2_ceil(gH)) break;
      const int d = 128 / (w * h);
      if (d > pow2_ceil(gD)) continue;
      const int64_t vol = sa_cdiv(gD, d) * d * sa_cdiv(gH, h) * h * sa_cdiv(gW, w) * w;
      if (best < 0 || vol < best || (vol == best && w > bw)) { best = vol; bd = d; bh = h; bw = w; }
    }
  }
  if (best < 0) return false;
  *td = bd; *th = bh; *tw = bw;
  return true;
}

int next_pow2_cols(int n) { int c = 32; while (c < n) c <<= 1; return c; }

std::once_flag g_attr_once;
int g_max_smem = 0;

int launch(TcConvParams& P, int batch, cudaStream_t st) {
  const size_t stage_bytes = (size_t)TC_M * 128 + (size_t)P.N * 128;
  std::call_once(g_attr_once, [] {
    int dev = 0; cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&g_max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    cudaFuncSetAttribute(tc_conv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, g_max_smem - 1024);
  });
  const size_t budget = (size_t)g_max_smem - 1024 /*static*/ - 1024 /*align*/;
  int stages = (int)(budget / stage_bytes);
  if (stages > TC_MAX_STAGES) stages = TC_MAX_STAGES;
  if (const char* e = getenv("SA_TC_MAX_STAGES")) {   // tuning knob: fewer stages => more co-resident CTAs per SM
    const int v = atoi(e);
    if (v >= 1 && v < stages) stages = v;
  }
  const int iters = P.ntaps * P.cchunks;
  if (stages > iters) stages = iters;
  if (stages < 1) { sa_set_error("tc_conv: stage does not fit shared memory"); return SA_ERR_UNSUPPORTED; }
  P.stages = stages;
  const size_t smem = stages * stage_bytes + 1024;
  const unsigned grid = (unsigned)((int64_t)P.ntd * P.nth * P.ntw * batch);
  tc_conv_kernel<<<grid, TC_THREADS, smem, st>>>(P);
  SA_LAUNCH_CHECK();
  return SA_OK;
}

int make_act_map(CUtensorMap* m, const void* base, int C, int D, int H, int W, int B, int sub, int pd, int ph, int pw,
                 int td, int th, int tw) {
  // view X[b, d*sub + pd, h*sub + ph, w*sub + pw, c]; dims innermost first
  const uint64_t dims[5] = {(uint64_t)C, (uint64_t)(W / sub), (uint64_t)(H / sub), (uint64_t)(D / sub), (uint64_t)B};
  const uint64_t es = 2;
  const uint64_t strides[5] = {es, (uint64_t)C * es * sub, (uint64_t)W * C * es * sub, (uint64_t)H * W * C * es * sub,
                               (uint64_t)D * H * W * C * es};
  const uint32_t box[5] = {64, (uint32_t)tw, (uint32_t)th, (uint32_t)td, 1};
  const uint8_t* p = (const uint8_t*)base + ((uint64_t)pd * H * W + (uint64_t)ph * W + pw) * C * es;
  return sa_make_tmap_bf16(m, p, 5, dims, strides, box);
}

}  // namespace

bool sa_tc_conv3d_supported(const sa_conv_desc* d) {
  if (d->act_dtype != SA_BF16) return false;
  if (d->c_in % 64 != 0 || d->c_out % 16 != 0 || d->c_out < 16 || d->c_out > 256) return false;
  const int k = d->ksize, s = d->stride, p = d->pad;
  int gD, gH, gW;
  if (s == 1) {
    // (transposed stride-1 is the same map with flipped taps / pad' = k-1-p)
    const int pe = d->transposed ? k - 1 - p : p;
    for (int i = 0; i < 3; ++i) if (d->out_dhw[i] != d->in_dhw[i] + 2 * pe - k + 1) return false;
    if (k * k * k > TC_MAX_TAPS) return false;
    gD = d->out_dhw[0]; gH = d->out_dhw[1]; gW = d->out_dhw[2];
  } else if (s == 2 && k == 4 && p == 1) {
    if (!d->transposed) {
      for (int i = 0; i < 3; ++i) if (d->in_dhw[i] % 2 || d->out_dhw[i] * 2 != d->in_dhw[i]) return false;
      gD = d->out_dhw[0]; gH = d->out_dhw[1]; gW = d->out_dhw[2];
    } else {
      for (int i = 0; i < 3; ++i) if (d->out_dhw[i] != d->in_dhw[i] * 2) return false;
      gD = d->in_dhw[0]; gH = d->in_dhw[1]; gW = d->in_dhw[2];
    }
  } else {
    return false;
  }
  int td, th, tw;
  return choose_tile(gD, gH, gW, &td, &th, &tw);
}

int sa_tc_conv3d_fwd(const sa_conv_desc* d, const void* x, const void* wp, const float* bias, const void* addend,
                     const void* mask, int relu, void* y, cudaStream_t st) {
  if (!sa_tc_conv3d_supported(d)) { sa_set_error("tc_conv: unsupported configuration"); return SA_ERR_UNSUPPORTED; }
  if (!sa_get_tmap_encode()) { sa_set_error("tc_conv: cuTensorMapEncodeTiled unavailable"); return SA_ERR_CUDA; }
  sa_note_path(SA_PATH_TCGEN05);
  const int k = d->ksize, s = d->stride, p = d->pad;
  const int iD = d->in_dhw[0], iH = d->in_dhw[1], iW = d->in_dhw[2];
  const int oD = d->out_dhw[0], oH = d->out_dhw[1], oW = d->out_dhw[2];
  const int taps_total = k * k * k;

  static thread_local TcConvParams P;   // 1.9 KB; avoid re-zeroing the maps on the stack for every call
  P.N = d->c_out;
  P.cchunks = d->c_in / TC_KC;
  P.tmem_cols = next_pow2_cols(P.N);
  P.relu = relu;
  P.bias = bias;
  P.addend = (const __nv_bfloat16*)addend;
  P.mask = (const __nv_bfloat16*)mask;
  P.y = (__nv_bfloat16*)y;
  P.oD = oD; P.oH = oH; P.oW = oW;

  // weight matrix: rows = taps * c_out, cols = c_in (K-major)
  {
    const uint64_t dims[2] = {(uint64_t)d->c_in, (uint64_t)taps_total * d->c_out};
    const uint64_t strides[2] = {2, (uint64_t)d->c_in * 2};
    const uint32_t box[2] = {64, (uint32_t)d->c_out};
    int rc = sa_make_tmap_bf16(&P.wmap, wp, 2, dims, strides, box);
    if (rc != SA_OK) return rc;
  }

  auto set_tiles = [&](int gD, int gH, int gW) {
    P.gD = gD; P.gH = gH; P.gW = gW;
    choose_tile(gD, gH, gW, &P.td, &P.th, &P.tw);
    P.ntd = (int)sa_cdiv(gD, P.td); P.nth = (int)sa_cdiv(gH, P.th); P.ntw = (int)sa_cdiv(gW, P.tw);
  };

  if (s == 1) {
    const int pe = d->transposed ? k - 1 - p : p;
    set_tiles(oD, oH, oW);
    P.os = 1; P.oqd = P.oqh = P.oqw = 0;
    int rc = make_act_map(&P.amap[0], x, d->c_in, iD, iH, iW, d->batch, 1, 0, 0, 0, P.td, P.th, P.tw);
    if (rc != SA_OK) return rc;
    P.ntaps = taps_total;
    for (int t = 0; t < taps_total; ++t) {
      const int tw_ = t % k, th_ = (t / k) % k, td_ = t / (k * k);
      const int tsrc = d->transposed ? taps_total - 1 - t : t;   // flipped taps for the transposed form
      P.taps[t] = TcTap{0, (int8_t)(td_ - pe), (int8_t)(th_ - pe), (int8_t)(tw_ - pe), tsrc * d->c_out};
    }
    return launch(P, d->batch, st);
  }

  if (!d->transposed) {
    // i = 2g - 1 + t : t even -> odd-parity view, t odd -> even-parity view
    set_tiles(oD, oH, oW);
    P.os = 1; P.oqd = P.oqh = P.oqw = 0;
    for (int m = 0; m < 8; ++m) {
      int rc = make_act_map(&P.amap[m], x, d->c_in, iD, iH, iW, d->batch, 2, (m >> 2) & 1, (m >> 1) & 1, m & 1, P.td,
                            P.th, P.tw);
      if (rc != SA_OK) return rc;
    }
    auto par = [](int t) { return (t & 1) ? 0 : 1; };
    auto off = [](int t) { return t == 0 ? -1 : (t == 3 ? 1 : 0); };
    P.ntaps = 64;
    for (int t = 0; t < 64; ++t) {
      const int tw_ = t % 4, th_ = (t / 4) % 4, td_ = t / 16;
      const int map = (par(td_) << 2) | (par(th_) << 1) | par(tw_);
      P.taps[t] = TcTap{(int8_t)map, (int8_t)off(td_), (int8_t)off(th_), (int8_t)off(tw_), t * d->c_out};
    }
    return launch(P, d->batch, st);
  }

  // transposed 4/2/1: output parity phase q -> two taps per dim:  q=0: (t=1, off 0), (t=3, off -1)
  //                                                              q=1: (t=0, off +1), (t=2, off 0)
  set_tiles(iD, iH, iW);
  P.os = 2;
  {
    int rc = make_act_map(&P.amap[0], x, d->c_in, iD, iH, iW, d->batch, 1, 0, 0, 0, P.td, P.th, P.tw);
    if (rc != SA_OK) return rc;
  }
  static const int tap_of[2][2] = {{1, 3}, {0, 2}};
  static const int off_of[2][2] = {{0, -1}, {1, 0}};
  for (int q = 0; q < 8; ++q) {
    const int qd = (q >> 2) & 1, qh = (q >> 1) & 1, qw = q & 1;
    P.oqd = qd; P.oqh = qh; P.oqw = qw;
    P.ntaps = 8;
    for (int j = 0; j < 8; ++j) {
      const int jd = (j >> 2) & 1, jh = (j >> 1) & 1, jw = j & 1;
      const int t = (tap_of[qd][jd] * 4 + tap_of[qh][jh]) * 4 + tap_of[qw][jw];
      P.taps[j] = TcTap{0, (int8_t)off_of[qd][jd], (int8_t)off_of[qh][jh], (int8_t)off_of[qw][jw], t * d->c_out};
    }
    int rc = launch(P, d->batch, st);
    if (rc != SA_OK) return rc;
  }
  return SA_OK;
}
