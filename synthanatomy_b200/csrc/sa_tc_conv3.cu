// tcgen05 implicit-GEMM for the stride-1 3x3x3 convolutions (and their data gradients) of the VQ-VAE, sm_100a.
//
// The general shift-GEMM kernel (sa_tc_conv.cu) reloads a 128-position activation box and a weight slice for every
// tap: 32 KB of L2 -> shared memory traffic per four 128x128x16 MMAs = 128 B / clk / SM, three times what the L2 can
// feed 148 SMs.  This kernel cuts the traffic per MMA by 2.2x:
//
//   * tile = 8 x 4 x 8 output positions (256 rows = TWO 128-row accumulators that share every weight slice),
//   * for a fixed (dh, dw) the three depth taps read ONE activation box that is two planes deeper (10 x 4 x 8): a
//     depth shift is a whole number of 32-row planes = 4 KB, so the shifted A operand is the same shared-memory box
//     at a 1024-byte-aligned offset -- no re-load, the 128B swizzle atoms stay intact,
//   * persistent CTAs (one per SM) with two TMEM accumulator stages (2 x 256 columns): the epilogue of tile i runs
//     under the MMAs of tile i+1.
//
// Per pipeline stage (one (dh, dw, 64-channel chunk)): A box 40 KB + 3 weight slices (3 x N x 128 B) for 24 MMAs.
// Warp roles (320 threads): warp 0 = TMA producer, warp 1 = TMEM owner + MMA issuer, warps 2-9 = epilogue
// (TMEM lane quadrant = warp & 3, accumulator half = (warp - 2) >> 2).
//
// Reference call sites replaced: cuDNN fprop / dgrad behind the 3x3x3 nn.Conv3d of ResidualLayer and of the
// encoder / decoder heads, /root/reference/src/networks/vqvae/baseline.py:153-160, 242-244, 258.
#include <mutex>
#include <stdlib.h>

#include "sa_tc_common.cuh"

using namespace satc;

namespace {

constexpr int C3_THREADS = 320;
constexpr int C3_TD = 8, C3_TH = 4, C3_TW = 8;
constexpr int C3_PLANE = C3_TH * C3_TW;                  // 32 rows
constexpr uint32_t C3_PLANE_BYTES = C3_PLANE * 128;      // 4 KB
constexpr uint32_t C3_A_BYTES = (C3_TD + 2) * C3_PLANE_BYTES;   // 40 KB
constexpr int C3_MAX_STAGES = 4;

// one pipeline iteration per (group, 64-channel chunk): ONE activation box (ndd + 7 planes deep) serves ndd depth taps
struct C3Group {
  int8_t map, od, oh, ow;   // activation view and box origin relative to the tile origin
  int32_t wrow[3];          // first row of each depth tap's [N][c_in] slice in the packed weight matrix
};

struct C3Params {
  CUtensorMap amap[8];
  CUtensorMap wmap;
  C3Group groups[32];
  int ngroups, ndd;
  int N, cchunks, stages;                              // N: output channels of THIS launch (<= 128)
  int ldy;                                             // channel stride of y / addend / mask (the layer's c_out)
  int gD, gH, gW, ntd, nth, ntw, batch, total_tiles;   // tile grid
  int oD, oH, oW, os, oqd, oqh, oqw;                   // output position = g * os + oq
  int relu;
  const float* bias;
  const void* addend;          // OutT
  const void* mask;            // OutT
  void* y;                     // OutT
};

// OutT = __nv_bfloat16: the bf16 training path.  OutT = float: the bf16x3 "parity" path (sa_x3.cu) -- the operands are
// hi / lo bf16 splits of fp32 tensors concatenated along the channel axis, the accumulator (= the fp32-class result)
// leaves as fp32 and the epilogue tensors are fp32.
template <typename OutT>
__global__ void __launch_bounds__(C3_THREADS, 1)
tc_conv3_kernel(const __grid_constant__ C3Params P) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ uint64_t full_bar[C3_MAX_STAGES], empty_bar[C3_MAX_STAGES];
  __shared__ uint64_t tmem_full_bar[2], tmem_empty_bar[2];
  __shared__ uint32_t tmem_base_slot;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t b_bytes = (uint32_t)P.N * 128;
  const int ndd = P.ndd;                                   // depth taps that share one A box
  const uint32_t a_bytes = (uint32_t)(C3_TD + ndd - 1) * C3_PLANE_BYTES;
  const uint32_t stage_bytes = a_bytes + (uint32_t)ndd * b_bytes;
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const int iters = P.ngroups * P.cchunks;

  if (warp == 0 && lane == 0) {
    for (int i = 0; i < P.stages; ++i) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&tmem_full_bar[i], 1); mbar_init(&tmem_empty_bar[i], 8); }
    fence_mbar_init();
    fence_proxy_async();
  }
  if (warp == 1) { tmem_alloc(&tmem_base_slot, 512); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_slot;

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer
    if (lane == 0) {
      prefetch_tmap(&P.wmap);
      prefetch_tmap(&P.amap[0]);
      int stage = 0; uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < P.total_tiles; tile += gridDim.x) {
        int t = tile;
        const int tw_i = t % P.ntw; t /= P.ntw;
        const int th_i = t % P.nth; t /= P.nth;
        const int td_i = t % P.ntd; t /= P.ntd;
        const int b = t;
        const int g0d = td_i * C3_TD, g0h = th_i * C3_TH, g0w = tw_i * C3_TW;
        for (int it = 0; it < iters; ++it) {
          const int gi = it / P.cchunks, cc = it - gi * P.cchunks;
          const C3Group g = P.groups[gi];
          mbar_wait(&empty_bar[stage], phase ^ 1);
          mbar_expect_tx(&full_bar[stage], stage_bytes);
          uint8_t* sa = smem + (size_t)stage * stage_bytes;
          tma_load_5d(sa, &P.amap[g.map], &full_bar[stage], cc * 64, g0w + g.ow, g0h + g.oh, g0d + g.od, b);
          for (int dd = 0; dd < ndd; ++dd)
            tma_load_2d(sa + a_bytes + dd * b_bytes, &P.wmap, &full_bar[stage], cc * 64, g.wrow[dd]);
          if (++stage == P.stages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer
    if (lane == 0) {
      const uint32_t idesc = make_idesc_bf16(128, P.N, 0, 0);
      int stage = 0; uint32_t phase = 0;
      int acc = 0; uint32_t acc_phase = 0;
      for (int tile = blockIdx.x; tile < P.total_tiles; tile += gridDim.x) {
        mbar_wait(&tmem_empty_bar[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d0 = tmem_base + (uint32_t)(acc * 256);
        for (int it = 0; it < iters; ++it) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + (size_t)stage * stage_bytes);
          const uint32_t sb = sa + a_bytes;
          for (int dd = 0; dd < ndd; ++dd) {
#pragma unroll
            for (int half = 0; half < 2; ++half) {
              const uint32_t a0 = sa + (uint32_t)(dd + 4 * half) * C3_PLANE_BYTES;
#pragma unroll
              for (int k = 0; k < 4; ++k)
                umma_bf16(d0 + (uint32_t)(half * 128), make_smem_desc(a0 + k * 32, 16, 1024, 2),
                          make_smem_desc(sb + dd * b_bytes + k * 32, 16, 1024, 2), idesc, (it | dd | k) != 0);
            }
          }
          umma_commit(&empty_bar[stage]);
          if (++stage == P.stages) { stage = 0; phase ^= 1; }
        }
        umma_commit(&tmem_full_bar[acc]);
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else {
    // ------------------------------------------------------------ epilogue: TMEM -> regs -> bf16 NDHWC
    const int ew = warp - 2;
    const int quad = warp & 3;
    const int half = ew >> 2;
    const int r = quad * 32 + lane;                       // row of this half's accumulator
    const int dz = half * 4 + (r >> 5), hy = (r >> 3) & 3, wx = r & 7;
    int acc = 0; uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < P.total_tiles; tile += gridDim.x) {
      int t = tile;
      const int tw_i = t % P.ntw; t /= P.ntw;
      const int th_i = t % P.nth; t /= P.nth;
      const int td_i = t % P.ntd; t /= P.ntd;
      const int b = t;
      const int gd = td_i * C3_TD + dz, gh = th_i * C3_TH + hy, gw = tw_i * C3_TW + wx;
      const bool valid = gd < P.gD && gh < P.gH && gw < P.gW;
      const int64_t obase = ((((int64_t)b * P.oD + (gd * P.os + P.oqd)) * P.oH + (gh * P.os + P.oqh)) * P.oW +
                             (gw * P.os + P.oqw)) * P.ldy;
      mbar_wait(&tmem_full_bar[acc], acc_phase);
      tc_fence_after();
      for (int c0 = 0; c0 < P.N; c0 += 32) {
        uint32_t v[32];
        tmem_ld_32x32(tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(acc * 256 + half * 128 + c0), v);
        tmem_ld_wait();
        if (valid) {
          const int nc = min(32, P.N - c0);      // N is a multiple of 16
          float f[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[j]);
          if (P.bias) {
#pragma unroll
            for (int q = 0; q < 8; ++q) {
              if (q * 4 < nc) {
                const float4 bb = __ldg(reinterpret_cast<const float4*>(P.bias + c0) + q);
                f[q * 4 + 0] += bb.x; f[q * 4 + 1] += bb.y; f[q * 4 + 2] += bb.z; f[q * 4 + 3] += bb.w;
              }
            }
          }
          if constexpr (sizeof(OutT) == 4) {
            // ---- fp32 epilogue tensors (bf16x3 path)
            if (P.addend) {
              const float4* ap = reinterpret_cast<const float4*>(reinterpret_cast<const float*>(P.addend) + obase + c0);
#pragma unroll
              for (int q = 0; q < 8; ++q) {
                if (q * 4 < nc) {
                  const float4 u = __ldg(ap + q);
                  f[q * 4 + 0] += u.x; f[q * 4 + 1] += u.y; f[q * 4 + 2] += u.z; f[q * 4 + 3] += u.w;
                }
              }
            }
            if (P.relu) {
#pragma unroll
              for (int j = 0; j < 32; ++j) f[j] = fmaxf(f[j], 0.f);
            }
            if (P.mask) {
              const float4* mp = reinterpret_cast<const float4*>(reinterpret_cast<const float*>(P.mask) + obase + c0);
#pragma unroll
              for (int q = 0; q < 8; ++q) {
                if (q * 4 < nc) {
                  const float4 u = __ldg(mp + q);
                  if (!(u.x > 0.f)) f[q * 4 + 0] = 0.f;
                  if (!(u.y > 0.f)) f[q * 4 + 1] = 0.f;
                  if (!(u.z > 0.f)) f[q * 4 + 2] = 0.f;
                  if (!(u.w > 0.f)) f[q * 4 + 3] = 0.f;
                }
              }
            }
            float4* yp = reinterpret_cast<float4*>(reinterpret_cast<float*>(P.y) + obase + c0);
#pragma unroll
            for (int q = 0; q < 8; ++q) {
              if (q * 4 < nc) yp[q] = make_float4(f[q * 4 + 0], f[q * 4 + 1], f[q * 4 + 2], f[q * 4 + 3]);
            }
          } else {
          if (P.addend) {
            const uint4* ap = reinterpret_cast<const uint4*>(reinterpret_cast<const __nv_bfloat16*>(P.addend) + obase + c0);
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
            const uint4* mp = reinterpret_cast<const uint4*>(reinterpret_cast<const __nv_bfloat16*>(P.mask) + obase + c0);
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
          uint4* yp = reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(P.y) + obase + c0);
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
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty_bar[acc]);
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 512);
}

std::once_flag g_c3_once;
int g_c3_sms = 148;

}  // namespace

namespace {

int c3_act_map(CUtensorMap* m, const void* base, int C, int D, int H, int W, int B, int sub, int pd, int ph, int pw, int planes) {
  // view X[b, d*sub + pd, h*sub + ph, w*sub + pw, c]; dims innermost first
  const uint64_t dims[5] = {(uint64_t)C, (uint64_t)(W / sub), (uint64_t)(H / sub), (uint64_t)(D / sub), (uint64_t)B};
  const uint64_t es = 2;
  const uint64_t strides[5] = {es, (uint64_t)C * es * sub, (uint64_t)W * C * es * sub, (uint64_t)H * W * C * es * sub,
                               (uint64_t)D * H * W * C * es};
  const uint32_t box[5] = {64, C3_TW, C3_TH, (uint32_t)planes, 1};
  const uint8_t* p = (const uint8_t*)base + ((uint64_t)pd * H * W + (uint64_t)ph * W + pw) * C * es;
  return sa_make_tmap_bf16(m, p, 5, dims, strides, box);
}

int c3_launch(C3Params& P, int batch, bool out_f32, cudaStream_t st) {
  P.ntd = (int)sa_cdiv(P.gD, C3_TD); P.nth = (int)sa_cdiv(P.gH, C3_TH); P.ntw = (int)sa_cdiv(P.gW, C3_TW);
  P.batch = batch;
  const int64_t total = (int64_t)P.ntd * P.nth * P.ntw * batch;
  if (total >= (1LL << 31)) { sa_set_error("tc_conv3: too many tiles"); return SA_ERR_UNSUPPORTED; }
  P.total_tiles = (int)total;
  const size_t stage_bytes = (size_t)(C3_TD + P.ndd - 1) * C3_PLANE_BYTES + (size_t)P.ndd * P.N * 128;
  int stages = (int)((227 * 1024 - 2048 - 1024) / stage_bytes);
  if (stages > C3_MAX_STAGES) stages = C3_MAX_STAGES;
  if (stages > P.ngroups * P.cchunks) stages = P.ngroups * P.cchunks;
  if (stages < 1) { sa_set_error("tc_conv3: stage does not fit shared memory"); return SA_ERR_UNSUPPORTED; }
  P.stages = stages;
  const size_t smem = stages * stage_bytes + 1024;
  const unsigned grid = (unsigned)(P.total_tiles < g_c3_sms ? P.total_tiles : g_c3_sms);
  if (out_f32) tc_conv3_kernel<float><<<grid, C3_THREADS, smem, st>>>(P);
  else tc_conv3_kernel<__nv_bfloat16><<<grid, C3_THREADS, smem, st>>>(P);
  SA_LAUNCH_CHECK();
  return SA_OK;
}

}  // namespace

// stride-1 k = 3 / k = 1 convs (and their transposed forms), 4/2/1 strided convs and 4/2/1 transposed convs
bool sa_tc_conv3_supported(const sa_conv_desc* d) {
  if (d->act_dtype != SA_BF16) return false;
  // up to 128 output channels per launch; 256 = two launches over the two halves of the output channels
  // input channels: whole 64-channel TMA boxes, or ONE partial box (c_in = 16 / 32 / 48: the box is 64 wide, the
  // channels past c_in are zero-filled by the TMA unit on both operands) -- the 32-channel quantiser projections
  if ((d->c_in % 64 != 0 && !(d->c_in < 64 && d->c_in % 16 == 0)) || d->c_in < 16) return false;
  if (d->c_out % 16 != 0 || d->c_out < 16 || (d->c_out > 128 && d->c_out != 256)) return false;
  if (const char* e = getenv("SA_TC_CONV3")) { if (e[0] == '0') return false; }   // A/B switch for benchmarking
  if (d->c_out == 256) { if (const char* e = getenv("SA_TC_CONV3_N256")) { if (e[0] == '0') return false; } }
  if (!sa_get_tmap_encode()) return false;
  const int k = d->ksize;
  if (d->stride == 1 && (k == 3 || k == 1)) {
    const int pe = d->transposed ? k - 1 - d->pad : d->pad;
    if (pe < 0 || pe > k - 1) return false;
    for (int i = 0; i < 3; ++i) if (d->out_dhw[i] != d->in_dhw[i] + 2 * pe - (k - 1)) return false;
    return true;
  }
  if (d->stride == 2 && k == 4 && d->pad == 1) {
    if (const char* e = getenv("SA_TC_CONV3_S2")) { if (e[0] == '0') return false; }
    if (!d->transposed) {
      for (int i = 0; i < 3; ++i) if (d->in_dhw[i] % 2 || d->out_dhw[i] * 2 != d->in_dhw[i]) return false;
    } else {
      for (int i = 0; i < 3; ++i) if (d->out_dhw[i] != d->in_dhw[i] * 2) return false;
    }
    return true;
  }
  return false;
}

int sa_tc_conv3_fwd_ex(const sa_conv_desc* d, const void* x, const void* wp, const float* bias, const void* addend,
                       const void* mask, int relu, void* y, bool out_f32, cudaStream_t st);

int sa_tc_conv3_fwd(const sa_conv_desc* d, const void* x, const void* wp, const float* bias, const void* addend,
                    const void* mask, int relu, void* y, cudaStream_t st) {
  return sa_tc_conv3_fwd_ex(d, x, wp, bias, addend, mask, relu, y, false, st);
}

static int c3_fwd_part(const sa_conv_desc* d, const void* x, const void* wp, const float* bias, const void* addend,
                       const void* mask, int relu, void* y, bool out_f32, int n0, int N, cudaStream_t st);

// out_f32: y / addend / mask are fp32 tensors (x and wp stay bf16): the bf16x3 path of sa_x3.cu
int sa_tc_conv3_fwd_ex(const sa_conv_desc* d, const void* x, const void* wp, const float* bias, const void* addend,
                       const void* mask, int relu, void* y, bool out_f32, cudaStream_t st) {
  const int parts = d->c_out > 128 ? 2 : 1;
  const int N = d->c_out / parts;
  for (int h = 0; h < parts; ++h) {
    const int rc = c3_fwd_part(d, x, wp, bias, addend, mask, relu, y, out_f32, h * N, N, st);
    if (rc != SA_OK) return rc;
  }
  return SA_OK;
}

// output channels [n0, n0 + N) of the layer
static int c3_fwd_part(const sa_conv_desc* d, const void* x, const void* wp, const float* bias, const void* addend,
                       const void* mask, int relu, void* y, bool out_f32, int n0, int N, cudaStream_t st) {
  std::call_once(g_c3_once, [] {
    int dev = 0, v = 0;
    cudaGetDevice(&dev);
    if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && v > 0) g_c3_sms = v;
    cudaFuncSetAttribute(tc_conv3_kernel<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024 - 2048);
    cudaFuncSetAttribute(tc_conv3_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024 - 2048);
  });
  sa_note_path(SA_PATH_TCGEN05);
  if (bias && (reinterpret_cast<uintptr_t>(bias) & 15)) { sa_set_error("tc_conv3: bias not 16-byte aligned"); return SA_ERR_INVALID; }
  static thread_local C3Params P;
  const int iD = d->in_dhw[0], iH = d->in_dhw[1], iW = d->in_dhw[2];
  const int k = d->ksize, taps = k * k * k;
  P.N = N;
  P.ldy = d->c_out;
  P.cchunks = (d->c_in + 63) / 64;
  P.oD = d->out_dhw[0]; P.oH = d->out_dhw[1]; P.oW = d->out_dhw[2];
  const size_t es = out_f32 ? 4 : 2;
  P.relu = relu; P.bias = bias ? bias + n0 : nullptr;
  P.addend = addend ? (const uint8_t*)addend + (size_t)n0 * es : nullptr;
  P.mask = mask ? (const uint8_t*)mask + (size_t)n0 * es : nullptr;
  P.y = (uint8_t*)y + (size_t)n0 * es;
  {
    const uint64_t dims[2] = {(uint64_t)d->c_in, (uint64_t)taps * d->c_out};
    const uint64_t strides[2] = {2, (uint64_t)d->c_in * 2};
    const uint32_t box[2] = {64, (uint32_t)N};
    int rc = sa_make_tmap_bf16(&P.wmap, wp, 2, dims, strides, box);
    if (rc != SA_OK) return rc;
  }
  int rc;
  if (d->stride == 1) {
    // k x k (dh, dw) groups, each with the k depth taps in one box
    const int pe = d->transposed ? k - 1 - d->pad : d->pad;
    P.ndd = k; P.ngroups = k * k;
    P.gD = P.oD; P.gH = P.oH; P.gW = P.oW;
    P.os = 1; P.oqd = P.oqh = P.oqw = 0;
    if ((rc = c3_act_map(&P.amap[0], x, d->c_in, iD, iH, iW, d->batch, 1, 0, 0, 0, C3_TD + k - 1)) != SA_OK) return rc;
    for (int dh = 0; dh < k; ++dh)
      for (int dw = 0; dw < k; ++dw) {
        C3Group& g = P.groups[dh * k + dw];
        g.map = 0; g.od = (int8_t)(-pe); g.oh = (int8_t)(dh - pe); g.ow = (int8_t)(dw - pe);
        for (int dd = 0; dd < k; ++dd) {
          const int t = (dd * k + dh) * k + dw;
          g.wrow[dd] = (d->transposed ? taps - 1 - t : t) * d->c_out + n0;    // flipped taps: transposed form
        }
      }
    return c3_launch(P, d->batch, out_f32, st);
  }
  // ---- 4/2/1: per dimension tap t <-> (parity view, offset):  0: (odd, -1)  1: (even, 0)  2: (odd, 0)  3: (even, +1)
  auto par = [](int t) { return (t & 1) ? 0 : 1; };
  auto off = [](int t) { return t == 0 ? -1 : (t == 3 ? 1 : 0); };
  P.ndd = 2;
  if (!d->transposed) {
    // strided conv: i = 2 g - 1 + t.  The two depth taps of one parity ({0, 2}: offsets -1, 0; {1, 3}: offsets 0, +1) share a box
    P.gD = P.oD; P.gH = P.oH; P.gW = P.oW;
    P.os = 1; P.oqd = P.oqh = P.oqw = 0;
    for (int m = 0; m < 8; ++m)
      if ((rc = c3_act_map(&P.amap[m], x, d->c_in, iD, iH, iW, d->batch, 2, (m >> 2) & 1, (m >> 1) & 1, m & 1, C3_TD + 1)) != SA_OK)
        return rc;
    int gi = 0;
    for (int pd = 0; pd < 2; ++pd)
      for (int th = 0; th < 4; ++th)
        for (int tw = 0; tw < 4; ++tw) {
          C3Group& g = P.groups[gi++];
          const int td0 = pd ? 0 : 1, td1 = pd ? 2 : 3;       // odd-parity view: taps 0, 2   even: taps 1, 3
          g.map = (int8_t)((pd << 2) | (par(th) << 1) | par(tw));
          g.od = (int8_t)off(td0); g.oh = (int8_t)off(th); g.ow = (int8_t)off(tw);
          g.wrow[0] = ((td0 * 4 + th) * 4 + tw) * d->c_out + n0;
          g.wrow[1] = ((td1 * 4 + th) * 4 + tw) * d->c_out + n0;
          g.wrow[2] = 0;
        }
    P.ngroups = gi;
    return c3_launch(P, d->batch, out_f32, st);
  }
  // transposed 4/2/1: output parity phase q -> two taps per dim:  q=0: (t=3, off -1), (t=1, off 0)   q=1: (t=2, off 0), (t=0, off +1)
  P.gD = iD; P.gH = iH; P.gW = iW;
  P.os = 2;
  if ((rc = c3_act_map(&P.amap[0], x, d->c_in, iD, iH, iW, d->batch, 1, 0, 0, 0, C3_TD + 1)) != SA_OK) return rc;
  static const int tap_of[2][2] = {{3, 1}, {2, 0}};      // ordered by increasing offset
  static const int off_of[2][2] = {{-1, 0}, {0, 1}};
  for (int q = 0; q < 8; ++q) {
    const int qd = (q >> 2) & 1, qh = (q >> 1) & 1, qw = q & 1;
    P.oqd = qd; P.oqh = qh; P.oqw = qw;
    P.ngroups = 4;
    for (int j = 0; j < 4; ++j) {
      const int jh = (j >> 1) & 1, jw = j & 1;
      C3Group& g = P.groups[j];
      g.map = 0; g.od = (int8_t)off_of[qd][0]; g.oh = (int8_t)off_of[qh][jh]; g.ow = (int8_t)off_of[qw][jw];
      for (int dd = 0; dd < 2; ++dd)
        g.wrow[dd] = ((tap_of[qd][dd] * 4 + tap_of[qh][jh]) * 4 + tap_of[qw][jw]) * d->c_out + n0;
      g.wrow[2] = 0;
    }
    if ((rc = c3_launch(P, d->batch, out_f32, st)) != SA_OK) return rc;
  }
  return SA_OK;
}
