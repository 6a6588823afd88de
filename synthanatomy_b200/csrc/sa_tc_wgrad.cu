// tcgen05 weight-gradient kernel for sm_100a.
//
//   dWp[t][n][c] += sum_{b, o}  P[b, o, n] * Q[b, o*stride - pad + t, c]
//
// Per tap this is a GEMM with M = n (output channels of the primitive), N = c, K = positions.  Both operands
// are "MN-major" in memory (channels contiguous, positions strided), which tcgen05 reads directly from the
// 128B-swizzled tiles TMA produces: a (64 channels x KP positions) box per channel block, row = position.
// A CTA owns (a group of T = 512 / c_in taps) x (one 128-row half of n) x (a contiguous range of position tiles):
// the T accumulators [128 x c_in] fp32 fill the 512 TMEM columns, the P tile is loaded once per stage and
// shared by the T taps, and the split-K partials are reduced with fp32 red.global.add into dWp.
//
// Reference call sites replaced: cuDNN wgrad reached through autograd of the convs at
// src/networks/vqvae/baseline.py:153-156, 218-227, 283-293.
#include <mutex>
#include <stdlib.h>

#include "sa_tc_common.cuh"

using namespace satc;

namespace {

constexpr int WG_THREADS = 192;
constexpr int WG_KP = 32;            // positions per pipeline stage
constexpr int WG_MAX_TAPS = 64;
constexpr int WG_MAX_STAGES = 8;
constexpr int WG_BLOCK_BYTES = WG_KP * 128;   // one (64 ch x KP pos) swizzled block

struct WgTap {
  int8_t map, dd, dh, dw;
};

struct WgParams {
  CUtensorMap pmap;
  CUtensorMap qmap[8];
  WgTap taps[WG_MAX_TAPS];
  int ntaps, tpg, ngroups;     // taps, taps per group, groups
  int Cn, Cc;                  // channels of P (n) and Q (c)
  int nhalves, cblocks;        // Cn / 128, Cc / 64
  int stages;
  int td, th, tw, ntd, nth, ntw;
  int64_t tiles_total, tiles_per_split;
  float* dwp;
  float* parts;        // deterministic mode: split s stores its partial dWp at parts + s * out_size (summed in order afterwards)
  int64_t out_size;
};

__global__ void __launch_bounds__(WG_THREADS)
tc_wgrad_kernel(const __grid_constant__ WgParams P) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ uint64_t full_bar[WG_MAX_STAGES];
  __shared__ uint64_t empty_bar[WG_MAX_STAGES];
  __shared__ uint64_t tmem_full_bar;
  __shared__ uint32_t tmem_base_slot;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);

  const int group = blockIdx.x % P.ngroups;
  const int half = (blockIdx.x / P.ngroups) % P.nhalves;
  const int split = blockIdx.x / (P.ngroups * P.nhalves);
  const int tap0 = group * P.tpg;
  const int T = min(P.tpg, P.ntaps - tap0);
  const int64_t tile_beg = (int64_t)split * P.tiles_per_split;
  const int64_t tile_end = min(P.tiles_total, tile_beg + P.tiles_per_split);
  const int64_t ntiles = tile_end - tile_beg;

  const uint32_t a_bytes = 2 * WG_BLOCK_BYTES;                        // 128 n-channels
  const uint32_t b_bytes = (uint32_t)P.cblocks * WG_BLOCK_BYTES;      // per tap
  const uint32_t stage_bytes = a_bytes + (uint32_t)P.tpg * b_bytes;

  if (warp == 0 && lane == 0) {
    for (int i = 0; i < P.stages; ++i) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], 1); }
    mbar_init(&tmem_full_bar, 1);
    fence_mbar_init();
    fence_proxy_async();
  }
  if (warp == 1) {
    tmem_alloc(&tmem_base_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_slot;

  if (warp == 0) {
    if (lane == 0 && ntiles > 0) {
      prefetch_tmap(&P.pmap);
      int stage = 0; uint32_t phase = 0;
      for (int64_t tl = tile_beg; tl < tile_end; ++tl) {
        int64_t r = tl;
        const int tw_i = (int)(r % P.ntw); r /= P.ntw;
        const int th_i = (int)(r % P.nth); r /= P.nth;
        const int td_i = (int)(r % P.ntd); r /= P.ntd;
        const int b = (int)r;
        const int g0d = td_i * P.td, g0h = th_i * P.th, g0w = tw_i * P.tw;
        mbar_wait(&empty_bar[stage], phase ^ 1);
        mbar_expect_tx(&full_bar[stage], a_bytes + (uint32_t)T * b_bytes);
        uint8_t* sa = smem + (size_t)stage * stage_bytes;
        for (int h = 0; h < 2; ++h)
          tma_load_5d(sa + h * WG_BLOCK_BYTES, &P.pmap, &full_bar[stage], half * 128 + h * 64, g0w, g0h, g0d, b);
        for (int t = 0; t < T; ++t) {
          const WgTap tp = P.taps[tap0 + t];
          uint8_t* sb = sa + a_bytes + (size_t)t * b_bytes;
          for (int cb = 0; cb < P.cblocks; ++cb)
            tma_load_5d(sb + cb * WG_BLOCK_BYTES, &P.qmap[tp.map], &full_bar[stage], cb * 64, g0w + tp.dw, g0h + tp.dh,
                        g0d + tp.dd, b);
        }
        if (++stage == P.stages) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    if (lane == 0 && ntiles > 0) {
      const uint32_t idesc = make_idesc_bf16(128, P.Cc, 1, 1);   // both operands MN-major
      // MN-major SWIZZLE_128B: 64-channel blocks `lbo` apart, 8-position groups `sbo` apart
      const uint32_t lbo = WG_BLOCK_BYTES, sbo = 1024;
      int stage = 0; uint32_t phase = 0;
      for (int64_t tl = 0; tl < ntiles; ++tl) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        const uint32_t sa = smem_u32(smem + (size_t)stage * stage_bytes);
        for (int t = 0; t < T; ++t) {
          const uint32_t sb = sa + a_bytes + (uint32_t)t * b_bytes;
#pragma unroll
          for (int j = 0; j < WG_KP / 16; ++j) {
            const uint64_t da = make_smem_desc(sa + j * 2048, lbo, sbo, 2);
            const uint64_t db = make_smem_desc(sb + j * 2048, lbo, sbo, 2);
            umma_bf16(tmem_base + (uint32_t)(t * P.Cc), da, db, idesc, (tl | j) != 0);
          }
        }
        umma_commit(&empty_bar[stage]);
        if (++stage == P.stages) { stage = 0; phase ^= 1; }
      }
      umma_commit(&tmem_full_bar);
    }
  } else if (ntiles > 0) {
    const int quad = warp & 3;
    const int n = half * 128 + quad * 32 + lane;
    mbar_wait(&tmem_full_bar, 0);
    tc_fence_after();
    float* const out = P.parts ? P.parts + (int64_t)split * P.out_size : P.dwp;
    for (int t = 0; t < T; ++t) {
      float* dst = out + ((int64_t)(tap0 + t) * P.Cn + n) * P.Cc;
      for (int c0 = 0; c0 < P.Cc; c0 += 32) {
        uint32_t v[32];
        tmem_ld_32x32(tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(t * P.Cc + c0), v);
        tmem_ld_wait();
        if (P.parts) {
#pragma unroll
          for (int j = 0; j < 32; j += 4)
            *reinterpret_cast<float4*>(dst + c0 + j) = make_float4(__uint_as_float(v[j]), __uint_as_float(v[j + 1]),
                                                                   __uint_as_float(v[j + 2]), __uint_as_float(v[j + 3]));
        } else {                                   // rows of dWp are Cc floats: 16-byte aligned vector reductions
#pragma unroll
          for (int j = 0; j < 32; j += 4)
            sa_red_add_v4(dst + c0 + j, __uint_as_float(v[j]), __uint_as_float(v[j + 1]), __uint_as_float(v[j + 2]),
                          __uint_as_float(v[j + 3]));
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 512);
}

int pow2_ceil(int v) { int p = 1; while (p < v) p <<= 1; return p; }

bool choose_tile32(int gD, int gH, int gW, int* td, int* th, int* tw) {
  int64_t best = -1; int bd = 0, bh = 0, bw = 0;
  for (int w = 1; w <= WG_KP; w <<= 1) {
    if (w > pow2_ceil(gW)) break;
    for (int h = 1; h * w <= WG_KP; h <<= 1) {
      if (h > pow2_ceil(gH)) break;
      const int d = WG_KP / (w * h);
      if (d > pow2_ceil(gD)) continue;
      const int64_t vol = sa_cdiv(gD, d) * d * sa_cdiv(gH, h) * h * sa_cdiv(gW, w) * w;
      if (best < 0 || vol < best || (vol == best && w > bw)) { best = vol; bd = d; bh = h; bw = w; }
    }
  }
  if (best < 0) return false;
  *td = bd; *th = bh; *tw = bw;
  return true;
}

std::once_flag g_once;
int g_max_smem = 0;

int make_map(CUtensorMap* m, const void* base, int C, int D, int H, int W, int B, int sub, int pd, int ph, int pw, int td,
             int th, int tw) {
  const uint64_t dims[5] = {(uint64_t)C, (uint64_t)(W / sub), (uint64_t)(H / sub), (uint64_t)(D / sub), (uint64_t)B};
  const uint64_t es = 2;
  const uint64_t strides[5] = {es, (uint64_t)C * es * sub, (uint64_t)W * C * es * sub, (uint64_t)H * W * C * es * sub,
                               (uint64_t)D * H * W * C * es};
  const uint32_t box[5] = {64, (uint32_t)tw, (uint32_t)th, (uint32_t)td, 1};
  const uint8_t* p = (const uint8_t*)base + ((uint64_t)pd * H * W + (uint64_t)ph * W + pw) * C * es;
  return sa_make_tmap_bf16(m, p, 5, dims, strides, box);
}

}  // namespace

bool sa_tc_wgrad_supported(const sa_conv_desc* d) {
  if (d->act_dtype != SA_BF16 || d->transposed) return false;
  if (d->c_out % 128 != 0 || d->c_out > 256) return false;
  if (!(d->c_in == 64 || d->c_in == 128 || d->c_in == 256)) return false;
  const int k = d->ksize, s = d->stride, p = d->pad;
  if (k * k * k > WG_MAX_TAPS) return false;
  if (s == 1) {
    for (int i = 0; i < 3; ++i) if (d->out_dhw[i] != d->in_dhw[i] + 2 * p - k + 1) return false;
  } else if (s == 2 && k == 4 && p == 1) {
    for (int i = 0; i < 3; ++i) if (d->in_dhw[i] % 2 || d->out_dhw[i] * 2 != d->in_dhw[i]) return false;
  } else {
    return false;
  }
  int td, th, tw;
  return choose_tile32(d->out_dhw[0], d->out_dhw[1], d->out_dhw[2], &td, &th, &tw);
}

int sa_tc_conv3d_wgrad(const sa_conv_desc* d, const void* p, const void* q, float* dwp, cudaStream_t st) {
  if (!sa_tc_wgrad_supported(d)) { sa_set_error("tc_wgrad: unsupported configuration"); return SA_ERR_UNSUPPORTED; }
  if (!sa_get_tmap_encode()) { sa_set_error("tc_wgrad: cuTensorMapEncodeTiled unavailable"); return SA_ERR_CUDA; }
  sa_note_path(SA_PATH_TCGEN05);
  std::call_once(g_once, [] {
    int dev = 0; cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&g_max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    cudaFuncSetAttribute(tc_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, g_max_smem - 1024);
  });
  const int k = d->ksize, s = d->stride, pad = d->pad;
  const int iD = d->in_dhw[0], iH = d->in_dhw[1], iW = d->in_dhw[2];
  const int oD = d->out_dhw[0], oH = d->out_dhw[1], oW = d->out_dhw[2];

  static thread_local WgParams P;
  P.Cn = d->c_out; P.Cc = d->c_in;
  P.nhalves = d->c_out / 128; P.cblocks = d->c_in / 64;
  P.ntaps = k * k * k;
  P.tpg = 512 / d->c_in;
  if (P.tpg > P.ntaps) P.tpg = P.ntaps;
  P.ngroups = (int)sa_cdiv(P.ntaps, P.tpg);
  P.dwp = dwp;
  choose_tile32(oD, oH, oW, &P.td, &P.th, &P.tw);
  P.ntd = (int)sa_cdiv(oD, P.td); P.nth = (int)sa_cdiv(oH, P.th); P.ntw = (int)sa_cdiv(oW, P.tw);
  P.tiles_total = (int64_t)d->batch * P.ntd * P.nth * P.ntw;

  int rc = make_map(&P.pmap, p, d->c_out, oD, oH, oW, d->batch, 1, 0, 0, 0, P.td, P.th, P.tw);
  if (rc != SA_OK) return rc;
  if (s == 1) {
    rc = make_map(&P.qmap[0], q, d->c_in, iD, iH, iW, d->batch, 1, 0, 0, 0, P.td, P.th, P.tw);
    if (rc != SA_OK) return rc;
    for (int t = 0; t < P.ntaps; ++t) {
      const int tw_ = t % k, th_ = (t / k) % k, td_ = t / (k * k);
      P.taps[t] = WgTap{0, (int8_t)(td_ - pad), (int8_t)(th_ - pad), (int8_t)(tw_ - pad)};
    }
  } else {
    for (int m = 0; m < 8; ++m) {
      rc = make_map(&P.qmap[m], q, d->c_in, iD, iH, iW, d->batch, 2, (m >> 2) & 1, (m >> 1) & 1, m & 1, P.td, P.th, P.tw);
      if (rc != SA_OK) return rc;
    }
    auto par = [](int t) { return (t & 1) ? 0 : 1; };
    auto off = [](int t) { return t == 0 ? -1 : (t == 3 ? 1 : 0); };
    for (int t = 0; t < 64; ++t) {
      const int tw_ = t % 4, th_ = (t / 4) % 4, td_ = t / 16;
      P.taps[t] = WgTap{(int8_t)((par(td_) << 2) | (par(th_) << 1) | par(tw_)), (int8_t)off(td_), (int8_t)off(th_),
                        (int8_t)off(tw_)};
    }
  }

  const size_t stage_bytes = (size_t)2 * WG_BLOCK_BYTES + (size_t)P.tpg * P.cblocks * WG_BLOCK_BYTES;
  const size_t budget = (size_t)g_max_smem - 1024 - 1024;
  int stages = (int)(budget / stage_bytes);
  if (stages > WG_MAX_STAGES) stages = WG_MAX_STAGES;
  if (stages < 2) { sa_set_error("tc_wgrad: stage does not fit shared memory"); return SA_ERR_UNSUPPORTED; }
  P.stages = stages;

  // one CTA per SM is resident (the stages fill shared memory): size the grid to at most 2 full waves of 148
  const int64_t base_ctas = (int64_t)P.ngroups * P.nhalves;
  int64_t splits = (148 * 2) / base_ctas;
  const int64_t max_splits = sa_cdiv(P.tiles_total, 8);
  if (splits > max_splits) splits = max_splits;
  if (splits < 1) splits = 1;
  P.tiles_per_split = sa_cdiv(P.tiles_total, splits);
  splits = sa_cdiv(P.tiles_total, P.tiles_per_split);
  const unsigned grid = (unsigned)(base_ctas * splits);
  P.out_size = (int64_t)P.ntaps * P.Cn * P.Cc;
  P.parts = sa_parts_alloc(splits, P.out_size, st);
  tc_wgrad_kernel<<<grid, WG_THREADS, stages * stage_bytes + 1024, st>>>(P);
  SA_LAUNCH_CHECK();
  if (P.parts) {
    if ((rc = sa_parts_reduce(P.parts, splits, P.out_size, P.out_size, dwp, st)) != SA_OK) return rc;
    return sa_parts_free(P.parts, st);
  }
  return SA_OK;
}
