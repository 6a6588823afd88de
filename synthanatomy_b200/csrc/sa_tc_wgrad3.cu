// tcgen05 weight gradient of the stride-1 3x3x3 convolutions, sm_100a (depth-fused variant of sa_tc_wgrad.cu).
//
//   dWp[t][n][c] += sum_{b, o}  P[b, o, n] * Q[b, o + t - pad, c]          t = (dd, dh, dw)
//
// Per tap this is a GEMM with M = n, N = c, K = positions; both operands are MN-major 128B-swizzled TMA boxes
// (row = position, 64 channels = 128 B).  A CTA owns one (dh, dw) pair, one 128-row half of n and a contiguous range of
// 128-position tiles (4 x 4 x 8).  For a fixed (dh, dw) the three depth taps read ONE Q box that is two planes deeper
// (6 x 4 x 8 positions): a depth shift is a whole 32-row plane = 4 KB, so the shifted B operand is the same
// shared-memory box at a 1024-byte-aligned offset.  Per stage: P 32 KB + Q 24 KB per channel block for 24 MMAs
// (53 B / clk / SM instead of 80, in four large TMA boxes instead of ten small ones); the three accumulators
// [128 x c_in] live in TMEM, split-K partials are reduced with fp32 red.global.add into dWp.
//
// Reference call sites replaced: cuDNN wgrad reached through autograd of the 3x3x3 convs at
// /root/reference/src/networks/vqvae/baseline.py:153-156, 242-244, 258.
#include <mutex>
#include <stdlib.h>

#include "sa_tc_common.cuh"

using namespace satc;

namespace {

constexpr int W3_THREADS = 192;
constexpr int W3_TD = 4, W3_TH = 4, W3_TW = 8;
constexpr int W3_POS = W3_TD * W3_TH * W3_TW;                  // 128 positions per tile
constexpr uint32_t W3_PLANE_BYTES = W3_TH * W3_TW * 128;       // 4 KB
constexpr uint32_t W3_P_BLOCK = W3_POS * 128;                  // 16 KB: 64 channels x 128 positions
constexpr uint32_t W3_Q_BLOCK = (W3_TD + 2) * W3_PLANE_BYTES;  // 24 KB: 64 channels x 192 positions
constexpr int W3_MAX_STAGES = 4;

// a CTA owns one group: ONE Q box (ndd + 3 planes deep) serves ndd depth taps
struct W3Group {
  int8_t map, od, oh, ow;   // Q view and box origin relative to the tile origin
  int32_t tap[3];           // tap index (row block of dWp) of each depth tap
};

struct W3Params {
  CUtensorMap pmap;
  CUtensorMap qmap[8];
  W3Group groups[32];
  int ngroups, ndd;
  int Cn, Cc, nhalves, cblocks, stages;
  int ntd, nth, ntw;
  int64_t tiles_total, tiles_per_split;
  float* dwp;
  float* parts;        // deterministic mode: split s stores its partial dWp at parts + s * out_size (summed in order afterwards)
  int64_t out_size;
};

__global__ void __launch_bounds__(W3_THREADS, 1)
tc_wgrad3_kernel(const __grid_constant__ W3Params P) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ uint64_t full_bar[W3_MAX_STAGES], empty_bar[W3_MAX_STAGES];
  __shared__ uint64_t tmem_full_bar;
  __shared__ uint32_t tmem_base_slot;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);

  const int gi = blockIdx.x % P.ngroups;
  const int half = (blockIdx.x / P.ngroups) % P.nhalves;
  const int split = blockIdx.x / (P.ngroups * P.nhalves);
  const W3Group grp = P.groups[gi];
  const int ndd = P.ndd;
  const uint32_t q_block = (uint32_t)(W3_TD + ndd - 1) * W3_PLANE_BYTES;      // 64 channels x (4 + ndd - 1) planes
  const int64_t tile_beg = (int64_t)split * P.tiles_per_split;
  const int64_t tile_end = min(P.tiles_total, tile_beg + P.tiles_per_split);
  const int64_t ntiles = tile_end - tile_beg;
  const uint32_t a_bytes = 2 * W3_P_BLOCK;
  const uint32_t stage_bytes = a_bytes + (uint32_t)P.cblocks * q_block;

  if (warp == 0 && lane == 0) {
    for (int i = 0; i < P.stages; ++i) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], 1); }
    mbar_init(&tmem_full_bar, 1);
    fence_mbar_init();
    fence_proxy_async();
  }
  if (warp == 1) { tmem_alloc(&tmem_base_slot, 512); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_slot;

  if (warp == 0) {
    if (lane == 0 && ntiles > 0) {
      prefetch_tmap(&P.pmap);
      prefetch_tmap(&P.qmap[grp.map]);
      int stage = 0; uint32_t phase = 0;
      for (int64_t tl = tile_beg; tl < tile_end; ++tl) {
        int64_t r = tl;
        const int tw_i = (int)(r % P.ntw); r /= P.ntw;
        const int th_i = (int)(r % P.nth); r /= P.nth;
        const int td_i = (int)(r % P.ntd); r /= P.ntd;
        const int b = (int)r;
        const int g0d = td_i * W3_TD, g0h = th_i * W3_TH, g0w = tw_i * W3_TW;
        mbar_wait(&empty_bar[stage], phase ^ 1);
        mbar_expect_tx(&full_bar[stage], stage_bytes);
        uint8_t* sa = smem + (size_t)stage * stage_bytes;
        for (int h = 0; h < 2; ++h)
          tma_load_5d(sa + h * W3_P_BLOCK, &P.pmap, &full_bar[stage], half * 128 + h * 64, g0w, g0h, g0d, b);
        for (int cb = 0; cb < P.cblocks; ++cb)
          tma_load_5d(sa + a_bytes + cb * q_block, &P.qmap[grp.map], &full_bar[stage], cb * 64, g0w + grp.ow, g0h + grp.oh,
                      g0d + grp.od, b);
        if (++stage == P.stages) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    if (lane == 0 && ntiles > 0) {
      const uint32_t idesc = make_idesc_bf16(128, P.Cc, 1, 1);   // both operands MN-major
      int stage = 0; uint32_t phase = 0;
      for (int64_t tl = 0; tl < ntiles; ++tl) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        const uint32_t sa = smem_u32(smem + (size_t)stage * stage_bytes);
        const uint32_t sb = sa + a_bytes;
        for (int dd = 0; dd < ndd; ++dd) {
#pragma unroll
          for (int j = 0; j < W3_POS / 16; ++j) {
            // MN-major SWIZZLE_128B: 64-channel blocks LBO apart, 8-position groups 1024 B apart, a K step = 16 positions
            const uint64_t da = make_smem_desc(sa + j * 2048, W3_P_BLOCK, 1024, 2);
            const uint64_t db = make_smem_desc(sb + dd * W3_PLANE_BYTES + j * 2048, q_block, 1024, 2);
            umma_bf16(tmem_base + (uint32_t)(dd * P.Cc), da, db, idesc, (tl | j) != 0);
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
    for (int dd = 0; dd < ndd; ++dd) {
      const int tap = grp.tap[dd];
      float* dst = out + ((int64_t)tap * P.Cn + n) * P.Cc;
      for (int c0 = 0; c0 < P.Cc; c0 += 32) {
        uint32_t v[32];
        tmem_ld_32x32(tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(dd * P.Cc + c0), v);
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

std::once_flag g_w3_once;
int g_w3_sms = 148;

int w3_map(CUtensorMap* m, const void* base, int C, int D, int H, int W, int B, int sub, int pd, int ph, int pw, int planes) {
  const uint64_t dims[5] = {(uint64_t)C, (uint64_t)(W / sub), (uint64_t)(H / sub), (uint64_t)(D / sub), (uint64_t)B};
  const uint64_t es = 2;
  const uint64_t strides[5] = {es, (uint64_t)C * es * sub, (uint64_t)W * C * es * sub, (uint64_t)H * W * C * es * sub,
                               (uint64_t)D * H * W * C * es};
  const uint32_t box[5] = {64, W3_TW, W3_TH, (uint32_t)planes, 1};
  const uint8_t* p = (const uint8_t*)base + ((uint64_t)pd * H * W + (uint64_t)ph * W + pw) * C * es;
  return sa_make_tmap_bf16(m, p, 5, dims, strides, box);
}

}  // namespace

// stride-1 3x3x3 convs and 4/2/1 strided convs (FORM_CONV indexing)
bool sa_tc_wgrad3_supported(const sa_conv_desc* d) {
  if (d->act_dtype != SA_BF16 || d->transposed) return false;
  if (d->c_out % 128 != 0 || !(d->c_in == 64 || d->c_in == 128)) return false;
  if (const char* e = getenv("SA_TC_WGRAD3")) { if (e[0] == '0') return false; }   // A/B switch for benchmarking
  if (!sa_get_tmap_encode()) return false;
  if (d->ksize == 3 && d->stride == 1) {
    if (d->pad < 0 || d->pad > 2) return false;
    for (int i = 0; i < 3; ++i) if (d->out_dhw[i] != d->in_dhw[i] + 2 * d->pad - 2) return false;
    return true;
  }
  if (d->ksize == 4 && d->stride == 2 && d->pad == 1) {
    for (int i = 0; i < 3; ++i) if (d->in_dhw[i] % 2 || d->out_dhw[i] * 2 != d->in_dhw[i]) return false;
    return true;
  }
  return false;
}

int sa_tc_conv3d_wgrad3(const sa_conv_desc* d, const void* p, const void* q, float* dwp, cudaStream_t st) {
  std::call_once(g_w3_once, [] {
    int dev = 0, v = 0;
    cudaGetDevice(&dev);
    if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && v > 0) g_w3_sms = v;
    cudaFuncSetAttribute(tc_wgrad3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024 - 2048);
  });
  sa_note_path(SA_PATH_TCGEN05);
  const int iD = d->in_dhw[0], iH = d->in_dhw[1], iW = d->in_dhw[2];
  const int oD = d->out_dhw[0], oH = d->out_dhw[1], oW = d->out_dhw[2];
  static thread_local W3Params P;
  P.Cn = d->c_out; P.Cc = d->c_in;
  P.nhalves = d->c_out / 128; P.cblocks = d->c_in / 64;
  P.dwp = dwp;
  P.ntd = (int)sa_cdiv(oD, W3_TD); P.nth = (int)sa_cdiv(oH, W3_TH); P.ntw = (int)sa_cdiv(oW, W3_TW);
  P.tiles_total = (int64_t)d->batch * P.ntd * P.nth * P.ntw;
  int rc = w3_map(&P.pmap, p, d->c_out, oD, oH, oW, d->batch, 1, 0, 0, 0, W3_TD);
  if (rc != SA_OK) return rc;
  if (d->stride == 1) {
    P.ndd = 3; P.ngroups = 9;
    if ((rc = w3_map(&P.qmap[0], q, d->c_in, iD, iH, iW, d->batch, 1, 0, 0, 0, W3_TD + 2)) != SA_OK) return rc;
    for (int dh = 0; dh < 3; ++dh)
      for (int dw = 0; dw < 3; ++dw) {
        W3Group& g = P.groups[dh * 3 + dw];
        g.map = 0; g.od = (int8_t)(-d->pad); g.oh = (int8_t)(dh - d->pad); g.ow = (int8_t)(dw - d->pad);
        for (int dd = 0; dd < 3; ++dd) g.tap[dd] = (dd * 3 + dh) * 3 + dw;
      }
  } else {
    // i = 2 o - 1 + t: tap t <-> (parity view, offset): 0: (odd, -1)  1: (even, 0)  2: (odd, 0)  3: (even, +1);
    // the two depth taps of one parity share a Q box that is one plane deeper
    auto par = [](int t) { return (t & 1) ? 0 : 1; };
    auto off = [](int t) { return t == 0 ? -1 : (t == 3 ? 1 : 0); };
    P.ndd = 2; P.ngroups = 32;
    for (int m = 0; m < 8; ++m)
      if ((rc = w3_map(&P.qmap[m], q, d->c_in, iD, iH, iW, d->batch, 2, (m >> 2) & 1, (m >> 1) & 1, m & 1, W3_TD + 1)) != SA_OK)
        return rc;
    int gi = 0;
    for (int pd = 0; pd < 2; ++pd)
      for (int th = 0; th < 4; ++th)
        for (int tw = 0; tw < 4; ++tw) {
          W3Group& g = P.groups[gi++];
          const int td0 = pd ? 0 : 1, td1 = pd ? 2 : 3;
          g.map = (int8_t)((pd << 2) | (par(th) << 1) | par(tw));
          g.od = (int8_t)off(td0); g.oh = (int8_t)off(th); g.ow = (int8_t)off(tw);
          g.tap[0] = (td0 * 4 + th) * 4 + tw; g.tap[1] = (td1 * 4 + th) * 4 + tw; g.tap[2] = 0;
        }
  }
  const size_t stage_bytes = (size_t)2 * W3_P_BLOCK + (size_t)P.cblocks * (W3_TD + P.ndd - 1) * W3_PLANE_BYTES;
  int stages = (int)((227 * 1024 - 2048 - 1024) / stage_bytes);
  if (stages > W3_MAX_STAGES) stages = W3_MAX_STAGES;
  if (stages < 2) { sa_set_error("tc_wgrad3: stage does not fit shared memory"); return SA_ERR_UNSUPPORTED; }
  P.stages = stages;
  // one CTA per SM is resident: one wave, as many splits over the position tiles as SMs allow
  const int64_t base_ctas = (int64_t)P.ngroups * P.nhalves;
  int64_t splits = g_w3_sms / base_ctas;
  if (splits > P.tiles_total) splits = P.tiles_total;
  if (splits < 1) splits = 1;
  P.tiles_per_split = sa_cdiv(P.tiles_total, splits);
  splits = sa_cdiv(P.tiles_total, P.tiles_per_split);
  const unsigned grid = (unsigned)(base_ctas * splits);
  P.out_size = (int64_t)(d->ksize * d->ksize * d->ksize) * d->c_out * d->c_in;
  P.parts = sa_parts_alloc(splits, P.out_size, st);
  tc_wgrad3_kernel<<<grid, W3_THREADS, stages * stage_bytes + 1024, st>>>(P);
  SA_LAUNCH_CHECK();
  if (P.parts) {
    if ((rc = sa_parts_reduce(P.parts, splits, P.out_size, P.out_size, dwp, st)) != SA_OK) return rc;
    return sa_parts_free(P.parts, st);
  }
  return SA_OK;
}
