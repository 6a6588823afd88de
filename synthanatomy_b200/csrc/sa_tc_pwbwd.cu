// Fused backward of the pointwise (1x1x1) convolution of a ResidualLayer, tcgen05 / sm_100a.
//
//   y = relu(x + W1 h + b1),  h = relu(conv3(x) + b3)        (/root/reference/src/networks/vqvae/baseline.py:153-160)
//
// Given g = dL/d(pre-activation of y) [M positions][n] and the saved h [M][c] (both bf16 NDHWC, flattened), ONE pass
// over g and h produces
//   dh[pos][c]  = (sum_n g[pos][n] W1[n][c]) * (h[pos][c] > 0)        data gradient with the ReLU mask of h
//   dW1[n][c]  += sum_pos g[pos][n] h[pos][c]                         weight gradient
//   db1[n]     += sum_pos g[pos][n]                                   bias gradient
// instead of three kernels that each stream g (and h) from HBM (1x1 dgrad, 1x1 wgrad, bias reduction: 2.5 ms at level 1).
// The same 128B-swizzled g tile in shared memory is read as a K-major operand (g W1: contraction over channels) and as an
// MN-major operand (g^T h and g^T 1: contraction over positions).  Persistent CTAs (one per SM): two-stage TMA ring of
// (g, h) tiles of 128 positions, two TMEM buffers for dh so that its epilogue (mask from the h tile already in shared
// memory, bf16, TMA store) overlaps the next tile's MMAs, dW1 / db1 accumulate in TMEM over all tiles of the CTA and are
// reduced with fp32 red.global at the end.  HBM-bound: 3 x 256 bytes per position.
//
// Warp roles (320 threads): warp 0 = TMA producer, warp 1 = TMEM owner + MMA issuer, warps 2-9 = epilogue.
#include <mutex>

#include "sa_tc_common.cuh"

using namespace satc;

namespace {

constexpr int PW_THREADS = 320;
constexpr int PW_M = 128;                     // positions per tile
constexpr uint32_t PW_BLK = PW_M * 128;       // [128 x 64] bf16 block
constexpr int PW_STAGES = 2;
constexpr int PW_PART = 128 * 128 + 128 + 128;   // floats per CTA partial in deterministic mode

struct PwParams {
  CUtensorMap gmap, hmap, wmap, omap;
  long long m;
  int tiles;
  float* dwp;
  float* dbias;
  float* dbias_h;        // MODE 0, optional: column sums of dh (the bias gradient of the conv that produced h)
  float* parts;          // deterministic mode (MODE 0): CTA b stores [dW 128 x 128 | db 128 | dbh 128] at parts + b * PW_PART
  const float* bias;     // forward mode
  int relu;
};

__device__ __forceinline__ void pw_tma_store_2d(const CUtensorMap* m, const void* smem_src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(
                   reinterpret_cast<uint64_t>(m)),
               "r"(smem_u32(smem_src)), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void pw_bar_epi() { asm volatile("bar.sync 1, 256;" ::: "memory"); }

// MODE 0: the fused backward above.   MODE 1: the forward  y = relu(h W1^T + b1 + x)  with the same streaming structure
// ("g" tile = h, "h" tile = the residual addend x; no weight / bias gradient products).
template <int MODE>
__global__ void __launch_bounds__(PW_THREADS, 1)
tc_pw_kernel(const __grid_constant__ PwParams P) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ uint64_t w_full, full_bar[PW_STAGES], empty_bar[PW_STAGES], dh_full[2], dh_empty[2], final_full;
  __shared__ uint32_t tmem_base_slot;
  __shared__ float s_colsum[4][128];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* Ws = smem;                          // W1^T: 2 blocks [128 c rows][64 n cols]                 32 KB
  uint8_t* Ones = Ws + 2 * PW_BLK;             // [128 pos][64 cols], columns 0..15 = 1                   16 KB
  uint8_t* Out = Ones + PW_BLK;                // dh staging: 2 blocks [128 pos][64 c cols]               32 KB
  uint8_t* Ring = Out + 2 * PW_BLK;            // stages x (g: 2 blocks | h: 2 blocks)                    2 x 64 KB
  constexpr uint32_t STAGE = 4 * PW_BLK;

  if (threadIdx.x == 0) {
    mbar_init(&w_full, 1); mbar_init(&final_full, 1);
    for (int i = 0; i < PW_STAGES; ++i) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], 1 + 8); }
    for (int i = 0; i < 2; ++i) { mbar_init(&dh_full[i], 1); mbar_init(&dh_empty[i], 8); }
    fence_mbar_init();
    fence_proxy_async();
  }
  if (warp >= 2) {   // the "ones" operand: element (pos, col) = 1 for col < 16 (chunks 0 and 1 of every row, swizzled)
    const int t = threadIdx.x - 64;
    for (int idx = t; idx < PW_M * 8; idx += 256) {
      const int r = idx >> 3, ch = idx & 7;
      const uint32_t one2 = 0x3F803F80u;      // two bf16 ones
      const uint4 u = ch < 2 ? make_uint4(one2, one2, one2, one2) : make_uint4(0, 0, 0, 0);
      *reinterpret_cast<uint4*>(Ones + (r >> 3) * 1024 + (r & 7) * 128 + ((ch ^ (r & 7)) << 4)) = u;
    }
    fence_proxy_async();
  }
  if (warp == 1) { tmem_alloc(&tmem_base_slot, 512); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_slot;
  const uint32_t tDh = tmem_base, tDw = tmem_base + 256, tDb = tmem_base + 384;
  int my_tiles = 0;
  for (int tile = blockIdx.x; tile < P.tiles; tile += gridDim.x) ++my_tiles;

  if (warp == 0) {
    if (lane == 0) {
      prefetch_tmap(&P.gmap); prefetch_tmap(&P.hmap); prefetch_tmap(&P.wmap);
      mbar_expect_tx(&w_full, 2 * PW_BLK);
      tma_load_2d(Ws, &P.wmap, &w_full, 0, 0);
      tma_load_2d(Ws + PW_BLK, &P.wmap, &w_full, 64, 0);
      int stage = 0; uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < P.tiles; tile += gridDim.x) {
        mbar_wait(&empty_bar[stage], phase ^ 1);
        mbar_expect_tx(&full_bar[stage], STAGE);
        uint8_t* sg = Ring + stage * STAGE;
        tma_load_2d(sg, &P.gmap, &full_bar[stage], 0, tile * PW_M);
        tma_load_2d(sg + PW_BLK, &P.gmap, &full_bar[stage], 64, tile * PW_M);
        tma_load_2d(sg + 2 * PW_BLK, &P.hmap, &full_bar[stage], 0, tile * PW_M);
        tma_load_2d(sg + 3 * PW_BLK, &P.hmap, &full_bar[stage], 64, tile * PW_M);
        if (++stage == PW_STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    if (lane == 0 && my_tiles > 0) {
      const uint32_t wa = smem_u32(Ws), oa = smem_u32(Ones);
      const uint32_t idesc_kk = make_idesc_bf16(128, 128, 0, 0);
      const uint32_t idesc_mm = make_idesc_bf16(128, 128, 1, 1);
      const uint32_t idesc_m1 = make_idesc_bf16(128, 16, 1, 1);
      mbar_wait(&w_full, 0);
      int stage = 0; uint32_t phase = 0;
      for (int it = 0; it < my_tiles; ++it) {
        const int bsel = it & 1;
        mbar_wait(&full_bar[stage], phase);
        mbar_wait(&dh_empty[bsel], (uint32_t)(((it >> 1) & 1) ^ 1));
        tc_fence_after();
        const uint32_t ga = smem_u32(Ring + stage * STAGE), ha = ga + 2 * PW_BLK;
        // dh = g W1: g K-major (contraction over the n channels: 2 blocks x 4 k-steps), W1^T K-major [c rows][n cols]
#pragma unroll
        for (int kb = 0; kb < 2; ++kb)
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_bf16(tDh + (uint32_t)(bsel * 128), make_smem_desc(ga + kb * PW_BLK + k * 32, 16, 1024, 2),
                      make_smem_desc(wa + kb * PW_BLK + k * 32, 16, 1024, 2), idesc_kk, (kb | k) != 0);
        umma_commit(&dh_full[bsel]);
        if (MODE == 0) {
          // dW1 += g^T h, db1 += g^T 1: both operands MN-major (contraction over the 128 positions: 8 k-steps)
#pragma unroll
          for (int j = 0; j < PW_M / 16; ++j) {
            const uint64_t da = make_smem_desc(ga + j * 2048, PW_BLK, 1024, 2);
            umma_bf16(tDw, da, make_smem_desc(ha + j * 2048, PW_BLK, 1024, 2), idesc_mm, (it | j) != 0);
            umma_bf16(tDb, da, make_smem_desc(oa + j * 2048, PW_BLK, 1024, 2), idesc_m1, (it | j) != 0);
          }
        }
        umma_commit(&empty_bar[stage]);
        if (++stage == PW_STAGES) { stage = 0; phase ^= 1; }
      }
      umma_commit(&final_full);
    }
  } else {
    const int ew = warp - 2;
    const int quad = warp & 3, half = ew >> 2;
    const int r = quad * 32 + lane;
    const uint32_t tlane = (uint32_t)(quad * 32) << 16;
    int stage = 0;
    int it = 0;
    // column sums of this thread's rows of dh over all tiles of the CTA (its 64 columns), reduced over the warp at the end
    float colsum[64];
#pragma unroll
    for (int j = 0; j < 64; ++j) colsum[j] = 0.f;
    const bool want_dbh = MODE == 0 && P.dbias_h != nullptr;
    for (int tile = blockIdx.x; tile < P.tiles; tile += gridDim.x, ++it) {
      const int bsel = it & 1;
      mbar_wait(&full_bar[stage], (uint32_t)((it >> 1) & 1));       // the h tile (mask) is in shared memory
      mbar_wait(&dh_full[bsel], (uint32_t)((it >> 1) & 1));
      tc_fence_after();
      if (it > 0) {                                                  // the previous tile's store has read the staging
        if (threadIdx.x == 64) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        pw_bar_epi();
      }
      const uint8_t* hrow = Ring + stage * STAGE + (2 + half) * PW_BLK + (r >> 3) * 1024 + (r & 7) * 128;
      uint8_t* orow = Out + half * PW_BLK + (r >> 3) * 1024 + (r & 7) * 128;
#pragma unroll
      for (int hh = 0; hh < 2; ++hh) {
        uint32_t v[32];
        tmem_ld_32x32(tDh + tlane + (uint32_t)(bsel * 128 + half * 64 + hh * 32), v);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int ch = hh * 4 + i;
          const uint4 hm = *reinterpret_cast<const uint4*>(hrow + ((ch ^ (r & 7)) << 4));
          const uint32_t w[4] = {hm.x, hm.y, hm.z, hm.w};
          float f[8];
          if (MODE == 0) {
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              f[2 * e] = bf16lo(w[e]) > 0.f ? __uint_as_float(v[i * 8 + 2 * e]) : 0.f;
              f[2 * e + 1] = bf16hi(w[e]) > 0.f ? __uint_as_float(v[i * 8 + 2 * e + 1]) : 0.f;
            }
          } else {
            const int c0 = half * 64 + ch * 8;
            const float4 b0 = P.bias ? __ldg(reinterpret_cast<const float4*>(P.bias + c0)) : make_float4(0.f, 0.f, 0.f, 0.f);
            const float4 b1 = P.bias ? __ldg(reinterpret_cast<const float4*>(P.bias + c0 + 4)) : make_float4(0.f, 0.f, 0.f, 0.f);
            const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              f[2 * e] = __uint_as_float(v[i * 8 + 2 * e]) + bb[2 * e] + bf16lo(w[e]);
              f[2 * e + 1] = __uint_as_float(v[i * 8 + 2 * e + 1]) + bb[2 * e + 1] + bf16hi(w[e]);
            }
            if (P.relu) {
#pragma unroll
              for (int e = 0; e < 8; ++e) f[e] = fmaxf(f[e], 0.f);
            }
          }
          uint4 u;
          u.x = pack_bf16x2(f[0], f[1]); u.y = pack_bf16x2(f[2], f[3]);
          u.z = pack_bf16x2(f[4], f[5]); u.w = pack_bf16x2(f[6], f[7]);
          *reinterpret_cast<uint4*>(orow + ((ch ^ (r & 7)) << 4)) = u;
          if (want_dbh) {          // the bf16-rounded values, i.e. exactly what a reduction over the stored dh would add
            const uint32_t ww[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              colsum[ch * 8 + 2 * e] += bf16lo(ww[e]);
              colsum[ch * 8 + 2 * e + 1] += bf16hi(ww[e]);
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) { mbar_arrive(&dh_empty[bsel]); mbar_arrive(&empty_bar[stage]); }
      fence_proxy_async();
      pw_bar_epi();
      if (threadIdx.x == 64) {
        pw_tma_store_2d(&P.omap, Out, 0, tile * PW_M);
        pw_tma_store_2d(&P.omap, Out + PW_BLK, 64, tile * PW_M);
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      }
      if (++stage == PW_STAGES) stage = 0;
    }
    if (threadIdx.x == 64) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    float* const part = P.parts ? P.parts + (long long)blockIdx.x * PW_PART : nullptr;   // summed in CTA order afterwards
    if (want_dbh) {
      // the four lane quadrants of a column half meet in shared memory and leave as ONE addition per column (a fixed
      // order inside the CTA; across CTAs the per-CTA partials are summed in CTA order)
#pragma unroll
      for (int j = 0; j < 64; ++j) {
        float v = colsum[j];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (lane == 0) s_colsum[quad][half * 64 + j] = v;
      }
      pw_bar_epi();
      const int c = threadIdx.x - 64;
      if (c < 128 && my_tiles > 0) {
        const float v = (s_colsum[0][c] + s_colsum[1][c]) + (s_colsum[2][c] + s_colsum[3][c]);
        if (part) part[128 * 128 + 128 + c] = v;
        else atomicAdd(P.dbias_h + c, v);
      }
    }
    if (MODE == 0 && my_tiles > 0) {
      // dW1 / db1 partials of this CTA: lane = n
      mbar_wait(&final_full, 0);
      tc_fence_after();
      const int n = quad * 32 + lane;
      float* dst = (part ? part : P.dwp) + (long long)n * 128 + half * 64;
#pragma unroll
      for (int hh = 0; hh < 2; ++hh) {
        uint32_t v[32];
        tmem_ld_32x32(tDw + tlane + (uint32_t)(half * 64 + hh * 32), v);
        tmem_ld_wait();
        if (part) {
#pragma unroll
          for (int j = 0; j < 32; ++j) dst[hh * 32 + j] = __uint_as_float(v[j]);
        } else {
#pragma unroll
          for (int j = 0; j < 32; j += 4)
            sa_red_add_v4(dst + hh * 32 + j, __uint_as_float(v[j]), __uint_as_float(v[j + 1]), __uint_as_float(v[j + 2]),
                          __uint_as_float(v[j + 3]));
        }
      }
      if (half == 0) {
        uint32_t v[32];
        tmem_ld_32x32(tDb + tlane, v);       // columns 0..15 hold the sum, the rest of the 32 are never written
        tmem_ld_wait();
        if (part) part[128 * 128 + n] = __uint_as_float(v[0]);
        else atomicAdd(P.dbias + n, __uint_as_float(v[0]));
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 512);
}

std::once_flag g_pw_once;
int g_pw_sms = 148;

int pw_map(CUtensorMap* m, const void* base, uint64_t cols, uint64_t rows, uint32_t box_rows) {
  const uint64_t dims[2] = {cols, rows};
  const uint64_t strides[2] = {2, cols * 2};
  const uint32_t box[2] = {64, box_rows};
  return sa_make_tmap_bf16(m, base, 2, dims, strides, box);
}

}  // namespace

// g [m][128], h [m][128], wp_t [128 (c)][128 (n)] bf16 (pack_weight(w1, transpose = 1)); dh [m][128] bf16;
// dwp [128 (n)][128 (c)] fp32 and dbias [128] fp32 are ACCUMULATED into (zero them first).
extern "C" int sa_conv1x1_bwd_fused_dbh(int64_t m, int c_out, int c_in, const void* g, const void* h, const void* wp_t,
                                        void* dh, float* dwp, float* dbias, float* dbias_h, void* stream);

extern "C" int sa_conv1x1_bwd_fused(int64_t m, int c_out, int c_in, const void* g, const void* h, const void* wp_t, void* dh,
                                    float* dwp, float* dbias, void* stream) {
  return sa_conv1x1_bwd_fused_dbh(m, c_out, c_in, g, h, wp_t, dh, dwp, dbias, nullptr, stream);
}

// the same pass with one more product: dbias_h[c] += sum_pos dh[pos][c] (may be NULL), the bias gradient of the 3x3x3 conv
// that produced h -- saves a separate streaming reduction over dh
extern "C" int sa_conv1x1_bwd_fused_dbh(int64_t m, int c_out, int c_in, const void* g, const void* h, const void* wp_t,
                                        void* dh, float* dwp, float* dbias, float* dbias_h, void* stream) {
  SA_CHECK_ARG(g && h && wp_t && dh && dwp && dbias, "null pointer");
  SA_CHECK_ARG(m > 0, "bad sizes");
  SA_CHECK_ARG((reinterpret_cast<uintptr_t>(dwp) & 15) == 0, "dwp must be 16-byte aligned");
  SA_UNSUPPORTED(c_out != 128 || c_in != 128, "the fused pointwise backward is built for 128 -> 128 channels");
  SA_UNSUPPORTED(m >= (1LL << 31) - 256, "too many positions");
  if (!sa_get_tmap_encode()) { sa_set_error("cuTensorMapEncodeTiled unavailable"); return SA_ERR_CUDA; }
  std::call_once(g_pw_once, [] {
    int dev = 0, v = 0;
    cudaGetDevice(&dev);
    if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && v > 0) g_pw_sms = v;
    cudaFuncSetAttribute(tc_pw_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024 - 4096);
    cudaFuncSetAttribute(tc_pw_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024 - 4096);
  });
  sa_note_path(SA_PATH_TCGEN05);
  static thread_local PwParams P;
  P.m = m;
  P.tiles = (int)sa_cdiv(m, PW_M);
  P.dwp = dwp; P.dbias = dbias; P.dbias_h = dbias_h;
  int rc;
  if ((rc = pw_map(&P.gmap, g, 128, (uint64_t)m, PW_M)) != SA_OK) return rc;
  if ((rc = pw_map(&P.hmap, h, 128, (uint64_t)m, PW_M)) != SA_OK) return rc;
  if ((rc = pw_map(&P.omap, dh, 128, (uint64_t)m, PW_M)) != SA_OK) return rc;
  if ((rc = pw_map(&P.wmap, wp_t, 128, 128, 128)) != SA_OK) return rc;
  const size_t smem = (size_t)(2 + 1 + 2 + PW_STAGES * 4) * PW_BLK + 1024;
  const unsigned grid = (unsigned)(P.tiles < g_pw_sms ? P.tiles : g_pw_sms);
  P.bias = nullptr; P.relu = 0;
  cudaStream_t st = sa_stream(stream);
  P.parts = sa_parts_alloc(grid, PW_PART, st);
  tc_pw_kernel<0><<<grid, PW_THREADS, smem, st>>>(P);
  SA_LAUNCH_CHECK();
  if (P.parts) {
    if ((rc = sa_parts_reduce(P.parts, grid, PW_PART, 128 * 128, dwp, st)) != SA_OK) return rc;
    if ((rc = sa_parts_reduce(P.parts + 128 * 128, grid, PW_PART, 128, dbias, st)) != SA_OK) return rc;
    if (dbias_h && (rc = sa_parts_reduce(P.parts + 128 * 128 + 128, grid, PW_PART, 128, dbias_h, st)) != SA_OK) return rc;
    return sa_parts_free(P.parts, st);
  }
  return SA_OK;
}

// y [m][128] = relu?(x W^T + bias + addend):  x [m][128], wp [128 (c_out)][128 (c_in)] bf16 (sa_pack_weight, no transpose),
// addend [m][128] bf16 (the residual input).  The streaming forward twin of sa_conv1x1_bwd_fused.
extern "C" int sa_conv1x1_fwd_fused(int64_t m, int c_out, int c_in, const void* x, const void* wp, const float* bias,
                                    const void* addend, int relu, void* y, void* stream) {
  SA_CHECK_ARG(x && wp && addend && y, "null pointer");
  SA_CHECK_ARG(m > 0, "bad sizes");
  SA_UNSUPPORTED(c_out != 128 || c_in != 128, "the fused pointwise forward is built for 128 -> 128 channels");
  SA_UNSUPPORTED(m >= (1LL << 31) - 256, "too many positions");
  SA_CHECK_ARG(!bias || (reinterpret_cast<uintptr_t>(bias) & 15) == 0, "bias not 16-byte aligned");
  if (!sa_get_tmap_encode()) { sa_set_error("cuTensorMapEncodeTiled unavailable"); return SA_ERR_CUDA; }
  std::call_once(g_pw_once, [] {
    int dev = 0, v = 0;
    cudaGetDevice(&dev);
    if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && v > 0) g_pw_sms = v;
    cudaFuncSetAttribute(tc_pw_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024 - 4096);
    cudaFuncSetAttribute(tc_pw_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024 - 4096);
  });
  sa_note_path(SA_PATH_TCGEN05);
  static thread_local PwParams P;
  P.m = m;
  P.tiles = (int)sa_cdiv(m, PW_M);
  P.dwp = nullptr; P.dbias = nullptr; P.dbias_h = nullptr; P.bias = bias; P.relu = relu; P.parts = nullptr;
  int rc;
  if ((rc = pw_map(&P.gmap, x, 128, (uint64_t)m, PW_M)) != SA_OK) return rc;
  if ((rc = pw_map(&P.hmap, addend, 128, (uint64_t)m, PW_M)) != SA_OK) return rc;
  if ((rc = pw_map(&P.omap, y, 128, (uint64_t)m, PW_M)) != SA_OK) return rc;
  if ((rc = pw_map(&P.wmap, wp, 128, 128, 128)) != SA_OK) return rc;
  const size_t smem = (size_t)(2 + 1 + 2 + PW_STAGES * 4) * PW_BLK + 1024;
  const unsigned grid = (unsigned)(P.tiles < g_pw_sms ? P.tiles : g_pw_sms);
  tc_pw_kernel<1><<<grid, PW_THREADS, smem, sa_stream(stream)>>>(P);
  SA_LAUNCH_CHECK();
  return SA_OK;
}
