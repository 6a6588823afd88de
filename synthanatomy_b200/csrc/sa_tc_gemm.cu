// tcgen05 GEMMs of the Performer path (sm_100a): bf16 operands staged by TMA into 128B-swizzled shared memory,
// fp32 accumulators in TMEM, fused epilogues out of TMEM.
//
//   NT  v[i][j] = sum_k A[i][k] B[j][k]          dense layers and their data gradients (both operands K-major)
//       persistent, one CTA per SM; 128 x BN output tiles (BN <= 256), 64-wide k blocks, 4-stage TMA ring,
//       TWO accumulator stages in TMEM (2 x BN columns) so the epilogue of tile t overlaps the MMAs of tile t+1.
//       warp 0 = TMA producer, warp 1 = TMEM owner + MMA issuer, warps 2-9 = epilogue (two warps per TMEM lane
//       quadrant, each taking half of the tile's columns).
//   TN  D[i][j] += s * sum_r A[r][i] B[r][j]    weight gradients (both operands MN-major: rows of the activation
//       matrices are the contraction index).  128 x BN tiles, 64-row k blocks, split-K over the rows with fp32
//       red.global.add into D.
//
// Reference call sites replaced: cuBLAS behind nn.Linear forward / backward of performer-pytorch SelfAttention
// (to_q/to_k/to_v/to_out), FeedForward (w1, w2) and /root/reference/src/networks/transformers/performer.py:221,286.
#include <mutex>
#include <stdlib.h>

#include "sa_pf_common.cuh"
#include "sa_tc_common.cuh"

using namespace satc;

namespace {

constexpr int G_THREADS = 320;
constexpr int G_BM = 128;
constexpr int G_BK = 64;
constexpr int G_STAGES = 4;
constexpr int G_STAGES_MAX = 6;               // CTA-pair kernel: 32 KB stages

struct NtParams {
  CUtensorMap amap, bmap;
  CUtensorMap omap_act, omap_b;   // TMA-store maps: out_act (bf16, 64 x 32 boxes); `pre` (bf16) or out_f32 (fp32, 32 x 32 boxes)
  CUtensorMap imap_a, imap_b;     // TMA-load maps of epilogue inputs: `pre` and `dot_with` (bf16 blocks) or `resid` (fp32 chunks)
  int stages, tma_out;            // pipeline stages; 1 = outputs leave through shared-memory staging + TMA stores
  int tma_in;                     // 1: GELU' epilogue reads pre / dot_with blocks through TMA; 2: residual chunks through TMA
  int row_out;                    // 1: fp32 output with an unaligned leading dimension (the logits, n = 2049): 32 x 32 chunks are
                                  //    transposed through shared memory so that a warp stores whole row segments
  long long m;
  int n, k;
  int BN, m_tiles, n_tiles, kblocks;
  int vec;          // epilogue pointers / leading dimension allow 16-byte accesses
  SaEpi e;
};

__device__ __forceinline__ void ld32_bf16(const __nv_bfloat16* p, float (&f)[32]) {
  const uint4* q = reinterpret_cast<const uint4*>(p);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const uint4 u = q[i];
    f[i * 8 + 0] = bf16lo(u.x); f[i * 8 + 1] = bf16hi(u.x); f[i * 8 + 2] = bf16lo(u.y); f[i * 8 + 3] = bf16hi(u.y);
    f[i * 8 + 4] = bf16lo(u.z); f[i * 8 + 5] = bf16hi(u.z); f[i * 8 + 6] = bf16lo(u.w); f[i * 8 + 7] = bf16hi(u.w);
  }
}
__device__ __forceinline__ void st32_bf16(__nv_bfloat16* p, const float (&f)[32]) {
  uint4* q = reinterpret_cast<uint4*>(p);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    uint4 u;
    u.x = pack_bf16x2(f[i * 8 + 0], f[i * 8 + 1]); u.y = pack_bf16x2(f[i * 8 + 2], f[i * 8 + 3]);
    u.z = pack_bf16x2(f[i * 8 + 4], f[i * 8 + 5]); u.w = pack_bf16x2(f[i * 8 + 6], f[i * 8 + 7]);
    q[i] = u;
  }
}
__device__ __forceinline__ void ld32_f32(const float* p, float (&f)[32]) {
  const float4* q = reinterpret_cast<const float4*>(p);
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const float4 u = q[i];
    f[i * 4 + 0] = u.x; f[i * 4 + 1] = u.y; f[i * 4 + 2] = u.z; f[i * 4 + 3] = u.w;
  }
}
__device__ __forceinline__ void st32_f32(float* p, const float (&f)[32]) {
  float4* q = reinterpret_cast<float4*>(p);
#pragma unroll
  for (int i = 0; i < 8; ++i) q[i] = make_float4(f[i * 4 + 0], f[i * 4 + 1], f[i * 4 + 2], f[i * 4 + 3]);
}

__device__ __forceinline__ void tma_store_2d(const CUtensorMap* m, const void* smem_src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(
                   reinterpret_cast<uint64_t>(m)),
               "r"(smem_u32(smem_src)), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read1() { asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
// 32 bf16 columns [c0, c0 + 32) of row r of a 128B-swizzled [32 rows][64 cols] staging block
__device__ __forceinline__ void stage_bf16_32(uint8_t* block, int r, int c0, const float (&f)[32]) {
  uint8_t* row = block + r * 128;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int ch = (c0 >> 3) + i;
    uint4 u;
    u.x = pack_bf16x2(f[i * 8 + 0], f[i * 8 + 1]); u.y = pack_bf16x2(f[i * 8 + 2], f[i * 8 + 3]);
    u.z = pack_bf16x2(f[i * 8 + 4], f[i * 8 + 5]); u.w = pack_bf16x2(f[i * 8 + 6], f[i * 8 + 7]);
    *reinterpret_cast<uint4*>(row + ((ch ^ (r & 7)) << 4)) = u;
  }
}
// 32 fp32 columns of row r of a 128B-swizzled [32 rows][32 cols] staging block
__device__ __forceinline__ void stage_f32_32(uint8_t* block, int r, const float (&f)[32]) {
  uint8_t* row = block + r * 128;
#pragma unroll
  for (int i = 0; i < 8; ++i)
    *reinterpret_cast<float4*>(row + ((i ^ (r & 7)) << 4)) = make_float4(f[i * 4], f[i * 4 + 1], f[i * 4 + 2], f[i * 4 + 3]);
}

// read back 32 bf16 / fp32 columns of row r of a staged block (inverse of stage_bf16_32 / stage_f32_32)
__device__ __forceinline__ void unstage_bf16_32(const uint8_t* block, int r, int c0, float (&f)[32]) {
  const uint8_t* row = block + r * 128;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int ch = (c0 >> 3) + i;
    const uint4 u = *reinterpret_cast<const uint4*>(row + ((ch ^ (r & 7)) << 4));
    f[i * 8 + 0] = bf16lo(u.x); f[i * 8 + 1] = bf16hi(u.x); f[i * 8 + 2] = bf16lo(u.y); f[i * 8 + 3] = bf16hi(u.y);
    f[i * 8 + 4] = bf16lo(u.z); f[i * 8 + 5] = bf16hi(u.z); f[i * 8 + 6] = bf16lo(u.w); f[i * 8 + 7] = bf16hi(u.w);
  }
}
__device__ __forceinline__ void unstage_f32_32(const uint8_t* block, int r, float (&f)[32]) {
  const uint8_t* row = block + r * 128;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const float4 u = *reinterpret_cast<const float4*>(row + ((i ^ (r & 7)) << 4));
    f[i * 4] = u.x; f[i * 4 + 1] = u.y; f[i * 4 + 2] = u.z; f[i * 4 + 3] = u.w;
  }
}

// GELU(x) = x Phi(x) and GELU'(x) = Phi(x) + x phi(x) with Phi through erf(|x| / sqrt 2) = 1 - poly(t) e^{-x^2/2},
// t = 1 / (1 + p |x| / sqrt 2)  (Abramowitz & Stegun 7.1.26, |error| < 1.5e-7): one reciprocal, one ex2 and a
// handful of FMAs instead of libm erff (+ expf for the derivative, which shares the exponential here).  Used by the
// bf16 tensor-core epilogue only (outputs are rounded to bf16, 2^-9); the fp32 parity path keeps erff.
__device__ __forceinline__ void fast_phi(float x, float& cdf, float& pdf) {
  const float ax = fabsf(x);
  const float z = ax * 0.70710678118654752f;
  float t;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(fmaf(0.3275911f, z, 1.0f)));
  float e;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(-0.72134752044448170f * x * x));     // e^{-x^2/2}
  float poly = fmaf(1.061405429f, t, -1.453152027f);
  poly = fmaf(poly, t, 1.421413741f);
  poly = fmaf(poly, t, -0.284496736f);
  poly = fmaf(poly, t, 0.254829592f);
  poly *= t;
  const float erf_abs = fmaf(-poly, e, 1.0f);                 // erf(|x| / sqrt 2)
  const float half = 0.5f * erf_abs;
  cdf = x >= 0.f ? 0.5f + half : 0.5f - half;
  pdf = 0.39894228040143268f * e;
}
__device__ __forceinline__ float fast_gelu(float x) {
  float c, p;
  fast_phi(x, c, p);
  return x * c;
}
__device__ __forceinline__ float fast_gelu_grad(float x) {
  float c, p;
  fast_phi(x, c, p);
  return fmaf(x, p, c);
}
// gelu(x) and gelu'(x) from one evaluation of Phi / phi (SA_ACT_GELU_FWD_D: the forward pass stores the derivative, so
// that the backward epilogue is one multiplication instead of a second erf evaluation)
__device__ __forceinline__ float fast_gelu_with_grad(float x, float& grad) {
  float c, p;
  fast_phi(x, c, p);
  grad = fmaf(x, p, c);
  return x * c;
}
__host__ __device__ __forceinline__ bool act_is_fwd(int a) { return a == SA_ACT_GELU_FWD || a == SA_ACT_GELU_FWD_D; }
__host__ __device__ __forceinline__ bool act_is_bwd(int a) { return a == SA_ACT_GELU_BWD || a == SA_ACT_MUL_PRE; }

// CG = 1: one CTA per 128 x BN tile.  CG = 2: a CTA pair (cluster of 2, cta_group::2) per 256 x BN tile -- CTA r stages
// rows r * 128 .. of the A tile and rows r * BN/2 .. of the B tile, the even CTA issues the 256-row MMAs, and each CTA
// drains its own 128 accumulator rows.  Per k block an SM then pulls 16 KB + BN/2 * 128 B from L2 instead of
// 16 KB + BN * 128 B for the same flops, which is what bounds the single-CTA kernel.
template <int CG>
__global__ void __launch_bounds__(G_THREADS, 1)
tc_gemm_nt_kernel(const __grid_constant__ NtParams P) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ uint64_t full_bar[G_STAGES_MAX];
  __shared__ uint64_t empty_bar[G_STAGES_MAX];
  __shared__ uint64_t tmem_full_bar[2];
  __shared__ uint64_t tmem_empty_bar[2];
  __shared__ uint64_t in_bar[8];            // one per epilogue warp: TMA loads of epilogue inputs
  __shared__ uint32_t tmem_base_slot;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t a_bytes = G_BM * 128;
  const uint32_t b_bytes = (uint32_t)(P.BN / CG) * 128;
  const uint32_t stage_bytes = a_bytes + b_bytes;
  const uint32_t rank = (CG == 2) ? cluster_ctarank() : 0u;
  const int tile0 = blockIdx.x / CG, tile_step = gridDim.x / CG;
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const int total_tiles = P.m_tiles * P.n_tiles;

  if (warp == 0 && lane == 0) {
    for (int i = 0; i < P.stages; ++i) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&tmem_full_bar[i], 1); mbar_init(&tmem_empty_bar[i], 8 * CG); }
    for (int i = 0; i < 8; ++i) mbar_init(&in_bar[i], 1);
    fence_mbar_init();
    fence_proxy_async();
  }
  if (CG == 2) cluster_sync_all();            // the peer's barriers exist before anything is signalled across the pair
  if (warp == 1) {
    if (CG == 2) { tmem_alloc_pair(&tmem_base_slot, 512); tmem_relinquish_pair(); }
    else { tmem_alloc(&tmem_base_slot, 512); tmem_relinquish(); }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_slot;

  if (warp == 0) {
    if (lane == 0) {
      prefetch_tmap(&P.amap);
      prefetch_tmap(&P.bmap);
      int stage = 0; uint32_t phase = 0;
      for (int tile = tile0; tile < total_tiles; tile += tile_step) {
        const int nt = tile % P.n_tiles, mt = tile / P.n_tiles;
        const int arow = (mt * CG + (int)rank) * G_BM;
        const int brow = nt * P.BN + (int)rank * (P.BN / CG);
        for (int kb = 0; kb < P.kblocks; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = smem + (size_t)stage * stage_bytes;
          if (CG == 2) {
            // both CTAs' bytes are credited to the leader's barrier (it may see the peer's bytes before its own expect_tx)
            if (rank == 0) mbar_expect_tx(&full_bar[stage], 2 * stage_bytes);
            const uint32_t lbar = mapa_u32(&full_bar[stage], 0);
            tma_load_2d_pair(sa, &P.amap, lbar, kb * G_BK, arow);
            tma_load_2d_pair(sa + a_bytes, &P.bmap, lbar, kb * G_BK, brow);
          } else {
            mbar_expect_tx(&full_bar[stage], stage_bytes);
            tma_load_2d(sa, &P.amap, &full_bar[stage], kb * G_BK, arow);
            tma_load_2d(sa + a_bytes, &P.bmap, &full_bar[stage], kb * G_BK, brow);
          }
          if (++stage == P.stages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0 && rank == 0) {
      const uint32_t idesc = make_idesc_bf16(G_BM * CG, P.BN, 0, 0);
      int stage = 0; uint32_t phase = 0;
      int acc = 0; uint32_t acc_phase = 0;
      for (int tile = tile0; tile < total_tiles; tile += tile_step) {
        mbar_wait(&tmem_empty_bar[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * P.BN);
        for (int kb = 0; kb < P.kblocks; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + (size_t)stage * stage_bytes);
          const uint32_t sb = sa + a_bytes;
#pragma unroll
          for (int k = 0; k < G_BK / 16; ++k) {
            const uint64_t da = make_smem_desc(sa + k * 32, 16, 1024, 2);
            const uint64_t db = make_smem_desc(sb + k * 32, 16, 1024, 2);
            if (CG == 2) umma_bf16_pair(d_tmem, da, db, idesc, (kb | k) != 0);
            else umma_bf16(d_tmem, da, db, idesc, (kb | k) != 0);
          }
          if (CG == 2) umma_commit_pair(&empty_bar[stage]); else umma_commit(&empty_bar[stage]);
          if (++stage == P.stages) { stage = 0; phase ^= 1; }
        }
        if (CG == 2) umma_commit_pair(&tmem_full_bar[acc]); else umma_commit(&tmem_full_bar[acc]);
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else {
    // ------------------------------------------------------------ epilogue
    using T = __nv_bfloat16;
    const SaEpi& e = P.e;
    const int ew = warp - 2;                  // 0..7
    const int quad = warp & 3;                // TMEM lane quadrant this warp may access
    const int half = ew >> 2;                 // which half of the tile's columns
    const int half_cols = P.BN / 2;           // BN is a multiple of 16 -> halves are multiples of 8
    const float st = e.scale * (e.scale_dev ? __ldg(e.scale_dev) : 1.0f);
    uint8_t* stgA = smem + (size_t)P.stages * stage_bytes + (size_t)ew * 8192;   // out_act block   [32 rows][64 bf16]
    uint8_t* stgB = stgA + 4096;                                               // pre block / out_f32 chunk
    float dot = 0.f;
    uint32_t in_phase = 0;
    int acc = 0; uint32_t acc_phase = 0;
    for (int tile = tile0; tile < total_tiles; tile += tile_step) {
      const int nt = tile % P.n_tiles, mt = tile / P.n_tiles;
      const int row0 = (mt * CG + (int)rank) * G_BM + quad * 32;
      const long long row = (long long)row0 + lane;
      const bool row_ok = row < P.m;
      mbar_wait(&tmem_full_bar[acc], acc_phase);
      tc_fence_after();
      for (int cc = 0; cc < half_cols; cc += 32) {
        const int ct = half * half_cols + cc;             // column inside the tile
        const int col = nt * P.BN + ct;                   // global column
        uint32_t v[32];
        tmem_ld_32x32(tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(acc * P.BN + ct), v);
        tmem_ld_wait();
        const int nc = min(min(32, half_cols - cc), P.n - col);
        if (nc <= 0) continue;                             // warp-uniform
        if (!P.tma_out && !P.row_out && !row_ok) continue;  // (the staging paths need every lane)
        const long long o = row * e.ldo + col;
        if (P.tma_out) {
          // Outputs leave through 128B-swizzled shared-memory staging blocks and TMA stores: the per-thread stores
          // (thread = row) would touch 32 different 128-byte lines per instruction; the bulk stores write whole rows.
          // Rows past m / nothing past n (n % 32 == 0 on this path) are clipped by the TMA unit.
          float f[32], t[32];
          const int ci = cc >> 5;
#pragma unroll
          for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[j]);
          if (P.tma_in == 1) {
            // GELU' epilogue (du = g (dx W2^T) gelu'(u), dot = sum (dx W2^T) h): the u and h blocks of this warp arrive by
            // TMA in stgA / stgB, du is formed in place over u and leaves by TMA
            if ((ci & 1) == 0) {
              if (lane == 0) {
                bulk_wait_read0();
                mbar_expect_tx(&in_bar[ew], e.dot_with ? 8192u : 4096u);
                tma_load_2d(stgA, &P.imap_a, &in_bar[ew], col, row0);
                if (e.dot_with) tma_load_2d(stgB, &P.imap_b, &in_bar[ew], col, row0);
              }
              mbar_wait(&in_bar[ew], in_phase);
              in_phase ^= 1;
            }
            if (e.bias) {
              ld32_f32(e.bias + col, t);
#pragma unroll
              for (int j = 0; j < 32; ++j) f[j] += t[j];
            }
            if (e.dot_with) {
              unstage_bf16_32(stgB, lane, (ci & 1) * 32, t);
              if (row_ok) {
#pragma unroll
                for (int j = 0; j < 32; ++j) dot = fmaf(f[j], t[j], dot);
              }
            }
            unstage_bf16_32(stgA, lane, (ci & 1) * 32, t);
            if (e.act == SA_ACT_MUL_PRE) {
#pragma unroll
              for (int j = 0; j < 32; ++j) f[j] = f[j] * st * t[j];
            } else {
#pragma unroll
              for (int j = 0; j < 32; ++j) f[j] = f[j] * st * fast_gelu_grad(t[j]);
            }
            stage_bf16_32(stgA, lane, (ci & 1) * 32, f);
            if (ci & 1) {
              fence_proxy_async();
              __syncwarp();
              if (lane == 0) { tma_store_2d(&P.omap_act, stgA, col - 32, row0); bulk_commit(); }
            }
            continue;
          }
          if (e.bias) {
            ld32_f32(e.bias + col, t);
#pragma unroll
            for (int j = 0; j < 32; ++j) f[j] += t[j];
          }
          if (e.dot_with && row_ok) {
            ld32_bf16(reinterpret_cast<const T*>(e.dot_with) + o, t);
#pragma unroll
            for (int j = 0; j < 32; ++j) dot = fmaf(f[j], t[j], dot);
          }
#pragma unroll
          for (int j = 0; j < 32; ++j) f[j] *= st;
          if (act_is_fwd(e.act)) {
            if ((ci & 1) == 0) { if (lane == 0) bulk_wait_read0(); __syncwarp(); }
            if (e.act == SA_ACT_GELU_FWD_D) {          // `pre` receives gelu'(v)
#pragma unroll
              for (int j = 0; j < 32; ++j) f[j] = fast_gelu_with_grad(f[j], t[j]);
              stage_bf16_32(stgB, lane, (ci & 1) * 32, t);
            } else {
              stage_bf16_32(stgB, lane, (ci & 1) * 32, f);
            }
            if (ci & 1) {
              fence_proxy_async();
              __syncwarp();
              if (lane == 0) { tma_store_2d(&P.omap_b, stgB, col - 32, row0); bulk_commit(); }
            }
            if (e.act == SA_ACT_GELU_FWD) {
#pragma unroll
              for (int j = 0; j < 32; ++j) f[j] = fast_gelu(f[j]);
            }
          } else if (act_is_bwd(e.act)) {
            if (row_ok) {
              ld32_bf16(reinterpret_cast<const T*>(e.pre) + o, t);
              if (e.act == SA_ACT_MUL_PRE) {
#pragma unroll
                for (int j = 0; j < 32; ++j) f[j] *= t[j];
              } else {
#pragma unroll
                for (int j = 0; j < 32; ++j) f[j] *= fast_gelu_grad(t[j]);
              }
            }
          }
          if (P.tma_in == 2) {
            // residual chunk of this warp through TMA into stgB; out_f32 is then formed in place
            if (lane == 0) {
              bulk_wait_read0();
              mbar_expect_tx(&in_bar[ew], 4096);
              tma_load_2d(stgB, &P.imap_a, &in_bar[ew], col, row0);
            }
            mbar_wait(&in_bar[ew], in_phase);
            in_phase ^= 1;
            unstage_f32_32(stgB, lane, t);
#pragma unroll
            for (int j = 0; j < 32; ++j) f[j] += t[j];
          } else if (e.resid && row_ok) {
            ld32_f32(e.resid + o, t);
#pragma unroll
            for (int j = 0; j < 32; ++j) f[j] += t[j];
          }
          if (e.out_f32) {
            if (P.tma_in != 2) { if (lane == 0) bulk_wait_read0(); }
            __syncwarp();
            stage_f32_32(stgB, lane, f);
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) { tma_store_2d(&P.omap_b, stgB, col, row0); bulk_commit(); }
          }
          if (e.out_act) {
            // single output: the two staging blocks alternate (wait only for the store issued two blocks ago)
            const bool solo = !e.out_f32 && !act_is_fwd(e.act);
            uint8_t* sbuf = (solo && (ci & 2)) ? stgB : stgA;
            if ((ci & 1) == 0 && solo) { if (lane == 0) bulk_wait_read1(); __syncwarp(); }
            stage_bf16_32(sbuf, lane, (ci & 1) * 32, f);
            if (ci & 1) {
              fence_proxy_async();
              __syncwarp();
              if (lane == 0) { tma_store_2d(&P.omap_act, sbuf, col - 32, row0); bulk_commit(); }
            }
          }
        } else if (P.vec && nc == 32) {
          float f[32], t[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[j]);
          if (e.bias) {
            ld32_f32(e.bias + col, t);
#pragma unroll
            for (int j = 0; j < 32; ++j) f[j] += t[j];
          }
          if (e.dot_with) {
            ld32_bf16(reinterpret_cast<const T*>(e.dot_with) + o, t);
#pragma unroll
            for (int j = 0; j < 32; ++j) dot = fmaf(f[j], t[j], dot);
          }
#pragma unroll
          for (int j = 0; j < 32; ++j) f[j] *= st;
          if (e.act == SA_ACT_GELU_FWD) {
            st32_bf16(reinterpret_cast<T*>(e.pre) + o, f);
#pragma unroll
            for (int j = 0; j < 32; ++j) f[j] = fast_gelu(f[j]);
          } else if (e.act == SA_ACT_GELU_FWD_D) {
#pragma unroll
            for (int j = 0; j < 32; ++j) f[j] = fast_gelu_with_grad(f[j], t[j]);
            st32_bf16(reinterpret_cast<T*>(e.pre) + o, t);
          } else if (e.act == SA_ACT_GELU_BWD) {
            ld32_bf16(reinterpret_cast<const T*>(e.pre) + o, t);
#pragma unroll
            for (int j = 0; j < 32; ++j) f[j] *= fast_gelu_grad(t[j]);
          } else if (e.act == SA_ACT_MUL_PRE) {
            ld32_bf16(reinterpret_cast<const T*>(e.pre) + o, t);
#pragma unroll
            for (int j = 0; j < 32; ++j) f[j] *= t[j];
          }
          if (e.resid) {
            ld32_f32(e.resid + o, t);
#pragma unroll
            for (int j = 0; j < 32; ++j) f[j] += t[j];
          }
          if (e.out_f32) st32_f32(e.out_f32 + o, f);
          if (e.out_act) st32_bf16(reinterpret_cast<T*>(e.out_act) + o, f);
        } else if (P.row_out) {
          // (acc + bias) * scale -> fp32, no other epilogue tensor: thread = row would store 4 bytes to 32 different
          // lines per instruction; through the staging block a warp writes 128 contiguous bytes of ONE row at a time
          float f[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) f[j] = (__uint_as_float(v[j]) + ((e.bias && j < nc) ? __ldg(e.bias + col + j) : 0.f)) * st;
          stage_f32_32(stgB, lane, f);
          __syncwarp();
          const int rows_here = (int)min(32LL, P.m - (long long)row0);
          if (lane < nc) {
            float* dst = e.out_f32 + (long long)row0 * e.ldo + col + lane;
            for (int i = 0; i < rows_here; ++i)
              dst[(long long)i * e.ldo] =
                  *reinterpret_cast<const float*>(stgB + i * 128 + ((((lane >> 2) ^ (i & 7))) << 4) + (lane & 3) * 4);
          }
          __syncwarp();
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (j < nc) dot += sa_epi_elem<T>(e, row, col + j, __uint_as_float(v[j]), st);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {                       // the accumulator stage is free once BOTH CTAs have drained their rows
        if (CG == 2 && rank != 0) mbar_arrive_cluster(mapa_u32(&tmem_empty_bar[acc], 0));
        else mbar_arrive(&tmem_empty_bar[acc]);
      }
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
    if (e.dot_out) {
      dot = sa_warp_sum(dot);
      if (lane == 0) atomicAdd(e.dot_out, dot);
    }
    if (P.tma_out && lane == 0) bulk_wait0();
  }

  tc_fence_before();
  if (CG == 2) {
    cluster_sync_all();                       // neither CTA leaves (or frees TMEM) while the other may still signal it
    if (warp == 1) tmem_dealloc_pair(tmem_base, 512);
  } else {
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, 512);
  }
}

// ------------------------------------------------------------------------------------------------ TN (wgrad)
constexpr int W_THREADS = 192;
constexpr int W_KP = 64;                       // contraction rows per stage
constexpr int W_BLOCK = W_KP * 128;            // one (64 columns x 64 rows) swizzled block
constexpr int W_STAGES = 4;

struct TnParams {
  CUtensorMap amap, bmap;
  long long m;
  int na, nb;
  int BN, bblocks;             // tile columns (multiple of 64, <= 256) and BN / 64
  int a_tiles, b_tiles;
  long long kblocks_total, kblocks_per_split;
  int tmem_cols;
  const float* scale_dev;
  float scale;
  float* d;
  float* colsum;               // optional: colsum[i] += sum_r A[r][i] (unscaled) -- the bias gradient that goes with D
  int red4;                    // D rows are 16-byte aligned: vector reductions
  float* parts;                // deterministic mode: split s stores its partial [na x nb | na column sums] at
  long long part_stride;       //   parts + s * part_stride (summed in split order afterwards)
};

__global__ void __launch_bounds__(W_THREADS, 1)
tc_gemm_tn_kernel(const __grid_constant__ TnParams P) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ uint64_t full_bar[W_STAGES];
  __shared__ uint64_t empty_bar[W_STAGES];
  __shared__ uint64_t tmem_full_bar;
  __shared__ uint32_t tmem_base_slot;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const int tile = blockIdx.x % (P.a_tiles * P.b_tiles);
  const int split = blockIdx.x / (P.a_tiles * P.b_tiles);
  const int bt = tile % P.b_tiles, at = tile / P.b_tiles;
  const long long kb_beg = (long long)split * P.kblocks_per_split;
  const long long kb_end = min(P.kblocks_total, kb_beg + P.kblocks_per_split);
  const long long nkb = kb_end - kb_beg;
  const uint32_t a_bytes = 2 * W_BLOCK;
  const uint32_t b_bytes = (uint32_t)P.bblocks * W_BLOCK;
  const uint32_t stage_bytes = a_bytes + b_bytes;
  // column sums of A ride along as one more product, A^T 1: a [64 rows x 64 cols] "ones" operand (columns 0..15 = 1)
  // after the ring, a 16-column accumulator after the tile's; only the CTAs of the first column tile do it
  const bool do_colsum = P.colsum != nullptr && bt == 0;
  uint8_t* Ones = smem + (size_t)W_STAGES * stage_bytes;
  if (do_colsum && warp >= 2) {
    for (int idx = threadIdx.x - 64; idx < W_KP * 8; idx += W_THREADS - 64) {
      const int r = idx >> 3, ch = idx & 7;
      const uint32_t one2 = 0x3F803F80u;      // two bf16 ones
      const uint4 u = ch < 2 ? make_uint4(one2, one2, one2, one2) : make_uint4(0, 0, 0, 0);
      *reinterpret_cast<uint4*>(Ones + (r >> 3) * 1024 + (r & 7) * 128 + ((ch ^ (r & 7)) << 4)) = u;
    }
    fence_proxy_async();
  }

  if (warp == 0 && lane == 0) {
    for (int i = 0; i < W_STAGES; ++i) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], 1); }
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

  if (warp == 0) {
    if (lane == 0 && nkb > 0) {
      prefetch_tmap(&P.amap);
      prefetch_tmap(&P.bmap);
      int stage = 0; uint32_t phase = 0;
      for (long long kb = kb_beg; kb < kb_end; ++kb) {
        mbar_wait(&empty_bar[stage], phase ^ 1);
        mbar_expect_tx(&full_bar[stage], stage_bytes);
        uint8_t* sa = smem + (size_t)stage * stage_bytes;
        const int r0 = (int)(kb * W_KP);
        for (int h = 0; h < 2; ++h) tma_load_2d(sa + h * W_BLOCK, &P.amap, &full_bar[stage], at * 128 + h * 64, r0);
        for (int c = 0; c < P.bblocks; ++c)
          tma_load_2d(sa + a_bytes + c * W_BLOCK, &P.bmap, &full_bar[stage], bt * P.BN + c * 64, r0);
        if (++stage == W_STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    if (lane == 0 && nkb > 0) {
      const uint32_t idesc = make_idesc_bf16(128, P.BN, 1, 1);   // both operands MN-major
      const uint32_t idesc_1 = make_idesc_bf16(128, 16, 1, 1);
      const uint32_t ones_a = smem_u32(Ones);
      int stage = 0; uint32_t phase = 0;
      for (long long kb = 0; kb < nkb; ++kb) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        const uint32_t sa = smem_u32(smem + (size_t)stage * stage_bytes);
        const uint32_t sb = sa + a_bytes;
#pragma unroll
        for (int j = 0; j < W_KP / 16; ++j) {
          // MN-major SWIZZLE_128B: 64-column blocks W_BLOCK apart (LBO), 8-row groups 1024 B apart (SBO),
          // a K step of 16 rows = 2048 B
          const uint64_t da = make_smem_desc(sa + j * 2048, W_BLOCK, 1024, 2);
          const uint64_t db = make_smem_desc(sb + j * 2048, W_BLOCK, 1024, 2);
          umma_bf16(tmem_base, da, db, idesc, (kb | j) != 0);
          if (do_colsum)
            umma_bf16(tmem_base + (uint32_t)P.BN, da, make_smem_desc(ones_a + j * 2048, W_BLOCK, 1024, 2), idesc_1, (kb | j) != 0);
        }
        umma_commit(&empty_bar[stage]);
        if (++stage == W_STAGES) { stage = 0; phase ^= 1; }
      }
      umma_commit(&tmem_full_bar);
    }
  } else if (nkb > 0) {
    const int quad = warp & 3;
    const int i = at * 128 + quad * 32 + lane;
    const float st = P.scale * (P.scale_dev ? __ldg(P.scale_dev) : 1.0f);
    mbar_wait(&tmem_full_bar, 0);
    tc_fence_after();
    float* const out = P.parts ? P.parts + (long long)split * P.part_stride : P.d;
    for (int c0 = 0; c0 < P.BN; c0 += 32) {
      uint32_t v[32];
      tmem_ld_32x32(tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)c0, v);
      tmem_ld_wait();
      const int col = bt * P.BN + c0;
      if (i < P.na) {
        float* dst = out + (long long)i * P.nb + col;
        const int nc = min(32, P.nb - col);
        if (P.parts) {
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (j < nc) dst[j] = st * __uint_as_float(v[j]);
        } else if (P.red4 && nc == 32) {
#pragma unroll
          for (int j = 0; j < 32; j += 4)
            sa_red_add_v4(dst + j, st * __uint_as_float(v[j]), st * __uint_as_float(v[j + 1]), st * __uint_as_float(v[j + 2]),
                          st * __uint_as_float(v[j + 3]));
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (j < nc) atomicAdd(dst + j, st * __uint_as_float(v[j]));
        }
      }
    }
    if (do_colsum) {
      uint32_t v[32];
      tmem_ld_32x32(tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)P.BN, v);   // columns 0..15 all hold the sum
      tmem_ld_wait();
      if (i < P.na) {
        if (P.parts) out[(long long)P.na * P.nb + i] = __uint_as_float(v[0]);
        else atomicAdd(P.colsum + i, __uint_as_float(v[0]));
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, (uint32_t)P.tmem_cols);
}

std::once_flag g_once;
int g_max_smem = 0;
int g_sms = 148;

void init_once() {
  std::call_once(g_once, [] {
    int dev = 0; cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&g_max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    cudaDeviceGetAttribute(&g_sms, cudaDevAttrMultiProcessorCount, dev);
    cudaFuncSetAttribute(tc_gemm_nt_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, g_max_smem - 2048);
    cudaFuncSetAttribute(tc_gemm_nt_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, g_max_smem - 2048);
    cudaFuncSetAttribute(tc_gemm_tn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, g_max_smem - 2048);
  });
}

bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

int make_2d(CUtensorMap* m, const void* base, uint64_t cols, uint64_t rows, uint64_t ld, uint32_t box_cols, uint32_t box_rows) {
  const uint64_t dims[2] = {cols, rows};
  const uint64_t strides[2] = {2, ld * 2};
  const uint32_t box[2] = {box_cols, box_rows};
  return sa_make_tmap_bf16(m, base, 2, dims, strides, box);
}

}  // namespace

bool sa_tc_gemm_nt_supported(int64_t m, int n, int k, int dtype, const void* a, int64_t lda, const void* b, int64_t ldb) {
  if (dtype != SA_BF16) return false;
  if (m < 1 || n < 8 || k < 16) return false;
  if ((lda & 7) || (ldb & 7) || !aligned16(a) || !aligned16(b)) return false;
  if (m >= (1LL << 31) - 256) return false;
  return sa_get_tmap_encode() != nullptr;
}

int sa_tc_gemm_nt(int64_t m, int n, int k, const void* a, int64_t lda, const void* b, int64_t ldb, const SaEpi& e,
                  cudaStream_t st) {
  init_once();
  sa_note_path(SA_PATH_TCGEN05);
  static thread_local NtParams P;
  P.m = m; P.n = n; P.k = k;
  P.BN = n >= 256 ? 256 : (int)sa_cdiv(n, 16) * 16;
  // CTA pairs for the big launches (256-row tiles); SA_GEMM_PAIR=0 keeps the single-CTA kernel
  // (measured at m = 84 000: +6 % at k = 1024 / 2048, -5 % at k = 512, where the tile count per CTA matters more)
  bool pair = m >= 1024 && k >= 1024 && P.BN % 32 == 0 && g_sms >= 2;
  if (const char* env = getenv("SA_GEMM_PAIR")) { pair = env[0] != '0' && m >= 1024 && P.BN % 32 == 0 && g_sms >= 2; }
  const int cg = pair ? 2 : 1;
  P.m_tiles = (int)sa_cdiv(m, G_BM * cg);
  P.n_tiles = (int)sa_cdiv(n, P.BN);
  P.kblocks = (int)sa_cdiv(k, G_BK);
  P.e = e;
  bool vec = (e.ldo & 7) == 0;
  const void* ptrs[] = {e.bias, e.dot_with, e.pre, e.resid, e.out_f32, e.out_act};
  for (const void* p : ptrs) vec = vec && aligned16(p);
  P.vec = vec ? 1 : 0;
  int rc = make_2d(&P.amap, a, (uint64_t)k, (uint64_t)m, (uint64_t)lda, G_BK, G_BM);
  if (rc != SA_OK) return rc;
  rc = make_2d(&P.bmap, b, (uint64_t)k, (uint64_t)n, (uint64_t)ldb, G_BK, (uint32_t)(P.BN / cg));
  if (rc != SA_OK) return rc;
  // TMA-store epilogue: whole 64-column blocks per warp, at most one "second" output (pre or out_f32)
  P.tma_out = (vec && P.BN % 128 == 0 && n % 64 == 0 && (e.out_act || e.out_f32) &&
               !(e.out_f32 && act_is_fwd(e.act))) ? 1 : 0;
  if (const char* env = getenv("SA_GEMM_TMA_OUT")) { if (env[0] == '0') P.tma_out = 0; }
  const size_t stage_bytes = G_BM * 128 + (size_t)(P.BN / cg) * 128;
  P.stages = pair ? G_STAGES_MAX : G_STAGES;
  P.row_out = (!P.tma_out && !vec && e.out_f32 && !e.out_act && !e.pre && !e.dot_with && !e.resid && e.act == SA_ACT_NONE) ? 1 : 0;
  if (const char* env = getenv("SA_GEMM_ROW_OUT")) { if (env[0] == '0') P.row_out = 0; }
  const size_t stg_bytes = (P.tma_out || P.row_out) ? 8 * 8192 : 0;
  while (P.stages > 2 && (size_t)P.stages * stage_bytes + stg_bytes + 1024 > (size_t)g_max_smem - 2048) --P.stages;
  if (P.tma_out) {
    const uint64_t dims[2] = {(uint64_t)n, (uint64_t)m};
    if (e.out_act) {
      const uint64_t strides[2] = {2, (uint64_t)e.ldo * 2};
      const uint32_t box[2] = {64, 32};
      if ((rc = sa_make_tmap(&P.omap_act, SA_BF16, e.out_act, 2, dims, strides, box)) != SA_OK) return rc;
    }
    if (act_is_fwd(e.act)) {
      const uint64_t strides[2] = {2, (uint64_t)e.ldo * 2};
      const uint32_t box[2] = {64, 32};
      if ((rc = sa_make_tmap(&P.omap_b, SA_BF16, e.pre, 2, dims, strides, box)) != SA_OK) return rc;
    } else if (e.out_f32) {
      const uint64_t strides[2] = {4, (uint64_t)e.ldo * 4};
      const uint32_t box[2] = {32, 32};
      if ((rc = sa_make_tmap(&P.omap_b, SA_F32, e.out_f32, 2, dims, strides, box)) != SA_OK) return rc;
    }
  }
  P.tma_in = 0;
  if (P.tma_out) {
    const uint64_t dims[2] = {(uint64_t)n, (uint64_t)m};
    const bool want = !getenv("SA_GEMM_TMA_IN") || getenv("SA_GEMM_TMA_IN")[0] != '0';
    if (want && act_is_bwd(e.act) && e.out_act && !e.out_f32 && !e.resid) {
      const uint64_t strides[2] = {2, (uint64_t)e.ldo * 2};
      const uint32_t box[2] = {64, 32};
      if ((rc = sa_make_tmap(&P.imap_a, SA_BF16, e.pre, 2, dims, strides, box)) != SA_OK) return rc;
      if (e.dot_with && (rc = sa_make_tmap(&P.imap_b, SA_BF16, e.dot_with, 2, dims, strides, box)) != SA_OK) return rc;
      P.tma_in = 1;
    } else if (want && e.resid && e.out_f32 && e.act == SA_ACT_NONE) {
      const uint64_t strides[2] = {4, (uint64_t)e.ldo * 4};
      const uint32_t box[2] = {32, 32};
      if ((rc = sa_make_tmap(&P.imap_a, SA_F32, e.resid, 2, dims, strides, box)) != SA_OK) return rc;
      P.tma_in = 2;
    }
  }
  const size_t smem = (size_t)P.stages * stage_bytes + stg_bytes + 1024;
  const int total = P.m_tiles * P.n_tiles;
  if (pair) {
    const int pairs = g_sms / 2;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(2 * (total < pairs ? total : pairs)));
    cfg.blockDim = dim3(G_THREADS);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    const cudaError_t err = cudaLaunchKernelEx(&cfg, tc_gemm_nt_kernel<2>, P);
    if (err != cudaSuccess) { sa_set_error("tc_gemm_nt (CTA pairs): %s", cudaGetErrorString(err)); return SA_ERR_CUDA; }
  } else {
    const unsigned grid = (unsigned)(total < g_sms ? total : g_sms);
    tc_gemm_nt_kernel<1><<<grid, G_THREADS, smem, st>>>(P);
  }
  SA_LAUNCH_CHECK();
  return SA_OK;
}

bool sa_tc_gemm_tn_supported(int64_t m, int na, int nb, int dtype, const void* a, int64_t lda, const void* b, int64_t ldb) {
  if (dtype != SA_BF16) return false;
  if (m < 64 || na < 8 || nb < 8) return false;
  if ((lda & 7) || (ldb & 7) || !aligned16(a) || !aligned16(b)) return false;
  if (m >= (1LL << 31) - 256) return false;
  return sa_get_tmap_encode() != nullptr;
}

int sa_tc_gemm_tn_colsum(int64_t m, int na, int nb, const void* a, int64_t lda, const void* b, int64_t ldb,
                         const float* scale_dev, float scale, float* d, float* colsum, cudaStream_t st);

int sa_tc_gemm_tn(int64_t m, int na, int nb, const void* a, int64_t lda, const void* b, int64_t ldb,
                  const float* scale_dev, float scale, float* d, cudaStream_t st) {
  return sa_tc_gemm_tn_colsum(m, na, nb, a, lda, b, ldb, scale_dev, scale, d, nullptr, st);
}

int sa_tc_gemm_tn_colsum(int64_t m, int na, int nb, const void* a, int64_t lda, const void* b, int64_t ldb,
                         const float* scale_dev, float scale, float* d, float* colsum, cudaStream_t st) {
  init_once();
  sa_note_path(SA_PATH_TCGEN05);
  static thread_local TnParams P;
  P.m = m; P.na = na; P.nb = nb;
  P.BN = nb >= 256 ? 256 : (int)sa_cdiv(nb, 64) * 64;
  P.bblocks = P.BN / 64;
  P.a_tiles = (int)sa_cdiv(na, 128);
  P.b_tiles = (int)sa_cdiv(nb, P.BN);
  P.kblocks_total = sa_cdiv(m, W_KP);
  const int acc_cols = P.BN + (colsum ? 32 : 0);       // the column-sum accumulator: 16 columns after the tile's
  P.tmem_cols = acc_cols <= 64 ? 64 : (acc_cols <= 128 ? 128 : (acc_cols <= 256 ? 256 : 512));
  P.scale_dev = scale_dev; P.scale = scale; P.d = d; P.colsum = colsum;
  const int64_t tiles = (int64_t)P.a_tiles * P.b_tiles;
  int64_t splits = g_sms / tiles;
  const int64_t max_splits = sa_cdiv(P.kblocks_total, 8);
  if (splits > max_splits) splits = max_splits;
  if (splits < 1) splits = 1;
  P.kblocks_per_split = sa_cdiv(P.kblocks_total, splits);
  splits = sa_cdiv(P.kblocks_total, P.kblocks_per_split);
  int rc = make_2d(&P.amap, a, (uint64_t)na, (uint64_t)m, (uint64_t)lda, 64, W_KP);
  if (rc != SA_OK) return rc;
  rc = make_2d(&P.bmap, b, (uint64_t)nb, (uint64_t)m, (uint64_t)ldb, 64, W_KP);
  if (rc != SA_OK) return rc;
  const size_t smem = (size_t)W_STAGES * (2 * W_BLOCK + (size_t)P.bblocks * W_BLOCK) + W_BLOCK + 1024;
  P.red4 = ((nb & 3) == 0 && aligned16(d)) ? 1 : 0;
  P.part_stride = (long long)na * nb + (colsum ? na : 0);
  P.parts = sa_parts_alloc(splits, P.part_stride, st);
  tc_gemm_tn_kernel<<<(unsigned)(tiles * splits), W_THREADS, smem, st>>>(P);
  SA_LAUNCH_CHECK();
  if (P.parts) {
    if ((rc = sa_parts_reduce(P.parts, splits, P.part_stride, (long long)na * nb, d, st)) != SA_OK) return rc;
    if (colsum && (rc = sa_parts_reduce(P.parts + (long long)na * nb, splits, P.part_stride, na, colsum, st)) != SA_OK) return rc;
    return sa_parts_free(P.parts, st);
  }
  return SA_OK;
}
