// tcgen05 local-window causal attention for sm_100a (bf16 operands, fp32 accumulation), flash style: the score tensor
// never leaves the SM.
//
//   query p attends keys j with lo(p) <= j <= p,  lo(p) = max(0, floor(p / w) - 1) * w      (window w, look-back 1)
//   S = q k^T * d^-1/2   P = softmax(S)   O = P v          (q, k already carry the rotary position term: sa_rotary)
//
// forward   one CTA per (batch, head, 128-query tile); key tiles of 64.  S = Q K^T (double-buffered in TMEM) and
//           O += P V on tcgen05 with O resident in TMEM: the softmax warps (thread = query row) read an S row once, keep
//           a reference maximum that is raised only when a tile exceeds it by 2^8 (then they rescale their O lanes with
//           tcgen05.ld / st), and hand P over as the TMEM-resident A operand of the second product (TS form) -- they
//           never wait for a P V product, so the tensor pipe simply runs one tile behind.
// backward  dq kernel: per 128-query tile, loop key tiles:  S, dP = dO V^T  ->  dS = P (dP - delta) / sqrt(d)  ->  dQ += dS K;
//                      S / dP of tile t + 1 are issued as soon as the softmax warps have read tile t's out of TMEM
//           dkv kernel: per 128-key tile, loop 64-query tiles, transposed roles (TMEM lane = key):
//                       S^T = K Q^T, dP^T = V dO^T  ->  P^T, dS^T (written in place over S^T / dP^T)  ->  dV += P^T dO,
//                       dK += dS^T Q
//           Causal + window masks are one visible interval per row: boundary tiles overwrite hidden scores with -inf
//           through a 64-bit column mask and then run the interior-tile arithmetic.
// Warp roles: warps 0-3 softmax / epilogue (TMEM lane quadrants 0-3), warp 4 MMA issuer, warp 5 TMA producer.
//
// Replaces local_attention.LocalAttention.forward (+ autograd) called by performer-pytorch SelfAttention for the local
// heads, reached from /root/reference/src/networks/transformers/performer.py:270 (ctor args :199-200).
#include <mutex>
#include <stdlib.h>

#include "sa_pf_common.cuh"
#include "sa_tc_common.cuh"

using namespace satc;

namespace {

constexpr int L_THREADS = 192;
constexpr int L_STAGES_FWD = 3;
constexpr int L_STAGES_BWD = 2;
constexpr int L_STAGES_DQ = 2;            // dq kernel: the second dS buffer takes the third K/V stage's shared memory
constexpr float LOG2E = 1.4426950408889634f;

struct LcParams {
  CUtensorMap qmap, kmap, vmap, domap;   // 2-D maps over the [rows][ld] buffers, based at head 0 of each block
  int B, N, H, W;
  int ld, out_ld;
  int fast;                  // 1: warp-uniform interior-tile fast path (SA_LOCAL_FASTMASK=0 turns it off)
  float scale;
  const __nv_bfloat16* out;
  const __nv_bfloat16* dout;
  __nv_bfloat16* o_out;      // forward output
  float* lse;
  __nv_bfloat16* dq;
  __nv_bfloat16* dk;
  __nv_bfloat16* dv;
  float* delta_ws;           // optional [batch * heads][seq]: delta = dO . O per query, written by the dq kernel for the dk/dv kernel
  const float2* rot;         // optional [seq][32] (cos, sin): q / k carry the rotary term, dq / dk leave through its transpose
};

__device__ __forceinline__ int lc_lo(int p, int W) { const int w = p / W - 1; return (w > 0 ? w : 0) * W; }
__device__ __forceinline__ float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// bit c of (lo | hi << 32) = column c of a 64-column score tile is visible, for the visible interval [c_lo, c_hi] of a row
// (causal + window masks are one interval per row): boundary tiles overwrite their hidden scores with -inf through one
// bit test + select per score and then run the same arithmetic as interior tiles (exp2(-inf) = 0 exactly)
__device__ __forceinline__ void lc_colmask(int c_lo, int c_hi, uint32_t& lo, uint32_t& hi) {
  c_lo = max(c_lo, 0);
  c_hi = min(c_hi, 63);
  unsigned long long m = 0ull;
  if (c_hi >= c_lo) m = (~0ull >> (63 - c_hi)) & (~0ull << c_lo);
  lo = (uint32_t)m;
  hi = (uint32_t)(m >> 32);
}
__device__ __forceinline__ void lc_apply_mask(uint32_t (&v)[32], uint32_t bits) {
#pragma unroll
  for (int c = 0; c < 32; ++c) v[c] = (bits >> c) & 1u ? v[c] : 0xff800000u;      // -inf
}
__device__ __forceinline__ void bar_softmax() { asm volatile("bar.sync 1, 128;" ::: "memory"); }

// write 32 consecutive columns [c0, c0+32) of row r of a K-major SWIZZLE_128B bf16 block ([rows][64 cols], 128 B rows)
__device__ __forceinline__ void st_sw128_32(uint8_t* block, int r, int c0, const float (&f)[32]) {
  uint8_t* row = block + (r >> 3) * 1024 + (r & 7) * 128;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int ch = (c0 >> 3) + i;
    uint4 u;
    u.x = pack_bf16x2(f[i * 8 + 0], f[i * 8 + 1]); u.y = pack_bf16x2(f[i * 8 + 2], f[i * 8 + 3]);
    u.z = pack_bf16x2(f[i * 8 + 4], f[i * 8 + 5]); u.w = pack_bf16x2(f[i * 8 + 6], f[i * 8 + 7]);
    *reinterpret_cast<uint4*>(row + ((ch ^ (r & 7)) << 4)) = u;
  }
}

// D[tmem] (+)= A[128 x 64, K-major block] * B^T, B a K-major [n x 64] block  (contraction over the 64 columns)
__device__ __forceinline__ void mma_kk(uint32_t d_tmem, uint32_t a_addr, uint32_t b_addr, uint32_t idesc, bool accumulate) {
#pragma unroll
  for (int k = 0; k < 4; ++k)
    umma_bf16(d_tmem, make_smem_desc(a_addr + k * 32, 16, 1024, 2), make_smem_desc(b_addr + k * 32, 16, 1024, 2), idesc,
              (accumulate || k != 0) ? 1u : 0u);
}
// D[tmem] (+)= A[128 x 64, K-major block] * B, B an MN-major [64 rows (contraction) x 64 cols] block
__device__ __forceinline__ void mma_km(uint32_t d_tmem, uint32_t a_addr, uint32_t b_addr, uint32_t idesc, bool accumulate) {
#pragma unroll
  for (int k = 0; k < 4; ++k)
    umma_bf16(d_tmem, make_smem_desc(a_addr + k * 32, 16, 1024, 2), make_smem_desc(b_addr + k * 2048, 8192, 1024, 2), idesc,
              (accumulate || k != 0) ? 1u : 0u);
}

__device__ __forceinline__ void store_row64_bf16(__nv_bfloat16* dst, const float (&f)[64]) {
  uint4* q = reinterpret_cast<uint4*>(dst);
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    uint4 u;
    u.x = pack_bf16x2(f[i * 8 + 0], f[i * 8 + 1]); u.y = pack_bf16x2(f[i * 8 + 2], f[i * 8 + 3]);
    u.z = pack_bf16x2(f[i * 8 + 4], f[i * 8 + 5]); u.w = pack_bf16x2(f[i * 8 + 6], f[i * 8 + 7]);
    q[i] = u;
  }
}

// transpose of the rotary map on one 64-wide gradient row (pairs (e, e + 32); table row = 32 x (cos, sin) of the position):
// the forward pass rotated q / k in place, so their gradients leave the attention kernels through the inverse rotation
// -- in registers, on the fp32 values, instead of a separate in-place pass over the stored bf16 gradients.  The table row
// is requested BEFORE the wait for the last accumulator (one CTA per SM: an L2 round trip in the CTA's tail is exposed).
__device__ __forceinline__ void lc_rot_fetch(float4 (&t)[16], const float2* __restrict__ tb) {
#pragma unroll
  for (int e4 = 0; e4 < 16; ++e4)        // (cos, sin) of frequencies 2 e4, 2 e4 + 1; volatile: stays above the mbarrier wait
    asm volatile("ld.global.nc.v4.f32 {%0, %1, %2, %3}, [%4];"
                 : "=f"(t[e4].x), "=f"(t[e4].y), "=f"(t[e4].z), "=f"(t[e4].w)
                 : "l"(reinterpret_cast<const float4*>(tb) + e4));
}
__device__ __forceinline__ void lc_unrotate64(float (&g)[64], const float4 (&t)[16]) {
#pragma unroll
  for (int e4 = 0; e4 < 16; ++e4) {
    const int e = 2 * e4;
    const float a0 = g[e], b0 = g[e + 32], a1 = g[e + 1], b1 = g[e + 33];
    g[e] = a0 * t[e4].x + b0 * t[e4].y;      g[e + 32] = b0 * t[e4].x - a0 * t[e4].y;
    g[e + 1] = a1 * t[e4].z + b1 * t[e4].w;  g[e + 33] = b1 * t[e4].z - a1 * t[e4].w;
  }
}

// delta = sum_e dO[e] * O[e] of one row (global reads, 2 x 128 B)
__device__ __forceinline__ float row_delta(const __nv_bfloat16* o, const __nv_bfloat16* d_o) {
  const uint4* a = reinterpret_cast<const uint4*>(o);
  const uint4* b = reinterpret_cast<const uint4*>(d_o);
  float acc = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const uint4 x = __ldg(a + i), y = __ldg(b + i);
    acc = fmaf(bf16lo(x.x), bf16lo(y.x), acc); acc = fmaf(bf16hi(x.x), bf16hi(y.x), acc);
    acc = fmaf(bf16lo(x.y), bf16lo(y.y), acc); acc = fmaf(bf16hi(x.y), bf16hi(y.y), acc);
    acc = fmaf(bf16lo(x.z), bf16lo(y.z), acc); acc = fmaf(bf16hi(x.z), bf16hi(y.z), acc);
    acc = fmaf(bf16lo(x.w), bf16lo(y.w), acc); acc = fmaf(bf16hi(x.w), bf16hi(y.w), acc);
  }
  return acc;
}

// ------------------------------------------------------------------------------------------------ forward
// O accumulates in TMEM (P V with accumulate) instead of in registers, so the softmax warps of tile t + 1 no longer wait
// for the P V product of tile t: they read the S row once (64 registers), keep a running maximum that is only RAISED
// when a tile's maximum exceeds it by more than 2^8 (the probabilities stay exact: the final O / l division uses the same
// reference maximum; P <= 256 is harmless in bf16), and in that rare case rescale their O lanes in TMEM
// (tcgen05.ld / st, warp-uniform).  P is double-buffered in shared memory; the tensor pipe runs one tile behind.
constexpr float L_RESCALE_THRESHOLD = 8.0f;       // log2 units

// TS: the probabilities / score gradients reach the second product as TMEM-resident A operands (tcgen05.st) instead of
// through 128B-swizzled shared memory
template <bool TS>
__global__ void __launch_bounds__(L_THREADS, 2)
tc_local_fwd2_kernel(const __grid_constant__ LcParams P) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ uint64_t q_full, s_full[2], p_full[2], o_done, o_final;
  constexpr int NST = TS ? L_STAGES_FWD + 2 : L_STAGES_FWD;    // TS: the two P buffers' shared memory holds K/V stages
  __shared__ uint64_t kv_full[NST], kv_empty[NST];
  __shared__ uint32_t tmem_base_slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* Qs = smem;                       // 16 KB
  uint8_t* Ps = Qs + 16384;                 // 2 x 16 KB: P of tile t in buffer t & 1
  uint8_t* KV = TS ? Ps : Ps + 2 * 16384;   // stages x (K 8 KB | V 8 KB)
  const int bh = blockIdx.y, b = bh / P.H, h = bh % P.H;
  const int i0 = blockIdx.x * 128;
  const int j_beg = (lc_lo(i0, P.W) / 64) * 64;
  const int j_last = min(P.N - 1, i0 + 127);
  const int ntiles = (j_last - j_beg) / 64 + 1;

  if (threadIdx.x == 0) {
    mbar_init(&q_full, 1); mbar_init(&o_done, 1); mbar_init(&o_final, 1);
    for (int i = 0; i < 2; ++i) { mbar_init(&s_full[i], 1); mbar_init(&p_full[i], 128); }
    for (int i = 0; i < NST; ++i) { mbar_init(&kv_full[i], 1); mbar_init(&kv_empty[i], 1); }
    fence_mbar_init();
    fence_proxy_async();
  }
  if (warp == 4) { tmem_alloc(&tmem_base_slot, 256); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_slot;
  const uint32_t tS0 = tmem_base, tO = tmem_base + 128;     // S buffers at columns 0 / 64, O at 128

  if (warp == 5) {
    if (lane == 0) {
      prefetch_tmap(&P.qmap); prefetch_tmap(&P.kmap); prefetch_tmap(&P.vmap);
      mbar_expect_tx(&q_full, 16384);
      tma_load_2d(Qs, &P.qmap, &q_full, h * 64, b * P.N + i0);
      int stage = 0; uint32_t phase = 0;
      for (int t = 0; t < ntiles; ++t) {
        mbar_wait(&kv_empty[stage], phase ^ 1);
        mbar_expect_tx(&kv_full[stage], 16384);
        uint8_t* ks = KV + stage * 16384;
        tma_load_2d(ks, &P.kmap, &kv_full[stage], h * 64, b * P.N + j_beg + t * 64);
        tma_load_2d(ks + 8192, &P.vmap, &kv_full[stage], h * 64, b * P.N + j_beg + t * 64);
        if (++stage == NST) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 4) {
    if (lane == 0) {
      const uint32_t idesc_s = make_idesc_bf16(128, 64, 0, 0);
      const uint32_t idesc_o = make_idesc_bf16(128, 64, 0, 1);
      const uint32_t qa = smem_u32(Qs), pa = smem_u32(Ps), kva = smem_u32(KV);
      mbar_wait(&q_full, 0);
      mbar_wait(&kv_full[0], 0);
      tc_fence_after();
      mma_kk(tS0, qa, kva, idesc_s, false);
      umma_commit(&s_full[0]);
      int stage = 0; uint32_t phase = 0;
      for (int t = 0; t < ntiles; ++t) {
        int nstage = stage + 1; uint32_t nphase = phase;
        if (nstage == NST) { nstage = 0; nphase ^= 1; }
        if (t + 1 < ntiles) {
          // S of the next tile.  Its buffer was last read for tile t - 1 (p_full waited below, one iteration ago); its
          // commit also covers P V of tile t - 1, which is what frees P buffer (t + 1) & 1 for the softmax warps.
          mbar_wait(&kv_full[nstage], nphase);
          tc_fence_after();
          mma_kk(tS0 + (uint32_t)(((t + 1) & 1) * 64), qa, kva + nstage * 16384, idesc_s, false);
          umma_commit(&s_full[(t + 1) & 1]);
        }
        mbar_wait(&p_full[t & 1], (uint32_t)((t >> 1) & 1));    // P_t is in shared memory, O lanes rescaled if needed
        tc_fence_after();
        if (TS) {
          // P as the TMEM-resident A operand (TS form): 64 keys = 32 columns of packed bf16 pairs, 8 columns per K step
          const uint32_t tp = tmem_base + 192 + (uint32_t)((t & 1) * 32);
          const uint32_t vb_addr = kva + stage * 16384 + 8192;
#pragma unroll
          for (int kk = 0; kk < 4; ++kk)
            umma_bf16_ts(tO, tp + (uint32_t)(kk * 8), make_smem_desc(vb_addr + kk * 2048, 8192, 1024, 2), idesc_o,
                         (t > 0 || kk != 0) ? 1u : 0u);
        } else {
          mma_km(tO, pa + (uint32_t)((t & 1) * 16384), kva + stage * 16384 + 8192, idesc_o, t > 0);
        }
        umma_commit(&o_done);
        umma_commit(&kv_empty[stage]);
        if (t == ntiles - 1) umma_commit(&o_final);
        stage = nstage; phase = nphase;
      }
    }
  } else {
    // ---------------------------------------------------------------- softmax warps: thread = query row
    const int r = warp * 32 + lane;
    const int p = i0 + r;
    const int lo = lc_lo(p, P.W);
    const uint32_t lane_addr = (uint32_t)(warp * 32) << 16;
    const float c2 = P.scale * LOG2E;
    const int w_p0 = i0 + warp * 32;                      // first row of this warp; lo() is non-decreasing in the row
    const int w_lo_max = lc_lo(w_p0 + 31, P.W);
    float m = -INFINITY, l = 0.f;                         // m: the reference maximum the stored probabilities use
    for (int t = 0; t < ntiles; ++t) {
      const int j0 = j_beg + t * 64;
      mbar_wait(&s_full[t & 1], (uint32_t)((t >> 1) & 1));
      tc_fence_after();
      const uint32_t ts = tS0 + lane_addr + (uint32_t)((t & 1) * 64);
      uint32_t va[32], vb[32];
      tmem_ld_32x32(ts, va);
      tmem_ld_32x32(ts + 32, vb);
      tmem_ld_wait();
      // interior tiles (every key of the tile visible to every row of this warp: a warp-uniform test, so no divergence)
      // skip the per-score mask arithmetic
      const bool full = P.fast && (j0 + 63 <= w_p0) && (j0 >= w_lo_max) && (w_p0 + 31 < P.N);
      if (!full) {                                   // boundary tile: hide the masked scores, then the same arithmetic
        uint32_t mlo, mhi;
        lc_colmask(lo - j0, (p < P.N) ? p - j0 : -1, mlo, mhi);
        lc_apply_mask(va, mlo);
        lc_apply_mask(vb, mhi);
      }
      float mt;
      {
        float m4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};        // independent chains: the warp is latency-bound
#pragma unroll
        for (int c = 0; c < 32; ++c) m4[c & 3] = fmaxf(m4[c & 3], fmaxf(__uint_as_float(va[c]), __uint_as_float(vb[c])));
        mt = fmaxf(fmaxf(m4[0], m4[1]), fmaxf(m4[2], m4[3])) * c2;
      }
      // raise the reference maximum only when it pays (warp-uniform: tcgen05.ld / st are warp-collective)
      if (__any_sync(0xffffffffu, mt > m + L_RESCALE_THRESHOLD)) {
        const float m_new = fmaxf(m, mt);
        const float alpha = (m == -INFINITY) ? 1.0f : ex2(m - m_new);     // m = -inf: nothing accumulated yet (O lane is 0)
        if (t > 0) {
          mbar_wait(&o_done, (uint32_t)((t - 1) & 1));                   // P V of tile t - 1 has landed in O
          tc_fence_after();
#pragma unroll
          for (int q4 = 0; q4 < 4; ++q4) {
            uint32_t v[16];
            tmem_ld_32x16(tO + lane_addr + (uint32_t)(q4 * 16), v);
            tmem_ld_wait();
#pragma unroll
            for (int e = 0; e < 16; ++e) v[e] = __float_as_uint(__uint_as_float(v[e]) * alpha);
            tmem_st_32x16(tO + lane_addr + (uint32_t)(q4 * 16), v);
          }
          tmem_st_wait();
        }
        l *= alpha;
        m = m_new;
      }
      const float m_use = (m == -INFINITY) ? 0.f : m;
      float ps4[4] = {0.f, 0.f, 0.f, 0.f};
      uint8_t* pbuf = Ps + (t & 1) * 16384;
      if (TS) {
        const uint32_t tp = tmem_base + 192 + (uint32_t)((t & 1) * 32) + lane_addr;
        uint32_t u[16];
#pragma unroll
        for (int c = 0; c < 32; c += 2) {
          const float f0 = ex2(fmaf(__uint_as_float(va[c]), c2, -m_use)), f1 = ex2(fmaf(__uint_as_float(va[c + 1]), c2, -m_use));
          ps4[c & 3] += f0; ps4[(c + 1) & 3] += f1;
          u[c >> 1] = pack_bf16x2(f0, f1);
        }
        tmem_st_32x16(tp, u);
#pragma unroll
        for (int c = 0; c < 32; c += 2) {
          const float f0 = ex2(fmaf(__uint_as_float(vb[c]), c2, -m_use)), f1 = ex2(fmaf(__uint_as_float(vb[c + 1]), c2, -m_use));
          ps4[c & 3] += f0; ps4[(c + 1) & 3] += f1;
          u[c >> 1] = pack_bf16x2(f0, f1);
        }
        tmem_st_32x16(tp + 16, u);
        tmem_st_wait();
      } else {
        float f[32];
#pragma unroll
        for (int c = 0; c < 32; ++c) { f[c] = ex2(fmaf(__uint_as_float(va[c]), c2, -m_use)); ps4[c & 3] += f[c]; }
        st_sw128_32(pbuf, r, 0, f);
#pragma unroll
        for (int c = 0; c < 32; ++c) { f[c] = ex2(fmaf(__uint_as_float(vb[c]), c2, -m_use)); ps4[c & 3] += f[c]; }
        st_sw128_32(pbuf, r, 32, f);
        fence_proxy_async();
      }
      tc_fence_before();
      mbar_arrive(&p_full[t & 1]);
      l += (ps4[0] + ps4[1]) + (ps4[2] + ps4[3]);
    }
    // (a parity wait names a phase only relative to the current one, and the softmax warps may be two P V products ahead
    //  of the tensor pipe here: the last product has its own barrier)
    mbar_wait(&o_final, 0);
    tc_fence_after();
    float o[64];
    {
      uint32_t v[32];
      tmem_ld_32x32(tO + lane_addr, v);
      tmem_ld_wait();
#pragma unroll
      for (int e = 0; e < 32; ++e) o[e] = __uint_as_float(v[e]);
      tmem_ld_32x32(tO + lane_addr + 32, v);
      tmem_ld_wait();
#pragma unroll
      for (int e = 0; e < 32; ++e) o[32 + e] = __uint_as_float(v[e]);
    }
    if (p < P.N) {
      const float inv = 1.0f / l;
#pragma unroll
      for (int e = 0; e < 64; ++e) o[e] *= inv;
      store_row64_bf16(P.o_out + ((long long)b * P.N + p) * P.out_ld + h * 64, o);
      P.lse[(long long)bh * P.N + p] = (m + log2f(l)) * 0.6931471805599453f;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 4) tmem_dealloc(tmem_base, 256);
}

// ------------------------------------------------------------------------------------------------ backward: dq
// TS: the probabilities / score gradients reach the second product as TMEM-resident A operands (tcgen05.st) instead of
// through 128B-swizzled shared memory
template <bool TS>
__global__ void __launch_bounds__(L_THREADS, 2)
tc_local_bwd_dq_kernel(const __grid_constant__ LcParams P) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ uint64_t q_full, sdp_full, sdp_free, ds_full[2], dq_full;
  constexpr int NST = TS ? L_STAGES_DQ + 2 : L_STAGES_DQ;      // TS: the two dS buffers' shared memory holds K/V stages
  __shared__ uint64_t kv_full[NST], kv_empty[NST];
  __shared__ uint32_t tmem_base_slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* Qs = smem;                       // 16 KB
  uint8_t* dOs = Qs + 16384;                // 16 KB
  uint8_t* dSs = dOs + 16384;               // 2 x 16 KB  [128 q x 64 keys], tile t in buffer t & 1
  uint8_t* KV = TS ? dSs : dSs + 2 * 16384; // stages x (K 8 KB | V 8 KB)
  const int bh = blockIdx.y, b = bh / P.H, h = bh % P.H;
  const int i0 = blockIdx.x * 128;
  const int j_beg = (lc_lo(i0, P.W) / 64) * 64;
  const int j_last = min(P.N - 1, i0 + 127);
  const int ntiles = (j_last - j_beg) / 64 + 1;

  if (threadIdx.x == 0) {
    mbar_init(&q_full, 1); mbar_init(&sdp_full, 1); mbar_init(&sdp_free, 128); mbar_init(&ds_full[0], 128); mbar_init(&ds_full[1], 128); mbar_init(&dq_full, 1);
    for (int i = 0; i < NST; ++i) { mbar_init(&kv_full[i], 1); mbar_init(&kv_empty[i], 1); }
    fence_mbar_init();
    fence_proxy_async();
  }
  if (warp == 4) { tmem_alloc(&tmem_base_slot, 256); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_slot;
  const uint32_t tS = tmem_base, tdP = tmem_base + 64, tdQ = tmem_base + 128;

  if (warp == 5) {
    if (lane == 0) {
      prefetch_tmap(&P.qmap); prefetch_tmap(&P.kmap); prefetch_tmap(&P.vmap); prefetch_tmap(&P.domap);
      mbar_expect_tx(&q_full, 32768);
      tma_load_2d(Qs, &P.qmap, &q_full, h * 64, b * P.N + i0);
      tma_load_2d(dOs, &P.domap, &q_full, h * 64, b * P.N + i0);
      int stage = 0; uint32_t phase = 0;
      for (int t = 0; t < ntiles; ++t) {
        mbar_wait(&kv_empty[stage], phase ^ 1);
        mbar_expect_tx(&kv_full[stage], 16384);
        uint8_t* ks = KV + stage * 16384;
        tma_load_2d(ks, &P.kmap, &kv_full[stage], h * 64, b * P.N + j_beg + t * 64);
        tma_load_2d(ks + 8192, &P.vmap, &kv_full[stage], h * 64, b * P.N + j_beg + t * 64);
        if (++stage == NST) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 4) {
    if (lane == 0) {
      const uint32_t idesc_kk = make_idesc_bf16(128, 64, 0, 0);
      const uint32_t idesc_km = make_idesc_bf16(128, 64, 0, 1);
      const uint32_t qa = smem_u32(Qs), doa = smem_u32(dOs), dsa = smem_u32(dSs), kva = smem_u32(KV);
      mbar_wait(&q_full, 0);
      mbar_wait(&kv_full[0], 0);
      tc_fence_after();
      mma_kk(tS, qa, kva, idesc_kk, false);                                // S = Q K^T
      mma_kk(tdP, doa, kva + 8192, idesc_kk, false);                       // dP = dO V^T
      umma_commit(&sdp_full);
      int stage = 0; uint32_t phase = 0;
      for (int t = 0; t < ntiles; ++t) {
        int nstage = stage + 1; uint32_t nphase = phase;
        if (nstage == NST) { nstage = 0; nphase ^= 1; }
        if (t + 1 < ntiles) {
          // S / dP of the next tile as soon as the softmax warps have READ this tile's (they still compute on registers)
          mbar_wait(&kv_full[nstage], nphase);
          mbar_wait(&sdp_free, (uint32_t)(t & 1));
          tc_fence_after();
          mma_kk(tS, qa, kva + nstage * 16384, idesc_kk, false);
          mma_kk(tdP, doa, kva + nstage * 16384 + 8192, idesc_kk, false);
          umma_commit(&sdp_full);
        }
        mbar_wait(&ds_full[t & 1], (uint32_t)((t >> 1) & 1));
        tc_fence_after();
        if (TS) {                                                          // dQ += dS K, dS the TMEM-resident A operand
          const uint32_t tds = tmem_base + 192 + (uint32_t)((t & 1) * 32);
#pragma unroll
          for (int kk = 0; kk < 4; ++kk)
            umma_bf16_ts(tdQ, tds + (uint32_t)(kk * 8), make_smem_desc(kva + stage * 16384 + kk * 2048, 8192, 1024, 2),
                         idesc_km, (t > 0 || kk != 0) ? 1u : 0u);
        } else {
          mma_km(tdQ, dsa + (uint32_t)((t & 1) * 16384), kva + stage * 16384, idesc_km, t > 0);   // dQ += dS K
        }
        umma_commit(&kv_empty[stage]);
        stage = nstage; phase = nphase;
      }
      umma_commit(&dq_full);
    }
  } else {
    const int r = warp * 32 + lane;
    const int p = i0 + r;
    const int lo = lc_lo(p, P.W);
    const bool row_ok = p < P.N;
    const uint32_t lane_addr = (uint32_t)(warp * 32) << 16;
    const float c2 = P.scale * LOG2E;
    float lse2 = 0.f, delta = 0.f;
    if (row_ok) {
      lse2 = P.lse[(long long)bh * P.N + p] * LOG2E;
      const long long ro = ((long long)b * P.N + p) * P.out_ld + h * 64;
      delta = row_delta(P.out + ro, P.dout + ro);
      if (P.delta_ws) P.delta_ws[(long long)bh * P.N + p] = delta;
    }
    const int w_p0 = i0 + warp * 32;
    const int w_lo_max = lc_lo(w_p0 + 31, P.W);
    for (int t = 0; t < ntiles; ++t) {
      const int j0 = j_beg + t * 64;
      mbar_wait(&sdp_full, (uint32_t)(t & 1));
      tc_fence_after();
      // interior tile for the whole warp (uniform): no mask arithmetic
      const bool full = P.fast && (j0 + 63 <= w_p0) && (j0 >= w_lo_max) && (w_p0 + 31 < P.N);
      uint32_t mlo = 0xffffffffu, mhi = 0xffffffffu;
      if (!full) lc_colmask(lo - j0, row_ok ? p - j0 : -1, mlo, mhi);
#pragma unroll
      for (int hh = 0; hh < 2; ++hh) {
        uint32_t vs[32], vd[32];
        tmem_ld_32x32(tS + lane_addr + (uint32_t)(hh * 32), vs);
        tmem_ld_32x32(tdP + lane_addr + (uint32_t)(hh * 32), vd);
        tmem_ld_wait();
        if (hh == 1) {                                      // S and dP of this tile are in registers: the next tile's may land
          tc_fence_before();
          mbar_arrive(&sdp_free);
        }
        float f[32];
        if (!full) lc_apply_mask(vs, hh ? mhi : mlo);       // boundary tile: hidden scores -> -inf -> probability 0
#pragma unroll
        for (int c = 0; c < 32; ++c)
          f[c] = ex2(fmaf(__uint_as_float(vs[c]), c2, -lse2)) * (__uint_as_float(vd[c]) - delta);   // d^-1/2: applied to dQ
        if (TS) {
          uint32_t u[16];
#pragma unroll
          for (int c = 0; c < 32; c += 2) u[c >> 1] = pack_bf16x2(f[c], f[c + 1]);
          tmem_st_32x16(tmem_base + 192 + (uint32_t)((t & 1) * 32 + hh * 16) + lane_addr, u);
        } else {
          st_sw128_32(dSs + (t & 1) * 16384, r, hh * 32, f);
        }
      }
      if (TS) tmem_st_wait();
      else fence_proxy_async();
      tc_fence_before();
      mbar_arrive(&ds_full[t & 1]);
    }
    float4 rt[16];
    if (P.rot && row_ok) lc_rot_fetch(rt, P.rot + (long long)p * 32);
    mbar_wait(&dq_full, 0);
    tc_fence_after();
    float g[64];
    {
      uint32_t v[32];
      tmem_ld_32x32(tdQ + lane_addr, v);
      tmem_ld_wait();
#pragma unroll
      for (int e = 0; e < 32; ++e) g[e] = __uint_as_float(v[e]);
      tmem_ld_32x32(tdQ + lane_addr + 32, v);
      tmem_ld_wait();
#pragma unroll
      for (int e = 0; e < 32; ++e) g[32 + e] = __uint_as_float(v[e]);
    }
    if (row_ok) {
#pragma unroll
      for (int e = 0; e < 64; ++e) g[e] *= P.scale;       // dS = P (dP - delta) d^-1/2: the factor commutes with the product
      if (P.rot) lc_unrotate64(g, rt);
      store_row64_bf16(P.dq + ((long long)b * P.N + p) * P.ld + h * 64, g);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 4) tmem_dealloc(tmem_base, 256);
}

// ------------------------------------------------------------------------------------------------ backward: dk, dv
// TS: the probabilities / score gradients reach the second product as TMEM-resident A operands (tcgen05.st) instead of
// through 128B-swizzled shared memory
template <bool TS>
__global__ void __launch_bounds__(L_THREADS, 2)
tc_local_bwd_dkv_kernel(const __grid_constant__ LcParams P) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ uint64_t kv_full, sdp_full, ds_full, acc_full;
  constexpr int NST = TS ? L_STAGES_BWD + 2 : L_STAGES_BWD;    // TS: the P^T / dS^T buffers' shared memory holds Q/dO stages
  __shared__ uint64_t qd_full[NST], qd_empty[NST];
  __shared__ uint32_t tmem_base_slot;
  __shared__ __align__(16) float s_lse2[2][64], s_delta[2][64];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* Ks = smem;                       // 16 KB [128 keys x 64]
  uint8_t* Vs = Ks + 16384;                 // 16 KB
  uint8_t* PTs = Vs + 16384;                // 16 KB [128 keys x 64 queries]
  uint8_t* dSTs = PTs + 16384;              // 16 KB
  uint8_t* QD = TS ? PTs : dSTs + 16384;    // stages x (Q 8 KB | dO 8 KB)
  const int bh = blockIdx.y, b = bh / P.H, h = bh % P.H;
  const int j0 = blockIdx.x * 128;
  const int j_hi = min(P.N - 1, j0 + 127);
  const int i_last = min(P.N - 1, (j_hi / P.W + 2) * P.W - 1);
  const int ntiles = (i_last - j0) / 64 + 1;         // query tiles of 64 starting at j0

  if (threadIdx.x == 0) {
    mbar_init(&kv_full, 1); mbar_init(&sdp_full, 1); mbar_init(&ds_full, 128); mbar_init(&acc_full, 1);
    for (int i = 0; i < NST; ++i) { mbar_init(&qd_full[i], 1); mbar_init(&qd_empty[i], 1); }
    fence_mbar_init();
    fence_proxy_async();
  }
  if (warp == 4) { tmem_alloc(&tmem_base_slot, 256); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_slot;
  const uint32_t tST = tmem_base, tdPT = tmem_base + 64, tdV = tmem_base + 128, tdK = tmem_base + 192;

  if (warp == 5) {
    if (lane == 0) {
      prefetch_tmap(&P.qmap); prefetch_tmap(&P.kmap); prefetch_tmap(&P.vmap); prefetch_tmap(&P.domap);
      mbar_expect_tx(&kv_full, 32768);
      tma_load_2d(Ks, &P.kmap, &kv_full, h * 64, b * P.N + j0);
      tma_load_2d(Vs, &P.vmap, &kv_full, h * 64, b * P.N + j0);
      int stage = 0; uint32_t phase = 0;
      for (int t = 0; t < ntiles; ++t) {
        mbar_wait(&qd_empty[stage], phase ^ 1);
        mbar_expect_tx(&qd_full[stage], 16384);
        uint8_t* qs = QD + stage * 16384;
        tma_load_2d(qs, &P.qmap, &qd_full[stage], h * 64, b * P.N + j0 + t * 64);
        tma_load_2d(qs + 8192, &P.domap, &qd_full[stage], h * 64, b * P.N + j0 + t * 64);
        if (++stage == NST) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 4) {
    if (lane == 0) {
      const uint32_t idesc_kk = make_idesc_bf16(128, 64, 0, 0);
      const uint32_t idesc_km = make_idesc_bf16(128, 64, 0, 1);
      const uint32_t ka = smem_u32(Ks), va = smem_u32(Vs), pta = smem_u32(PTs), dsta = smem_u32(dSTs), qda = smem_u32(QD);
      mbar_wait(&kv_full, 0);
      int stage = 0; uint32_t phase = 0;
      for (int t = 0; t < ntiles; ++t) {
        mbar_wait(&qd_full[stage], phase);
        tc_fence_after();
        mma_kk(tST, ka, qda + stage * 16384, idesc_kk, false);             // S^T = K Q^T
        mma_kk(tdPT, va, qda + stage * 16384 + 8192, idesc_kk, false);     // dP^T = V dO^T
        umma_commit(&sdp_full);
        mbar_wait(&ds_full, (uint32_t)(t & 1));
        tc_fence_after();
        if (TS) {
          // P^T / dS^T as TMEM-resident A operands, written in place over the first 32 columns of S^T / dP^T
#pragma unroll
          for (int kk = 0; kk < 4; ++kk)
            umma_bf16_ts(tdV, tST + (uint32_t)(kk * 8), make_smem_desc(qda + stage * 16384 + 8192 + kk * 2048, 8192, 1024, 2),
                         idesc_km, (t > 0 || kk != 0) ? 1u : 0u);
#pragma unroll
          for (int kk = 0; kk < 4; ++kk)
            umma_bf16_ts(tdK, tdPT + (uint32_t)(kk * 8), make_smem_desc(qda + stage * 16384 + kk * 2048, 8192, 1024, 2),
                         idesc_km, (t > 0 || kk != 0) ? 1u : 0u);
        } else {
          mma_km(tdV, pta, qda + stage * 16384 + 8192, idesc_km, t > 0);     // dV += P^T dO
          mma_km(tdK, dsta, qda + stage * 16384, idesc_km, t > 0);           // dK += dS^T Q
        }
        umma_commit(&qd_empty[stage]);
        if (++stage == NST) { stage = 0; phase ^= 1; }
      }
      umma_commit(&acc_full);
    }
  } else {
    const int r = warp * 32 + lane;          // key row
    const int j = j0 + r;
    // key j is seen by queries p with j <= p and floor(p / W) - 1 <= floor(j / W)  <=>  p <= (floor(j / W) + 2) W - 1
    // (one division per thread instead of one per score)
    const int p_hi = (j < P.N) ? min(P.N - 1, (j / P.W + 2) * P.W - 1) : -1;
    const uint32_t lane_addr = (uint32_t)(warp * 32) << 16;
    const float c2 = P.scale * LOG2E;
    // per-query statistics (log-sum-exp, delta = dO . O) of a query tile: threads 0..63, one query each.  With the
    // delta workspace (filled by the dq kernel) both are coalesced 4-byte loads and the next tile's pair is fetched
    // while this tile is processed; without it delta is recomputed from the two 128-byte rows.
    auto load_stats = [&](int tt, float& l2, float& dl) {
      l2 = 0.f; dl = 0.f;
      const int p = j0 + tt * 64 + r;
      if (r < 64 && tt < ntiles && p < P.N) {
        l2 = P.lse[(long long)bh * P.N + p] * LOG2E;
        if (P.delta_ws) {
          dl = P.delta_ws[(long long)bh * P.N + p];
        } else {
          const long long ro = ((long long)b * P.N + p) * P.out_ld + h * 64;
          dl = row_delta(P.out + ro, P.dout + ro);
        }
      }
    };
    const int w_j0 = j0 + warp * 32;                       // first key of this warp; p_hi is non-decreasing in the key
    const int w_p_hi_min = min(P.N - 1, (w_j0 / P.W + 2) * P.W - 1);
    float nl2, ndl;
    load_stats(0, nl2, ndl);
    for (int t = 0; t < ntiles; ++t) {
      const int q0 = j0 + t * 64;
      const int buf = t & 1;
      if (r < 64) { s_lse2[buf][r] = nl2; s_delta[buf][r] = ndl; }
      if (P.delta_ws) load_stats(t + 1, nl2, ndl);      // in flight during this tile
      bar_softmax();
      mbar_wait(&sdp_full, (uint32_t)(t & 1));
      tc_fence_after();
      // every (key of this warp, query of this tile) pair visible (warp-uniform): no mask arithmetic
      const bool full = P.fast && (w_j0 + 31 <= q0) && (q0 + 63 <= w_p_hi_min) && (w_j0 + 31 < P.N);
      uint32_t mlo = 0xffffffffu, mhi = 0xffffffffu;
      if (!full) lc_colmask(j - q0, p_hi - q0, mlo, mhi);          // queries p = q0 + c with j <= p <= p_hi
#pragma unroll
      for (int hh = 0; hh < 2; ++hh) {
        uint32_t vs[32], vd[32];
        tmem_ld_32x32(tST + lane_addr + (uint32_t)(hh * 32), vs);
        tmem_ld_32x32(tdPT + lane_addr + (uint32_t)(hh * 32), vd);
        tmem_ld_wait();
        float fp[32], fd[32];
        if (!full) lc_apply_mask(vs, hh ? mhi : mlo);       // boundary tile: hidden scores -> -inf -> probability 0
        // the tile's per-query statistics as 16-byte broadcast reads (one LDS.128 per four scores instead of two LDS.32
        // per score: the shared-memory pipe was 25 % busy with them)
        const float4* l4 = reinterpret_cast<const float4*>(&s_lse2[buf][hh * 32]);
        const float4* d4 = reinterpret_cast<const float4*>(&s_delta[buf][hh * 32]);
#pragma unroll
        for (int c4 = 0; c4 < 8; ++c4) {
          const float4 lq = l4[c4], dq4 = d4[c4];
          const float lv[4] = {lq.x, lq.y, lq.z, lq.w}, dv[4] = {dq4.x, dq4.y, dq4.z, dq4.w};
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const int c = c4 * 4 + e;
            const float pr = ex2(fmaf(__uint_as_float(vs[c]), c2, -lv[e]));
            fp[c] = pr;
            fd[c] = pr * (__uint_as_float(vd[c]) - dv[e]);                                       // d^-1/2: applied to dK
          }
        }
        if (TS) {
          uint32_t u[16];
#pragma unroll
          for (int c = 0; c < 32; c += 2) u[c >> 1] = pack_bf16x2(fp[c], fp[c + 1]);
          tmem_st_32x16(tST + lane_addr + (uint32_t)(hh * 16), u);
#pragma unroll
          for (int c = 0; c < 32; c += 2) u[c >> 1] = pack_bf16x2(fd[c], fd[c + 1]);
          tmem_st_32x16(tdPT + lane_addr + (uint32_t)(hh * 16), u);
        } else {
          st_sw128_32(PTs, r, hh * 32, fp);
          st_sw128_32(dSTs, r, hh * 32, fd);
        }
      }
      if (TS) tmem_st_wait();
      else fence_proxy_async();
      tc_fence_before();
      mbar_arrive(&ds_full);
      if (!P.delta_ws) load_stats(t + 1, nl2, ndl);
    }
    float4 rt[16];
    if (P.rot && j < P.N) lc_rot_fetch(rt, P.rot + (long long)j * 32);
    mbar_wait(&acc_full, 0);
    tc_fence_after();
    float g[64];
#pragma unroll
    for (int which = 0; which < 2; ++which) {
      const uint32_t ta = which ? tdK : tdV;
      uint32_t v[32];
      tmem_ld_32x32(ta + lane_addr, v);
      tmem_ld_wait();
#pragma unroll
      for (int e = 0; e < 32; ++e) g[e] = __uint_as_float(v[e]);
      tmem_ld_32x32(ta + lane_addr + 32, v);
      tmem_ld_wait();
#pragma unroll
      for (int e = 0; e < 32; ++e) g[32 + e] = __uint_as_float(v[e]);
      if (which) {
#pragma unroll
        for (int e = 0; e < 64; ++e) g[e] *= P.scale;
        if (P.rot && j < P.N) lc_unrotate64(g, rt);
      }
      if (j < P.N) store_row64_bf16((which ? P.dk : P.dv) + ((long long)b * P.N + j) * P.ld + h * 64, g);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 4) tmem_dealloc(tmem_base, 256);
}

// ------------------------------------------------------------------------------------------------ rotary (in place)
template <typename T>
__global__ void rotary_kernel(T* __restrict__ buf, long long ld, int B, int N, int H, int d, const float* __restrict__ inv_freq,
                              int inverse) {
  // one thread per (row, frequency): the sine / cosine pair is shared by all heads of the row
  const int half = d / 2;
  const long long total = (long long)B * N * half;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int dd = (int)(i % half);
    const long long row = i / half;
    const int n = (int)(row % N);
    float sn, cs;
    sincosf((float)n * inv_freq[dd], &sn, &cs);
    if (inverse) sn = -sn;
    for (int hh = 0; hh < H; ++hh) {
      const long long o = row * ld + hh * d + dd;
      const float x1 = sa_ld(buf, o), x2 = sa_ld(buf, o + half);
      sa_st(buf, o, x1 * cs - x2 * sn);
      sa_st(buf, o + half, x2 * cs + x1 * sn);
    }
  }
}

// q and k blocks in one launch: one thread per (row, frequency); the sine / cosine pair is shared by both blocks and all heads
template <typename T>
__global__ void rotary_qk_kernel(T* __restrict__ buf, long long ld, long long k_offset, int B, int N, int H, int d,
                                 const float* __restrict__ inv_freq, int inverse) {
  const int half = d / 2;
  const long long total = (long long)B * N * half;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int dd = (int)(i % half);
    const long long row = i / half;
    const int n = (int)(row % N);
    float sn, cs;
    sincosf((float)n * inv_freq[dd], &sn, &cs);
    if (inverse) sn = -sn;
#pragma unroll 2
    for (int blk = 0; blk < 2; ++blk) {
      const long long o0 = row * ld + (blk ? k_offset : 0) + dd;
      for (int hh = 0; hh < H; ++hh) {
        const long long o = o0 + hh * d;
        const float x1 = sa_ld(buf, o), x2 = sa_ld(buf, o + half);
        sa_st(buf, o, x1 * cs - x2 * sn);
        sa_st(buf, o + half, x2 * cs + x1 * sn);
      }
    }
  }
}

// bf16, d = 64: 8 rows per CTA; the 8 x 32 sine / cosine pairs go through shared memory, the data moves in 16-byte vectors
__global__ void __launch_bounds__(256)
rotary_qk_vec_kernel(__nv_bfloat16* __restrict__ buf, long long ld, long long k_offset, long long rows, int N, int H,
                     const float* __restrict__ inv_freq, int inverse) {
  __shared__ float s_cs[8][32], s_sn[8][32];
  const long long row0 = (long long)blockIdx.x * 8;
  {
    const int rr = threadIdx.x >> 5, dd = threadIdx.x & 31;
    const long long row = row0 + rr;
    float sn = 0.f, cs = 1.f;
    if (row < rows) sincosf((float)(int)(row % N) * inv_freq[dd], &sn, &cs);
    s_cs[rr][dd] = cs; s_sn[rr][dd] = inverse ? -sn : sn;
  }
  __syncthreads();
  const int per_row = 2 * H * 4;                 // vector items per row: (block, head, 8-frequency group)
  for (int e = threadIdx.x; e < 8 * per_row; e += 256) {
    const int rr = e / per_row, it = e - rr * per_row;
    const long long row = row0 + rr;
    if (row >= rows) continue;
    const int g8 = it & 3, hh = (it >> 2) % H, blk = it / (4 * H);
    __nv_bfloat16* p1 = buf + row * ld + (blk ? k_offset : 0) + hh * 64 + g8 * 8;
    uint4 u1 = *reinterpret_cast<const uint4*>(p1), u2 = *reinterpret_cast<const uint4*>(p1 + 32);
    uint32_t* a = reinterpret_cast<uint32_t*>(&u1);
    uint32_t* b = reinterpret_cast<uint32_t*>(&u2);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float c0 = s_cs[rr][g8 * 8 + 2 * j], c1 = s_cs[rr][g8 * 8 + 2 * j + 1];
      const float s0 = s_sn[rr][g8 * 8 + 2 * j], s1 = s_sn[rr][g8 * 8 + 2 * j + 1];
      const float x1l = satc::bf16lo(a[j]), x1h = satc::bf16hi(a[j]), x2l = satc::bf16lo(b[j]), x2h = satc::bf16hi(b[j]);
      a[j] = satc::pack_bf16x2(x1l * c0 - x2l * s0, x1h * c1 - x2h * s1);
      b[j] = satc::pack_bf16x2(x2l * c0 + x1l * s0, x2h * c1 + x1h * s1);
    }
    *reinterpret_cast<uint4*>(p1) = u1;
    *reinterpret_cast<uint4*>(p1 + 32) = u2;
  }
}

// (cos, sin)(n * inv_freq[dd]) with the same sincosf the rotary kernels evaluate
__global__ void rotary_table_kernel(const float* __restrict__ inv_freq, int N, int half, float2* __restrict__ table) {
  const long long total = (long long)N * half;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int dd = (int)(i % half);
    const int n = (int)(i / half);
    float sn, cs;
    sincosf((float)n * inv_freq[dd], &sn, &cs);
    table[i] = make_float2(cs, sn);
  }
}

std::once_flag g_once;
bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

constexpr size_t SMEM_FWD2 = 16384 * 3 + L_STAGES_FWD * 16384 + 1024;
constexpr size_t SMEM_DQ = 16384 * 4 + L_STAGES_DQ * 16384 + 1024;
constexpr size_t SMEM_DKV = 16384 * 4 + L_STAGES_BWD * 16384 + 1024;

void init_once() {
  std::call_once(g_once, [] {
    cudaFuncSetAttribute(tc_local_fwd2_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_FWD2);
    cudaFuncSetAttribute(tc_local_fwd2_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_FWD2);
    cudaFuncSetAttribute(tc_local_bwd_dq_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_DQ);
    cudaFuncSetAttribute(tc_local_bwd_dq_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_DQ);
    cudaFuncSetAttribute(tc_local_bwd_dkv_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_DKV);
    cudaFuncSetAttribute(tc_local_bwd_dkv_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_DKV);
  });
}

// A/B switch: SA_LOCAL_TS=0 stages P / dS through shared memory (the older SS form)
bool local_ts() { const char* env = getenv("SA_LOCAL_TS"); return !(env && env[0] == '0'); }

int make_map(CUtensorMap* m, const void* base, const sa_local_desc* d, int ld, uint32_t box_rows) {
  const uint64_t dims[2] = {(uint64_t)d->heads * 64, (uint64_t)d->batch * d->seq};
  const uint64_t strides[2] = {2, (uint64_t)ld * 2};
  const uint32_t box[2] = {64, box_rows};
  return sa_make_tmap_bf16(m, base, 2, dims, strides, box);
}

void fill_common(LcParams& P, const sa_local_desc* d) {
  P.B = d->batch; P.N = d->seq; P.H = d->heads; P.W = d->window; P.ld = d->ld; P.out_ld = d->out_ld;
  P.scale = 1.0f / sqrtf((float)d->dim_head);
  { const char* env = getenv("SA_LOCAL_FASTMASK"); P.fast = (env && env[0] == '0') ? 0 : 1; }
  P.out = nullptr; P.dout = nullptr; P.o_out = nullptr; P.lse = nullptr; P.dq = P.dk = P.dv = nullptr;
  P.delta_ws = nullptr; P.rot = nullptr;
}

}  // namespace

bool sa_tc_local_supported(const sa_local_desc* d, const void* q, const void* k, const void* v, const float* inv_freq) {
  if (d->act_dtype != SA_BF16 || d->dim_head != 64 || inv_freq != nullptr) return false;
  if ((d->ld & 7) || (d->out_ld & 7) || !aligned16(q) || !aligned16(k) || !aligned16(v)) return false;
  if ((long long)d->batch * d->heads > 65535 || d->window < 1) return false;
  return sa_get_tmap_encode() != nullptr;
}

int sa_tc_local_attn_fwd(const sa_local_desc* d, const void* q, const void* k, const void* v, void* out, float* lse,
                         cudaStream_t st) {
  init_once();
  sa_note_path(SA_PATH_TCGEN05);
  if (!aligned16(out)) { sa_set_error("tc_local_fwd: output not 16-byte aligned"); return SA_ERR_INVALID; }
  static thread_local LcParams P;
  fill_common(P, d);
  int rc;
  if ((rc = make_map(&P.qmap, q, d, d->ld, 128)) != SA_OK) return rc;
  if ((rc = make_map(&P.kmap, k, d, d->ld, 64)) != SA_OK) return rc;
  if ((rc = make_map(&P.vmap, v, d, d->ld, 64)) != SA_OK) return rc;
  P.o_out = (__nv_bfloat16*)out; P.lse = lse;
  dim3 grid((unsigned)sa_cdiv(d->seq, 128), (unsigned)(d->batch * d->heads));
  if (local_ts()) tc_local_fwd2_kernel<true><<<grid, L_THREADS, SMEM_FWD2, st>>>(P);
  else tc_local_fwd2_kernel<false><<<grid, L_THREADS, SMEM_FWD2, st>>>(P);
  SA_LAUNCH_CHECK();
  return SA_OK;
}

int sa_rotary_table_launch(const float* inv_freq, int seq, int half, float* table, cudaStream_t st) {
  long long blocks = sa_cdiv((long long)seq * half, 256);
  if (blocks > 148 * 8) blocks = 148 * 8;
  rotary_table_kernel<<<(unsigned)blocks, 256, 0, st>>>(inv_freq, seq, half, reinterpret_cast<float2*>(table));
  SA_LAUNCH_CHECK();
  return SA_OK;
}

int sa_tc_local_attn_bwd(const sa_local_desc* d, const void* q, const void* k, const void* v, const void* out,
                         const void* dout, const float* lse, void* dq, void* dk, void* dv, float* delta_ws,
                         const float* rot_table, cudaStream_t st) {
  init_once();
  sa_note_path(SA_PATH_TCGEN05);
  if (!aligned16(out) || !aligned16(dout) || !aligned16(dq) || !aligned16(dk) || !aligned16(dv)) {
    sa_set_error("tc_local_bwd: pointers not 16-byte aligned");
    return SA_ERR_INVALID;
  }
  static thread_local LcParams P;
  fill_common(P, d);
  P.out = (const __nv_bfloat16*)out; P.dout = (const __nv_bfloat16*)dout; P.lse = const_cast<float*>(lse);
  P.dq = (__nv_bfloat16*)dq; P.dk = (__nv_bfloat16*)dk; P.dv = (__nv_bfloat16*)dv;
  P.delta_ws = delta_ws;
  P.rot = reinterpret_cast<const float2*>(rot_table);
  int rc;
  // dq kernel: 128-row Q / dO boxes, 64-row K / V boxes
  if ((rc = make_map(&P.qmap, q, d, d->ld, 128)) != SA_OK) return rc;
  if ((rc = make_map(&P.domap, dout, d, d->out_ld, 128)) != SA_OK) return rc;
  if ((rc = make_map(&P.kmap, k, d, d->ld, 64)) != SA_OK) return rc;
  if ((rc = make_map(&P.vmap, v, d, d->ld, 64)) != SA_OK) return rc;
  dim3 grid((unsigned)sa_cdiv(d->seq, 128), (unsigned)(d->batch * d->heads));
  if (local_ts()) tc_local_bwd_dq_kernel<true><<<grid, L_THREADS, SMEM_DQ, st>>>(P);
  else tc_local_bwd_dq_kernel<false><<<grid, L_THREADS, SMEM_DQ, st>>>(P);
  SA_LAUNCH_CHECK();
  // dkv kernel: 128-row K / V boxes, 64-row Q / dO boxes
  if ((rc = make_map(&P.qmap, q, d, d->ld, 64)) != SA_OK) return rc;
  if ((rc = make_map(&P.domap, dout, d, d->out_ld, 64)) != SA_OK) return rc;
  if ((rc = make_map(&P.kmap, k, d, d->ld, 128)) != SA_OK) return rc;
  if ((rc = make_map(&P.vmap, v, d, d->ld, 128)) != SA_OK) return rc;
  if (local_ts()) tc_local_bwd_dkv_kernel<true><<<grid, L_THREADS, SMEM_DKV, st>>>(P);
  else tc_local_bwd_dkv_kernel<false><<<grid, L_THREADS, SMEM_DKV, st>>>(P);
  SA_LAUNCH_CHECK();
  return SA_OK;
}

int sa_rotary_launch(void* buf, int dtype, int64_t ld, int batch, int seq, int heads, int dim_head, const float* inv_freq,
                     int inverse, cudaStream_t st) {
  const long long total = (long long)batch * seq * (dim_head / 2);
  long long blocks = sa_cdiv(total, 256);
  if (blocks > 148 * 16) blocks = 148 * 16;
  if (dtype == SA_BF16)
    rotary_kernel<__nv_bfloat16><<<(unsigned)blocks, 256, 0, st>>>((__nv_bfloat16*)buf, ld, batch, seq, heads, dim_head,
                                                                   inv_freq, inverse);
  else
    rotary_kernel<float><<<(unsigned)blocks, 256, 0, st>>>((float*)buf, ld, batch, seq, heads, dim_head, inv_freq, inverse);
  SA_LAUNCH_CHECK();
  return SA_OK;
}

int sa_rotary_qk_launch(void* buf, int dtype, int64_t ld, int64_t k_offset, int batch, int seq, int heads, int dim_head,
                        const float* inv_freq, int inverse, cudaStream_t st) {
  const long long total = (long long)batch * seq * (dim_head / 2);
  long long blocks = sa_cdiv(total, 256);
  if (blocks > 148 * 32) blocks = 148 * 32;
  const bool vec_ok = dtype == SA_BF16 && dim_head == 64 && (ld % 8) == 0 && (k_offset % 8) == 0 &&
                      (reinterpret_cast<uintptr_t>(buf) & 15) == 0 && heads <= 64;
  if (vec_ok) {
    const long long rows = (long long)batch * seq;
    rotary_qk_vec_kernel<<<(unsigned)sa_cdiv(rows, 8), 256, 0, st>>>((__nv_bfloat16*)buf, ld, k_offset, rows, seq, heads,
                                                                     inv_freq, inverse);
  } else if (dtype == SA_BF16)
    rotary_qk_kernel<__nv_bfloat16><<<(unsigned)blocks, 256, 0, st>>>((__nv_bfloat16*)buf, ld, k_offset, batch, seq, heads,
                                                                      dim_head, inv_freq, inverse);
  else
    rotary_qk_kernel<float><<<(unsigned)blocks, 256, 0, st>>>((float*)buf, ld, k_offset, batch, seq, heads, dim_head, inv_freq,
                                                              inverse);
  SA_LAUNCH_CHECK();
  return SA_OK;
}
