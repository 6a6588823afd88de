// tcgen05 FAVOR+ (global heads) for sm_100a: bf16 operands, fp32 accumulation in TMEM.
//
// Replaces performer_pytorch.softmax_kernel / causal_linear_attention (+ autograd) and fast_transformers'
// CausalDotProduct, reached from /root/reference/src/networks/transformers/performer.py:270.  Same entry points and
// tensor layouts as the CUDA-core path in sa_pf_favor_simt.cu (which stays the fp32 parity path).
//
// Every kernel works on one (batch, head, 128-token chunk).  A [128 rows x 64 cols] bf16 block with 128-byte rows and
// the 128B swizzle is the unit of shared memory: the same bytes serve as a K-major operand (contraction over the 64
// columns) and as an MN-major operand (contraction over the rows), so one TMA load / one register re-staging feeds
// both kinds of product.  Feature tensors [bh][n][mp] are read as ceil(mp / 64) such blocks (TMA zero-fills past mp).
//
//   featmap_fwd   D = x P^T (MMA)  ->  r (exp(c D - c^2|x|^2/2 - stab) + eps)            (also the key-max pre-pass)
//   chunk_state   Z^T[e'][f] = sum_tok W_aug[tok][e'] F[tok][f]   W_aug = [v | 1]  or  [dout/den | -delta/den]
//   prefix        exclusive prefix (suffix) over chunks, fp32 -> bf16 states [80 rows][mp]; row 64 = the "1" column
//   scan<0>       A = q' k'^T -> tril -> bf16;  O = q' S + tril(A) v;   out = O / den
//   scan<1>       A^T = k' q'^T -> triu, * 1/den_i -> bf16;  dv = k' R + A^T dout
//   dqk<0>        B = dout v^T - delta -> tril;  dq' = (dout S - delta ksum + B k') / den
//   dqk<1>        B^T = (v dout^T - delta) / den -> triu;  dk' = v R + Rden + B^T q'
//   featmap_bwd   dD = dfeat (feat - r eps) (- row sum at the arg-max for queries);  dx = c dD P - c^2 s x
//   dqk_fb<0/1>   dqk<0/1> with featmap_bwd as its epilogue (dq' / dk' stay on chip): the training path's backward
//
// Warp roles (320 threads): warps 0-7 epilogue (TMEM lane quadrant = warp & 3, column half = warp >> 2),
// warp 8 MMA issuer, warp 9 TMA producer.
#include <mutex>

#include "sa_pf_common.cuh"
#include "sa_tc_common.cuh"

using namespace satc;

namespace {

constexpr int F_THREADS = 320;
constexpr int FC = 128;            // tokens per chunk
constexpr int F_MAXBLK = 5;        // feature blocks of 64 (mp <= 320)
constexpr int ST_ROWS = 80;        // rows of a bf16 state tile: 64 value columns, the "1" column, zero padding
constexpr int SUM_ROWS = 65;       // rows of the fp32 chunk sums
constexpr uint32_t BLK = 16384;    // [128 x 64] bf16 block
constexpr uint32_t ST_BLK = ST_ROWS * 128;   // [80 x 64] bf16 block

__device__ __forceinline__ unsigned int f2ord(float f) {
  const unsigned int b = __float_as_uint(f);
  return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ float ord2f(unsigned int o) {
  return __uint_as_float((o & 0x80000000u) ? (o ^ 0x80000000u) : ~o);
}
__device__ __forceinline__ float ex2f(float x) {
  float y;
  asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ void bar_epi() { asm volatile("bar.sync 1, 256;" ::: "memory"); }

__device__ __forceinline__ void tma_load_3d(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(
          smem_u32(smem_dst)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}

// shared -> global tensor store (bulk async group); rows / columns outside the tensor are clipped by the TMA unit
__device__ __forceinline__ void tma_store_3d(const CUtensorMap* m, const void* smem_src, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"(
                   reinterpret_cast<uint64_t>(m)),
               "r"(smem_u32(smem_src)), "r"(c0), "r"(c1), "r"(c2)
               : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

__device__ __forceinline__ uint8_t* sw_row(uint8_t* block, int r) { return block + (r >> 3) * 1024 + (r & 7) * 128; }

// write 32 / 16 consecutive columns starting at c0 (multiple of 16) of row r of a 128B-swizzled [rows][64] bf16 block
__device__ __forceinline__ void st_sw_32(uint8_t* block, int r, int c0, const float (&f)[32]) {
  uint8_t* row = sw_row(block, r);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int ch = (c0 >> 3) + i;
    uint4 u;
    u.x = pack_bf16x2(f[i * 8 + 0], f[i * 8 + 1]); u.y = pack_bf16x2(f[i * 8 + 2], f[i * 8 + 3]);
    u.z = pack_bf16x2(f[i * 8 + 4], f[i * 8 + 5]); u.w = pack_bf16x2(f[i * 8 + 6], f[i * 8 + 7]);
    *reinterpret_cast<uint4*>(row + ((ch ^ (r & 7)) << 4)) = u;
  }
}
__device__ __forceinline__ void st_sw_16(uint8_t* block, int r, int c0, const float (&f)[16]) {
  uint8_t* row = sw_row(block, r);
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const int ch = (c0 >> 3) + i;
    uint4 u;
    u.x = pack_bf16x2(f[i * 8 + 0], f[i * 8 + 1]); u.y = pack_bf16x2(f[i * 8 + 2], f[i * 8 + 3]);
    u.z = pack_bf16x2(f[i * 8 + 4], f[i * 8 + 5]); u.w = pack_bf16x2(f[i * 8 + 6], f[i * 8 + 7]);
    *reinterpret_cast<uint4*>(row + ((ch ^ (r & 7)) << 4)) = u;
  }
}
__device__ __forceinline__ float bf16_round(float x) { return __bfloat162float(__float2bfloat16_rn(x)); }

__device__ __forceinline__ void unpack8(const uint4& u, float* f) {
  f[0] = bf16lo(u.x); f[1] = bf16hi(u.x); f[2] = bf16lo(u.y); f[3] = bf16hi(u.y);
  f[4] = bf16lo(u.z); f[5] = bf16hi(u.z); f[6] = bf16lo(u.w); f[7] = bf16hi(u.w);
}
__device__ __forceinline__ uint4 pack8(const float* f) {
  uint4 u;
  u.x = pack_bf16x2(f[0], f[1]); u.y = pack_bf16x2(f[2], f[3]); u.z = pack_bf16x2(f[4], f[5]); u.w = pack_bf16x2(f[6], f[7]);
  return u;
}

// P [m][64] fp32 (global) -> bf16 128B-swizzled [mp rows][64] block image (rows >= m are zero)
__device__ __forceinline__ void stage_proj(uint8_t* Ps, const float* __restrict__ proj, int m, int mp, int tid) {
  for (int idx = tid; idx < mp * 8; idx += 256) {
    const int f = idx >> 3, ch = idx & 7;
    uint4 u = make_uint4(0, 0, 0, 0);
    if (f < m) {
      const float4 a = __ldg(reinterpret_cast<const float4*>(proj + f * 64 + ch * 8));
      const float4 b = __ldg(reinterpret_cast<const float4*>(proj + f * 64 + ch * 8 + 4));
      u.x = pack_bf16x2(a.x, a.y); u.y = pack_bf16x2(a.z, a.w); u.z = pack_bf16x2(b.x, b.y); u.w = pack_bf16x2(b.z, b.w);
    }
    *reinterpret_cast<uint4*>(sw_row(Ps, f) + ((ch ^ (f & 7)) << 4)) = u;
  }
}

// delta[bh][n] = dout[n] . out[n] and 1 / den[bh][n], once per backward pass (four kernels need them per row); optionally
// also dout_s = dout / den (bf16, dense [B * N][H * 64]): the W operand of the backward state scan, which then arrives by
// TMA like v does in the forward scan instead of being rebuilt per chunk by that kernel's running-state warps
__global__ void __launch_bounds__(256)
fv_delta_kernel(const __nv_bfloat16* __restrict__ out, const __nv_bfloat16* __restrict__ dout,
                const float* __restrict__ den, float2* __restrict__ dinv, int B, int N, int H, int out_ld,
                __nv_bfloat16* __restrict__ dout_s) {
  // eight lanes per (row, head): one 16-byte piece of the 64 columns each (coalesced), the dot product through shuffles
  const long long total = (long long)B * N * H * 8;
  const long long stride = (long long)gridDim.x * blockDim.x;          // a multiple of 32: a warp stays together
  for (long long i0 = (long long)blockIdx.x * blockDim.x + threadIdx.x; i0 - (threadIdx.x & 31) < total; i0 += stride) {
    const bool live = i0 < total;
    const long long i = live ? i0 : total - 1;
    const int q = (int)(i & 7);
    const long long rh = i >> 3;
    const int h = (int)(rh % H);
    const long long row = rh / H;                 // b * N + n
    const int n = (int)(row % N), b = (int)(row / N);
    const long long ro = row * out_ld + h * 64 + q * 8;
    const long long o = ((long long)b * H + h) * N + n;
    float fo[8], fd[8];
    unpack8(__ldg(reinterpret_cast<const uint4*>(out + ro)), fo);
    unpack8(__ldg(reinterpret_cast<const uint4*>(dout + ro)), fd);
    float acc = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) acc = fmaf(fo[j], fd[j], acc);
    acc += __shfl_xor_sync(0xffffffffu, acc, 1);
    acc += __shfl_xor_sync(0xffffffffu, acc, 2);
    acc += __shfl_xor_sync(0xffffffffu, acc, 4);
    const float inv = 1.0f / den[o];
    if (live && q == 0) dinv[o] = make_float2(acc, inv);
    if (live && dout_s) {
#pragma unroll
      for (int j = 0; j < 8; ++j) fd[j] *= inv;
      *reinterpret_cast<uint4*>(dout_s + row * (long long)(H * 64) + h * 64 + q * 8) = pack8(fd);
    }
  }
}

struct FvParams {
  CUtensorMap map_a, map_b, map_c, map_d, map_e, map_f;
  int B, N, H, m, mp, nblk, tail;    // tail = valid columns of the last feature block
  int ld, out_ld, nchunks, tmem_cols;
  float c, r, eps;
  const float* proj;
  const unsigned long long* kmax_in;
  unsigned long long* kmax_out;
  const __nv_bfloat16* x;            // featmap input (head 0, column 0 of the block)
  __nv_bfloat16* feat;               // featmap output / featmap_bwd: features
  const __nv_bfloat16* dfeat;
  int* argmax;
  float* gsum;
  float* gpart;          // deterministic mode: one partial of gsum per (CTA, lane quadrant), summed in order afterwards
  int is_query;
  const __nv_bfloat16* out;
  const __nv_bfloat16* dout;
  const float* den_in;
  float* den_out;
  __nv_bfloat16* o_out;              // scan<0>: out;  scan<1>: dv;  featmap_bwd: dx
  __nv_bfloat16* df_out;             // dqk: dq' / dk'
  float* sums;                       // chunk_state output
  const __nv_bfloat16* st_vec;       // dqk: the bf16 states (their row 64 is read directly)
  __nv_bfloat16* st_out;             // state_scan: the bf16 prefix / suffix states it writes
  const float2* dinv;                // backward: per (batch, head, position) {delta = dout . out, 1 / den}
  int w_tma;                         // state_scan<1>: the W operand (dout / den, prepared by fv_delta_kernel) arrives by TMA (map_b)
};

#define FV_PROLOGUE(NBAR_INIT)                                                              \
  extern __shared__ uint8_t smem_raw[];                                                     \
  __shared__ uint32_t tmem_base_slot;                                                       \
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;                               \
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023); \
  const int chunk = blockIdx.x, bh = blockIdx.y, b = bh / P.H, h = bh % P.H;                \
  const int n0 = chunk * FC;                                                                \
  (void)b; (void)h; (void)lane;                                                             \
  if (threadIdx.x == 0) { NBAR_INIT; fence_mbar_init(); fence_proxy_async(); }              \
  __syncthreads();

#define FV_ALLOC()                                                                          \
  if (warp == 8) { tmem_alloc(&tmem_base_slot, (uint32_t)P.tmem_cols); tmem_relinquish(); } \
  tc_fence_before();                                                                        \
  __syncthreads();                                                                          \
  tc_fence_after();                                                                         \
  const uint32_t tmem_base = tmem_base_slot;

#define FV_EPILOGUE()                                                                       \
  tc_fence_before();                                                                        \
  __syncthreads();                                                                          \
  if (warp == 8) tmem_dealloc(tmem_base, (uint32_t)P.tmem_cols);

// issue D[128 x mp] (+)= A * B over the column parts (<= 256 columns per instruction).
//   b_kmajor: B is [mp rows][64] K-major (rows contiguous, 128 B each);  else MN-major blocks `b_blk` bytes apart
__device__ __forceinline__ void mma_cols(uint32_t tD, uint64_t desc_a, uint32_t b_addr, bool b_kmajor, uint32_t b_blk,
                                         int mp, int a_mn, uint32_t accumulate) {
  const int n0 = mp < 256 ? mp : 256;
  if (b_kmajor) {
    umma_bf16(tD, desc_a, make_smem_desc(b_addr, 16, 1024, 2), make_idesc_bf16(128, n0, a_mn, 0), accumulate);
    if (mp > 256)
      umma_bf16(tD + 256, desc_a, make_smem_desc(b_addr + 256 * 128, 16, 1024, 2), make_idesc_bf16(128, mp - 256, a_mn, 0),
                accumulate);
  } else {
    umma_bf16(tD, desc_a, make_smem_desc(b_addr, b_blk, 1024, 2), make_idesc_bf16(128, n0, a_mn, 1), accumulate);
    if (mp > 256)
      umma_bf16(tD + 256, desc_a, make_smem_desc(b_addr + 4 * b_blk, b_blk, 1024, 2),
                make_idesc_bf16(128, mp - 256, a_mn, 1), accumulate);
  }
}

// ------------------------------------------------------------------------------------------------ feature map, forward
// MODE 0: key-max pre-pass   1: queries (row max stabiliser, arg-max saved)   2: keys (global stabiliser)
// Persistent: one CTA per SM stages P once, then walks (batch, head, chunk) tiles; the x tiles arrive through a
// two-stage TMA ring, so the load of tile i+1 overlaps the MMA / epilogue of tile i.
template <int MODE>
__global__ void __launch_bounds__(F_THREADS, 1)
tc_featmap_fwd_kernel(const __grid_constant__ FvParams P) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ uint32_t tmem_base_slot;
  __shared__ uint64_t x_full[2], x_empty[2], p_ready, d_full, d_empty;
  __shared__ float s_red[2][FC];
  __shared__ int s_arg[2][FC];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* Xs = smem;               // 2 x 16 KB
  uint8_t* Os = smem + 2 * BLK;     // nblk blocks: the bf16 feature tile on its way out (TMA store)
  uint8_t* Ps = Os + P.nblk * BLK;  // mp x 128 B
  const int total = P.nchunks * P.B * P.H;
  if (threadIdx.x == 0) {
    for (int i = 0; i < 2; ++i) { mbar_init(&x_full[i], 1); mbar_init(&x_empty[i], 256); }
    mbar_init(&p_ready, 256); mbar_init(&d_full, 1); mbar_init(&d_empty, 256);
    fence_mbar_init();
    fence_proxy_async();
  }
  __syncthreads();
  if (warp < 8) {
    stage_proj(Ps, P.proj, P.m, P.mp, threadIdx.x);
    fence_proxy_async();
    mbar_arrive(&p_ready);
  }
  FV_ALLOC();
  if (warp == 9) {
    if (lane == 0) {
      prefetch_tmap(&P.map_a);
      int it = 0;
      for (int tile = blockIdx.x; tile < total; tile += gridDim.x, ++it) {
        const int s = it & 1;
        const int bh = tile / P.nchunks, chunk = tile % P.nchunks;
        mbar_wait(&x_empty[s], (uint32_t)(((it >> 1) & 1) ^ 1));
        mbar_expect_tx(&x_full[s], BLK);
        tma_load_3d(Xs + s * BLK, &P.map_a, &x_full[s], (bh % P.H) * 64, chunk * FC, bh / P.H);
      }
    }
  } else if (warp == 8) {
    if (lane == 0) {
      mbar_wait(&p_ready, 0);
      const uint32_t pa = smem_u32(Ps);
      int it = 0;
      for (int tile = blockIdx.x; tile < total; tile += gridDim.x, ++it) {
        const int s = it & 1;
        mbar_wait(&x_full[s], (uint32_t)((it >> 1) & 1));
        mbar_wait(&d_empty, (uint32_t)((it & 1) ^ 1));
        tc_fence_after();
        const uint32_t xa = smem_u32(Xs + s * BLK);
#pragma unroll
        for (int k = 0; k < 4; ++k)
          mma_cols(tmem_base, make_smem_desc(xa + k * 32, 16, 1024, 2), pa + k * 32, true, 0, P.mp, 0, k > 0);
        umma_commit(&d_full);
      }
    }
  } else {
    const int q = warp & 3, hf = warp >> 2;
    const int r = q * 32 + lane;
    const uint32_t tbase = tmem_base + ((uint32_t)(q * 32) << 16);
    const int U = P.mp >> 4, U0 = U >> 1;
    const int u_beg = hf ? U0 : 0, u_end = hf ? U : U0;
    float stab_k = 0.f;
    if (MODE == 2) stab_k = ord2f((unsigned int)(P.kmax_in[0] >> 32));
    unsigned long long best_packed = 0ull;
    int it = 0;
    for (int tile = blockIdx.x; tile < total; tile += gridDim.x, ++it) {
      const int s = it & 1;
      const int bh = tile / P.nchunks, chunk = tile % P.nchunks;
      const int n = chunk * FC + r;
      const bool row_ok = n < P.N;
      mbar_wait(&x_full[s], (uint32_t)((it >> 1) & 1));
      float diag = 0.f;
      if (MODE != 0) {
        const uint8_t* row = sw_row(Xs + s * BLK, r);
#pragma unroll
        for (int ch = 0; ch < 8; ++ch) {
          float f[8];
          unpack8(*reinterpret_cast<const uint4*>(row + ((ch ^ (r & 7)) << 4)), f);
#pragma unroll
          for (int j = 0; j < 8; ++j) diag = fmaf(f[j], f[j], diag);
        }
        diag *= 0.5f * P.c * P.c;
      }
      mbar_wait(&d_full, (uint32_t)(it & 1));
      tc_fence_after();
      mbar_arrive(&x_empty[s]);        // the MMA has consumed the x tile and this thread has read its row
      // raw maximum of this thread's columns: unit-level bookkeeping (one FMNMX per element, the 16 values of the best
      // unit are kept in registers), the first arg-max is located afterwards.  (tcgen05.ld is warp-collective: its
      // address must not depend on the lane, so the best unit cannot simply be re-read.)
      float mx = -INFINITY;
      int am = 0;
      if (MODE != 2) {
        int au = u_beg;
        float best[16];
#pragma unroll
        for (int cix = 0; cix < 16; ++cix) best[cix] = -INFINITY;
        for (int u = u_beg; u < u_end; ++u) {
          uint32_t v[16];
          tmem_ld_32x16(tbase + (uint32_t)(u * 16), v);
          tmem_ld_wait();
          float w[16];
          if (u * 16 + 16 <= P.m) {
#pragma unroll
            for (int cix = 0; cix < 16; ++cix) w[cix] = __uint_as_float(v[cix]);
          } else {
#pragma unroll
            for (int cix = 0; cix < 16; ++cix) w[cix] = (u * 16 + cix < P.m) ? __uint_as_float(v[cix]) : -INFINITY;
          }
          float um = w[0];
#pragma unroll
          for (int cix = 1; cix < 16; ++cix) um = fmaxf(um, w[cix]);
          const bool better = um > mx;
          mx = better ? um : mx;
          au = better ? u : au;
#pragma unroll
          for (int cix = 0; cix < 16; ++cix) best[cix] = better ? w[cix] : best[cix];
        }
#pragma unroll
        for (int cix = 15; cix >= 0; --cix)
          if (best[cix] == mx) am = cix;
        am += au * 16;
      }
      if (MODE == 0) {
        if (row_ok && mx > -INFINITY) {
          const unsigned int flat = (unsigned int)(((long long)bh * P.N + n) * P.m + am);
          const unsigned long long packed = ((unsigned long long)f2ord(P.c * mx) << 32) | (unsigned long long)(0xFFFFFFFFu - flat);
          best_packed = packed > best_packed ? packed : best_packed;
        }
      } else {
        float stab;
        if (MODE == 1) {
          bar_epi();                     // the previous tile's readers of s_red / s_arg are done
          s_red[hf][r] = mx; s_arg[hf][r] = am;
        }
        if (threadIdx.x == 0) tma_store_wait_read();     // the previous tile's store has read the staging buffer
        bar_epi();
        if (MODE == 1) {
          const float m0 = s_red[0][r], m1 = s_red[1][r];
          stab = P.c * fmaxf(m0, m1);
          if (hf == 0 && row_ok) P.argmax[(long long)bh * P.N + n] = (m0 >= m1) ? s_arg[0][r] : s_arg[1][r];
        } else {
          stab = stab_k;
        }
        // r (exp(c D - off) + eps) = r 2^(D c log2e - off log2e) + r eps
        const float kc = P.c * 1.4426950408889634f;
        const float koff = (diag + stab) * 1.4426950408889634f;
        const float reps = P.r * P.eps;
        for (int u = u_beg; u < u_end; ++u) {
          uint32_t v[16];
          tmem_ld_32x16(tbase + (uint32_t)(u * 16), v);
          tmem_ld_wait();
          float f[16];
#pragma unroll
          for (int cix = 0; cix < 16; ++cix) f[cix] = fmaf(P.r, ex2f(fmaf(__uint_as_float(v[cix]), kc, -koff)), reps);
          if (u * 16 + 16 > P.m) {
#pragma unroll
            for (int cix = 0; cix < 16; ++cix) f[cix] = (u * 16 + cix < P.m) ? f[cix] : 0.f;
          }
          st_sw_16(Os + (u >> 2) * BLK, r, (u & 3) * 16, f);
        }
        fence_proxy_async();
        bar_epi();
        if (threadIdx.x == 0) {
          for (int cb = 0; cb < P.nblk; ++cb) tma_store_3d(&P.map_b, Os + cb * BLK, cb * 64, chunk * FC, bh);
          tma_store_commit();
        }
      }
      tc_fence_before();
      mbar_arrive(&d_empty);
    }
    if (MODE != 0 && threadIdx.x == 0) tma_store_wait_all();
    if (MODE == 0) {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const unsigned long long other = __shfl_xor_sync(0xffffffffu, best_packed, o);
        best_packed = other > best_packed ? other : best_packed;
      }
      if (lane == 0 && best_packed) atomicMax(P.kmax_out, best_packed);
    }
  }
  FV_EPILOGUE();
}

// ------------------------------------------------------------------------------------------------ chunk sums
// sums[bh][chunk][e'][f] = sum_tok W_aug[tok][e'] F[tok][f]      (e' < 65)
// MODE 0: W_aug = [v | 1]    MODE 1: W_aug = [dout / den | -(dout . out) / den]
template <int MODE>
__global__ void __launch_bounds__(F_THREADS, 2)
tc_chunk_state_kernel(const __grid_constant__ FvParams P) {
  __shared__ uint64_t f_full, w_ready, d_full;
  FV_PROLOGUE(mbar_init(&f_full, 1); mbar_init(&w_ready, 256); mbar_init(&d_full, 1));
  uint8_t* Ws = smem;                    // 2 blocks: W | aug
  uint8_t* Fs = smem + 2 * BLK;          // this CTA's feature blocks
  // gridDim.z CTAs split the feature blocks of a chunk (fewer TMEM columns / less shared memory per CTA: 2+ CTAs per SM)
  const int cb_beg = (gridDim.z > 1 && blockIdx.z == 1) ? 2 : 0;
  const int cb_end = (gridDim.z > 1 && blockIdx.z == 0) ? 2 : P.nblk;
  const int nb = cb_end - cb_beg;
  const int ncols = min(P.mp, cb_end * 64) - cb_beg * 64;       // feature columns of this CTA (multiple of 16, <= 192)
  if (warp == 9 && lane == 0) {
    prefetch_tmap(&P.map_a); prefetch_tmap(&P.map_b);
    mbar_expect_tx(&f_full, (uint32_t)nb * BLK + (MODE == 0 ? BLK : 0u));
    for (int cb = 0; cb < nb; ++cb) tma_load_3d(Fs + cb * BLK, &P.map_a, &f_full, (cb_beg + cb) * 64, n0, bh);
    if (MODE == 0) tma_load_3d(Ws, &P.map_b, &f_full, h * 64, n0, b);
  }
  if (warp < 8) {
    const int q = warp & 3, hf = warp >> 2;
    const int r = q * 32 + lane;
    const int n = n0 + r;
    // aug block: zero, then column 0
    {
      uint8_t* row = Ws + BLK + (r >> 3) * 1024 + (r & 7) * 128 + hf * 64;
#pragma unroll
      for (int i = 0; i < 4; ++i) *reinterpret_cast<uint4*>(row + i * 16) = make_uint4(0, 0, 0, 0);
    }
    bar_epi();
    float aug = 0.f;
    if (MODE == 0) {
      aug = 1.0f;
    } else {
      float f[32];
#pragma unroll
      for (int i = 0; i < 32; ++i) f[i] = 0.f;
      if (n < P.N) {
        const long long ro = ((long long)b * P.N + n) * P.out_ld + h * 64;
        const float2 di = __ldg(P.dinv + (long long)bh * P.N + n);
        const float inv = di.y;
        aug = -di.x * inv;
        const uint4* pd = reinterpret_cast<const uint4*>(P.dout + ro + hf * 32);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          unpack8(__ldg(pd + i), f + i * 8);
#pragma unroll
          for (int j = 0; j < 8; ++j) f[i * 8 + j] *= inv;
        }
      }
      st_sw_32(Ws, r, hf * 32, f);
    }
    if (hf == 0) {
      __nv_bfloat16* p0 = reinterpret_cast<__nv_bfloat16*>(sw_row(Ws + BLK, r) + (((0 ^ (r & 7))) << 4));
      *p0 = __float2bfloat16_rn(aug);
    }
    fence_proxy_async();
    mbar_arrive(&w_ready);
  }
  FV_ALLOC();
  if (warp == 8) {
    if (lane == 0) {
      mbar_wait(&f_full, 0);
      mbar_wait(&w_ready, 0);
      tc_fence_after();
      const uint32_t wa = smem_u32(Ws), fa = smem_u32(Fs);
#pragma unroll
      for (int k = 0; k < FC / 16; ++k)      // (two instructions per k-step when this CTA owns more than 256 columns)
        mma_cols(tmem_base, make_smem_desc(wa + k * 2048, BLK, 1024, 2), fa + k * 2048, false, BLK, ncols, 1, k > 0);
      umma_commit(&d_full);
    }
  } else if (warp < 8) {
    const int q = warp & 3, hf = warp >> 2;
    const int r = q * 32 + lane;       // e'
    const uint32_t tbase = tmem_base + ((uint32_t)(q * 32) << 16);
    const int U = ncols >> 4, U0 = U >> 1;
    const int u_beg = hf ? U0 : 0, u_end = hf ? U : U0;
    mbar_wait(&d_full, 0);
    tc_fence_after();
    if (q * 32 < SUM_ROWS) {           // warp-uniform: quadrants 0..2 hold rows < 65
      float* dst = P.sums + (((long long)bh * P.nchunks + chunk) * SUM_ROWS + r) * P.mp + cb_beg * 64;
      for (int u = u_beg; u < u_end; ++u) {
        uint32_t v[16];
        tmem_ld_32x16(tbase + (uint32_t)(u * 16), v);
        tmem_ld_wait();
        if (r < SUM_ROWS) {
#pragma unroll
          for (int i = 0; i < 4; ++i)
            reinterpret_cast<float4*>(dst + u * 16)[i] =
                make_float4(__uint_as_float(v[i * 4]), __uint_as_float(v[i * 4 + 1]), __uint_as_float(v[i * 4 + 2]),
                            __uint_as_float(v[i * 4 + 3]));
        }
      }
    }
  }
  FV_EPILOGUE();
}

// ------------------------------------------------------------------------------------------------ running state
// The chunk sums AND their exclusive prefix (suffix) in one kernel: a CTA owns (batch, head, up to two 64-column feature
// blocks) and walks the chunks of the sequence in order.  The tensor core computes the chunk sums
//   Z_c[e'][f] = sum_tok W_aug[tok][e'] F[tok][f]
// into two alternating TMEM accumulators (no dependency between chunks: TMA, MMA and the drain are pipelined), the
// epilogue warps keep the running state  S = sum_{earlier chunks} Z  in fp32 REGISTERS (thread = row e', 64 columns each):
// per chunk they write the state before the chunk is added -- the exclusive prefix, rounded to bf16, the tensor the
// scan / dqk kernels consume and the backward pass keeps -- and then add the chunk's sums out of TMEM.
// Replaces tc_chunk_state_kernel + fv_prefix_kernel: no fp32 chunk sums in HBM (373 MB per layer and direction), one
// pass instead of two, bound by the feature reads + state writes.  MODE / W_aug as in tc_chunk_state_kernel; MODE 1 walks
// the chunks backwards (suffix).  Roles: warps 0-7 build the W operand (MODE 1), drain TMEM and own the running state,
// warp 8 issues MMAs, warp 9 drives TMA (2-stage ring).
constexpr int SS_STAGES = 3;        // shared-memory ring (W | aug | F blocks per stage); the TMEM accumulators alternate

template <int MODE>
__global__ void __launch_bounds__(F_THREADS, 1)
tc_state_scan_kernel(const __grid_constant__ FvParams P) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ uint32_t tmem_base_slot;
  __shared__ uint64_t f_full[SS_STAGES], f_empty[SS_STAGES], w_ready[SS_STAGES], acc_full[2], acc_empty[2];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const int bh = blockIdx.y, b = bh / P.H, h = bh % P.H;
  const int cb_beg = blockIdx.x * 2;
  const int cb_end = min(cb_beg + 2, P.nblk);
  const int nb = cb_end - cb_beg;
  const int ncols = min(P.mp, cb_end * 64) - cb_beg * 64;     // feature columns of this CTA: a multiple of 16, <= 128
  const uint32_t stage_bytes = (uint32_t)(2 + nb) * BLK;      // W | aug | F blocks
  const int nch = P.nchunks;
  if (threadIdx.x == 0) {
    for (int i = 0; i < SS_STAGES; ++i) { mbar_init(&f_full[i], 1); mbar_init(&f_empty[i], 1); mbar_init(&w_ready[i], 256); }
    for (int i = 0; i < 2; ++i) { mbar_init(&acc_full[i], 1); mbar_init(&acc_empty[i], 256); }
    fence_mbar_init();
    fence_proxy_async();
  }
  if (warp < 8) {
    // aug blocks of every stage: all zero; MODE 0: column 0 = 1 (constant), MODE 1: column 0 is rewritten per chunk
    const int q = warp & 3, hf = warp >> 2;
    const int r = q * 32 + lane;
    for (int s = 0; s < SS_STAGES; ++s) {
      uint8_t* aug = smem + s * stage_bytes + BLK;
      uint8_t* row = aug + (r >> 3) * 1024 + (r & 7) * 128 + hf * 64;
#pragma unroll
      for (int i = 0; i < 4; ++i) *reinterpret_cast<uint4*>(row + i * 16) = make_uint4(0, 0, 0, 0);
    }
    bar_epi();
    if (MODE == 0 && hf == 0) {
      for (int s = 0; s < SS_STAGES; ++s) {
        __nv_bfloat16* p0 = reinterpret_cast<__nv_bfloat16*>(sw_row(smem + s * stage_bytes + BLK, r) + ((0 ^ (r & 7)) << 4));
        *p0 = __float2bfloat16_rn(1.0f);
      }
    }
    fence_proxy_async();
  }
  __syncthreads();
  FV_ALLOC();

  if (warp == 9) {
    // ------------------------------------------------------------ TMA producer (the last chunk in walking order is never added)
    if (lane == 0) {
      const bool w_tma = MODE == 0 || P.w_tma;
      prefetch_tmap(&P.map_a);
      if (w_tma) prefetch_tmap(&P.map_b);
      for (int it = 0; it + 1 < nch; ++it) {
        const int s = it % SS_STAGES;
        const int chunk = MODE ? nch - 1 - it : it;
        mbar_wait(&f_empty[s], (uint32_t)(((it / SS_STAGES) & 1) ^ 1));
        mbar_expect_tx(&f_full[s], (uint32_t)nb * BLK + (w_tma ? BLK : 0u));
        uint8_t* st = smem + s * stage_bytes;
        for (int cb = 0; cb < nb; ++cb) tma_load_3d(st + (2 + cb) * BLK, &P.map_a, &f_full[s], (cb_beg + cb) * 64, chunk * FC, bh);
        if (w_tma) tma_load_3d(st, &P.map_b, &f_full[s], h * 64, chunk * FC, b);
      }
    }
  } else if (warp == 8) {
    // ------------------------------------------------------------ MMA issuer: chunk sums into accumulator it & 1
    if (lane == 0) {
      for (int it = 0; it + 1 < nch; ++it) {
        const int s = it % SS_STAGES, a = it & 1;
        const uint32_t ph = (uint32_t)((it / SS_STAGES) & 1);
        mbar_wait(&f_full[s], ph);
        if (MODE == 1) mbar_wait(&w_ready[s], ph);
        mbar_wait(&acc_empty[a], (uint32_t)(((it >> 1) & 1) ^ 1));   // the drain of chunk it - 2 has left this accumulator
        tc_fence_after();
        const uint32_t wa = smem_u32(smem + s * stage_bytes), fa = wa + 2 * BLK;
        const uint32_t td = tmem_base + (uint32_t)(a * 128);
#pragma unroll
        for (int k = 0; k < FC / 16; ++k)
          mma_cols(td, make_smem_desc(wa + k * 2048, BLK, 1024, 2), fa + k * 2048, false, BLK, ncols, 1, (uint32_t)(k != 0));
        umma_commit(&f_empty[s]);
        umma_commit(&acc_full[a]);
      }
    }
  } else {
    // ------------------------------------------------------------ W operand (MODE 1) + running state
    const int q = warp & 3, hf = warp >> 2;
    const int r = q * 32 + lane;
    const int U = ncols >> 4, U0 = U >> 1;
    const int u_beg = hf ? U0 : 0, u_end = hf ? U : U0;
    const int nu = u_end - u_beg;                           // this thread's 16-column units (<= 4)
    const bool active = q * 32 < ST_ROWS;                   // warp-uniform: quadrants 0..2 hold rows < 80

    auto prep = [&](int it) {      // MODE 1: W = dout / den (64 columns) | aug = -delta / den, for the chunk walked at `it`
      const int s = it % SS_STAGES;
      const int chunk = nch - 1 - it;
      const int n = chunk * FC + r;
      uint8_t* Ws = smem + s * stage_bytes;
      if (P.w_tma) {               // W arrives by TMA: only the aug column is written here (one thread per row)
        if (hf == 0) {
          float aug = 0.f;
          if (n < P.N) {
            const float2 di = __ldg(P.dinv + (long long)bh * P.N + n);
            aug = -di.x * di.y;
          }
          __nv_bfloat16* p0 = reinterpret_cast<__nv_bfloat16*>(sw_row(Ws + BLK, r) + ((0 ^ (r & 7)) << 4));
          *p0 = __float2bfloat16_rn(aug);
        }
        fence_proxy_async();
        mbar_arrive(&w_ready[s]);
        return;
      }
      float f[32];
#pragma unroll
      for (int i = 0; i < 32; ++i) f[i] = 0.f;
      float aug = 0.f;
      if (n < P.N) {
        const long long ro = ((long long)b * P.N + n) * P.out_ld + h * 64;
        const float2 di = __ldg(P.dinv + (long long)bh * P.N + n);
        const float inv = di.y;
        aug = -di.x * inv;
        const uint4* pd = reinterpret_cast<const uint4*>(P.dout + ro + hf * 32);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          unpack8(__ldg(pd + i), f + i * 8);
#pragma unroll
          for (int j = 0; j < 8; ++j) f[i * 8 + j] *= inv;
        }
      }
      st_sw_32(Ws, r, hf * 32, f);
      if (hf == 0) {
        __nv_bfloat16* p0 = reinterpret_cast<__nv_bfloat16*>(sw_row(Ws + BLK, r) + ((0 ^ (r & 7)) << 4));
        *p0 = __float2bfloat16_rn(aug);
      }
      fence_proxy_async();
      mbar_arrive(&w_ready[s]);
    };

    float run[4][16];                                       // the running state of this thread's row and columns (fp32)
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 16; ++j) run[i][j] = 0.f;
    const float seed = (MODE == 0 && r == 64) ? P.eps : 0.f;              // forward: k_cumsum + eps

    if (MODE == 1) {
      for (int i = 0; i < SS_STAGES - 1 && i + 1 < nch; ++i) prep(i);
    }
    for (int it = 0; it < nch; ++it) {
      const int chunk = MODE ? nch - 1 - it : it;
      // (1) the exclusive prefix of this chunk
      if (active && r < ST_ROWS) {
        __nv_bfloat16* dst = P.st_out + (((long long)bh * nch + chunk) * ST_ROWS + r) * P.mp + cb_beg * 64;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          if (i < nu) {
            float f[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) f[j] = (r < SUM_ROWS) ? run[i][j] + seed : 0.f;
            uint4* d4 = reinterpret_cast<uint4*>(dst + (u_beg + i) * 16);
            d4[0] = pack8(f);
            d4[1] = pack8(f + 8);
          }
        }
      }
      if (it + 1 >= nch) break;
      // (2) MODE 1: the W operand of the chunk SS_STAGES - 1 ahead (its stage was last read by the MMAs of chunk it - 1,
      //     whose accumulator this thread drained in the previous iteration)
      if (MODE == 1 && it + SS_STAGES < nch) prep(it + SS_STAGES - 1);
      // (3) add this chunk's sums
      const int s = it & 1;                                 // accumulator of this chunk
      mbar_wait(&acc_full[s], (uint32_t)((it >> 1) & 1));
      tc_fence_after();
      if (active) {
        const uint32_t tbase = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(s * 128);
        uint32_t v[4][16];
#pragma unroll
        for (int i = 0; i < 4; ++i)
          if (i < nu) tmem_ld_32x16(tbase + (uint32_t)((u_beg + i) * 16), v[i]);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 4; ++i)
          if (i < nu) {
#pragma unroll
            for (int j = 0; j < 16; ++j) run[i][j] += __uint_as_float(v[i][j]);
          }
      }
      tc_fence_before();
      mbar_arrive(&acc_empty[s]);
    }
  }
  FV_EPILOGUE();
}

// ------------------------------------------------------------------------------------------------ prefix over chunks
// states[bh][chunk][80][mp] (bf16) = exclusive prefix (reverse: suffix) over chunks of sums[bh][chunk][65][mp];
// forward: the "1" row (64) is seeded with eps (k_cumsum + eps);  rows 65..79 are zero.
__global__ void fv_prefix_kernel(const float* __restrict__ sums, __nv_bfloat16* __restrict__ states, int nchunks, int mp,
                                 int reverse, float eps) {
  const int per = ST_ROWS * mp;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= per) return;
  const int bh = blockIdx.y;
  const int e = i / mp, f = i % mp;
  __nv_bfloat16* dst = states + (long long)bh * nchunks * per + i;
  if (e >= SUM_ROWS) {
    for (int cix = 0; cix < nchunks; ++cix) dst[(long long)cix * per] = __float2bfloat16_rn(0.f);
    return;
  }
  const long long sper = (long long)SUM_ROWS * mp;
  const float* src = sums + (long long)bh * nchunks * sper + (long long)e * mp + f;
  float run = (!reverse && e == 64) ? eps : 0.f;
  if (!reverse) {
    for (int cix = 0; cix < nchunks; ++cix) {
      const float v = src[cix * sper];
      dst[(long long)cix * per] = __float2bfloat16_rn(run);
      run += v;
    }
  } else {
    for (int cix = nchunks - 1; cix >= 0; --cix) {
      const float v = src[cix * sper];
      dst[(long long)cix * per] = __float2bfloat16_rn(run);
      run += v;
    }
  }
}

// ------------------------------------------------------------------------------------------------ scan / dv
// MODE 0 (forward):  X = q' (rows i), Y = k' (cols j), St = S, W = v:    out = (q' S + tril(q' k'^T) v) / den
// MODE 1 (dv):       X = k' (rows j), Y = q' (cols i), St = R, W = dout: dv = k' R + triu(k' q'^T * 1/den_i) dout
// maps: a = X features, b = Y features, c = states, d = W (head columns)
constexpr uint32_t SC_STAGE = 2 * BLK + ST_BLK;      // 43008
constexpr int SC_STAGES = 2;

template <int MODE>
__global__ void __launch_bounds__(F_THREADS, 2)
tc_scan_kernel(const __grid_constant__ FvParams P) {
  __shared__ uint64_t full[SC_STAGES], empty[SC_STAGES], w_full, a_full, am_ready, o_full;
  __shared__ float s_vec[FC];
  __shared__ float s_red[2][FC];
  FV_PROLOGUE(for (int i = 0; i < SC_STAGES; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
              mbar_init(&w_full, 1); mbar_init(&a_full, 1); mbar_init(&am_ready, 256); mbar_init(&o_full, 1));
  uint8_t* Ring = smem;
  uint8_t* Ws = smem + SC_STAGES * SC_STAGE;
  uint8_t* Am = smem;                    // 2 blocks, re-using stage 0 after the feature loop
  constexpr int NST = MODE == 0 ? ST_ROWS : 64;
  // the first loads go out BEFORE the TMEM allocation (which waits for a finishing CTA of this SM to release its columns):
  // the ring stages are free at this point, so the chunk's first blocks arrive while the allocation is pending
  auto issue_stage = [&](int cb, int stage) {
    mbar_expect_tx(&full[stage], SC_STAGE);
    uint8_t* s = Ring + stage * SC_STAGE;
    tma_load_3d(s, &P.map_a, &full[stage], cb * 64, n0, bh);
    tma_load_3d(s + BLK, &P.map_b, &full[stage], cb * 64, n0, bh);
    tma_load_2d(s + 2 * BLK, &P.map_c, &full[stage], cb * 64, (bh * P.nchunks + chunk) * ST_ROWS);
  };
  const int npre = P.nblk < SC_STAGES ? P.nblk : SC_STAGES;
  if (warp == 9 && lane == 0) {
    prefetch_tmap(&P.map_a); prefetch_tmap(&P.map_b); prefetch_tmap(&P.map_c); prefetch_tmap(&P.map_d);
    mbar_expect_tx(&w_full, BLK);
    tma_load_3d(Ws, &P.map_d, &w_full, h * 64, n0, b);
    for (int cb = 0; cb < npre; ++cb) issue_stage(cb, cb);
  }
  FV_ALLOC();
  const uint32_t tA = tmem_base, tO = tmem_base + 128;
  if (warp == 9) {
    if (lane == 0) {
      int stage = npre % SC_STAGES; uint32_t phase = npre == SC_STAGES ? 1u : 0u;
      for (int cb = npre; cb < P.nblk; ++cb) {
        mbar_wait(&empty[stage], phase ^ 1);
        mbar_expect_tx(&full[stage], SC_STAGE);
        uint8_t* s = Ring + stage * SC_STAGE;
        tma_load_3d(s, &P.map_a, &full[stage], cb * 64, n0, bh);
        tma_load_3d(s + BLK, &P.map_b, &full[stage], cb * 64, n0, bh);
        tma_load_2d(s + 2 * BLK, &P.map_c, &full[stage], cb * 64, (bh * P.nchunks + chunk) * ST_ROWS);
        if (++stage == SC_STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 8) {
    if (lane == 0) {
      const uint32_t idesc_a = make_idesc_bf16(128, 128, 0, 0);
      const uint32_t idesc_s = make_idesc_bf16(128, NST, 0, 0);
      const uint32_t idesc_w = make_idesc_bf16(128, 64, 0, 1);
      int stage = 0; uint32_t phase = 0;
      for (int cb = 0; cb < P.nblk; ++cb) {
        mbar_wait(&full[stage], phase);
        tc_fence_after();
        const uint32_t xa = smem_u32(Ring + stage * SC_STAGE), ya = xa + BLK, sa = xa + 2 * BLK;
        const int ksteps = (cb == P.nblk - 1) ? (P.tail >> 4) : 4;
        for (int k = 0; k < ksteps; ++k) {
          const uint64_t dx = make_smem_desc(xa + k * 32, 16, 1024, 2);
          umma_bf16(tA, dx, make_smem_desc(ya + k * 32, 16, 1024, 2), idesc_a, (cb | k) != 0);
          umma_bf16(tO, dx, make_smem_desc(sa + k * 32, 16, 1024, 2), idesc_s, (cb | k) != 0);
        }
        umma_commit(&empty[stage]);
        if (++stage == SC_STAGES) { stage = 0; phase ^= 1; }
      }
      umma_commit(&a_full);
      mbar_wait(&am_ready, 0);
      mbar_wait(&w_full, 0);
      tc_fence_after();
      const uint32_t ama = smem_u32(Am), wa = smem_u32(Ws);
#pragma unroll
      for (int k = 0; k < FC / 16; ++k)
        umma_bf16(tO, make_smem_desc(ama + (k >> 2) * BLK + (k & 3) * 32, 16, 1024, 2),
                  make_smem_desc(wa + k * 2048, BLK, 1024, 2), idesc_w, 1u);
      umma_commit(&o_full);
    }
  } else {
    const int q = warp & 3, hf = warp >> 2;
    const int r = q * 32 + lane;
    const int n = n0 + r;
    const bool row_ok = n < P.N;
    const uint32_t tlane = (uint32_t)(q * 32) << 16;
    if (MODE == 1) {
      if (hf == 0) s_vec[r] = row_ok ? __ldg(P.dinv + (long long)bh * P.N + n).y : 0.f;
      bar_epi();
    }
    mbar_wait(&a_full, 0);
    tc_fence_after();
    float rowsum = 0.f;
#pragma unroll
    for (int hh = 0; hh < 2; ++hh) {
      uint32_t v[32];
      tmem_ld_32x32(tA + tlane + (uint32_t)(hf * 64 + hh * 32), v);
      tmem_ld_wait();
      float f[32];
#pragma unroll
      for (int cix = 0; cix < 32; ++cix) {
        const int cc = hf * 64 + hh * 32 + cix;
        float val;
        if (MODE == 0) val = (cc <= r) ? __uint_as_float(v[cix]) : 0.f;
        else val = (cc >= r) ? __uint_as_float(v[cix]) * s_vec[cc] : 0.f;
        val = bf16_round(val);
        rowsum += val;
        f[cix] = val;
      }
      st_sw_32(Am + hf * BLK, r, hh * 32, f);
    }
    fence_proxy_async();
    tc_fence_before();
    mbar_arrive(&am_ready);
    if (MODE == 0) {
      s_red[hf][r] = rowsum;
      bar_epi();
      rowsum = s_red[0][r] + s_red[1][r];
    }
    mbar_wait(&o_full, 0);
    tc_fence_after();
    uint32_t v[32];
    tmem_ld_32x32(tO + tlane + (uint32_t)(hf * 32), v);
    tmem_ld_wait();
    float scale = 1.0f;
    if (MODE == 0) {
      uint32_t d16[16];
      tmem_ld_32x16(tO + tlane + 64u, d16);
      tmem_ld_wait();
      const float den = __uint_as_float(d16[0]) + rowsum;
      scale = 1.0f / den;
      if (hf == 0 && row_ok) P.den_out[(long long)bh * P.N + n] = den;
    }
    if (row_ok) {
      float f[32];
#pragma unroll
      for (int e = 0; e < 32; ++e) f[e] = __uint_as_float(v[e]) * scale;
      const int ldo = MODE == 0 ? P.out_ld : P.ld;
      uint4* dst = reinterpret_cast<uint4*>(P.o_out + ((long long)b * P.N + n) * ldo + h * 64 + hf * 32);
#pragma unroll
      for (int i = 0; i < 4; ++i) dst[i] = pack8(f + i * 8);
    }
  }
  FV_EPILOGUE();
}

// ------------------------------------------------------------------------------------------------ dq' / dk'
// MODE 0 (dq'): X = dout (rows i), Y = v (cols j), St = S, F = k':  dq' = (dout S - delta ksum + tril(dout v^T - delta) k') / den
// MODE 1 (dk'): X = v (rows j), Y = dout (cols i), St = R, F = q':  dk' = v R + Rden + triu((v dout^T - delta_i) / den_i) q'
// maps: a = X (head columns), b = Y (head columns), c = F features, d = states
// The feature blocks (64 features each) stream through a two-stage TMA ring together with the matching state block; each
// gets its own [128 x 64] accumulator (two TMEM buffers), so the epilogue of block c runs under the MMAs of block c + 1
// and the loads of block c + 2.  100 KB of shared memory and 256 TMEM columns: two CTAs per SM.
constexpr uint32_t DQ_STAGE = BLK + ST_BLK;      // 26 KB
constexpr int DQ_STAGES = 2;
constexpr size_t SMEM_DQK = 3 * BLK + DQ_STAGES * DQ_STAGE + 1024;

template <int MODE>
__global__ void __launch_bounds__(F_THREADS, 2)
tc_dqk_kernel(const __grid_constant__ FvParams P) {
  __shared__ uint64_t xy_full, b_full, bm_ready, ring_full[DQ_STAGES], ring_empty[DQ_STAGES], acc_full[2], acc_empty[2];
  __shared__ float s_delta[FC], s_inv[FC];
  FV_PROLOGUE(mbar_init(&xy_full, 1); mbar_init(&b_full, 1); mbar_init(&bm_ready, 256);
              for (int i = 0; i < DQ_STAGES; ++i) { mbar_init(&ring_full[i], 1); mbar_init(&ring_empty[i], 1); }
              for (int i = 0; i < 2; ++i) { mbar_init(&acc_full[i], 1); mbar_init(&acc_empty[i], 256); });
  uint8_t* Xs = smem;
  uint8_t* Bm = smem + BLK;              // 2 blocks; the first one holds Y until B' = X Y^T has been formed
  uint8_t* Ys = Bm;
  uint8_t* Ring = smem + 3 * BLK;
  const int npre = P.nblk < DQ_STAGES ? P.nblk : DQ_STAGES;
  if (warp == 9 && lane == 0) {
    prefetch_tmap(&P.map_a); prefetch_tmap(&P.map_b); prefetch_tmap(&P.map_c); prefetch_tmap(&P.map_d);
    mbar_expect_tx(&xy_full, 2 * BLK);
    tma_load_3d(Xs, &P.map_a, &xy_full, h * 64, n0, b);
    tma_load_3d(Ys, &P.map_b, &xy_full, h * 64, n0, b);
    for (int cb = 0; cb < npre; ++cb) {      // the free ring stages too: they arrive while the TMEM allocation is pending
      mbar_expect_tx(&ring_full[cb], DQ_STAGE);
      uint8_t* sp = Ring + cb * DQ_STAGE;
      tma_load_3d(sp, &P.map_c, &ring_full[cb], cb * 64, n0, bh);
      tma_load_2d(sp + BLK, &P.map_d, &ring_full[cb], cb * 64, (bh * P.nchunks + chunk) * ST_ROWS);
    }
  }
  FV_ALLOC();
  const uint32_t tB = tmem_base, tD = tmem_base + 128;
  if (warp == 9) {
    if (lane == 0) {       // (after the CTA-wide barrier of the TMEM allocation: this loop waits on the MMA warp)
      int stage = npre % DQ_STAGES; uint32_t phase = npre == DQ_STAGES ? 1u : 0u;
      for (int cb = npre; cb < P.nblk; ++cb) {
        mbar_wait(&ring_empty[stage], phase ^ 1);
        mbar_expect_tx(&ring_full[stage], DQ_STAGE);
        uint8_t* sp = Ring + stage * DQ_STAGE;
        tma_load_3d(sp, &P.map_c, &ring_full[stage], cb * 64, n0, bh);
        tma_load_2d(sp + BLK, &P.map_d, &ring_full[stage], cb * 64, (bh * P.nchunks + chunk) * ST_ROWS);
        if (++stage == DQ_STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 8) {
    if (lane == 0) {
      const uint32_t xa = smem_u32(Xs), ya = smem_u32(Ys), bma = smem_u32(Bm);
      mbar_wait(&xy_full, 0);
      tc_fence_after();
#pragma unroll
      for (int k = 0; k < 4; ++k)
        umma_bf16(tB, make_smem_desc(xa + k * 32, 16, 1024, 2), make_smem_desc(ya + k * 32, 16, 1024, 2),
                  make_idesc_bf16(128, 128, 0, 0), k > 0);
      umma_commit(&b_full);
      int stage = 0; uint32_t phase = 0;
      for (int cb = 0; cb < P.nblk; ++cb) {
        const int ncols = (cb == P.nblk - 1) ? P.tail : 64;
        const uint32_t idesc = make_idesc_bf16(128, ncols, 0, 1);
        const uint32_t td = tD + (uint32_t)((cb & 1) * 64);
        mbar_wait(&ring_full[stage], phase);
        mbar_wait(&acc_empty[cb & 1], (uint32_t)(((cb >> 1) & 1) ^ 1));
        tc_fence_after();
        const uint32_t fa = smem_u32(Ring + stage * DQ_STAGE), sa = fa + BLK;
#pragma unroll
        for (int k = 0; k < 4; ++k)      // X . St_c  (contraction over the 64 value columns = rows of the state block)
          umma_bf16(td, make_smem_desc(xa + k * 32, 16, 1024, 2), make_smem_desc(sa + k * 2048, ST_BLK, 1024, 2), idesc, k > 0);
        if (cb == 0) { mbar_wait(&bm_ready, 0); tc_fence_after(); }
#pragma unroll
        for (int k = 0; k < FC / 16; ++k)  // Bm . F_c  (contraction over the chunk's tokens)
          umma_bf16(td, make_smem_desc(bma + (k >> 2) * BLK + (k & 3) * 32, 16, 1024, 2),
                    make_smem_desc(fa + k * 2048, BLK, 1024, 2), idesc, 1u);
        umma_commit(&ring_empty[stage]);
        umma_commit(&acc_full[cb & 1]);
        if (++stage == DQ_STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp < 8) {
    const int q = warp & 3, hf = warp >> 2;
    const int r = q * 32 + lane;
    const int n = n0 + r;
    const bool row_ok = n < P.N;
    const uint32_t tlane = (uint32_t)(q * 32) << 16;
    if (hf == 0) {
      float dl = 0.f, iv = 0.f;
      if (row_ok) {
        const float2 di = __ldg(P.dinv + (long long)bh * P.N + n);
        dl = di.x; iv = di.y;
      }
      s_delta[r] = dl; s_inv[r] = iv;
    }
    bar_epi();
    const float my_delta = s_delta[r], my_inv = s_inv[r];
    mbar_wait(&b_full, 0);
    tc_fence_after();
#pragma unroll
    for (int hh = 0; hh < 2; ++hh) {
      uint32_t v[32];
      tmem_ld_32x32(tB + tlane + (uint32_t)(hf * 64 + hh * 32), v);
      tmem_ld_wait();
      float f[32];
#pragma unroll
      for (int cix = 0; cix < 32; ++cix) {
        const int cc = hf * 64 + hh * 32 + cix;
        if (MODE == 0) f[cix] = (cc <= r) ? __uint_as_float(v[cix]) - my_delta : 0.f;
        else f[cix] = (cc >= r) ? (__uint_as_float(v[cix]) - s_delta[cc]) * s_inv[cc] : 0.f;
      }
      st_sw_32(Bm + hf * BLK, r, hh * 32, f);
    }
    fence_proxy_async();
    tc_fence_before();
    mbar_arrive(&bm_ready);
    // the "1" row of the states (k_cumsum + eps / Rden), read from global memory (L2-resident: the TMA loads fetch it too)
    const __nv_bfloat16* vecrow = P.st_vec + ((long long)(bh * P.nchunks + chunk) * ST_ROWS + 64) * P.mp;
    __nv_bfloat16* dst = P.df_out + ((long long)bh * P.N + n) * P.mp;
    for (int cb = 0; cb < P.nblk; ++cb) {
      const int ncols = (cb == P.nblk - 1) ? P.tail : 64;
      const int c0 = cb * 64 + hf * 32;
      const int nv = hf * 32 < ncols ? (min(32, ncols - hf * 32) >> 3) : 0;      // 16-byte output vectors (8 columns each)
      uint4 vr[4];                           // the "1" row of this block: fetched before the accumulator is waited for
#pragma unroll
      for (int i = 0; i < 4; ++i) vr[i] = i < nv ? __ldg(reinterpret_cast<const uint4*>(vecrow + c0) + i) : make_uint4(0, 0, 0, 0);
      mbar_wait(&acc_full[cb & 1], (uint32_t)((cb >> 1) & 1));
      tc_fence_after();
      if (hf * 32 < ncols) {                 // warp-uniform
        uint32_t v[32];
        tmem_ld_32x32(tD + tlane + (uint32_t)((cb & 1) * 64 + hf * 32), v);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          if (i < nv) {
            float sv[8], f[8];
            unpack8(vr[i], sv);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const float acc = __uint_as_float(v[i * 8 + j]);
              f[j] = MODE == 0 ? my_inv * (acc - my_delta * sv[j]) : acc + sv[j];
            }
            if (row_ok) reinterpret_cast<uint4*>(dst + c0)[i] = pack8(f);
          }
        }
      }
      tc_fence_before();
      mbar_arrive(&acc_empty[cb & 1]);
    }
  }
  FV_EPILOGUE();
}

// ------------------------------------------------------------------------------------------------ dq' / dk' + feature map, backward
// tc_dqk_kernel with the feature-map backward as its epilogue: dq' / dk' never reach HBM.  Per feature block c the
// epilogue warps turn the finished [128 x 64] block of dq' (dk') -- still fp32, in registers -- into
//   dD = df (feat - r eps)          feat = the chunk's OWN features (q' for dq', k' for dk'), a third box of the ring stage
// write it as bf16 over that box (own row, conflict-free swizzled 16-byte accesses) and hand it to the MMA warp, which
// accumulates dD P over the blocks in its own 64 TMEM columns.  The last epilogue of a tile is tc_featmap_bwd_kernel's:
//   dx = c (dD P - s P[argmax]) - c^2 s x,   s = the row sum of dD.
// Against the two-kernel form this drops the write of df and the reads of df + feat by the second kernel (2 x 366 MB per
// tensor and layer at the benchmark size).  ~210 KB of shared memory, so ONE CTA per SM -- and therefore persistent: P
// is staged once, the three-stage ring runs across tile boundaries, X / Y of the next tile are requested as soon as the
// last df block of this tile has been issued, and B' = X Y^T of the next tile runs under this tile's last epilogue.
constexpr int DF_STAGES = 3;
constexpr uint32_t DF_STAGE = 2 * BLK + ST_BLK;  // F block | state block | own-feature block (-> dD): 42 KB

template <int MODE>
__global__ void __launch_bounds__(F_THREADS, 1)
tc_dqk_fb_kernel(const __grid_constant__ FvParams P) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ uint32_t tmem_base_slot;
  __shared__ uint64_t xy_full, xy_empty, b_full, bm_ready, x_full, ring_full[DF_STAGES], ring_empty[DF_STAGES],
      dd_ready[DF_STAGES], acc_full[2], acc_empty[2];
  __shared__ float s_delta[FC], s_inv[FC], s_red[2][FC];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* Xs = smem;
  uint8_t* Bm = smem + BLK;              // 2 blocks; the first one holds Y until B' = X Y^T has been formed
  uint8_t* Ys = Bm;
  uint8_t* Ring = smem + 3 * BLK;
  uint8_t* Ps = Ring + DF_STAGES * DF_STAGE;     // mp x 128 B
  const int total = P.nchunks * P.B * P.H;
  if (threadIdx.x == 0) {
    mbar_init(&xy_full, 1); mbar_init(&xy_empty, 1); mbar_init(&b_full, 1); mbar_init(&bm_ready, 256); mbar_init(&x_full, 1);
    for (int i = 0; i < DF_STAGES; ++i) { mbar_init(&ring_full[i], 1); mbar_init(&ring_empty[i], 1); mbar_init(&dd_ready[i], 256); }
    for (int i = 0; i < 2; ++i) { mbar_init(&acc_full[i], 1); mbar_init(&acc_empty[i], 256); }
    fence_mbar_init();
    fence_proxy_async();
  }
  __syncthreads();
  if (warp < 8) { stage_proj(Ps, P.proj, P.m, P.mp, threadIdx.x); fence_proxy_async(); }
  FV_ALLOC();
  const uint32_t tB = tmem_base, tD = tmem_base + 128, tX = tmem_base + 256;
  if (warp == 9) {
    if (lane == 0) {
      prefetch_tmap(&P.map_a); prefetch_tmap(&P.map_b); prefetch_tmap(&P.map_c); prefetch_tmap(&P.map_d); prefetch_tmap(&P.map_e);
      int stage = 0; uint32_t phase = 0;
      int it = 0;
      for (int tile = blockIdx.x; tile < total; tile += gridDim.x, ++it) {
        const int bh = tile / P.nchunks, chunk = tile % P.nchunks;
        const int b = bh / P.H, h = bh % P.H, n0 = chunk * FC;
        mbar_wait(&xy_empty, (uint32_t)((it & 1) ^ 1));     // the df MMAs of the previous tile no longer read Xs / Bm
        mbar_expect_tx(&xy_full, 2 * BLK);
        tma_load_3d(Xs, &P.map_a, &xy_full, h * 64, n0, b);
        tma_load_3d(Ys, &P.map_b, &xy_full, h * 64, n0, b);
        for (int cb = 0; cb < P.nblk; ++cb) {
          mbar_wait(&ring_empty[stage], phase ^ 1);
          mbar_expect_tx(&ring_full[stage], DF_STAGE);
          uint8_t* sp = Ring + stage * DF_STAGE;
          tma_load_3d(sp, &P.map_c, &ring_full[stage], cb * 64, n0, bh);
          tma_load_2d(sp + BLK, &P.map_d, &ring_full[stage], cb * 64, (bh * P.nchunks + chunk) * ST_ROWS);
          tma_load_3d(sp + BLK + ST_BLK, &P.map_e, &ring_full[stage], cb * 64, n0, bh);
          if (++stage == DF_STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 8) {
    if (lane == 0) {
      const uint32_t xa = smem_u32(Xs), ya = smem_u32(Ys), bma = smem_u32(Bm), pa = smem_u32(Ps);
      const uint32_t idesc_x = make_idesc_bf16(128, 64, 0, 1);
      int s1 = 0, s2 = 0; uint32_t ph1 = 0, ph2 = 0;
      int it = 0, gb = 0;            // gb: running count of df blocks (the two df accumulators alternate across tiles)
      for (int tile = blockIdx.x; tile < total; tile += gridDim.x, ++it, gb += P.nblk) {
        mbar_wait(&xy_full, (uint32_t)(it & 1));
        tc_fence_after();
#pragma unroll
        for (int k = 0; k < 4; ++k)
          umma_bf16(tB, make_smem_desc(xa + k * 32, 16, 1024, 2), make_smem_desc(ya + k * 32, 16, 1024, 2),
                    make_idesc_bf16(128, 128, 0, 0), k > 0);
        umma_commit(&b_full);
        for (int cb = 0; cb <= P.nblk; ++cb) {
          if (cb < P.nblk) {       // df block cb:  X . St_c + Bm . F_c
            const int g = gb + cb, ab = g & 1;
            const int ncols = (cb == P.nblk - 1) ? P.tail : 64;
            const uint32_t idesc = make_idesc_bf16(128, ncols, 0, 1);
            const uint32_t td = tD + (uint32_t)(ab * 64);
            mbar_wait(&ring_full[s1], ph1);
            mbar_wait(&acc_empty[ab], (uint32_t)(((g >> 1) & 1) ^ 1));
            tc_fence_after();
            const uint32_t fa = smem_u32(Ring + s1 * DF_STAGE), sa = fa + BLK;
#pragma unroll
            for (int k = 0; k < 4; ++k)
              umma_bf16(td, make_smem_desc(xa + k * 32, 16, 1024, 2), make_smem_desc(sa + k * 2048, ST_BLK, 1024, 2), idesc, k > 0);
            if (cb == 0) { mbar_wait(&bm_ready, (uint32_t)(it & 1)); tc_fence_after(); }
#pragma unroll
            for (int k = 0; k < FC / 16; ++k)
              umma_bf16(td, make_smem_desc(bma + (k >> 2) * BLK + (k & 3) * 32, 16, 1024, 2),
                        make_smem_desc(fa + k * 2048, BLK, 1024, 2), idesc, 1u);
            umma_commit(&acc_full[ab]);
            if (cb == P.nblk - 1) umma_commit(&xy_empty);
            if (++s1 == DF_STAGES) { s1 = 0; ph1 ^= 1; }
          }
          if (cb >= 1) {           // dx += dD_{cb-1} P_{cb-1}: issued after the MMAs of block cb, so it waits under them
            const int pb = cb - 1;
            mbar_wait(&dd_ready[s2], ph2);
            tc_fence_after();
            const uint32_t da = smem_u32(Ring + s2 * DF_STAGE + BLK + ST_BLK);
            const int ksteps = (pb == P.nblk - 1) ? (P.tail >> 4) : 4;
            for (int k = 0; k < ksteps; ++k)
              umma_bf16(tX, make_smem_desc(da + k * 32, 16, 1024, 2), make_smem_desc(pa + (pb * 4 + k) * 2048, 8192, 1024, 2),
                        idesc_x, (pb | k) != 0);
            umma_commit(&ring_empty[s2]);
            if (++s2 == DF_STAGES) { s2 = 0; ph2 ^= 1; }
          }
        }
        umma_commit(&x_full);
      }
    }
  } else if (warp < 8) {
    const int q = warp & 3, hf = warp >> 2;
    const int r = q * 32 + lane;
    const uint32_t tlane = (uint32_t)(q * 32) << 16;
    const float re = P.r * P.eps;
    float gacc = 0.f;
    int se = 0; uint32_t phe = 0;
    int it = 0, gb = 0;
    float2 di = make_float2(0.f, 0.f);       // {delta, 1 / den} of this thread's row: requested one tile ahead
    if (hf == 0 && (int)blockIdx.x < total) {
      const int bh0 = blockIdx.x / P.nchunks, nn = (blockIdx.x % P.nchunks) * FC + r;
      if (nn < P.N) di = __ldg(P.dinv + (long long)bh0 * P.N + nn);
    }
    for (int tile = blockIdx.x; tile < total; tile += gridDim.x, ++it, gb += P.nblk) {
      const int bh = tile / P.nchunks, chunk = tile % P.nchunks;
      const int b = bh / P.H, h = bh % P.H;
      const int n = chunk * FC + r;
      const bool row_ok = n < P.N;
      if (hf == 0) {
        s_delta[r] = di.x; s_inv[r] = di.y;
        di = make_float2(0.f, 0.f);
        const int nt = tile + gridDim.x;
        if (nt < total) {
          const int nn = (nt % P.nchunks) * FC + r;
          if (nn < P.N) di = __ldg(P.dinv + (long long)(nt / P.nchunks) * P.N + nn);
        }
      }
      // what the tile's last epilogue needs is requested now: the arg-max feature of the row (queries) and the x row
      int am = -1;
      if (MODE == 0 && row_ok) am = P.argmax[(long long)bh * P.N + n];
      const long long xo = ((long long)b * P.N + n) * P.ld + h * 64 + hf * 32;
      uint4 xr[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) xr[i] = row_ok ? __ldg(reinterpret_cast<const uint4*>(P.x + xo) + i) : make_uint4(0, 0, 0, 0);
      bar_epi();
      const float my_delta = s_delta[r], my_inv = s_inv[r];
      mbar_wait(&b_full, (uint32_t)(it & 1));
      tc_fence_after();
#pragma unroll
      for (int hh = 0; hh < 2; ++hh) {
        uint32_t v[32];
        tmem_ld_32x32(tB + tlane + (uint32_t)(hf * 64 + hh * 32), v);
        tmem_ld_wait();
        float f[32];
#pragma unroll
        for (int cix = 0; cix < 32; ++cix) {
          const int cc = hf * 64 + hh * 32 + cix;
          if (MODE == 0) f[cix] = (cc <= r) ? __uint_as_float(v[cix]) - my_delta : 0.f;
          else f[cix] = (cc >= r) ? (__uint_as_float(v[cix]) - s_delta[cc]) * s_inv[cc] : 0.f;
        }
        st_sw_32(Bm + hf * BLK, r, hh * 32, f);
      }
      fence_proxy_async();
      tc_fence_before();
      mbar_arrive(&bm_ready);
      const __nv_bfloat16* vecrow = P.st_vec + ((long long)(bh * P.nchunks + chunk) * ST_ROWS + 64) * P.mp;
      float part = 0.f;
      for (int cb = 0; cb < P.nblk; ++cb) {
        const int g = gb + cb, ab = g & 1;
        const int ncols = (cb == P.nblk - 1) ? P.tail : 64;
        const int c0 = cb * 64 + hf * 32;
        const int nv = hf * 32 < ncols ? (min(32, ncols - hf * 32) >> 3) : 0;      // 16-byte vectors (8 columns each)
        uint4 vr[4];                           // the "1" row of this block: fetched before the accumulator is waited for
#pragma unroll
        for (int i = 0; i < 4; ++i) vr[i] = i < nv ? __ldg(reinterpret_cast<const uint4*>(vecrow + c0) + i) : make_uint4(0, 0, 0, 0);
        mbar_wait(&ring_full[se], phe);        // the own-feature box of this stage (TMA writes become visible to this thread)
        mbar_wait(&acc_full[ab], (uint32_t)((g >> 1) & 1));
        tc_fence_after();
        uint32_t v[32];
        if (nv > 0) {                          // warp-uniform
          tmem_ld_32x32(tD + tlane + (uint32_t)(ab * 64 + hf * 32), v);
          tmem_ld_wait();
        }
        tc_fence_before();
        mbar_arrive(&acc_empty[ab]);
        if (nv > 0) {
          uint8_t* rowd = sw_row(Ring + se * DF_STAGE + BLK + ST_BLK, r);
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            if (i < nv) {
              float sv[8], a[8], gg[8];
              unpack8(vr[i], sv);
              uint4* pd = reinterpret_cast<uint4*>(rowd + (((hf * 4 + i) ^ (r & 7)) << 4));
              unpack8(*pd, a);
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                const float acc = __uint_as_float(v[i * 8 + j]);
                const float df = MODE == 0 ? my_inv * (acc - my_delta * sv[j]) : acc + sv[j];
                gg[j] = (row_ok && c0 + i * 8 + j < P.m) ? df * (a[j] - re) : 0.f;
                part += gg[j];
              }
              *pd = pack8(gg);
            }
          }
        }
        fence_proxy_async();
        mbar_arrive(&dd_ready[se]);
        if (++se == DF_STAGES) { se = 0; phe ^= 1; }
      }
      bar_epi();                                // readers of s_red of the previous tile are done
      s_red[hf][r] = part;
      bar_epi();
      const float ssum = row_ok ? s_red[0][r] + s_red[1][r] : 0.f;
      if (MODE == 1 && hf == 0) gacc += ssum;   // keys: the sum over all rows feeds the global stabiliser's gradient
      mbar_wait(&x_full, (uint32_t)(it & 1));
      tc_fence_after();
      uint32_t v[32];
      tmem_ld_32x32(tX + tlane + (uint32_t)(hf * 32), v);
      tmem_ld_wait();
      tc_fence_before();                        // (the next tile's dD P overwrites these columns after dd_ready)
      if (row_ok) {
        uint4* dst = reinterpret_cast<uint4*>(P.o_out + xo);
        const float c2s = P.c * P.c * ssum;
        const uint8_t* prow = am >= 0 ? sw_row(Ps, am) : nullptr;     // bf16 P row of the arg-max feature (64 columns)
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          float xv[8], f[8], pv[8];
          unpack8(xr[i], xv);
          if (prow) {
            unpack8(*reinterpret_cast<const uint4*>(prow + (((hf * 4 + i) ^ (am & 7)) << 4)), pv);
          } else {
#pragma unroll
            for (int j = 0; j < 8; ++j) pv[j] = 0.f;
          }
#pragma unroll
          for (int j = 0; j < 8; ++j) f[j] = P.c * (__uint_as_float(v[i * 8 + j]) - ssum * pv[j]) - c2s * xv[j];
          dst[i] = pack8(f);
        }
      }
    }
    if (MODE == 1 && hf == 0 && (P.gsum || P.gpart)) {
      const float w = sa_warp_sum(gacc);
      if (lane == 0) {
        if (P.gpart) P.gpart[blockIdx.x * 4 + q] = w;
        else atomicAdd(P.gsum, w);
      }
    }
  }
  FV_EPILOGUE();
}

// ------------------------------------------------------------------------------------------------ feature map, backward
// Persistent (one CTA per SM, P staged once), pipelined over the feature blocks: the producer warp streams (feat, dfeat)
// blocks of 64 features through a four-stage TMA ring; the epilogue warps turn each dfeat block IN PLACE into
// dD = dfeat (feat - r eps) (bf16, own row: conflict-free swizzled 16-byte accesses) and hand it to the MMA warp, which
// accumulates dD P over the blocks in one of two TMEM buffers; when a tile's last block is in, the epilogue warps finish
// dx = c (dD P - s P[argmax]) - c^2 s x  (the arg-max term of the non-detached query stabiliser is applied here, with the
// same bf16 P row the MMA used) while the blocks of the next tile are already streaming.
constexpr int FB_STAGES = 4;
constexpr uint32_t FB_STAGE = 2 * BLK;       // feat block | dfeat block

__global__ void __launch_bounds__(F_THREADS, 1)
tc_featmap_bwd_kernel(const __grid_constant__ FvParams P) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ uint32_t tmem_base_slot;
  __shared__ uint64_t full[FB_STAGES], ready[FB_STAGES], empty[FB_STAGES], d_full[2], d_empty[2];
  __shared__ float s_red[2][FC];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* Ring = smem;                                  // FB_STAGES x (feat block | dfeat block -> dD block)
  uint8_t* Ps = smem + FB_STAGES * FB_STAGE;             // mp x 128 B
  const int total = P.nchunks * P.B * P.H;
  if (threadIdx.x == 0) {
    for (int i = 0; i < FB_STAGES; ++i) { mbar_init(&full[i], 1); mbar_init(&ready[i], 256); mbar_init(&empty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&d_full[i], 1); mbar_init(&d_empty[i], 256); }
    fence_mbar_init();
    fence_proxy_async();
  }
  __syncthreads();
  if (warp < 8) { stage_proj(Ps, P.proj, P.m, P.mp, threadIdx.x); fence_proxy_async(); }
  FV_ALLOC();
  if (warp == 9) {
    if (lane == 0) {
      prefetch_tmap(&P.map_a); prefetch_tmap(&P.map_b);
      int stage = 0; uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < total; tile += gridDim.x) {
        const int bh = tile / P.nchunks, chunk = tile % P.nchunks;
        for (int cb = 0; cb < P.nblk; ++cb) {
          mbar_wait(&empty[stage], phase ^ 1);
          mbar_expect_tx(&full[stage], FB_STAGE);
          uint8_t* sp = Ring + stage * FB_STAGE;
          tma_load_3d(sp, &P.map_a, &full[stage], cb * 64, chunk * FC, bh);
          tma_load_3d(sp + BLK, &P.map_b, &full[stage], cb * 64, chunk * FC, bh);
          if (++stage == FB_STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 8) {
    if (lane == 0) {
      const uint32_t pa = smem_u32(Ps);
      const uint32_t idesc = make_idesc_bf16(128, 64, 0, 1);
      int stage = 0; uint32_t phase = 0;
      int it = 0;
      for (int tile = blockIdx.x; tile < total; tile += gridDim.x, ++it) {
        const int bsel = it & 1;
        mbar_wait(&d_empty[bsel], (uint32_t)(((it >> 1) & 1) ^ 1));
        for (int cb = 0; cb < P.nblk; ++cb) {
          mbar_wait(&ready[stage], phase);
          tc_fence_after();
          const uint32_t da = smem_u32(Ring + stage * FB_STAGE + BLK);
          const int ksteps = (cb == P.nblk - 1) ? (P.tail >> 4) : 4;
          for (int k = 0; k < ksteps; ++k)
            umma_bf16(tmem_base + (uint32_t)(bsel * 64), make_smem_desc(da + k * 32, 16, 1024, 2),
                      make_smem_desc(pa + (cb * 4 + k) * 2048, 8192, 1024, 2), idesc, (cb | k) != 0);
          umma_commit(&empty[stage]);
          if (++stage == FB_STAGES) { stage = 0; phase ^= 1; }
        }
        umma_commit(&d_full[bsel]);
      }
    }
  } else if (warp < 8) {
    const int q = warp & 3, hf = warp >> 2;
    const int r = q * 32 + lane;
    const float re = P.r * P.eps;
    float gacc = 0.f;
    int stage = 0; uint32_t phase = 0;
    int it = 0;
    for (int tile = blockIdx.x; tile < total; tile += gridDim.x, ++it) {
      const int bsel = it & 1;
      const int bh = tile / P.nchunks, chunk = tile % P.nchunks;
      const int b = bh / P.H, h = bh % P.H;
      const int n = chunk * FC + r;
      const bool row_ok = n < P.N;
      int am = -1;
      if (P.is_query && row_ok) am = P.argmax[(long long)bh * P.N + n];
      // the x row of the final epilogue is fetched now: its latency hides behind the feature blocks
      const long long xo = ((long long)b * P.N + n) * P.ld + h * 64 + hf * 32;
      uint4 xr[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) xr[i] = row_ok ? __ldg(reinterpret_cast<const uint4*>(P.x + xo) + i) : make_uint4(0, 0, 0, 0);
      float part = 0.f;
      for (int cb = 0; cb < P.nblk; ++cb) {
        mbar_wait(&full[stage], phase);
        uint8_t* sp = Ring + stage * FB_STAGE;
        const int ncols = (cb == P.nblk - 1) ? P.tail : 64;
        if (hf * 32 < ncols) {
          const uint8_t* rowa = sw_row(sp, r);
          uint8_t* rowd = sw_row(sp + BLK, r);
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int col0 = hf * 32 + i * 8;                 // column inside the block
            if (col0 < ncols) {
              const int ch = col0 >> 3;
              uint4* pd = reinterpret_cast<uint4*>(rowd + ((ch ^ (r & 7)) << 4));
              float a[8], d[8], g[8];
              unpack8(*reinterpret_cast<const uint4*>(rowa + ((ch ^ (r & 7)) << 4)), a);
              unpack8(*pd, d);
              const int f0 = cb * 64 + col0;
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                g[j] = (f0 + j < P.m) ? d[j] * (a[j] - re) : 0.f;
                part += g[j];
              }
              *pd = pack8(g);
            }
          }
        }
        fence_proxy_async();
        mbar_arrive(&ready[stage]);
        if (++stage == FB_STAGES) { stage = 0; phase ^= 1; }
      }
      bar_epi();                                // readers of s_red of the previous tile are done
      s_red[hf][r] = part;
      bar_epi();
      const float ssum = row_ok ? s_red[0][r] + s_red[1][r] : 0.f;
      if (!P.is_query && hf == 0) gacc += ssum;
      mbar_wait(&d_full[bsel], (uint32_t)((it >> 1) & 1));
      tc_fence_after();
      uint32_t v[32];
      tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(bsel * 64 + hf * 32), v);
      tmem_ld_wait();
      tc_fence_before();
      mbar_arrive(&d_empty[bsel]);
      if (row_ok) {
        uint4* dst = reinterpret_cast<uint4*>(P.o_out + xo);
        const float c2s = P.c * P.c * ssum;
        const uint8_t* prow = am >= 0 ? sw_row(Ps, am) : nullptr;     // bf16 P row of the arg-max feature (64 columns)
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          float xv[8], f[8], pv[8];
          unpack8(xr[i], xv);
          if (prow) {
            unpack8(*reinterpret_cast<const uint4*>(prow + (((hf * 4 + i) ^ (am & 7)) << 4)), pv);
          } else {
#pragma unroll
            for (int j = 0; j < 8; ++j) pv[j] = 0.f;
          }
#pragma unroll
          for (int j = 0; j < 8; ++j) f[j] = P.c * (__uint_as_float(v[i * 8 + j]) - ssum * pv[j]) - c2s * xv[j];
          dst[i] = pack8(f);
        }
      }
    }
    if (!P.is_query && hf == 0) {
      const float w = sa_warp_sum(gacc);
      if (lane == 0) {
        if (P.gpart) P.gpart[blockIdx.x * 4 + q] = w;
        else atomicAdd(P.gsum, w);
      }
    }
  }
  FV_EPILOGUE();
}

// ------------------------------------------------------------------------------------------------ host side
std::once_flag g_once;
bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }
size_t round_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

int nblk_of(int mp) { return (mp + 63) / 64; }
int sa_sm_count() {
  static int n = 0;
  if (n == 0) {
    int dev = 0, v = 0;
    if (cudaGetDevice(&dev) == cudaSuccess && cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && v > 0) n = v;
    else n = 148;
  }
  return n;
}
size_t smem_featmap(int mp) { return (size_t)(2 + nblk_of(mp)) * BLK + round_up((size_t)mp * 128, 1024) + 1024; }
size_t smem_state(int mp) { return (size_t)(2 + nblk_of(mp)) * BLK + 1024; }
constexpr size_t SMEM_SCAN = SC_STAGES * SC_STAGE + BLK + 1024;
size_t smem_dqk(int mp) { return (size_t)4 * BLK + (size_t)nblk_of(mp) * (BLK + ST_BLK) + 1024; }
size_t smem_dqk_fb(int mp) { return (size_t)3 * BLK + (size_t)DF_STAGES * DF_STAGE + round_up((size_t)mp * 128, 1024) + 1024; }
size_t smem_fbwd(int mp) { return (size_t)FB_STAGES * FB_STAGE + round_up((size_t)mp * 128, 1024) + 1024; }
constexpr size_t SMEM_MAX = 227 * 1024 - 4096;   // opt-in ceiling minus the static shared memory of the kernels

void init_once() {
  std::call_once(g_once, [] {
    const int big = (int)SMEM_MAX;
    cudaFuncSetAttribute(tc_featmap_fwd_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, big);
    cudaFuncSetAttribute(tc_featmap_fwd_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, big);
    cudaFuncSetAttribute(tc_featmap_fwd_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, big);
    cudaFuncSetAttribute(tc_chunk_state_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, big);
    cudaFuncSetAttribute(tc_chunk_state_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, big);
    cudaFuncSetAttribute(tc_state_scan_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, big);
    cudaFuncSetAttribute(tc_state_scan_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, big);
    cudaFuncSetAttribute(tc_scan_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_SCAN);
    cudaFuncSetAttribute(tc_scan_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_SCAN);
    cudaFuncSetAttribute(tc_dqk_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_DQK);
    cudaFuncSetAttribute(tc_dqk_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_DQK);
    cudaFuncSetAttribute(tc_featmap_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, big);
    cudaFuncSetAttribute(tc_dqk_fb_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, big);
    cudaFuncSetAttribute(tc_dqk_fb_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, big);
  });
}

int tmem_cols_for(int need) {
  int c = 32;
  while (c < need) c <<= 1;
  return c;
}

void fill_common(FvParams& P, const sa_favor_desc* d, int out_ld, float eps) {
  P.B = d->batch; P.N = d->seq; P.H = d->heads; P.m = d->m; P.mp = d->mp; P.nblk = nblk_of(d->mp);
  P.tail = d->mp - 64 * (P.nblk - 1);
  P.ld = d->ld; P.out_ld = out_ld; P.nchunks = (int)sa_cdiv(d->seq, FC); P.tmem_cols = 32;
  P.c = powf((float)d->dim_head, -0.25f);
  P.r = powf((float)d->m, -0.5f);
  P.eps = eps;
  P.proj = nullptr; P.kmax_in = nullptr; P.kmax_out = nullptr; P.x = nullptr; P.feat = nullptr; P.dfeat = nullptr;
  P.argmax = nullptr; P.gsum = nullptr; P.gpart = nullptr; P.is_query = 0; P.out = nullptr; P.dout = nullptr; P.den_in = nullptr;
  P.den_out = nullptr; P.o_out = nullptr; P.df_out = nullptr; P.sums = nullptr; P.st_vec = nullptr; P.dinv = nullptr;
  P.st_out = nullptr;
}

// [bh][n][mp] feature tensor: box = 64 columns x 128 tokens of one (batch, head)
int feat_map(CUtensorMap* m, const void* base, const sa_favor_desc* d) {
  const uint64_t dims[3] = {(uint64_t)d->mp, (uint64_t)d->seq, (uint64_t)d->batch * d->heads};
  const uint64_t strides[3] = {2, (uint64_t)d->mp * 2, (uint64_t)d->seq * d->mp * 2};
  const uint32_t box[3] = {64, FC, 1};
  return sa_make_tmap_bf16(m, base, 3, dims, strides, box);
}
// head columns of a row-major [batch * seq][ld] buffer (base = column 0 of head 0)
int head_map(CUtensorMap* m, const void* base, const sa_favor_desc* d, int ld) {
  const uint64_t dims[3] = {(uint64_t)d->heads * 64, (uint64_t)d->seq, (uint64_t)d->batch};
  const uint64_t strides[3] = {2, (uint64_t)ld * 2, (uint64_t)d->seq * ld * 2};
  const uint32_t box[3] = {64, FC, 1};
  return sa_make_tmap_bf16(m, base, 3, dims, strides, box);
}
// bf16 states [bh * nchunks * 80][mp]: box = 64 columns x 80 rows
int state_map(CUtensorMap* m, const void* base, const sa_favor_desc* d, int nchunks) {
  const uint64_t dims[2] = {(uint64_t)d->mp, (uint64_t)d->batch * d->heads * nchunks * ST_ROWS};
  const uint64_t strides[2] = {2, (uint64_t)d->mp * 2};
  const uint32_t box[2] = {64, ST_ROWS};
  return sa_make_tmap_bf16(m, base, 2, dims, strides, box);
}

size_t sums_bytes(const sa_favor_desc* d) {
  return round_up((size_t)d->batch * d->heads * sa_cdiv(d->seq, FC) * SUM_ROWS * d->mp * sizeof(float), 1024);
}
size_t states_bytes(const sa_favor_desc* d) {
  return round_up((size_t)d->batch * d->heads * sa_cdiv(d->seq, FC) * ST_ROWS * d->mp * sizeof(__nv_bfloat16), 1024);
}

// persistent kernels: one CTA per SM walks the (batch, head, chunk) tiles
dim3 fb_grid(const sa_favor_desc* d) {
  const long long total = sa_cdiv(d->seq, FC) * (long long)d->batch * d->heads;
  return dim3((unsigned)(total < sa_sm_count() ? total : sa_sm_count()));
}
dim3 fv_grid(const sa_favor_desc* d) { return dim3((unsigned)sa_cdiv(d->seq, FC), (unsigned)(d->batch * d->heads)); }

// chunk sums + prefix:  mode 0: (kf, v) forward prefix;  mode 1: (qf, dout / den) suffix
int launch_states(const sa_favor_desc* d, int mode, const void* feat, const void* w, const void* out, const void* dout,
                  int out_ld, const float* den, float eps, float* sums, void* states, cudaStream_t st,
                  const float2* dinv = nullptr, const void* dout_s = nullptr) {
  static thread_local FvParams P;
  fill_common(P, d, out_ld, eps);
  int rc;
  if ((rc = feat_map(&P.map_a, feat, d)) != SA_OK) return rc;
  if (mode == 0 && (rc = head_map(&P.map_b, w, d, d->ld)) != SA_OK) return rc;
  P.w_tma = 0;
  if (mode == 1 && dout_s) {
    if ((rc = head_map(&P.map_b, dout_s, d, d->heads * 64)) != SA_OK) return rc;
    P.w_tma = 1;
  }
  P.out = (const __nv_bfloat16*)out; P.dout = (const __nv_bfloat16*)dout; P.den_in = den; P.sums = sums; P.dinv = dinv;
  bool serial = true;                                  // SA_FAVOR_STATE_SCAN=0: the two-kernel form (chunk sums, then prefix)
  if (const char* e = getenv("SA_FAVOR_STATE_SCAN")) serial = e[0] != '0';
  if (serial) {
    P.st_out = (__nv_bfloat16*)states;
    const int nb = P.nblk < 2 ? P.nblk : 2;
    const dim3 sgrid((unsigned)((P.nblk + 1) / 2), (unsigned)(d->batch * d->heads));
    const size_t ssmem = (size_t)SS_STAGES * (2 + nb) * BLK + 1024;
    P.tmem_cols = 256;                                 // two chunk-sum accumulators of up to 128 columns
    if (mode == 0) tc_state_scan_kernel<0><<<sgrid, F_THREADS, ssmem, st>>>(P);
    else tc_state_scan_kernel<1><<<sgrid, F_THREADS, ssmem, st>>>(P);
    SA_LAUNCH_CHECK();
    return SA_OK;
  }
  dim3 grid = fv_grid(d);
  size_t smem = smem_state(d->mp);
  P.tmem_cols = tmem_cols_for(d->mp);
  if (P.nblk >= 3) {        // split the feature blocks {0, 1} | {2 ..} over two CTAs
    grid.z = 2;
    const int nbmax = P.nblk - 2 > 2 ? P.nblk - 2 : 2;
    smem = (size_t)(2 + nbmax) * BLK + 1024;
    const int c1 = d->mp - 128;
    P.tmem_cols = tmem_cols_for(c1 > 128 ? c1 : 128);
  }
  if (mode == 0) tc_chunk_state_kernel<0><<<grid, F_THREADS, smem, st>>>(P);
  else tc_chunk_state_kernel<1><<<grid, F_THREADS, smem, st>>>(P);
  SA_LAUNCH_CHECK();
  const int per = ST_ROWS * d->mp;
  dim3 pgrid((unsigned)sa_cdiv(per, 256), (unsigned)(d->batch * d->heads));
  fv_prefix_kernel<<<pgrid, 256, 0, st>>>(sums, (__nv_bfloat16*)states, P.nchunks, d->mp, mode, eps);
  SA_LAUNCH_CHECK();
  return SA_OK;
}

}  // namespace

bool sa_tc_favor_supported(const sa_favor_desc* d) {
  if (d->act_dtype != SA_BF16 || d->dim_head != 64) return false;
  if (d->mp % 16 != 0 || d->mp < 16 || d->mp > 64 * F_MAXBLK || d->m > d->mp || d->m < 1) return false;
  if (d->ld % 8 != 0) return false;
  if ((long long)d->batch * d->heads > 65535) return false;
  if ((long long)d->batch * d->heads * d->seq * d->m >= (1LL << 32)) return false;
  return sa_get_tmap_encode() != nullptr;
}

size_t dinv_bytes(const sa_favor_desc* d) {
  return round_up((size_t)d->batch * d->heads * d->seq * sizeof(float2), 1024);
}
size_t sa_tc_favor_scan_workspace(const sa_favor_desc* d, int backward) {
  return sums_bytes(d) + (backward ? 2 * states_bytes(d) + dinv_bytes(d) : states_bytes(d));
}

int sa_tc_favor_featmap_fwd(const sa_favor_desc* d, int mode, const void* x, const float* proj,
                            const unsigned long long* kmax_in, unsigned long long* kmax_out, float eps, void* feat,
                            int32_t* argmax, cudaStream_t st) {
  init_once();
  sa_note_path(SA_PATH_TCGEN05);
  if (!aligned16(x) || !aligned16(proj) || (feat && !aligned16(feat))) {
    sa_set_error("tc_favor_featmap_fwd: pointers not 16-byte aligned");
    return SA_ERR_INVALID;
  }
  static thread_local FvParams P;
  fill_common(P, d, 0, eps);
  int rc;
  if ((rc = head_map(&P.map_a, x, d, d->ld)) != SA_OK) return rc;
  if (feat && (rc = feat_map(&P.map_b, feat, d)) != SA_OK) return rc;
  P.proj = proj; P.kmax_in = kmax_in; P.kmax_out = kmax_out; P.feat = (__nv_bfloat16*)feat; P.argmax = argmax;
  P.tmem_cols = tmem_cols_for(d->mp);
  const size_t smem = smem_featmap(d->mp);
  const int total = P.nchunks * d->batch * d->heads;
  const dim3 grid((unsigned)(total < sa_sm_count() ? total : sa_sm_count()));
  if (mode == 0) tc_featmap_fwd_kernel<0><<<grid, F_THREADS, smem, st>>>(P);
  else if (mode == 1) tc_featmap_fwd_kernel<1><<<grid, F_THREADS, smem, st>>>(P);
  else tc_featmap_fwd_kernel<2><<<grid, F_THREADS, smem, st>>>(P);
  SA_LAUNCH_CHECK();
  return SA_OK;
}

int sa_tc_favor_featmap_bwd(const sa_favor_desc* d, const void* x, const float* proj, int is_query, float eps,
                            const void* feat, const void* dfeat, const int32_t* argmax, void* dx, float* gsum,
                            cudaStream_t st) {
  init_once();
  sa_note_path(SA_PATH_TCGEN05);
  if (!aligned16(x) || !aligned16(proj) || !aligned16(feat) || !aligned16(dfeat) || !aligned16(dx)) {
    sa_set_error("tc_favor_featmap_bwd: pointers not 16-byte aligned");
    return SA_ERR_INVALID;
  }
  static thread_local FvParams P;
  fill_common(P, d, 0, eps);
  P.proj = proj; P.x = (const __nv_bfloat16*)x; P.feat = (__nv_bfloat16*)const_cast<void*>(feat);
  P.dfeat = (const __nv_bfloat16*)dfeat; P.argmax = const_cast<int32_t*>(argmax); P.gsum = gsum; P.is_query = is_query;
  P.o_out = (__nv_bfloat16*)dx;
  int rc;
  if ((rc = feat_map(&P.map_a, feat, d)) != SA_OK) return rc;
  if ((rc = feat_map(&P.map_b, dfeat, d)) != SA_OK) return rc;
  P.tmem_cols = 128;
  const int total = P.nchunks * d->batch * d->heads;
  const int ctas = total < sa_sm_count() ? total : sa_sm_count();
  P.gpart = (gsum && !is_query) ? sa_partial_slot(ctas * 4, st) : nullptr;
  tc_featmap_bwd_kernel<<<dim3((unsigned)ctas), F_THREADS, smem_fbwd(d->mp), st>>>(P);
  SA_LAUNCH_CHECK();
  if (P.gpart) return sa_ordered_sum(P.gpart, ctas * 4, gsum, st);
  return SA_OK;
}

size_t sa_tc_favor_states_bytes(const sa_favor_desc* d) { return states_bytes(d); }

int sa_tc_favor_scan_fwd(const sa_favor_desc* d, const void* qf, const void* kf, const void* v, float eps, void* out,
                         int out_ld, float* den, void* ws, size_t ws_bytes, void* states_out, cudaStream_t st) {
  init_once();
  sa_note_path(SA_PATH_TCGEN05);
  if (ws_bytes < sa_tc_favor_scan_workspace(d, 0)) { sa_set_error("tc_favor_scan_fwd: workspace too small"); return SA_ERR_WORKSPACE; }
  if (!aligned16(qf) || !aligned16(kf) || !aligned16(v) || !aligned16(out) || !aligned16(ws) || (out_ld & 7)) {
    sa_set_error("tc_favor_scan_fwd: pointers / leading dimensions not 16-byte aligned");
    return SA_ERR_INVALID;
  }
  float* sums = (float*)ws;
  uint8_t* states = states_out ? (uint8_t*)states_out : (uint8_t*)ws + sums_bytes(d);
  if (!aligned16(states)) { sa_set_error("tc_favor_scan_fwd: states buffer not 16-byte aligned"); return SA_ERR_INVALID; }
  int rc;
  if ((rc = launch_states(d, 0, kf, v, nullptr, nullptr, out_ld, nullptr, eps, sums, states, st)) != SA_OK) return rc;
  static thread_local FvParams P;
  fill_common(P, d, out_ld, eps);
  if ((rc = feat_map(&P.map_a, qf, d)) != SA_OK) return rc;
  if ((rc = feat_map(&P.map_b, kf, d)) != SA_OK) return rc;
  if ((rc = state_map(&P.map_c, states, d, P.nchunks)) != SA_OK) return rc;
  if ((rc = head_map(&P.map_d, v, d, d->ld)) != SA_OK) return rc;
  P.o_out = (__nv_bfloat16*)out; P.den_out = den;
  P.tmem_cols = 256;
  tc_scan_kernel<0><<<fv_grid(d), F_THREADS, SMEM_SCAN, st>>>(P);
  SA_LAUNCH_CHECK();
  return SA_OK;
}

namespace {
// the feature-map backward folded into the dq' / dk' kernels (tc_dqk_fb_kernel): what tc_featmap_bwd_kernel would be given
struct FbArgs {
  const void* x_q; const void* x_k;      // q / k head columns the features were computed from (leading dimension d->ld)
  const float* proj; float eps_feature;
  const int32_t* argq;                   // arg-max feature per query row (the non-detached row stabiliser)
  void* dx_q; void* dx_k;                // gradients of the q / k head columns (leading dimension d->ld)
  float* gsum;                           // += sum of dD over all key rows (the global key stabiliser's gradient)
};

int scan_bwd_impl(const sa_favor_desc* d, const void* qf, const void* kf, const void* v, float eps, const void* out,
                  const void* dout, int out_ld, const float* den, void* dqf, void* dkf, void* dv, void* ws,
                  size_t ws_bytes, const void* states_in, const FbArgs* fb, cudaStream_t st) {
  init_once();
  sa_note_path(SA_PATH_TCGEN05);
  if (ws_bytes < sa_tc_favor_scan_workspace(d, 1)) { sa_set_error("tc_favor_scan_bwd: workspace too small"); return SA_ERR_WORKSPACE; }
  if (!aligned16(qf) || !aligned16(kf) || !aligned16(v) || !aligned16(out) || !aligned16(dout) || !aligned16(dv) ||
      !aligned16(ws) || (out_ld & 7)) {
    sa_set_error("tc_favor_scan_bwd: pointers / leading dimensions not 16-byte aligned");
    return SA_ERR_INVALID;
  }
  if (fb ? (!aligned16(fb->x_q) || !aligned16(fb->x_k) || !aligned16(fb->dx_q) || !aligned16(fb->dx_k) || !aligned16(fb->proj))
         : (!aligned16(dqf) || !aligned16(dkf))) {
    sa_set_error("tc_favor_scan_bwd: pointers not 16-byte aligned");
    return SA_ERR_INVALID;
  }
  float* sums = (float*)ws;
  uint8_t* stS = (uint8_t*)ws + sums_bytes(d);
  uint8_t* stR = stS + states_bytes(d);
  int rc;
  if (states_in) {      // prefix states saved by the forward call: no recomputation
    if (!aligned16(states_in)) { sa_set_error("tc_favor_scan_bwd: states buffer not 16-byte aligned"); return SA_ERR_INVALID; }
    stS = (uint8_t*)const_cast<void*>(states_in);
  } else if ((rc = launch_states(d, 0, kf, v, nullptr, nullptr, out_ld, nullptr, eps, sums, stS, st)) != SA_OK) {
    return rc;
  }
  float2* dinv = reinterpret_cast<float2*>(stR + states_bytes(d));
  // dout / den for the backward state scan lives in the chunk-sum region of the workspace (unused by the state-scan form)
  __nv_bfloat16* dout_s = nullptr;
  {
    bool serial = true;
    if (const char* e = getenv("SA_FAVOR_STATE_SCAN")) serial = e[0] != '0';
    if (const char* e = getenv("SA_FAVOR_W_TMA")) { if (e[0] == '0') serial = false; }       // A/B switch
    const size_t need = (size_t)d->batch * d->seq * d->heads * 64 * sizeof(__nv_bfloat16);
    if (serial && need <= sums_bytes(d)) dout_s = reinterpret_cast<__nv_bfloat16*>(sums);
  }
  {
    const long long total = (long long)d->batch * d->seq * d->heads * 8;
    long long blocks = sa_cdiv(total, 256);
    if (blocks > 148 * 32) blocks = 148 * 32;
    fv_delta_kernel<<<(unsigned)blocks, 256, 0, st>>>((const __nv_bfloat16*)out, (const __nv_bfloat16*)dout, den, dinv, d->batch,
                                                      d->seq, d->heads, out_ld, dout_s);
    SA_LAUNCH_CHECK();
  }
  if ((rc = launch_states(d, 1, qf, nullptr, out, dout, out_ld, den, eps, sums, stR, st, dinv, dout_s)) != SA_OK) return rc;
  static thread_local FvParams P;
  const size_t smem = SMEM_DQK;
  // dq'
  fill_common(P, d, out_ld, eps);
  if ((rc = head_map(&P.map_a, dout, d, out_ld)) != SA_OK) return rc;
  if ((rc = head_map(&P.map_b, v, d, d->ld)) != SA_OK) return rc;
  if ((rc = feat_map(&P.map_c, kf, d)) != SA_OK) return rc;
  if ((rc = state_map(&P.map_d, stS, d, P.nchunks)) != SA_OK) return rc;
  P.out = (const __nv_bfloat16*)out; P.dout = (const __nv_bfloat16*)dout; P.den_in = den; P.df_out = (__nv_bfloat16*)dqf;
  P.st_vec = (const __nv_bfloat16*)stS;
  P.dinv = dinv;
  P.tmem_cols = 256;
  if (fb) {
    if ((rc = feat_map(&P.map_e, qf, d)) != SA_OK) return rc;
    P.proj = fb->proj; P.eps = fb->eps_feature; P.is_query = 1; P.argmax = const_cast<int32_t*>(fb->argq);
    P.x = (const __nv_bfloat16*)fb->x_q; P.o_out = (__nv_bfloat16*)fb->dx_q;
    P.tmem_cols = 512;                   // B' (128) | two df accumulators (2 x 64) | dx (64): one CTA per SM
    tc_dqk_fb_kernel<0><<<fb_grid(d), F_THREADS, smem_dqk_fb(d->mp), st>>>(P);
    P.tmem_cols = 256;
  } else {
    tc_dqk_kernel<0><<<fv_grid(d), F_THREADS, smem, st>>>(P);
  }
  SA_LAUNCH_CHECK();
  // dk'
  if ((rc = head_map(&P.map_a, v, d, d->ld)) != SA_OK) return rc;
  if ((rc = head_map(&P.map_b, dout, d, out_ld)) != SA_OK) return rc;
  if ((rc = feat_map(&P.map_c, qf, d)) != SA_OK) return rc;
  if ((rc = state_map(&P.map_d, stR, d, P.nchunks)) != SA_OK) return rc;
  P.df_out = (__nv_bfloat16*)dkf;
  P.st_vec = (const __nv_bfloat16*)stR;
  if (fb) {
    if ((rc = feat_map(&P.map_e, kf, d)) != SA_OK) return rc;
    const dim3 g = fb_grid(d);
    const int nparts = (int)g.x * 4;
    P.tmem_cols = 512;
    P.is_query = 0; P.argmax = nullptr; P.gsum = fb->gsum;
    P.gpart = fb->gsum ? sa_partial_slot(nparts, st) : nullptr;
    P.x = (const __nv_bfloat16*)fb->x_k; P.o_out = (__nv_bfloat16*)fb->dx_k;
    tc_dqk_fb_kernel<1><<<g, F_THREADS, smem_dqk_fb(d->mp), st>>>(P);
    SA_LAUNCH_CHECK();
    P.tmem_cols = 256;
    if (P.gpart && (rc = sa_ordered_sum(P.gpart, nparts, fb->gsum, st)) != SA_OK) return rc;
  } else {
    tc_dqk_kernel<1><<<fv_grid(d), F_THREADS, smem, st>>>(P);
    SA_LAUNCH_CHECK();
  }
  // dv
  fill_common(P, d, out_ld, eps);
  if ((rc = feat_map(&P.map_a, kf, d)) != SA_OK) return rc;
  if ((rc = feat_map(&P.map_b, qf, d)) != SA_OK) return rc;
  if ((rc = state_map(&P.map_c, stR, d, P.nchunks)) != SA_OK) return rc;
  if ((rc = head_map(&P.map_d, dout, d, out_ld)) != SA_OK) return rc;
  P.den_in = den; P.o_out = (__nv_bfloat16*)dv; P.dinv = dinv;
  P.tmem_cols = 256;
  tc_scan_kernel<1><<<fv_grid(d), F_THREADS, SMEM_SCAN, st>>>(P);
  SA_LAUNCH_CHECK();
  return SA_OK;
}
}  // namespace

int sa_tc_favor_scan_bwd(const sa_favor_desc* d, const void* qf, const void* kf, const void* v, float eps, const void* out,
                         const void* dout, int out_ld, const float* den, void* dqf, void* dkf, void* dv, void* ws,
                         size_t ws_bytes, const void* states_in, cudaStream_t st) {
  return scan_bwd_impl(d, qf, kf, v, eps, out, dout, out_ld, den, dqf, dkf, dv, ws, ws_bytes, states_in, nullptr, st);
}

// backward scan + feature-map backward of both tensors: dq' / dk' stay on chip
int sa_tc_favor_scan_bwd_fused(const sa_favor_desc* d, const void* qf, const void* kf, const void* v, const void* x_q,
                               const void* x_k, const float* proj, float eps, float eps_feature, const void* out,
                               const void* dout, int out_ld, const float* den, const int32_t* argq, void* dx_q, void* dx_k,
                               void* dv, float* gsum, void* ws, size_t ws_bytes, const void* states_in, cudaStream_t st) {
  const FbArgs fb{x_q, x_k, proj, eps_feature, argq, dx_q, dx_k, gsum};
  return scan_bwd_impl(d, qf, kf, v, eps, out, dout, out_ld, den, nullptr, nullptr, dv, ws, ws_bytes, states_in, &fb, st);
}
