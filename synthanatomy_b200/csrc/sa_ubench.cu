// Micro-benchmarks used while tuning the tensor-core kernels (not on the product path; no header entry):
//   sa_ubench_mma : tcgen05.mma issue/execute rate from resident shared-memory tiles (K-major vs MN-major operands)
//   sa_ubench_tma : TMA 5-D box load rate / latency for a single producer thread per CTA
#include "sa_tc_common.cuh"

using namespace satc;

namespace {

__global__ void __launch_bounds__(128)
ubench_mma_kernel(int a_mn, int b_mn, int N, int iters, int dep_commit, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < (16384 + 32768) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_mbar_init(); fence_proxy_async(); }
  if (warp == 0) { tmem_alloc(&tmem_slot, 512); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  if (warp == 1 && lane == 0) {
    const uint32_t idesc = make_idesc_bf16(128, N, a_mn, b_mn);
    const uint32_t sa = smem_u32(smem), sb = sa + 16384;
    const long long t0 = clock64();
    uint32_t phase = 0;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const uint64_t da = a_mn ? make_smem_desc(sa + (k & 1) * 2048, 4096, 1024, 2) : make_smem_desc(sa + k * 32, 16, 1024, 2);
        const uint64_t db = b_mn ? make_smem_desc(sb + (k & 1) * 2048, 4096, 1024, 2) : make_smem_desc(sb + k * 32, 16, 1024, 2);
        umma_bf16(tmem, da, db, idesc, 1);
      }
      if (dep_commit) { umma_commit(&bar); mbar_wait(&bar, phase); phase ^= 1; }
    }
    if (!dep_commit) { umma_commit(&bar); mbar_wait(&bar, 0); }
    const long long t1 = clock64();
    if (blockIdx.x == 0) out[0] = t1 - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}

struct TmaBenchParams {
  CUtensorMap map;
  int loads_per_stage, stages, iters;
  int box_bytes;
  int gW, gH, gD, tw, th, td;   // walk tiles like the conv kernel does
  int cblocks;
};

__global__ void __launch_bounds__(64)
ubench_tma_kernel(const __grid_constant__ TmaBenchParams P, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ uint64_t full_bar[8], empty_bar[8];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int i = 0; i < P.stages; ++i) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], 1); }
    fence_mbar_init(); fence_proxy_async();
  }
  __syncthreads();
  const uint32_t stage_bytes = (uint32_t)P.loads_per_stage * P.box_bytes;
  const int ntw = P.gW / P.tw, nth = P.gH / P.th, ntd = P.gD / P.td;
  if (warp == 0 && lane == 0) {
    const long long t0 = clock64();
    long long issue_cycles = 0;
    int stage = 0; uint32_t phase = 0;
    int tile = blockIdx.x;
    for (int it = 0; it < P.iters; ++it) {
      int r = tile % (ntw * nth * ntd);
      const int g0w = (r % ntw) * P.tw; r /= ntw;
      const int g0h = (r % nth) * P.th; r /= nth;
      const int g0d = r * P.td;
      mbar_wait(&empty_bar[stage], phase ^ 1);
      mbar_expect_tx(&full_bar[stage], stage_bytes);
      const long long i0 = clock64();
      for (int l = 0; l < P.loads_per_stage; ++l)
        tma_load_5d(smem + (size_t)stage * stage_bytes + (size_t)l * P.box_bytes, &P.map, &full_bar[stage],
                    (l % P.cblocks) * 64, g0w + (l % 3) - 1, g0h + ((l / 3) % 3) - 1, g0d, 0);
      issue_cycles += clock64() - i0;
      if (++stage == P.stages) { stage = 0; phase ^= 1; }
      if ((it & 7) == 7) tile += gridDim.x;
    }
    const long long t1 = clock64();
    if (blockIdx.x == 0) { out[0] = t1 - t0; out[1] = issue_cycles; }
  } else if (warp == 1 && lane == 0) {
    int stage = 0; uint32_t phase = 0;
    for (int it = 0; it < P.iters; ++it) {
      mbar_wait(&full_bar[stage], phase);
      mbar_arrive(&empty_bar[stage]);
      if (++stage == P.stages) { stage = 0; phase ^= 1; }
    }
    if (blockIdx.x == 0) out[2] = clock64();
  }
}

}  // namespace

extern "C" int sa_ubench_mma(int a_mn, int b_mn, int N, int iters, int dep_commit, int grid, long long* out_dev,
                             void* stream) {
  cudaFuncSetAttribute(ubench_mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  ubench_mma_kernel<<<grid, 128, 16384 + 32768 + 2048, sa_stream(stream)>>>(a_mn, b_mn, N, iters, dep_commit, out_dev);
  SA_LAUNCH_CHECK();
  return SA_OK;
}

extern "C" int sa_ubench_tma(const void* x, int C, int D, int H, int W, int tw, int th, int td, int loads_per_stage,
                             int stages, int iters, int grid, long long* out_dev, void* stream) {
  static TmaBenchParams P;
  const uint64_t dims[5] = {(uint64_t)C, (uint64_t)W, (uint64_t)H, (uint64_t)D, 1};
  const uint64_t strides[5] = {2, (uint64_t)C * 2, (uint64_t)W * C * 2, (uint64_t)H * W * C * 2, (uint64_t)D * H * W * C * 2};
  const uint32_t box[5] = {64, (uint32_t)tw, (uint32_t)th, (uint32_t)td, 1};
  int rc = sa_make_tmap_bf16(&P.map, x, 5, dims, strides, box);
  if (rc != SA_OK) return rc;
  P.loads_per_stage = loads_per_stage; P.stages = stages; P.iters = iters;
  P.box_bytes = tw * th * td * 128;
  P.gW = W; P.gH = H; P.gD = D; P.tw = tw; P.th = th; P.td = td;
  P.cblocks = C / 64;
  const size_t smem = (size_t)stages * loads_per_stage * P.box_bytes + 1024;
  cudaFuncSetAttribute(ubench_tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
  if (smem > 220 * 1024) { sa_set_error("ubench_tma: too much smem"); return SA_ERR_INVALID; }
  ubench_tma_kernel<<<grid, 64, smem, sa_stream(stream)>>>(P, out_dev);
  SA_LAUNCH_CHECK();
  return SA_OK;
}
