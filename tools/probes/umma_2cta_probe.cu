// Probe: tcgen05.mma.cta_group::2 (M = 256 across a CTA pair, B split by rows between the two CTAs' shared memories).
// Each CTA loads its own A tile (128 x 64, K-major SW128) and its half of B (N/2 x 64); the leader issues the MMAs,
// a multicast commit signals both CTAs, each CTA reads its 128 accumulator rows from its own TMEM.
#include "ptx.cuh"
#include <cstdio>
#include <vector>
using namespace hpri;
namespace hpri { long long g_launch_count = 0; }

constexpr int N = 128;

__device__ __forceinline__ void cluster_sync_all_() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1)
probe(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, float* out, int variant) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = raw + ((1024u - (smem_u32(raw) & 1023u)) & 1023u);
  uint8_t* sA = smem;                 // 16 KB
  uint8_t* sB = smem + 16384;         // N/2 x 128 B = 8 KB
  uint64_t* ld_bar = reinterpret_cast<uint64_t*>(smem + 32768);
  uint64_t* done = ld_bar + 1;
  uint32_t* tptr = reinterpret_cast<uint32_t*>(smem + 32768 + 64);
  uint32_t rank;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) { mbar_init(ld_bar, 1); mbar_init(done, 1); fence_mbar_init(); fence_proxy_async_smem(); }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tptr)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tc_fence_before(); __syncthreads(); cluster_sync_all_(); tc_fence_after();
  const uint32_t tb = *tptr;
  if (threadIdx.x == 0) {
    mbar_arrive_expect_tx(ld_bar, 16384 + (N / 2) * 128);
    tma_load_2d(sA, &tmA, ld_bar, 0, rank * 128);            // this CTA's 128 rows of A (global A is 256 x 64)
    tma_load_2d(sB, &tmB, ld_bar, 0, rank * (N / 2));        // this CTA's half of the B rows
    mbar_wait(ld_bar, 0);
  }
  __syncthreads();
  cluster_sync_all_();                                        // both CTAs' operands are in shared memory
  if (rank == 0 && threadIdx.x == 0) {
    tc_fence_after();
    const uint32_t idesc = make_idesc_16(256, N, 0, 0, DT_F16, DT_F16);
    const uint64_t da = make_smem_desc_sw128(smem_u32(sA), 16, 1024);
    const uint64_t db = make_smem_desc_sw128(smem_u32(sB), 16, 1024);
    for (int k = 0; k < 4; ++k) {
      const uint32_t acc = k > 0;
      asm volatile(
          "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
          "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tb),
          "l"(da + 2 * k), "l"(db + 2 * k), "r"(idesc), "r"(acc)
          : "memory");
    }
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                     smem_u32(done)),
                 "h"(static_cast<uint16_t>(3))
                 : "memory");
  }
  mbar_wait(done, 0);
  tc_fence_after();
  for (int c = 0; c < N / 32; ++c) {
    uint32_t v[32];
    tmem_ld32(tb + (static_cast<uint32_t>(warp * 32) << 16) + c * 32, v);
    tmem_ld_wait();
    for (int j = 0; j < 32; ++j) out[(rank * 128 + warp * 32 + lane) * N + c * 32 + j] = __uint_as_float(v[j]);
  }
  tc_fence_before(); __syncthreads(); cluster_sync_all_();
  if (warp == 0) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tb), "r"(512u) : "memory");
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static int map2d(EncodeTiledFn enc, CUtensorMap* m, void* p, int rows, int box_rows) {
  cuuint64_t d[2] = {64, (cuuint64_t)rows}; cuuint64_t s[1] = {128}; cuuint32_t b[2] = {64, (cuuint32_t)box_rows}, e[2] = {1, 1};
  return enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, p, d, s, b, e, CU_TENSOR_MAP_INTERLEAVE_NONE,
             CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS;
}
int main() {
  void* sym = nullptr; cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &q);
  EncodeTiledFn enc = (EncodeTiledFn)sym;
  std::vector<__half> ha(256 * 64), hb(N * 64);
  std::vector<float> fa(256 * 64), fb(N * 64);
  for (int m = 0; m < 256; ++m) for (int k = 0; k < 64; ++k) { fa[m * 64 + k] = (float)((m * 7 + k * 3) % 5 - 2); ha[m * 64 + k] = __float2half(fa[m * 64 + k]); }
  for (int n = 0; n < N; ++n) for (int k = 0; k < 64; ++k) { fb[n * 64 + k] = (float)((n * 5 + k) % 7 - 3); hb[n * 64 + k] = __float2half(fb[n * 64 + k]); }
  __half *da, *db; float* dout;
  cudaMalloc(&da, ha.size() * 2); cudaMalloc(&db, hb.size() * 2); cudaMalloc(&dout, 256 * N * 4);
  cudaMemcpy(da, ha.data(), ha.size() * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(db, hb.data(), hb.size() * 2, cudaMemcpyHostToDevice);
  cudaMemset(dout, 0, 256 * N * 4);
  CUtensorMap ma, mb;
  if (map2d(enc, &ma, da, 256, 128) || map2d(enc, &mb, db, N, N / 2)) { printf("map fail\n"); return 1; }
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  probe<<<2, 128, 64 * 1024>>>(ma, mb, dout, 0);
  cudaError_t e = cudaDeviceSynchronize();
  printf("launch: %s\n", cudaGetErrorString(e));
  if (e != cudaSuccess) return 2;
  std::vector<float> ho(256 * N);
  cudaMemcpy(ho.data(), dout, 256 * N * 4, cudaMemcpyDeviceToHost);
  int bad = 0;
  for (int m = 0; m < 256; ++m) for (int n = 0; n < N; ++n) {
    float r = 0; for (int k = 0; k < 64; ++k) r += fa[m * 64 + k] * fb[n * 64 + k];
    if (ho[m * N + n] != r) { if (bad < 8) printf("  mismatch m=%d n=%d got %g want %g\n", m, n, ho[m * N + n], r); ++bad; }
  }
  printf("cta_group::2 M=256 N=%d K=64: %s (%d wrong of %d)\n", N, bad ? "MISMATCH" : "exact", bad, 256 * N);
  return 0;
}
