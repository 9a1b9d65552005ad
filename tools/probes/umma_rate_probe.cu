// Probe: issue-bound vs pipe-bound rate of tcgen05.mma kind::f16 (M=128, K=16) with both operands in shared memory,
// for N = 64 / 128 / 256, on every SM at once.  No loads: operands are whatever is in shared memory.
#include "ptx.cuh"
#include <cstdio>
#include <vector>
using namespace hpri;
namespace hpri { long long g_launch_count = 0; }

template <int N, int DISTINCT>
__global__ void __launch_bounds__(128, 1) rate(long long* cycles, int iters, uint32_t a_off16, uint32_t sbo) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = raw + ((1024u - (smem_u32(raw) & 1023u)) & 1023u);
  uint64_t* done = reinterpret_cast<uint64_t*>(smem + 200 * 1024);
  uint32_t* tptr = reinterpret_cast<uint32_t*>(smem + 200 * 1024 + 64);
  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
  if (threadIdx.x == 0) { mbar_init(done, 1); fence_mbar_init(); }
  if (warp == 0) { tmem_alloc(tptr, 512); tmem_relinquish(); }
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tb = uniform_u32(*tptr);
  if (warp == 1) {
    const uint32_t idesc = make_idesc_16(128, N, 0, 0, DT_F16, DT_F16);
    const uint64_t d0 = make_smem_desc_sw128(0, 16, 1024), dA = make_smem_desc_sw128(0, 16, sbo);
    const uint32_t hi = static_cast<uint32_t>(d0 >> 32), hiA = static_cast<uint32_t>(dA >> 32);
    const uint32_t base = uniform_u32(smem_u32(smem));
    const uint32_t a_lo = (static_cast<uint32_t>(d0) | (base >> 4)) + a_off16, b_lo = static_cast<uint32_t>(d0) | ((base + 65536) >> 4);
    const long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
      // 8 MMAs per iteration; DISTINCT: walk 8 different 16 KB A tiles / B tiles (no operand reuse between MMAs)
      umma_f16_off_w<0, 0>(tb, a_lo, hiA, b_lo, hi, idesc, 1u);
      umma_f16_off_w<2, 2>(tb, a_lo, hiA, b_lo, hi, idesc, 1u);
      umma_f16_off_w<4, 4>(tb, a_lo, hiA, b_lo, hi, idesc, 1u);
      umma_f16_off_w<6, 6>(tb, a_lo, hiA, b_lo, hi, idesc, 1u);
      umma_f16_off_w<DISTINCT * 1024 + 0, DISTINCT * 2048 + 0>(tb + N, a_lo, hiA, b_lo, hi, idesc, 1u);
      umma_f16_off_w<DISTINCT * 1024 + 2, DISTINCT * 2048 + 2>(tb + N, a_lo, hiA, b_lo, hi, idesc, 1u);
      umma_f16_off_w<DISTINCT * 1024 + 4, DISTINCT * 2048 + 4>(tb + N, a_lo, hiA, b_lo, hi, idesc, 1u);
      umma_f16_off_w<DISTINCT * 1024 + 6, DISTINCT * 2048 + 6>(tb + N, a_lo, hiA, b_lo, hi, idesc, 1u);
    }
    umma_commit_w(uniform_u32(smem_u32(done)));
    mbar_wait_w(uniform_u32(smem_u32(done)), 0);
    const long long t1 = clock64();
    if ((threadIdx.x & 31) == 0) cycles[blockIdx.x] = t1 - t0;
  }
  tc_fence_before(); __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(tb, 512); }
}

template <int N, int DISTINCT>
static void run(const char* tag, uint32_t a_off16 = 0, uint32_t sbo = 1024) {
  long long* d; cudaMalloc(&d, 148 * 8);
  cudaFuncSetAttribute(rate<N, DISTINCT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 210 * 1024);
  const int iters = 4000;
  for (int rep = 0; rep < 2; ++rep) rate<N, DISTINCT><<<148, 128, 210 * 1024>>>(d, iters, a_off16, sbo);
  cudaError_t e = cudaDeviceSynchronize();
  std::vector<long long> h(148);
  cudaMemcpy(h.data(), d, 148 * 8, cudaMemcpyDeviceToHost);
  double avg = 0; for (auto v : h) avg += v; avg /= 148;
  printf("%-28s N=%3d: %.1f cycles per MMA (pipe floor %d)  [%s]\n", tag, N, avg / (iters * 8.0), N / 2, cudaGetErrorString(e));
  cudaFree(d);
}
int main() {
  run<64, 0>("same operands"); run<128, 0>("same operands"); run<256, 0>("same operands");
  run<64, 1>("two operand tiles"); run<128, 1>("two operand tiles"); run<256, 1>("two operand tiles");
  // halo-style A descriptors: start at an arbitrary 128-byte row, 8-row groups 1280 B apart (conv3x3_halo_kernel)
  run<64, 0>("A +1 row, SBO 1024", 8, 1024); run<128, 0>("A +1 row, SBO 1024", 8, 1024);
  run<64, 0>("A +0, SBO 1280", 0, 1280); run<128, 0>("A +0, SBO 1280", 0, 1280);
  run<64, 0>("A +11 rows, SBO 1280", 88, 1280); run<128, 0>("A +11 rows, SBO 1280", 88, 1280);
  run<64, 0>("A +4 rows, SBO 2304", 32, 2304); run<128, 0>("A +4 rows, SBO 2304", 32, 2304);
  return 0;
}
