// Probe: can a tcgen05.mma K-major SWIZZLE_128B A-descriptor start at an arbitrary 128-byte row of a TMA-written
// block, and can the 8-row group stride (SBO) be a non-multiple of 1024 B?  (Needed to read all nine 3x3 taps of
// a conv from ONE halo block in shared memory.)   nvcc -gencode arch=compute_100a,code=sm_100a -I hyperpri_b200/csrc
#include "ptx.cuh"
#include <cstdio>
#include <cstdlib>
#include <vector>
using namespace hpri;
namespace hpri { long long g_launch_count = 0; }

constexpr int ROWS = 224;
__global__ void __launch_bounds__(128, 1)
probe(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmB, float* out, int off_rows,
      int sbo_bytes, int base_off) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = raw + ((1024u - (smem_u32(raw) & 1023u)) & 1023u);
  uint8_t* sX = smem;                 // 224 rows x 128 B = 28 KB
  uint8_t* sB = smem + 32768;         // 64 x 128 B
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 49152);
  uint64_t* done = bar + 1;
  uint32_t* tptr = reinterpret_cast<uint32_t*>(smem + 49152 + 64);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) { mbar_init(bar, 1); mbar_init(done, 1); fence_mbar_init(); fence_proxy_async_smem(); }
  if (warp == 0) { tmem_alloc(tptr, 64); tmem_relinquish(); }
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tb = *tptr;
  if (threadIdx.x == 0) {
    mbar_arrive_expect_tx(bar, ROWS * 128 + 64 * 128);
    tma_load_2d(sX, &tmX, bar, 0, 0);
    tma_load_2d(sB, &tmB, bar, 0, 0);
    mbar_wait(bar, 0);
    tc_fence_after();
    const uint32_t idesc = make_idesc_16(128, 64, 0, 0, DT_F16, DT_F16);
    uint64_t da = make_smem_desc_sw128(smem_u32(sX) + off_rows * 128, 16, sbo_bytes);
    da |= static_cast<uint64_t>(base_off & 7) << 49;
    const uint64_t db = make_smem_desc_sw128(smem_u32(sB), 16, 1024);
    for (int k = 0; k < 4; ++k) umma_bf16(tb, da + 2 * k, db + 2 * k, idesc, k > 0);
    umma_commit(done);
  }
  mbar_wait(done, 0);
  tc_fence_after();
  uint32_t v[32];
  for (int hh = 0; hh < 2; ++hh) {
    tmem_ld32(tb + (static_cast<uint32_t>(warp * 32) << 16) + hh * 32, v);
    tmem_ld_wait();
    for (int j = 0; j < 32; ++j) out[(warp * 32 + lane) * 64 + hh * 32 + j] = __uint_as_float(v[j]);
  }
  tc_fence_before(); __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(tb, 64); }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static int map2d(EncodeTiledFn enc, CUtensorMap* m, void* p, int rows) {
  cuuint64_t d[2] = {64, (cuuint64_t)rows}; cuuint64_t s[1] = {128}; cuuint32_t b[2] = {64, (cuuint32_t)rows}, e[2] = {1, 1};
  return enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, p, d, s, b, e, CU_TENSOR_MAP_INTERLEAVE_NONE,
             CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS;
}
int main() {
  void* sym = nullptr; cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &q);
  EncodeTiledFn enc = (EncodeTiledFn)sym;
  std::vector<__half> hx(ROWS * 64), hb(64 * 64);
  auto xval = [](int r, int c) { return c == 0 ? (float)r : (float)(c + (r & 15) * 64); };
  for (int r = 0; r < ROWS; ++r) for (int c = 0; c < 64; ++c) hx[r * 64 + c] = __float2half(xval(r, c));
  for (int n = 0; n < 64; ++n) for (int k = 0; k < 64; ++k) hb[n * 64 + k] = __float2half(n == k ? 1.f : 0.f);
  __half *dx, *db; float* dout;
  cudaMalloc(&dx, hx.size() * 2); cudaMalloc(&db, hb.size() * 2); cudaMalloc(&dout, 128 * 64 * 4);
  cudaMemcpy(dx, hx.data(), hx.size() * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(db, hb.data(), hb.size() * 2, cudaMemcpyHostToDevice);
  CUtensorMap mx, mb;
  if (map2d(enc, &mx, dx, ROWS) || map2d(enc, &mb, db, 64)) { printf("map fail\n"); return 1; }
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  const int cfg[][3] = {{0, 1024, 0}, {8, 1024, 0}, {1, 1024, 0}, {1, 1024, 1}, {3, 1024, 0}, {3, 1024, 3},
                        {0, 1280, 0}, {11, 1280, 0}, {11, 1280, 3}, {1, 1280, 1}, {0, 1152, 0}, {5, 2304, 0}};
  std::vector<float> ho(128 * 64);
  for (auto& c : cfg) {
    cudaMemset(dout, 0, 128 * 64 * 4);
    probe<<<1, 128, 64 * 1024>>>(mx, mb, dout, c[0], c[1], c[2]);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("cfg off=%d sbo=%d bo=%d: CUDA error %s\n", c[0], c[1], c[2], cudaGetErrorString(e)); return 2; }
    cudaMemcpy(ho.data(), dout, 128 * 64 * 4, cudaMemcpyDeviceToHost);
    int bad_rows = 0, bad_chunks = 0;
    for (int m = 0; m < 128; ++m) {
      const int want = c[0] + (m / 8) * (c[1] / 128) + (m % 8);
      bool row_ok = true;
      for (int col = 0; col < 64; ++col) if (ho[m * 64 + col] != xval(want, col)) { row_ok = false; if (col % 8 == 1) ++bad_chunks; }
      if (!row_ok) ++bad_rows;
    }
    printf("off=%2d sbo=%4d base_off=%d : %s (rows wrong %d, chunks wrong %d)", c[0], c[1], c[2], bad_rows ? "MISMATCH" : "exact", bad_rows, bad_chunks);
    if (bad_rows) {
      printf("  | m: got row(c0), chunk-src of cols 8.. :");
      for (int m : {0, 1, 7, 8, 9}) printf("  m%d->r%d c9=%g", m, (int)ho[m * 64], ho[m * 64 + 9]);
    }
    printf("\n");
  }
  return 0;
}
