// Streaming-kernel design probe (B200): which thread mapping / staging reaches the HBM roofline for the
// BatchNorm-backward "apply" pattern (read x, read dy, write dx; per-channel constants) and the BatchNorm-apply
// pattern (read x, write y)?  Prints achieved GB/s per variant; algorithmic bytes = each tensor once.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o stream_probe stream_probe.cu
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e_), __LINE__); exit(1); } } while (0)

struct Consts { const float *sc, *sh, *ca, *cb, *cc; };

__device__ __forceinline__ float2 h2f(uint32_t v) { return __half22float2(*reinterpret_cast<__half2*>(&v)); }
__device__ __forceinline__ uint32_t f2h(float a, float b) { __half2 h = __floats2half2_rn(a, b); return *reinterpret_cast<uint32_t*>(&h); }

__device__ __forceinline__ uint4 bwd_math(const uint4 xr, const uint4 dr, const float* sc, const float* sh, const float* ca,
                                          const float* cb, const float* cc) {
  const uint32_t xs[4] = {xr.x, xr.y, xr.z, xr.w}, ds[4] = {dr.x, dr.y, dr.z, dr.w};
  uint32_t o[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const float2 x = h2f(xs[j]), d = h2f(ds[j]);
    const float dz0 = fmaf(x.x, sc[2 * j], sh[2 * j]) > 0.f ? d.x : 0.f;
    const float dz1 = fmaf(x.y, sc[2 * j + 1], sh[2 * j + 1]) > 0.f ? d.y : 0.f;
    o[j] = f2h(fmaf(ca[2 * j], dz0, fmaf(cb[2 * j], x.x, cc[2 * j])), fmaf(ca[2 * j + 1], dz1, fmaf(cb[2 * j + 1], x.y, cc[2 * j + 1])));
  }
  return make_uint4(o[0], o[1], o[2], o[3]);
}
__device__ __forceinline__ uint4 fwd_math(const uint4 xr, const float* sc, const float* sh) {
  const uint32_t xs[4] = {xr.x, xr.y, xr.z, xr.w};
  uint32_t o[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const float2 x = h2f(xs[j]);
    o[j] = f2h(fmaxf(fmaf(x.x, sc[2 * j], sh[2 * j]), 0.f), fmaxf(fmaf(x.y, sc[2 * j + 1], sh[2 * j + 1]), 0.f));
  }
  return make_uint4(o[0], o[1], o[2], o[3]);
}

template <int HINT>
__device__ __forceinline__ uint4 ldv(const uint4* p) {
  if (HINT == 1) return __ldcs(p);
  if (HINT == 2) {
    uint4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    return v;
  }
  return __ldg(p);
}
template <int HINT>
__device__ __forceinline__ void stv(uint4* p, uint4 v) {
  if (HINT >= 1) __stcs(p, v);
  else *p = v;
}

// ---- variant A: the round-1 mapping.  thread = (channel group, pixel slot); UNROLL pixels in flight, each
// `stride` pixels apart (a block touches UNROLL separate 4 KB runs per tensor)
template <int UNROLL, int MINB, int HINT, int MODE>      // MODE 0: bwd apply (2 in, 1 out), 1: fwd apply (1 in, 1 out), 2: copy
__global__ void __launch_bounds__(256, MINB)
k_strided(const uint16_t* __restrict__ x, const uint16_t* __restrict__ dy, uint16_t* __restrict__ out, long long npix, int C, Consts k) {
  const int CG = C / 8, slots = 256 / CG;
  const int cg = threadIdx.x % CG, slot = threadIdx.x / CG;
  if (slot >= slots) return;
  const int c0 = cg * 8;
  float sc[8], sh[8], ca[8], cb[8], cc[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) { sc[i] = k.sc[c0 + i]; sh[i] = k.sh[c0 + i]; ca[i] = k.ca[c0 + i]; cb[i] = k.cb[c0 + i]; cc[i] = k.cc[c0 + i]; }
  const long long stride = (long long)gridDim.x * slots;
  for (long long p0 = (long long)blockIdx.x * slots + slot; p0 < npix; p0 += UNROLL * stride) {
    uint4 xr[UNROLL], dr[UNROLL];
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) {
      const long long pp = p0 + u * stride;
      xr[u] = make_uint4(0, 0, 0, 0); dr[u] = make_uint4(0, 0, 0, 0);
      if (pp < npix) {
        xr[u] = ldv<HINT>(reinterpret_cast<const uint4*>(x + pp * C + c0));
        if (MODE == 0) dr[u] = ldv<HINT>(reinterpret_cast<const uint4*>(dy + pp * C + c0));
      }
    }
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) {
      const long long pp = p0 + u * stride;
      if (pp >= npix) break;
      const uint4 o = MODE == 0 ? bwd_math(xr[u], dr[u], sc, sh, ca, cb, cc) : MODE == 1 ? fwd_math(xr[u], sc, sh) : xr[u];
      stv<HINT>(reinterpret_cast<uint4*>(out + pp * C + c0), o);
    }
  }
}

// ---- variant B: block-contiguous.  Vector index v = chunk base + u*blockDim + tid: a block instruction covers
// 4 KB contiguous, UNROLL of them back to back (16 KB contiguous per block per tensor per iteration)
template <int UNROLL, int MINB, int HINT, int MODE>
__global__ void __launch_bounds__(256, MINB)
k_contig(const uint4* __restrict__ x, const uint4* __restrict__ dy, uint4* __restrict__ out, long long nvec, int C, Consts k) {
  const int CG = C / 8;
  const int c0 = (threadIdx.x % CG) * 8;          // 256 % CG == 0: a thread's channel group never changes
  float sc[8], sh[8], ca[8], cb[8], cc[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) { sc[i] = k.sc[c0 + i]; sh[i] = k.sh[c0 + i]; ca[i] = k.ca[c0 + i]; cb[i] = k.cb[c0 + i]; cc[i] = k.cc[c0 + i]; }
  const long long step = (long long)gridDim.x * 256 * UNROLL;
  for (long long v0 = (long long)blockIdx.x * 256 * UNROLL + threadIdx.x; v0 < nvec; v0 += step) {
    uint4 xr[UNROLL], dr[UNROLL];
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) {
      const long long v = v0 + u * 256;
      xr[u] = make_uint4(0, 0, 0, 0); dr[u] = make_uint4(0, 0, 0, 0);
      if (v < nvec) {
        xr[u] = ldv<HINT>(x + v);
        if (MODE == 0) dr[u] = ldv<HINT>(dy + v);
      }
    }
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) {
      const long long v = v0 + u * 256;
      if (v >= nvec) break;
      const uint4 o = MODE == 0 ? bwd_math(xr[u], dr[u], sc, sh, ca, cb, cc) : MODE == 1 ? fwd_math(xr[u], sc, sh) : xr[u];
      stv<HINT>(out + v, o);
    }
  }
}

// ---- variant B32: block-contiguous with 256-bit loads / stores (LDG.256 / STG.256 on sm_100a): 16 channels per thread
struct U8 { uint4 lo, hi; };
template <int EF>
__device__ __forceinline__ U8 ld32(const void* p) {
  U8 v;
  if (EF) asm volatile("ld.global.L2::evict_first.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];" : "=r"(v.lo.x), "=r"(v.lo.y), "=r"(v.lo.z), "=r"(v.lo.w), "=r"(v.hi.x), "=r"(v.hi.y), "=r"(v.hi.z), "=r"(v.hi.w) : "l"(p));
  else asm volatile("ld.global.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];" : "=r"(v.lo.x), "=r"(v.lo.y), "=r"(v.lo.z), "=r"(v.lo.w), "=r"(v.hi.x), "=r"(v.hi.y), "=r"(v.hi.z), "=r"(v.hi.w) : "l"(p));
  return v;
}
template <int EF>
__device__ __forceinline__ void st32(void* p, const U8& v) {
  if (EF) asm volatile("st.global.L2::evict_first.v8.b32 [%8], {%0,%1,%2,%3,%4,%5,%6,%7};" ::"r"(v.lo.x), "r"(v.lo.y), "r"(v.lo.z), "r"(v.lo.w), "r"(v.hi.x), "r"(v.hi.y), "r"(v.hi.z), "r"(v.hi.w), "l"(p) : "memory");
  else asm volatile("st.global.v8.b32 [%8], {%0,%1,%2,%3,%4,%5,%6,%7};" ::"r"(v.lo.x), "r"(v.lo.y), "r"(v.lo.z), "r"(v.lo.w), "r"(v.hi.x), "r"(v.hi.y), "r"(v.hi.z), "r"(v.hi.w), "l"(p) : "memory");
}
template <int UNROLL, int MINB, int EF, int MODE>
__global__ void __launch_bounds__(256, MINB)
k_contig32(const uint8_t* __restrict__ x, const uint8_t* __restrict__ dy, uint8_t* __restrict__ out, long long nv32, int C, Consts k) {
  const int CG = C / 16;
  const int c0 = (threadIdx.x % CG) * 16;
  float sc[16], sh[16], ca[16], cb[16], cc[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) { sc[i] = k.sc[c0 + i]; sh[i] = k.sh[c0 + i]; ca[i] = k.ca[c0 + i]; cb[i] = k.cb[c0 + i]; cc[i] = k.cc[c0 + i]; }
  const long long step = (long long)gridDim.x * 256 * UNROLL;
  for (long long v0 = (long long)blockIdx.x * 256 * UNROLL + threadIdx.x; v0 < nv32; v0 += step) {
    U8 xr[UNROLL], dr[UNROLL];
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) {
      const long long v = v0 + u * 256;
      if (v < nv32) {
        xr[u] = ld32<EF>(x + v * 32);
        if (MODE == 0) dr[u] = ld32<EF>(dy + v * 32);
      }
    }
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) {
      const long long v = v0 + u * 256;
      if (v >= nv32) break;
      U8 o;
      if (MODE == 0) { o.lo = bwd_math(xr[u].lo, dr[u].lo, sc, sh, ca, cb, cc); o.hi = bwd_math(xr[u].hi, dr[u].hi, sc + 8, sh + 8, ca + 8, cb + 8, cc + 8); }
      else if (MODE == 1) { o.lo = fwd_math(xr[u].lo, sc, sh); o.hi = fwd_math(xr[u].hi, sc + 8, sh + 8); }
      else o = xr[u];
      st32<EF>(out + v * 32, o);
    }
  }
}

// ---- variant C: 1-D bulk-copy (TMA) ring.  Warp 0 streams TILE-byte runs of x and dy into a STAGES-deep shared
// memory ring; 8 consumer warps compute from shared memory and store straight to global.
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, uint32_t n) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(n)); }
__device__ __forceinline__ void mbar_expect(uint64_t* b, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t* b) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(b)) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t parity) {
  asm volatile("{\n.reg .pred p;\nWAIT_%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra DONE_%=;\nbra WAIT_%=;\nDONE_%=:\n}" ::"r"(smem_u32(b)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_load(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void bulk_store(void* dst, const void* src, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(smem_u32(src)), "r"(bytes) : "memory");
}

template <int STAGES, int TILE, int MODE, int BSTORE>      // TILE bytes per tensor per stage
__global__ void __launch_bounds__(288, 1)
k_bulk(const uint8_t* __restrict__ x, const uint8_t* __restrict__ dy, uint8_t* __restrict__ out, long long nbytes, int C, Consts k) {
  extern __shared__ __align__(128) uint8_t sm[];
  constexpr int NIN = MODE == 0 ? 2 : 1;
  uint8_t* ring = sm;                                     // [STAGES][NIN][TILE]
  uint8_t* ost = sm + STAGES * NIN * TILE;                // [2][TILE] output staging (BSTORE)
  uint64_t* full = reinterpret_cast<uint64_t*>(ost + (BSTORE ? 2 * TILE : 0));
  uint64_t* empty = full + STAGES;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 8); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const long long ntiles = (nbytes + TILE - 1) / TILE;
  if (warp == 0) {
    if (lane == 0) {
      uint32_t s = 0, ph = 0;
      for (long long t = blockIdx.x; t < ntiles; t += gridDim.x) {
        mbar_wait(&empty[s], ph ^ 1);
        const long long off = t * TILE;
        const uint32_t bytes = (uint32_t)((nbytes - off) < TILE ? (nbytes - off) : TILE);
        mbar_expect(&full[s], bytes * NIN);
        bulk_load(ring + (s * NIN) * TILE, x + off, bytes, &full[s]);
        if (NIN == 2) bulk_load(ring + (s * NIN + 1) * TILE, dy + off, bytes, &full[s]);
        if (++s == STAGES) { s = 0; ph ^= 1; }
      }
    }
  } else {
    const int ct = threadIdx.x - 32;                       // 0..255
    const int CG = C / 8;
    const int c0 = (ct % CG) * 8;
    float sc[8], sh[8], ca[8], cb[8], cc[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { sc[i] = k.sc[c0 + i]; sh[i] = k.sh[c0 + i]; ca[i] = k.ca[c0 + i]; cb[i] = k.cb[c0 + i]; cc[i] = k.cc[c0 + i]; }
    uint32_t s = 0, ph = 0, ob = 0;
    constexpr int VPT = TILE / 16 / 256;                   // vectors per thread per tile
    for (long long t = blockIdx.x; t < ntiles; t += gridDim.x) {
      const long long off = t * TILE;
      const int nv = (int)(((nbytes - off) < TILE ? (nbytes - off) : TILE) / 16);
      mbar_wait(&full[s], ph);
      const uint4* xs = reinterpret_cast<const uint4*>(ring + (s * NIN) * TILE);
      const uint4* ds = reinterpret_cast<const uint4*>(ring + (s * NIN + 1) * TILE);
      uint4 o[VPT];
#pragma unroll
      for (int u = 0; u < VPT; ++u) {
        const int v = u * 256 + ct;
        if (v < nv) o[u] = MODE == 0 ? bwd_math(xs[v], ds[v], sc, sh, ca, cb, cc) : MODE == 1 ? fwd_math(xs[v], sc, sh) : xs[v];
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty[s]);               // this warp has read its part of the stage
      if (BSTORE) {
        uint4* od = reinterpret_cast<uint4*>(ost + ob * TILE);
        if (ct == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");   // the store that last read this buffer
        asm volatile("bar.sync 1, 256;" ::: "memory");
#pragma unroll
        for (int u = 0; u < VPT; ++u) { const int v = u * 256 + ct; if (v < nv) od[v] = o[u]; }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        asm volatile("bar.sync 1, 256;" ::: "memory");
        if (ct == 0) { bulk_store(out + off, od, nv * 16); asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
        ob ^= 1;
      } else {
        uint4* og = reinterpret_cast<uint4*>(out + off);
#pragma unroll
        for (int u = 0; u < VPT; ++u) { const int v = u * 256 + ct; if (v < nv) og[v] = o[u]; }
      }
      if (++s == STAGES) { s = 0; ph ^= 1; }
    }
    if (BSTORE && ct == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  }
}

static std::vector<uint16_t> ref_out;

template <typename F>
static void run(const char* name, F launch, int mode, double bytes, uint16_t* out, size_t n_el, const std::vector<uint16_t>& ref, bool check) {
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  for (int i = 0; i < 3; ++i) launch(i);
  CK(cudaDeviceSynchronize());
  const int iters = 10;
  float best = 1e30f, tot = 0;
  for (int i = 0; i < iters; ++i) {
    CK(cudaEventRecord(e0));
    launch(i);
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    float ms;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    if (ms < best) best = ms;
    tot += ms;
  }
  CK(cudaGetLastError());
  const char* ok = "";
  if (check) {
    std::vector<uint16_t> h(4096);
    CK(cudaMemcpy(h.data(), out + n_el - 4096, 4096 * 2, cudaMemcpyDeviceToHost));
    bool same = true;
    for (int i = 0; i < 4096; ++i) same &= h[i] == ref[n_el - 4096 + i];
    CK(cudaMemcpy(h.data(), out, 4096 * 2, cudaMemcpyDeviceToHost));
    for (int i = 0; i < 4096; ++i) same &= h[i] == ref[i];
    ok = same ? " ok" : " MISMATCH";
  }
  printf("%-44s mode %d  best %7.1f us  %7.1f GB/s   mean %7.1f GB/s%s\n", name, mode, best * 1e3, bytes / best * 1e-6, bytes / (tot / iters) * 1e-6, ok);
  fflush(stdout);
}

int main(int argc, char** argv) {
  const int C = argc > 1 ? atoi(argv[1]) : 64;
  const long long npix = argc > 2 ? atoll(argv[2]) : 2LL * 608 * 968;
  const size_t n_el = (size_t)npix * C;
  const long long nvec = (long long)n_el / 8;
  printf("C=%d npix=%lld  tensor %.1f MB\n", C, npix, n_el * 2 / 1e6);
  // two input sets so consecutive launches never re-read what the previous one left in L2 (each tensor > L2 anyway)
  uint16_t *x[2], *dy[2], *out[2];
  std::vector<uint16_t> hx(n_el), hd(n_el);
  uint32_t rs = 12345;
  for (size_t i = 0; i < n_el; ++i) {
    rs = rs * 1664525u + 1013904223u;
    hx[i] = __half_as_ushort(__float2half(((rs >> 8) & 0xffff) / 32768.f - 1.f));
    rs = rs * 1664525u + 1013904223u;
    hd[i] = __half_as_ushort(__float2half(((rs >> 8) & 0xffff) / 32768.f - 1.f));
  }
  for (int i = 0; i < 2; ++i) {
    CK(cudaMalloc(&x[i], n_el * 2)); CK(cudaMalloc(&dy[i], n_el * 2)); CK(cudaMalloc(&out[i], n_el * 2));
    CK(cudaMemcpy(x[i], hx.data(), n_el * 2, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dy[i], hd.data(), n_el * 2, cudaMemcpyHostToDevice));
  }
  std::vector<float> hc(5 * C);
  for (int i = 0; i < 5 * C; ++i) hc[i] = 0.25f + 0.01f * (i % 37) - (i % 3 == 0 ? 0.4f : 0.f);
  float* dc;
  CK(cudaMalloc(&dc, 5 * C * 4));
  CK(cudaMemcpy(dc, hc.data(), 5 * C * 4, cudaMemcpyHostToDevice));
  Consts k{dc, dc + C, dc + 2 * C, dc + 3 * C, dc + 4 * C};
  const double b3 = 3.0 * n_el * 2, b2 = 2.0 * n_el * 2;

  // reference result of mode 0 from the round-1 mapping
  std::vector<uint16_t> ref0(n_el), ref1(n_el);
  k_strided<4, 2, 0, 0><<<148 * 6, 256>>>(x[0], dy[0], out[0], npix, C, k);
  CK(cudaDeviceSynchronize());
  CK(cudaMemcpy(ref0.data(), out[0], n_el * 2, cudaMemcpyDeviceToHost));
  k_strided<4, 2, 0, 1><<<148 * 6, 256>>>(x[0], dy[0], out[0], npix, C, k);
  CK(cudaDeviceSynchronize());
  CK(cudaMemcpy(ref1.data(), out[0], n_el * 2, cudaMemcpyDeviceToHost));

#define RUN(NAME, MODE, EXPR)                                                                                      \
  run(NAME, [&](int i) { const int b = i & 1; (void)b; EXPR; }, MODE, (MODE) == 0 ? b3 : b2, out[1], n_el, (MODE) == 0 ? ref0 : ref1, (MODE) != 2)
  // cudaMemcpy D2D as the copy yardstick
  run("cudaMemcpyAsync d2d", [&](int i) { CK(cudaMemcpyAsync(out[1], x[i & 1], n_el * 2, cudaMemcpyDeviceToDevice)); }, 2, b2, out[1], n_el, ref1, false);

  for (int mode = 0; mode < 3; ++mode) {
    printf("---- mode %d (%s)\n", mode, mode == 0 ? "bwd apply: 2 in 1 out" : mode == 1 ? "fwd apply: 1 in 1 out" : "copy");
#define BOTH(K, NAME, GRID, ...)                                                                                   \
    if (mode == 0) RUN(NAME, 0, (K<__VA_ARGS__, 0><<<GRID, 256>>>(ARGS)));                                         \
    else if (mode == 1) RUN(NAME, 1, (K<__VA_ARGS__, 1><<<GRID, 256>>>(ARGS)));                                    \
    else RUN(NAME, 2, (K<__VA_ARGS__, 2><<<GRID, 256>>>(ARGS)));
#define ARGS x[b], dy[b], out[1], npix, C, k
    BOTH(k_strided, "strided u4 minb2 g148x6 (round 1)", 148 * 6, 4, 2, 0)
    BOTH(k_strided, "strided u4 minb2 g148x2", 148 * 2, 4, 2, 0)
    BOTH(k_strided, "strided u4 minb4 g148x4", 148 * 4, 4, 4, 0)
    BOTH(k_strided, "strided u4 minb4 g148x8", 148 * 8, 4, 4, 0)
    BOTH(k_strided, "strided u8 minb2 g148x2", 148 * 2, 8, 2, 0)
    BOTH(k_strided, "strided u4 minb2 g148x6 cs", 148 * 6, 4, 2, 1)
    BOTH(k_strided, "strided u4 minb2 g148x6 L1 no_allocate", 148 * 6, 4, 2, 2)
#undef ARGS
#define ARGS reinterpret_cast<const uint4*>(x[b]), reinterpret_cast<const uint4*>(dy[b]), reinterpret_cast<uint4*>(out[1]), nvec, C, k
    BOTH(k_contig, "contig u4 minb2 g148x2", 148 * 2, 4, 2, 0)
    BOTH(k_contig, "contig u4 minb2 g148x6", 148 * 6, 4, 2, 0)
    BOTH(k_contig, "contig u4 minb4 g148x4", 148 * 4, 4, 4, 0)
    BOTH(k_contig, "contig u4 minb4 g148x8", 148 * 8, 4, 4, 0)
    BOTH(k_contig, "contig u4 minb4 g148x16", 148 * 16, 4, 4, 0)
    BOTH(k_contig, "contig u2 minb4 g148x8", 148 * 8, 2, 4, 0)
    BOTH(k_contig, "contig u8 minb2 g148x2", 148 * 2, 8, 2, 0)
    BOTH(k_contig, "contig u8 minb2 g148x4", 148 * 4, 8, 2, 0)
    BOTH(k_contig, "contig u4 minb4 g148x4 cs", 148 * 4, 4, 4, 1)
    BOTH(k_contig, "contig u4 minb4 g148x4 L1 no_allocate", 148 * 4, 4, 4, 2)
    BOTH(k_contig, "contig u8 minb2 g148x2 cs", 148 * 2, 8, 2, 1)
    {
      // one block per 16 KB chunk, no grid-stride loop (hardware block scheduler does the balancing)
      const long long nb = (nvec + 256 * 4 - 1) / (256 * 4);
      BOTH(k_contig, "contig u4 minb4 one-chunk-per-block", (unsigned)nb, 4, 4, 0)
      BOTH(k_contig, "contig u4 minb4 one-chunk-per-block cs", (unsigned)nb, 4, 4, 1)
      const long long nb8 = (nvec + 256 * 8 - 1) / (256 * 8);
      BOTH(k_contig, "contig u8 minb2 one-chunk-per-block", (unsigned)nb8, 8, 2, 0)
    }
#undef ARGS
#define ARGS (const uint8_t*)x[b], (const uint8_t*)dy[b], (uint8_t*)out[1], nvec / 2, C, k
    BOTH(k_contig32, "contig32 u2 minb2 g148x2", 148 * 2, 2, 2, 0)
    BOTH(k_contig32, "contig32 u2 minb2 g148x4", 148 * 4, 2, 2, 0)
    BOTH(k_contig32, "contig32 u4 minb2 g148x2", 148 * 2, 4, 2, 0)
    BOTH(k_contig32, "contig32 u2 minb3 g148x3", 148 * 3, 2, 3, 0)
    BOTH(k_contig32, "contig32 u2 minb2 g148x2 evict_first", 148 * 2, 2, 2, 1)
    BOTH(k_contig32, "contig32 u4 minb2 g148x2 evict_first", 148 * 2, 4, 2, 1)
#undef ARGS
#define BULK(NAME, STAGES, TILE, BST)                                                                                   \
    {                                                                                                                   \
      const size_t smem = (size_t)STAGES * (mode == 0 ? 2 : 1) * TILE + (BST ? 2 * TILE : 0) + 2 * STAGES * 8 + 128;     \
      if (mode == 0) { CK(cudaFuncSetAttribute(k_bulk<STAGES, TILE, 0, BST>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
        RUN(NAME, 0, (k_bulk<STAGES, TILE, 0, BST><<<148, 288, smem>>>((const uint8_t*)x[b], (const uint8_t*)dy[b], (uint8_t*)out[1], (long long)n_el * 2, C, k))); } \
      else if (mode == 1) { CK(cudaFuncSetAttribute(k_bulk<STAGES, TILE, 1, BST>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
        RUN(NAME, 1, (k_bulk<STAGES, TILE, 1, BST><<<148, 288, smem>>>((const uint8_t*)x[b], (const uint8_t*)dy[b], (uint8_t*)out[1], (long long)n_el * 2, C, k))); } \
      else { CK(cudaFuncSetAttribute(k_bulk<STAGES, TILE, 2, BST>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
        RUN(NAME, 2, (k_bulk<STAGES, TILE, 2, BST><<<148, 288, smem>>>((const uint8_t*)x[b], (const uint8_t*)dy[b], (uint8_t*)out[1], (long long)n_el * 2, C, k))); } \
    }
    BULK("bulk ring 4 x 16 KB, direct stores", 4, 16384, 0)
    BULK("bulk ring 6 x 16 KB, direct stores", 6, 16384, 0)
    BULK("bulk ring 8 x 8 KB, direct stores", 8, 8192, 0)
    BULK("bulk ring 4 x 16 KB, bulk stores", 4, 16384, 1)
    BULK("bulk ring 8 x 8 KB, bulk stores", 8, 8192, 1)
    BULK("bulk ring 5 x 16 KB, bulk stores", 5, 16384, 1)
  }
  printf("done\n");
  return 0;
}
