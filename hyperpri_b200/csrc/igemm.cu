// Implicit-GEMM convolution / GEMM family on tcgen05 (sm_100a).
//
// One warp-specialised kernel template covers every dense contraction on the hot path:
//   MODE_FWD   D[pixels, n] = sum_k A[pixels (+tap shift), k] * B[n, k]        (K-major operands)
//              1x1 / Linear (models.py:108), ConvTranspose2d k2 s2 fprop (pixel-shuffle store) and
//              dgrad (2x2 stride-2 gather) (model_parts.py:63-64), 3x3 convs whose tiles do not fit the
//              halo kernel below.
//   MODE_WGRAD dW[n, (tap, c)] += sum_pixels X[pixel (+tap shift), c] * dY[pixel, n]   (MN-major operands)
// and conv3x3_halo_kernel runs the 3x3 pad-1 conv fprop and dgrad (model_parts.py:22,25).
//
// A operand tiles arrive by TMA straight from NHWC 16-bit activations: a 4-D box {64 ch, tw, th, 1}
// at (c0, w0+dw, h0+dh, n).  Out-of-bounds box elements are zero-filled by the TMA unit, which
// implements the conv's zero padding, ragged channel counts (238 -> 240 -> 4 chunks of 64) and
// partial tiles at the image border for free.  Rows of 64 elements = 128 B land in the 128B-swizzled
// layout tcgen05.mma consumes.  Accumulators live in TMEM (128 lanes x BLOCK_N fp32 columns).
//
// Warp roles (320 threads): warp 0 = TMA producer, warp 1 = TMEM allocator + MMA issuer (one
// elected thread), warps 2..9 = two epilogue groups of four warps.  A group turns a 128-row x 64-column
// accumulator chunk into 16-bit values in a 128B-swizzled staging tile (TMEM -> registers -> smem),
// one elected thread hands the tile to the TMA unit (cp.async.bulk.tensor store: clipping at the image
// border and at n_store comes from the tensor map), and meanwhile all its threads add the tile's
// per-channel BatchNorm statistics to register accumulators kept across tiles.  Two groups work on
// different chunks at the same time: one warp per scheduler cannot hide the TMEM / smem latencies
// (measured: a 4-warp epilogue took 3.8 k cycles per chunk and bounded every Cout=64 layer).
#include "ptx.cuh"
#include "hyperpri_b200.h"

#include <cstdlib>
#include <mutex>

namespace hpri {

enum { MODE_FWD = 0, MODE_WGRAD = 1 };
enum { TAP_NONE = 0, TAP_3X3 = 1, TAP_UP2 = 2 };

struct IgemmArgs {
  int N, H, W;            // pixel grid walked by M (fwd) or K (wgrad)
  int th, tw, ltw, tiles_h, tiles_w;      // tw is a power of two, ltw = log2(tw)
  int taps, tap_mode, kchunks;
  int n_total;            // logical extent of the GEMM N dimension
  // ---- fwd epilogue
  int up2, cout;          // ConvT pixel shuffle: n = (a*2+b)*cout + co, one output map per (a, b)
  const float* bias;      // [n_total] (up2: [cout]) or null
  int accum;              // 1: y += result (TMA reduce-add in the destination format)
  double* stats;          // [n_total][2] sum / sumsq or null
  // ---- wgrad epilogue
  float* dw;              // [n_total][dw_ld] fp32, accumulated with red.add
  int dw_ld, splits, total_chunks;
  int a_dt, b_dt, out_dt; // element formats (DT_BF16 / DT_F16) of the A, B operands and the fwd output
  // ---- halo kernel pipeline shape (runtime: depends on the tile aspect)
  int stages, a_bytes, a_stages;
  // ---- fused train-mode BatchNorm finalisation (has_fin): done by the last CTA to flush its statistics
  int has_fin;
  hpri_bn_fin_t fin;
  // ---- fused BatchNorm-backward reduction (dgrad launches of the halo kernel): the tile just produced is dy of the
  // layer below; with a TMA-loaded tile of that layer's raw conv output x the epilogue accumulates
  // sum(dz) and sum(dz * x), dz = dy * [x*scale+shift > 0], i.e. pass 1 of hpri_bn_relu_bwd_reduce
  const float *bw_scale, *bw_shift, *bw_mean, *bw_invstd;
  double* bw_sums;        // [n_total][3]
  int xstage_off;         // byte offset of the two 16 KB x staging tiles (after the rings)
};

constexpr int kThreads = 320;
constexpr int kEpiThreads = 256;
constexpr int kMaxStatCh = 2048;   // per-CTA BatchNorm partial sums live in smem for the whole kernel
constexpr int kStageTile = 16384;  // 128 rows x 64 columns x 2 B epilogue staging, one per epilogue group

template <int DT>
__device__ __forceinline__ uint32_t pack2_t(float lo, float hi) {
  if (DT == DT_F16) {
    __half2 h = __floats2half2_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&h);
  }
  return pack_bf16x2(lo, hi);
}
// 32 fp32 accumulator columns -> 16-bit pairs -> swizzled staging row (zeros for invalid rows)
template <int DT>
__device__ __forceinline__ void stage_row32(uint8_t* stage, int row, int hh, const uint32_t (&v)[32], bool valid) {
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    uint4 o = make_uint4(0, 0, 0, 0);
    if (valid) {
      o.x = pack2_t<DT>(__uint_as_float(v[8 * i + 0]), __uint_as_float(v[8 * i + 1]));
      o.y = pack2_t<DT>(__uint_as_float(v[8 * i + 2]), __uint_as_float(v[8 * i + 3]));
      o.z = pack2_t<DT>(__uint_as_float(v[8 * i + 4]), __uint_as_float(v[8 * i + 5]));
      o.w = pack2_t<DT>(__uint_as_float(v[8 * i + 6]), __uint_as_float(v[8 * i + 7]));
    }
    const int chunk = hh * 4 + i;
    *reinterpret_cast<uint4*>(stage + row * 128 + (((chunk ^ row) & 7) << 4)) = o;
  }
}
// Column sums of 32 staged rows (rows r0..r0+31, r0 a multiple of 8): lane l owns channels 2l, 2l+1.
template <int DT>
__device__ __forceinline__ void stats_rows32(const uint8_t* stage, int r0, int lane, float2& s, float2& q) {
  const int chunk = lane >> 2;
  const uint8_t* base = stage + r0 * 128 + (lane & 3) * 4;
#pragma unroll
  for (int i = 0; i < 32; ++i) {
    const uint32_t u = *reinterpret_cast<const uint32_t*>(base + i * 128 + (((chunk ^ i) & 7) << 4));
    acc_sum_sq2(s, q, unpack2_t<DT>(u));
  }
}
// Same walk over two staged tiles (dy and the raw conv output x of the layer below): lane l owns channels 2l, 2l+1;
// s1 += dz, s2 += dz * x with dz = dy where x*scale+shift > 0 (the ReLU mask of the forward pass), else 0.
template <int DT>
__device__ __forceinline__ void bw_stats_rows32(const uint8_t* stage, const uint8_t* xst, int r0, int lane, float2 sc,
                                                float2 sh, float2& s1, float2& s2) {
  const int chunk = lane >> 2;
  const int off0 = r0 * 128 + (lane & 3) * 4;
#pragma unroll
  for (int i = 0; i < 32; ++i) {
    const int off = off0 + i * 128 + (((chunk ^ i) & 7) << 4);
    const float2 d = unpack2_t<DT>(*reinterpret_cast<const uint32_t*>(stage + off));
    const float2 x = unpack2_t<DT>(*reinterpret_cast<const uint32_t*>(xst + off));
    const float dz0 = fmaf(x.x, sc.x, sh.x) > 0.f ? d.x : 0.f;
    const float dz1 = fmaf(x.y, sc.y, sh.y) > 0.f ? d.y : 0.f;
    s1.x += dz0; s1.y += dz1;
    s2.x = fmaf(dz0, x.x, s2.x); s2.y = fmaf(dz1, x.y, s2.y);
  }
}
// TMEM (this warp's 32 lanes, 64 columns at taddr) -> (+bias) -> 16-bit -> staging rows
__device__ __forceinline__ void chunk_to_stage(uint32_t taddr, uint8_t* stage, int row, bool valid, int out_dt,
                                               const float* bias64) {
  uint32_t v[64];
  tmem_ld64(taddr, v);                       // all 64 columns of the chunk in one TMEM round trip
  tmem_ld_wait();
  if (bias64 != nullptr) {
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      const float4 b = __ldg(reinterpret_cast<const float4*>(bias64) + j);
      v[4 * j + 0] = __float_as_uint(__uint_as_float(v[4 * j + 0]) + b.x);
      v[4 * j + 1] = __float_as_uint(__uint_as_float(v[4 * j + 1]) + b.y);
      v[4 * j + 2] = __float_as_uint(__uint_as_float(v[4 * j + 2]) + b.z);
      v[4 * j + 3] = __float_as_uint(__uint_as_float(v[4 * j + 3]) + b.w);
    }
  }
#pragma unroll
  for (int chunk = 0; chunk < 8; ++chunk) {
    uint4 o = make_uint4(0, 0, 0, 0);
    if (valid) {
      if (out_dt == DT_F16) {
        o.x = pack2_t<DT_F16>(__uint_as_float(v[8 * chunk + 0]), __uint_as_float(v[8 * chunk + 1]));
        o.y = pack2_t<DT_F16>(__uint_as_float(v[8 * chunk + 2]), __uint_as_float(v[8 * chunk + 3]));
        o.z = pack2_t<DT_F16>(__uint_as_float(v[8 * chunk + 4]), __uint_as_float(v[8 * chunk + 5]));
        o.w = pack2_t<DT_F16>(__uint_as_float(v[8 * chunk + 6]), __uint_as_float(v[8 * chunk + 7]));
      } else {
        o.x = pack2_t<DT_BF16>(__uint_as_float(v[8 * chunk + 0]), __uint_as_float(v[8 * chunk + 1]));
        o.y = pack2_t<DT_BF16>(__uint_as_float(v[8 * chunk + 2]), __uint_as_float(v[8 * chunk + 3]));
        o.z = pack2_t<DT_BF16>(__uint_as_float(v[8 * chunk + 4]), __uint_as_float(v[8 * chunk + 5]));
        o.w = pack2_t<DT_BF16>(__uint_as_float(v[8 * chunk + 6]), __uint_as_float(v[8 * chunk + 7]));
      }
    }
    *reinterpret_cast<uint4*>(stage + row * 128 + (((chunk ^ row) & 7) << 4)) = o;
  }
}
// per-thread statistics (channels ch, ch+1) -> the CTA's smem partial sums
__device__ __forceinline__ void flush_stats(float* ssum, float* ssq, int ch, int n_total, float2& s, float2& q) {
  if (ch < n_total && (s.x != 0.f || q.x != 0.f)) { atomicAdd(&ssum[ch], s.x); atomicAdd(&ssq[ch], q.x); }
  if (ch + 1 < n_total && (s.y != 0.f || q.y != 0.f)) { atomicAdd(&ssum[ch + 1], s.y); atomicAdd(&ssq[ch + 1], q.y); }
  s = make_float2(0.f, 0.f);
  q = make_float2(0.f, 0.f);
}

// Deterministic variant (hpri_bn_fin_t::partials set): the NW warps that own the same channels add their registers to
// the CTA's partial sums one after the other (plain read-modify-write, lanes own distinct channels) instead of through
// shared-memory atomics in arrival order.  Called by all NW * 32 threads of barrier `bar`.
template <int NW>
__device__ __forceinline__ void flush_stats_ordered(float* ssum, float* ssq, int ch, int n_total, float2& s, float2& q,
                                                    int my_warp, int bar) {
#pragma unroll 1
  for (int w = 0; w < NW; ++w) {
    if (my_warp == w) {
      if (ch < n_total) { ssum[ch] += s.x; ssq[ch] += q.x; }
      if (ch + 1 < n_total) { ssum[ch + 1] += s.y; ssq[ch + 1] += q.y; }
    }
    named_bar_sync(bar, NW * 32);
  }
  s = make_float2(0.f, 0.f);
  q = make_float2(0.f, 0.f);
}

// Called by the 256 epilogue threads after their CTA's statistics went to global memory: the last CTA of the grid
// (ticket counter) turns (sum, sumsq) into scale / shift / saved mean / invstd and the running-stat update -- the
// arithmetic of bn_finalize_k -- and leaves stats and the counter zeroed for the next step.
__device__ __forceinline__ void bn_finalize_tail(const IgemmArgs& p, int et, int* flag_smem) {
  __threadfence();
  named_bar_sync(3, kEpiThreads);
  if (et == 0) *flag_smem = atomicAdd(p.fin.counter, 1u) == gridDim.x - 1 ? 1 : 0;
  named_bar_sync(3, kEpiThreads);
  if (*flag_smem == 0) return;
  __threadfence();
  const hpri_bn_fin_t& f = p.fin;
  for (int c = et; c < p.n_total; c += kEpiThreads) {
    double s, ss;
    if (f.partials != nullptr) {            // deterministic: the CTAs' slots in CTA order
      s = 0.0; ss = 0.0;
      for (unsigned b = 0; b < gridDim.x; ++b) {
        const float2 v = __ldcg(reinterpret_cast<const float2*>(f.partials) + static_cast<size_t>(b) * p.n_total + c);
        s += static_cast<double>(v.x);
        ss += static_cast<double>(v.y);
      }
    } else {
      s = __ldcg(p.stats + 2 * c); ss = __ldcg(p.stats + 2 * c + 1);
    }
    const double mean = s / (double)f.count;
    double var = ss / (double)f.count - mean * mean;
    if (var < 0) var = 0;
    const float g = f.gamma ? f.gamma[c] : 1.f, b = f.beta ? f.beta[c] : 0.f;
    const float cb = f.conv_bias ? f.conv_bias[c] : 0.f;
    const float inv = (float)(1.0 / sqrt(var + (double)f.eps));
    const float sc = g * inv;
    f.scale[c] = sc;
    f.shift[c] = b - (float)mean * sc;
    if (f.save_mean) f.save_mean[c] = (float)mean;
    if (f.save_invstd) f.save_invstd[c] = inv;
    if (f.running_mean) f.running_mean[c] = (1.f - f.momentum) * f.running_mean[c] + f.momentum * ((float)mean + cb);
    if (f.running_var) {
      const double unb = f.count > 1 ? var * (double)f.count / (double)(f.count - 1) : var;
      f.running_var[c] = (1.f - f.momentum) * f.running_var[c] + f.momentum * (float)unb;
    }
    p.stats[2 * c] = 0.0;
    p.stats[2 * c + 1] = 0.0;
  }
  if (et == 0) {
    if (f.num_batches_tracked) *f.num_batches_tracked += 1;
    *f.counter = 0u;
  }
}

// fwd: k-block = 64 channels of one tap: A 128 pixels x 128 B, B BLOCK_N rows x 128 B.
// wgrad: k-block = 128 pixels: A 2 channel chunks x 128 pixel rows x 128 B, B BLOCK_N/64 chunks likewise.
template <int BLOCK_N, int STAGES, int MODE>
struct SmemLayout {
  static constexpr int KPIX = 128;                                         // wgrad pixels per k-block
  static constexpr int A_BYTES = MODE == MODE_FWD ? 16384 : 2 * KPIX * 128;
  static constexpr int B_BYTES = MODE == MODE_FWD ? BLOCK_N * 128 : (BLOCK_N / 64) * KPIX * 128;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int PIPE_BYTES = STAGES * STAGE_BYTES;
  static constexpr int STAGING_OFF = PIPE_BYTES;                 // two epilogue staging tiles (fwd)
  static constexpr int STAGING_BYTES = MODE == MODE_FWD ? 2 * kStageTile : 0;
  static constexpr int BAR_OFF = STAGING_OFF + STAGING_BYTES;    // full[S], empty[S], tmem_full[2], tmem_empty[2]
  static constexpr int TMEMPTR_OFF = BAR_OFF + (2 * STAGES + 4) * 8;
  static constexpr int SSUM_OFF = TMEMPTR_OFF + 8;               // float[kMaxStatCh] sum, float[kMaxStatCh] sumsq
  static constexpr int TOTAL = SSUM_OFF + (MODE == MODE_FWD ? 2 * kMaxStatCh * 4 : 0);
  static constexpr int ALLOC = TOTAL + 1024;                     // slack for manual 1024 B alignment
  static_assert(ALLOC <= 227 * 1024, "shared memory budget exceeded");
};

// Persistent kernel: grid = min(tiles, #SMs); each CTA walks tiles blockIdx.x, +gridDim.x, ...
// The smem ring runs continuously across tiles; the accumulator is double-buffered in TMEM
// (2 x BLOCK_N columns) so the epilogue of tile i overlaps the MMAs of tile i+1.
template <int BLOCK_N, int STAGES, int MODE>
__global__ void __launch_bounds__(kThreads, 1)
igemm_kernel(const __grid_constant__ CUtensorMap tmA0, const __grid_constant__ CUtensorMap tmA1,
             const __grid_constant__ CUtensorMap tmB0, const __grid_constant__ CUtensorMap tmB1,
             const __grid_constant__ CUtensorMap tmO0, const __grid_constant__ CUtensorMap tmO1,
             const __grid_constant__ CUtensorMap tmO2, const __grid_constant__ CUtensorMap tmO3,
             const IgemmArgs p) {
  using L = SmemLayout<BLOCK_N, STAGES, MODE>;
  constexpr int A_BYTES = L::A_BYTES;
  constexpr int NCH = BLOCK_N / 64;                 // 64-column chunks per accumulator
  constexpr int GROUPS = MODE == MODE_WGRAD ? 2 : (NCH >= 2 ? 2 : 1);   // epilogue groups with work
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + L::BAR_OFF);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full = empty_bar + STAGES;      // [2]
  uint64_t* tmem_empty = tmem_full + 2;          // [2]
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(smem + L::TMEMPTR_OFF);
  float* ssum = reinterpret_cast<float*>(smem + L::SSUM_OFF);
  float* ssq = ssum + kMaxStatCh;

  // lane-0 broadcast: makes the role dispatch below provably warp-uniform, which is what lets ptxas keep the producer
  // and MMA warps on the uniform datapath (with a plain threadIdx.x >> 5 it guards every UTCHMMA with ELECT + R2UR)
  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
  const int lane = threadIdx.x & 31;

  // ---------------- tile space
  const int n_tiles = (p.n_total + BLOCK_N - 1) / BLOCK_N;
  const int tiles_per_img = p.tiles_h * p.tiles_w;
  const long long T = static_cast<long long>(p.N) * tiles_per_img;        // pixel tiles
  const int m_tiles_w = (p.total_chunks + 1) / 2;                          // wgrad: M tiles (pairs of 64-ch chunks)
  const long long total_tiles =
      MODE == MODE_FWD ? T * n_tiles : static_cast<long long>(m_tiles_w) * n_tiles * p.splits;

  // ---------------- one-time setup
  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmA0);
    tma_prefetch_desc(&tmB0);
    if (MODE == MODE_FWD) tma_prefetch_desc(&tmO0);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(&tmem_full[0], 1);
    mbar_init(&tmem_full[1], 1);
    mbar_init(&tmem_empty[0], 4 * GROUPS);     // one arrival per epilogue warp that reads the accumulator
    mbar_init(&tmem_empty[1], 4 * GROUPS);
    fence_mbar_init();
    fence_proxy_async_smem();
  }
  if (warp == 1) {
    tmem_alloc(tmem_ptr, 2 * BLOCK_N);
    tmem_relinquish();
  }
  if (MODE == MODE_FWD)
    for (int i = threadIdx.x; i < 2 * kMaxStatCh; i += kThreads) ssum[i] = 0.f;
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  grid_dep_launch();
  grid_dep_wait();            // everything above overlapped the previous kernel's tail; its outputs are visible from here

  // decode a tile index into (m_tile, n_tile, [kb_begin, kb_end))
  auto decode = [&](long long tile, int& m_tile, int& n_tile, int& kb_begin, int& kb_end) {
    if (MODE == MODE_FWD) {
      n_tile = static_cast<int>(tile % n_tiles);
      m_tile = static_cast<int>(tile / n_tiles);
      kb_begin = 0;
      kb_end = p.taps * p.kchunks;
    } else {
      const int per_split = m_tiles_w * n_tiles;
      const int split = static_cast<int>(tile / per_split);
      const int r = static_cast<int>(tile - static_cast<long long>(split) * per_split);
      n_tile = r % n_tiles;
      m_tile = r / n_tiles;
      kb_begin = static_cast<int>(T * split / p.splits);
      kb_end = static_cast<int>(T * (split + 1) / p.splits);
    }
  };

  if (warp == 0) {
    // =========================== TMA producer (whole warp converged, one elected lane issues) ===========
    {
      const uint32_t pipe_u = uniform_u32(smem_u32(smem));
      const uint32_t full_u = uniform_u32(smem_u32(full_bar)), empty_u = full_u + STAGES * 8;
      uint32_t s = 0, ph = 0;
      for (long long tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        int m_tile, n_tile, kb_begin, kb_end;
        decode(tile, m_tile, n_tile, kb_begin, kb_end);
        const int n0 = n_tile * BLOCK_N;
        int img = 0, h0 = 0, w0 = 0;
        if (MODE == MODE_FWD) {
          img = m_tile / tiles_per_img;
          const int r = m_tile - img * tiles_per_img;
          h0 = (r / p.tiles_w) * p.th;
          w0 = (r % p.tiles_w) * p.tw;
        }
        for (int kb = kb_begin; kb < kb_end; ++kb) {
          mbar_wait_w(empty_u + s * 8, ph ^ 1);
          const uint32_t fb = full_u + s * 8;
          mbar_arrive_expect_tx_w(fb, L::STAGE_BYTES);
          const uint32_t sA = pipe_u + s * L::STAGE_BYTES;
          const uint32_t sB = sA + A_BYTES;
          if (MODE == MODE_FWD) {
            const int tap = kb / p.kchunks;
            const int cc = kb - tap * p.kchunks;
            if (p.tap_mode == TAP_UP2) {
              // ConvT dgrad: gather dy[2h+a, 2w+b]; 5-D view (c, w, a, h, n), one map per b
              tma_load_5d_w(sA, (tap & 1) ? &tmA1 : &tmA0, fb, cc * 64, w0, tap >> 1, h0, img);
            } else {
              int dh = 0, dw = 0;
              if (p.tap_mode == TAP_3X3) { dh = tap / 3 - 1; dw = tap % 3 - 1; }
              tma_load_4d_w(sA, &tmA0, fb, cc * 64, w0 + dw, h0 + dh, img);
            }
            tma_load_2d_w(sB, &tmB0, fb, kb * 64, n0);
          } else {
            // wgrad: k-block = one pixel tile of 128 pixels
            const int im = kb / tiles_per_img;
            const int r = kb - im * tiles_per_img;
            const int ph0 = (r / p.tiles_w) * p.th;
            const int pw0 = (r % p.tiles_w) * p.tw;
#pragma unroll
            for (int half = 0; half < 2; ++half) {
              const int q = 2 * m_tile + half;
              int tap = q / p.kchunks;
              int cc = q - tap * p.kchunks;
              int dh = 0, dw = 0;
              if (p.tap_mode == TAP_3X3) { dh = tap / 3 - 1; dw = tap % 3 - 1; }
              if (q >= p.total_chunks) cc = 0x100000;   // fully out of bounds -> zero fill
              tma_load_4d_w(sA + half * 16384, &tmA0, fb, cc * 64, pw0 + dw, ph0 + dh, im);
            }
#pragma unroll
            for (int j = 0; j < BLOCK_N / 64; ++j) {
              if (p.tap_mode == TAP_UP2) {
                const int ab = n0 / p.cout;
                const int co0 = n0 - ab * p.cout;
                tma_load_5d_w(sB + j * 16384, (ab & 1) ? &tmB1 : &tmB0, fb, co0 + j * 64, pw0, ab >> 1, ph0, im);
              } else {
                tma_load_4d_w(sB + j * 16384, &tmB0, fb, n0 + j * 64, pw0, ph0, im);
              }
            }
          }
          if (++s == STAGES) { s = 0; ph ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // =========================== MMA issuer (whole warp converged, one elected lane issues) ============
    {
      const uint32_t idesc = make_idesc_16(128, BLOCK_N, MODE == MODE_WGRAD, MODE == MODE_WGRAD, p.a_dt, p.b_dt);
      const uint32_t pipe_u = uniform_u32(smem_u32(smem));
      const uint32_t full_u = uniform_u32(smem_u32(full_bar)), empty_u = full_u + STAGES * 8;
      const uint32_t tfull_u = empty_u + STAGES * 8, tempty_u = tfull_u + 16;
      const uint32_t tmem_u = uniform_u32(tmem_base);
      uint32_t s = 0, ph = 0, lt = 0;
      for (long long tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        int m_tile, n_tile, kb_begin, kb_end;
        decode(tile, m_tile, n_tile, kb_begin, kb_end);
        if (kb_end <= kb_begin) continue;
        const uint32_t as = lt & 1, aph = (lt >> 1) & 1;
        ++lt;
        mbar_wait_w(tempty_u + as * 8, aph ^ 1);   // epilogue has drained this accumulator buffer
        tc_fence_after();
        const uint32_t tmem_d = tmem_u + as * BLOCK_N;
        for (int kb = kb_begin; kb < kb_end; ++kb) {
          mbar_wait_w(full_u + s * 8, ph);
          tc_fence_after();
          const uint32_t a_addr = pipe_u + s * L::STAGE_BYTES;
          const uint32_t b_addr = a_addr + A_BYTES;
          // descriptor low words (start address >> 4, LBO) advance per UMMA_K = 16; high words are constant
          const uint64_t d0 = MODE == MODE_FWD ? make_smem_desc_sw128(0, 16, 1024)
                                               : make_smem_desc_sw128(0, 16384, 1024);   // wgrad: LBO = next 64-channel group
          const uint32_t d_hi = static_cast<uint32_t>(d0 >> 32);
          const uint32_t a_lo = static_cast<uint32_t>(d0) | (a_addr >> 4), b_lo = static_cast<uint32_t>(d0) | (b_addr >> 4);
          constexpr uint32_t kstep = MODE == MODE_FWD ? (32 >> 4) : (2048 >> 4);   // 16 elements / 16 pixel rows
          const uint32_t acc0 = kb > kb_begin ? 1u : 0u;
          umma_f16_off_w<0 * kstep, 0 * kstep>(tmem_d, a_lo, d_hi, b_lo, d_hi, idesc, acc0);
          umma_f16_off_w<1 * kstep, 1 * kstep>(tmem_d, a_lo, d_hi, b_lo, d_hi, idesc, 1u);
          umma_f16_off_w<2 * kstep, 2 * kstep>(tmem_d, a_lo, d_hi, b_lo, d_hi, idesc, 1u);
          umma_f16_off_w<3 * kstep, 3 * kstep>(tmem_d, a_lo, d_hi, b_lo, d_hi, idesc, 1u);
          if (MODE == MODE_WGRAD) {
            umma_f16_off_w<4 * kstep, 4 * kstep>(tmem_d, a_lo, d_hi, b_lo, d_hi, idesc, 1u);
            umma_f16_off_w<5 * kstep, 5 * kstep>(tmem_d, a_lo, d_hi, b_lo, d_hi, idesc, 1u);
            umma_f16_off_w<6 * kstep, 6 * kstep>(tmem_d, a_lo, d_hi, b_lo, d_hi, idesc, 1u);
            umma_f16_off_w<7 * kstep, 7 * kstep>(tmem_d, a_lo, d_hi, b_lo, d_hi, idesc, 1u);
          }
          umma_commit_w(empty_u + s * 8);  // frees the smem slot once these MMAs have read it
          if (++s == STAGES) { s = 0; ph ^= 1; }
        }
        umma_commit_w(tfull_u + as * 8);   // accumulator complete
      }
    }
  } else {
    // =========================== epilogue (warps 2..9, two groups) ===========================
    const int g = (warp - 2) >> 2;          // group: takes the 64-column chunks c64 with (c64 & 1) == g
    const int q = warp & 3;                 // TMEM lane quarter this warp may read
    const int row = q * 32 + lane;
    const int gw = (warp - 2) & 3;          // warp within the group
    const bool elected = gw == 0 && lane == 0;
    uint8_t* stage = smem + L::STAGING_OFF + g * kStageTile;
    const int bar_id = 1 + g;
    const bool det = p.has_fin && p.fin.partials != nullptr;     // deterministic BatchNorm statistics
    uint32_t lt = 0;
    if (MODE == MODE_FWD) {
      constexpr int MYCH = (NCH + 1) / 2;   // chunks of one accumulator this group handles (at most)
      float2 ssv[MYCH], sqv[MYCH];
#pragma unroll
      for (int j = 0; j < MYCH; ++j) { ssv[j] = make_float2(0.f, 0.f); sqv[j] = make_float2(0.f, 0.f); }
      int acc_n0 = -1;
      if (g < GROUPS) {
        for (long long tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
          int m_tile, n_tile, kb_begin, kb_end;
          decode(tile, m_tile, n_tile, kb_begin, kb_end);
          if (kb_end <= kb_begin) continue;
          const uint32_t as = lt & 1, aph = (lt >> 1) & 1;
          ++lt;
          const int n0 = n_tile * BLOCK_N;
          if (p.stats != nullptr && n0 != acc_n0) {       // this CTA moved to other output channels: flush
            if (acc_n0 >= 0) {
#pragma unroll
              for (int j = 0; j < MYCH; ++j) {
                if (det) flush_stats_ordered<4>(ssum, ssq, acc_n0 + (2 * j + g) * 64 + 2 * lane, p.n_total, ssv[j], sqv[j], gw, bar_id);
                else flush_stats(ssum, ssq, acc_n0 + (2 * j + g) * 64 + 2 * lane, p.n_total, ssv[j], sqv[j]);
              }
            }
            acc_n0 = n0;
          }
          const uint32_t tmem_acc = tmem_base + as * BLOCK_N + (static_cast<uint32_t>(q * 32) << 16);
          const int img = m_tile / tiles_per_img;
          const int rr = m_tile - img * tiles_per_img;
          const int h0 = (rr / p.tiles_w) * p.th;
          const int w0 = (rr % p.tiles_w) * p.tw;
          const bool valid = (h0 + (row >> p.ltw) < p.H) && (w0 + (row & (p.tw - 1)) < p.W);
          mbar_wait(&tmem_full[as], aph);
          tc_fence_after();
#pragma unroll
          for (int j = 0; j < MYCH; ++j) {
            const int c64 = 2 * j + g;
            if (c64 < NCH) {
              const int nch = n0 + c64 * 64;            // first GEMM column of this chunk
              int co = nch, ab = 0;
              if (p.up2) { ab = nch / p.cout; co = nch - ab * p.cout; }
              if (elected) bulk_wait_read0();            // the previous store has finished reading the staging tile
              named_bar_sync(bar_id, 128);
              chunk_to_stage(tmem_acc + c64 * 64, stage, row, valid, p.out_dt, p.bias != nullptr ? p.bias + co : nullptr);
              if (c64 + 2 >= NCH) {                      // this warp's last TMEM read of the tile: release the buffer
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&tmem_empty[as]);
              }
              fence_proxy_async_smem();
              named_bar_sync(bar_id, 128);
              if (elected && co < (p.up2 ? p.cout : p.n_total)) {
                const CUtensorMap* mo = ab == 0 ? &tmO0 : ab == 1 ? &tmO1 : ab == 2 ? &tmO2 : &tmO3;
                if (p.accum) tma_reduce_add_4d(mo, stage, co, w0, h0, img);
                else tma_store_4d(mo, stage, co, w0, h0, img);
                bulk_commit();
              }
              if (p.stats != nullptr) {
                if (p.out_dt == DT_F16) stats_rows32<DT_F16>(stage, gw * 32, lane, ssv[j], sqv[j]);
                else stats_rows32<DT_BF16>(stage, gw * 32, lane, ssv[j], sqv[j]);
              }
            }
          }
        }
        if (p.stats != nullptr && acc_n0 >= 0) {
#pragma unroll
          for (int j = 0; j < MYCH; ++j) {
            if (det) flush_stats_ordered<4>(ssum, ssq, acc_n0 + (2 * j + g) * 64 + 2 * lane, p.n_total, ssv[j], sqv[j], gw, bar_id);
            else flush_stats(ssum, ssq, acc_n0 + (2 * j + g) * 64 + 2 * lane, p.n_total, ssv[j], sqv[j]);
          }
        }
        if (elected) bulk_wait_read0();
      }
    } else {
      // wgrad: rows = (chunk half, channel j); columns = n.  fp32 red.add into dW[n][k]; the groups split the
      // 32-column slices of the accumulator.
      for (long long tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        int m_tile, n_tile, kb_begin, kb_end;
        decode(tile, m_tile, n_tile, kb_begin, kb_end);
        if (kb_end <= kb_begin) continue;
        const uint32_t as = lt & 1, aph = (lt >> 1) & 1;
        ++lt;
        const int n0 = n_tile * BLOCK_N;
        const uint32_t tmem_acc = tmem_base + as * BLOCK_N + (static_cast<uint32_t>(q * 32) << 16);
        mbar_wait(&tmem_full[as], aph);
        tc_fence_after();
        const int qc = 2 * m_tile + (row >> 6);
        const bool row_ok = qc < p.total_chunks;
        const long long kidx = static_cast<long long>(qc) * 64 + (row & 63);
#pragma unroll 1
        for (int c = g; c < BLOCK_N / 32; c += 2) {
          uint32_t v[32];
          tmem_ld32(tmem_acc + c * 32, v);
          tmem_ld_wait();
          if (c + 2 >= BLOCK_N / 32) {
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tmem_empty[as]);
          }
          if (row_ok) {
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              const int n = n0 + c * 32 + j;
              if (n < p.n_total) red_add_f32(p.dw + static_cast<long long>(n) * p.dw_ld + kidx, __uint_as_float(v[j]));
            }
          }
        }
      }
    }
    if (MODE == MODE_FWD && p.stats != nullptr) {
      // one flush per CTA: per-channel partial sums of every tile this CTA produced -> fp64 global atomics
      named_bar_sync(3, kEpiThreads);
      for (int ch = threadIdx.x - 64; ch < p.n_total; ch += kEpiThreads) {
        const float a = ssum[ch], b = ssq[ch];
        if (det) {
          reinterpret_cast<float2*>(p.fin.partials)[static_cast<size_t>(blockIdx.x) * p.n_total + ch] = make_float2(a, b);
        } else if (a != 0.f || b != 0.f) {
          atomicAdd(p.stats + 2 * ch, static_cast<double>(a));
          atomicAdd(p.stats + 2 * ch + 1, static_cast<double>(b));
        }
      }
      if (p.has_fin) bn_finalize_tail(p, threadIdx.x - 64, reinterpret_cast<int*>(tmem_ptr + 1));
    }
  }
  __syncwarp();
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 2 * BLOCK_N);
  }
}

// =====================================================================================
// 3x3 convolution with full halo reuse (fprop and dgrad of the U-Net blocks).
//
// The generic kernel above re-reads the activation tile from L2 once per filter tap (9x) and the
// weight tile once per 128 output pixels; measured on B200 every layer then sits on the L2->SM
// bandwidth (10-13 TB/s) instead of the tensor pipe.  Here one CTA owns a 256-pixel tile made of two
// M=128 halves of 16 rows x 8 columns (stacked: 32x8 tile, or side by side: 16x16 tile) and, per
// 64-channel chunk, loads the (th+2) x (tw+2) halo block ONCE.  All nine taps are descriptor start
// offsets into that block: pixel (y+r, x+s) of the block is smem row (y+r)*(tw+2) + x+s, the eight
// pixels of one tile row are eight consecutive 128-byte rows (one swizzle group), and consecutive tile
// rows are (tw+2)*128 B apart (the descriptor's SBO).  tcgen05.mma applies the 128B swizzle to absolute
// shared-memory address bits (tools/probes/umma_rowshift_probe.cu: any 128-byte row offset and
// SBO = 1152 / 1280 read back exactly), so the TMA-written block is consumed in place.
// Activation bytes per MAC drop 3x against per-shift loads, 6.6x against per-tap loads.
// Two TMA rings: the A ring holds halo blocks (one per channel chunk, alive for nine taps), the
// B ring one weight tap tile per slot.
// PAIR = true (the default launch): a cluster of two CTAs on one TPC walks the list of pixel-tile pairs together and
// the leader issues tcgen05.mma.cta_group::2 with M = 256 -- rows 0..127 are CTA 0's pixels, 128..255 CTA 1's, each
// landing in its own CTA's TMEM.  Each CTA stages only HALF of every weight tile (rows rank*N/2 ...), so the shared-
// memory operand reads per MMA drop from 4 KB + N*32 B to 4 KB + N*16 B per SM: the N = 64 MMA goes from 48 to 43
// cycles (tools/probes/umma2_rate_probe.cu) and the weight fill traffic halves.  Both CTAs' loads complete on the
// leader's full barriers (cp.async.bulk.tensor ... cta_group::2 with the barrier address mapped to rank 0), commits are
// multicast to the same barrier offset in both CTAs, the peer's epilogue warps arrive remotely on the leader's
// tmem_empty.  The cluster rank must ride in the same lane-0 broadcast as the warp index: read on its own it is not
// treated as warp-uniform and every UTCHMMA ends up in an ELECT / BRA.U.ANY loop again.  Accumulators: 2 buffers x 2 halves x BLOCK_N TMEM
// columns; epilogue group g owns half g.
// Shared memory: [staging 2 x 16 KB][barriers][stats][A ring: 2 x a_bytes][B ring: p.stages x 3 weight tiles].
// =====================================================================================
constexpr int kHaloStatCh = 1024;
constexpr int kHaloAStages = 2;        // default depth of the halo-block ring
constexpr int kHaloMaxAStages = 3;     // barrier slots reserved (IgemmArgs::a_stages picks 2 or 3 per launch)
constexpr int kHaloMaxBStages = 12;     // weight ring slots: ONE tap tile each (fine-grained: more bytes in flight)

template <int BLOCK_N>
struct HaloSmem {
  static constexpr int B_TILE = BLOCK_N * 128;
  static constexpr int B_STAGE = B_TILE;
  static constexpr int STAGING_OFF = 0;
  static constexpr int BAR_OFF = 2 * kStageTile;       // a_full[2], a_empty[2], b_full[12], b_empty[12], tmem_full[2], tmem_empty[2], x_full[2]
  static constexpr int TMEMPTR_OFF = BAR_OFF + (2 * kHaloMaxAStages + 2 * kHaloMaxBStages + 6) * 8;
  static constexpr int SSUM_OFF = TMEMPTR_OFF + 8;
  static constexpr int PIPE_OFF = (SSUM_OFF + 2 * kHaloStatCh * 4 + 1023) / 1024 * 1024;
  static constexpr int BUDGET = 227 * 1024 - 1024;                       // after manual 1024 B alignment
  static int b_stages_for(int a_bytes, int extra = 0) {
    int s = (BUDGET - PIPE_OFF - kHaloAStages * a_bytes - extra) / B_STAGE;
    return s > kHaloMaxBStages ? kHaloMaxBStages : s;
  }
};

struct HaloGeom {
  int wb;                 // halo block width in pixels (tw + 2)
  int hr[2], hc[2];       // tile-relative origin (row, column) of the two 16 x 8 halves
};

template <bool PAIR, uint32_t OFF_A, uint32_t OFF_B>
__device__ __forceinline__ void halo_mma(uint32_t tmem_d, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi,
                                         uint32_t idesc, uint32_t accumulate) {
  if constexpr (PAIR) umma2_f16_off_w<OFF_A, OFF_B>(tmem_d, a_lo, a_hi, b_lo, b_hi, idesc, accumulate);
  else umma_f16_off_w<OFF_A, OFF_B>(tmem_d, a_lo, a_hi, b_lo, b_hi, idesc, accumulate);
}
template <bool PAIR>
__device__ __forceinline__ void halo_commit(uint32_t bar) {
  if constexpr (PAIR) umma2_commit_w(bar);
  else umma_commit_w(bar);
}

template <int BLOCK_N, bool PAIR>
__global__ void __launch_bounds__(kThreads, 1)
conv3x3_halo_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                    const __grid_constant__ CUtensorMap tmO, const __grid_constant__ CUtensorMap tmX,
                    const IgemmArgs p, const HaloGeom geo) {
  using L = HaloSmem<BLOCK_N>;
  constexpr int NCH = BLOCK_N / 64;
  constexpr uint32_t B_SLOT = PAIR ? L::B_TILE / 2 : L::B_TILE;   // a CTA of a pair holds half of the weight rows
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint64_t* a_full = reinterpret_cast<uint64_t*>(smem + L::BAR_OFF);
  uint64_t* a_empty = a_full + kHaloMaxAStages;
  uint64_t* b_full = a_empty + kHaloMaxAStages;
  const int AST = p.a_stages;                 // halo-block ring depth of this launch (2 or 3)
  uint64_t* b_empty = b_full + kHaloMaxBStages;
  uint64_t* tmem_full = b_empty + kHaloMaxBStages;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(smem + L::TMEMPTR_OFF);
  float* ssum = reinterpret_cast<float*>(smem + L::SSUM_OFF);
  float* ssq = ssum + kHaloStatCh;
  const uint32_t a_bytes = static_cast<uint32_t>(p.a_bytes);   // halo block bytes rounded up to 1024
  uint8_t* a_ring = smem + L::PIPE_OFF;

  // lane-0 broadcast: makes the role dispatch below provably warp-uniform, which is what lets ptxas keep the producer
  // and MMA warps on the uniform datapath (with a plain threadIdx.x >> 5 it guards every UTCHMMA with ELECT + R2UR)
  // (the cluster rank rides in the same broadcast: a separately read %cluster_ctarank is not treated as warp-uniform)
  const uint32_t wr = __shfl_sync(0xffffffffu, (threadIdx.x >> 5) | (PAIR ? cluster_ctarank() << 8 : 0u), 0);
  const int warp = static_cast<int>(wr & 0xffu);
  const int lane = threadIdx.x & 31;
  const int n_tiles = (p.n_total + BLOCK_N - 1) / BLOCK_N;
  const int tiles_per_img = p.tiles_h * p.tiles_w;
  // PAIR: the two CTAs of a cluster walk the list of pixel-tile PAIRS together (CTA r owns pixel tile 2*pair + r; an
  // odd tile count leaves the last pair's second tile past the last image: zero-filled loads, nothing stored)
  const uint32_t crank = wr >> 8;
  const long long m_tiles = static_cast<long long>(p.N) * tiles_per_img;
  const long long total_tiles = (PAIR ? (m_tiles + 1) / 2 : m_tiles) * n_tiles;
  const long long tile_first = PAIR ? (blockIdx.x >> 1) : blockIdx.x, tile_step = PAIR ? (gridDim.x >> 1) : gridDim.x;
  const int BST = p.stages;
  const uint32_t a_tx = static_cast<uint32_t>((p.th + 2) * geo.wb * 128);

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    tma_prefetch_desc(&tmO);
    for (int s = 0; s < kHaloMaxAStages; ++s) { mbar_init(&a_full[s], 1); mbar_init(&a_empty[s], 1); }
    for (int s = 0; s < BST; ++s) { mbar_init(&b_full[s], 1); mbar_init(&b_empty[s], 1); }
    mbar_init(&tmem_full[0], 1);
    mbar_init(&tmem_full[1], 1);
    mbar_init(&tmem_empty[0], PAIR ? 16 : 8);     // the leader's MMA warp waits for both CTAs' epilogue warps
    mbar_init(&tmem_empty[1], PAIR ? 16 : 8);
    mbar_init(&tmem_empty[2], 1);              // x_full[0], x_full[1]: the fused BN-backward x tiles of the two groups
    mbar_init(&tmem_empty[3], 1);
    fence_mbar_init();
    fence_proxy_async_smem();
  }
  if (warp == 1) {
    if (PAIR) { tmem_alloc2(tmem_ptr, 4 * BLOCK_N); tmem_relinquish2(); }
    else { tmem_alloc(tmem_ptr, 4 * BLOCK_N); tmem_relinquish(); }
  }
  for (int i = threadIdx.x; i < 2 * kHaloStatCh; i += kThreads) ssum[i] = 0.f;
  tc_fence_before();
  __syncthreads();
  if (PAIR) cluster_sync_all();           // the peer's barriers are initialised before anything signals them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  grid_dep_launch();
  grid_dep_wait();            // everything above overlapped the previous kernel's tail; its outputs are visible from here

  if (warp == 0) {
    // =========================== TMA producer (whole warp converged) ===========================
    {
      const uint32_t a_ring_u = uniform_u32(smem_u32(a_ring)), b_ring_u = a_ring_u + AST * a_bytes;
      const uint32_t afull_u = uniform_u32(smem_u32(a_full)), aempty_u = afull_u + kHaloMaxAStages * 8;
      const uint32_t bfull_u = aempty_u + kHaloMaxAStages * 8, bempty_u = bfull_u + kHaloMaxBStages * 8;
      // loads of both CTAs complete on the LEADER's full barriers, which expect the bytes of both
      const uint32_t afull_l = PAIR ? mapa_u32(afull_u, 0) : afull_u, bfull_l = PAIR ? mapa_u32(bfull_u, 0) : bfull_u;
      const bool expect = !PAIR || crank == 0;
      uint32_t sa = 0, pha = 0, sb = 0, phb = 0;
      for (long long tile = tile_first; tile < total_tiles; tile += tile_step) {
        const int n_tile = static_cast<int>(tile % n_tiles);
        const int m_tile = static_cast<int>(tile / n_tiles) * (PAIR ? 2 : 1) + static_cast<int>(crank);
        const int img = m_tile / tiles_per_img;
        const int rr = m_tile - img * tiles_per_img;
        const int h0 = (rr / p.tiles_w) * p.th, w0 = (rr % p.tiles_w) * p.tw;
        const int n0 = n_tile * BLOCK_N;
        for (int cc = 0; cc < p.kchunks; ++cc) {
          mbar_wait_w(aempty_u + sa * 8, pha ^ 1);
          if (expect) mbar_arrive_expect_tx_w(afull_u + sa * 8, PAIR ? 2 * a_tx : a_tx);
          if (PAIR) tma_load_4d_w2(a_ring_u + sa * a_bytes, &tmA, afull_l + sa * 8, cc * 64, w0 - 1, h0 - 1, img);
          else tma_load_4d_w(a_ring_u + sa * a_bytes, &tmA, afull_u + sa * 8, cc * 64, w0 - 1, h0 - 1, img);
          if (++sa == static_cast<uint32_t>(AST)) { sa = 0; pha ^= 1; }
#pragma unroll 1
          for (int tap = 0; tap < 9; ++tap) {
            mbar_wait_w(bempty_u + sb * 8, phb ^ 1);
            if (expect) mbar_arrive_expect_tx_w(bfull_u + sb * 8, L::B_TILE);
            if (PAIR) tma_load_2d_w2(b_ring_u + sb * B_SLOT, &tmB, bfull_l + sb * 8, (tap * p.kchunks + cc) * 64,
                                     n0 + static_cast<int>(crank) * (BLOCK_N / 2));
            else tma_load_2d_w(b_ring_u + sb * B_SLOT, &tmB, bfull_u + sb * 8, (tap * p.kchunks + cc) * 64, n0);
            if (++sb == static_cast<uint32_t>(BST)) { sb = 0; phb ^= 1; }
          }
        }
      }
      if (PAIR) {      // tail: every multicast commit aimed at this CTA's empty barriers has landed before it may exit
        for (int i = 0; i < AST; ++i) {
          mbar_wait_w(aempty_u + sa * 8, pha ^ 1);
          if (++sa == static_cast<uint32_t>(AST)) { sa = 0; pha ^= 1; }
        }
        for (int i = 0; i < BST; ++i) {
          mbar_wait_w(bempty_u + sb * 8, phb ^ 1);
          if (++sb == static_cast<uint32_t>(BST)) { sb = 0; phb ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // =========================== MMA issuer (whole warp converged; the leader CTA of a pair) ===========================
    if (!PAIR || crank == 0) {
      const uint32_t idesc = make_idesc_16(PAIR ? 256 : 128, BLOCK_N, 0, 0, p.a_dt, p.b_dt);
      const uint32_t sbo = static_cast<uint32_t>(geo.wb * 128);
      const uint32_t a_ring_u = uniform_u32(smem_u32(a_ring)), b_ring_u = a_ring_u + AST * a_bytes;
      const uint32_t afull_u = uniform_u32(smem_u32(a_full)), aempty_u = afull_u + kHaloMaxAStages * 8;
      const uint32_t bfull_u = aempty_u + kHaloMaxAStages * 8, bempty_u = bfull_u + kHaloMaxBStages * 8;
      const uint32_t tfull_u = bempty_u + kHaloMaxBStages * 8, tempty_u = tfull_u + 16;
      const uint32_t tmem_u = uniform_u32(tmem_base);
      // descriptor offsets (16-byte units) of the two halves' first pixel inside the halo block
      const uint32_t half_off0 = static_cast<uint32_t>((geo.hr[0] * geo.wb + geo.hc[0]) * 8);
      const uint32_t half_off1 = static_cast<uint32_t>((geo.hr[1] * geo.wb + geo.hc[1]) * 8);
      const uint64_t dA = make_smem_desc_sw128(0, 16, sbo), dB = make_smem_desc_sw128(0, 16, 1024);
      const uint32_t a_hi = static_cast<uint32_t>(dA >> 32), b_hi = static_cast<uint32_t>(dB >> 32);
      uint32_t sa = 0, pha = 0, sb = 0, phb = 0, lt = 0;
      for (long long tile = tile_first; tile < total_tiles; tile += tile_step, ++lt) {
        const uint32_t as = lt & 1, aph = (lt >> 1) & 1;
        mbar_wait_w(tempty_u + as * 8, aph ^ 1);
        tc_fence_after();
        const uint32_t tmem_d = tmem_u + as * (2 * BLOCK_N);
        for (int cc = 0; cc < p.kchunks; ++cc) {
          mbar_wait_w(afull_u + sa * 8, pha);
          const uint32_t a_lo0 = static_cast<uint32_t>(dA) | ((a_ring_u + sa * a_bytes) >> 4);
#pragma unroll 1
          for (int r = 0; r < 3; ++r) {
            // two base descriptors per filter row (halo block row r for each half); the horizontal shift
            // (8 x 16 B per pixel) and the K step are compile-time offsets; one weight-ring slot per tap
            const uint32_t a_r = a_lo0 + static_cast<uint32_t>(r * geo.wb * 8);
            const uint32_t aH0 = a_r + half_off0, aH1 = a_r + half_off1;
            const uint32_t t0 = tmem_d, t1 = tmem_d + BLOCK_N;
            const uint32_t acc0 = (cc > 0 || r > 0) ? 1u : 0u;
#define HPRI_TAP(S, ACC)                                                                          \
            {                                                                                     \
              mbar_wait_w(bfull_u + sb * 8, phb);                                                 \
              tc_fence_after();                                                                   \
              const uint32_t b_lo = static_cast<uint32_t>(dB) | ((b_ring_u + sb * B_SLOT) >> 4); \
              halo_mma<PAIR, S * 8 + 0, 0>(t0, aH0, a_hi, b_lo, b_hi, idesc, ACC);                  \
              halo_mma<PAIR, S * 8 + 2, 2>(t0, aH0, a_hi, b_lo, b_hi, idesc, 1u);                   \
              halo_mma<PAIR, S * 8 + 4, 4>(t0, aH0, a_hi, b_lo, b_hi, idesc, 1u);                   \
              halo_mma<PAIR, S * 8 + 6, 6>(t0, aH0, a_hi, b_lo, b_hi, idesc, 1u);                   \
              halo_mma<PAIR, S * 8 + 0, 0>(t1, aH1, a_hi, b_lo, b_hi, idesc, ACC);                  \
              halo_mma<PAIR, S * 8 + 2, 2>(t1, aH1, a_hi, b_lo, b_hi, idesc, 1u);                   \
              halo_mma<PAIR, S * 8 + 4, 4>(t1, aH1, a_hi, b_lo, b_hi, idesc, 1u);                   \
              halo_mma<PAIR, S * 8 + 6, 6>(t1, aH1, a_hi, b_lo, b_hi, idesc, 1u);                   \
              halo_commit<PAIR>(bempty_u + sb * 8);                                                         \
              if (++sb == static_cast<uint32_t>(BST)) { sb = 0; phb ^= 1; }                       \
            }
            HPRI_TAP(0, acc0)
            HPRI_TAP(1, 1u)
            HPRI_TAP(2, 1u)
#undef HPRI_TAP
          }
          halo_commit<PAIR>(aempty_u + sa * 8);   // the halo block is free once all nine taps have read it
          if (++sa == static_cast<uint32_t>(AST)) { sa = 0; pha ^= 1; }
        }
        halo_commit<PAIR>(tfull_u + as * 8);
      }
    }
  } else {
    // =========================== epilogue (warps 2..9): group g owns accumulator half g ===========================
    const int g = (warp - 2) >> 2;
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const int gw = (warp - 2) & 3;
    const bool elected = gw == 0 && lane == 0;
    uint8_t* stage = smem + L::STAGING_OFF + g * kStageTile;
    const int bar_id = 1 + g;
    const int my_r = geo.hr[g] + (row >> 3), my_c = geo.hc[g] + (row & 7);   // this thread's pixel inside the tile
    const bool bw = p.bw_sums != nullptr;          // fused BatchNorm-backward reduction (dgrad launches)
    const bool acc_on = p.stats != nullptr || bw;
    const bool det = !bw && p.has_fin && p.fin.partials != nullptr;     // deterministic BatchNorm statistics
    const uint8_t* xst = smem + p.xstage_off + g * kStageTile;
    uint64_t* x_full = tmem_empty + 2 + g;
    uint32_t xph = 0;
    float2 ssv[NCH], sqv[NCH];
#pragma unroll
    for (int j = 0; j < NCH; ++j) { ssv[j] = make_float2(0.f, 0.f); sqv[j] = make_float2(0.f, 0.f); }
    int acc_n0 = -1;
    uint32_t lt = 0;
    for (long long tile = tile_first; tile < total_tiles; tile += tile_step, ++lt) {
      const uint32_t as = lt & 1, aph = (lt >> 1) & 1;
      const int n_tile = static_cast<int>(tile % n_tiles);
      const int m_tile = static_cast<int>(tile / n_tiles) * (PAIR ? 2 : 1) + static_cast<int>(crank);
      const int img = m_tile / tiles_per_img;
      const int rr = m_tile - img * tiles_per_img;
      const int h0 = (rr / p.tiles_w) * p.th, w0 = (rr % p.tiles_w) * p.tw;
      const int n0 = n_tile * BLOCK_N;
      if (acc_on && n0 != acc_n0) {
        if (acc_n0 >= 0) {
#pragma unroll
          for (int j = 0; j < NCH; ++j) {
            if (det) flush_stats_ordered<8>(ssum, ssq, acc_n0 + j * 64 + 2 * lane, p.n_total, ssv[j], sqv[j], warp - 2, 3);
            else flush_stats(ssum, ssq, acc_n0 + j * 64 + 2 * lane, p.n_total, ssv[j], sqv[j]);
          }
        }
        acc_n0 = n0;
      }
      const bool valid = (img < p.N) && (h0 + my_r < p.H) && (w0 + my_c < p.W);
      const uint32_t tmem_acc = tmem_base + as * (2 * BLOCK_N) + g * BLOCK_N + (static_cast<uint32_t>(q * 32) << 16);
      mbar_wait(&tmem_full[as], aph);
      tc_fence_after();
#pragma unroll
      for (int c64 = 0; c64 < NCH; ++c64) {
        if (elected) bulk_wait_read0();
        named_bar_sync(bar_id, 128);           // staging (and the x tile) of the previous chunk are no longer read
        if (bw && elected) {                   // the matching tile of the layer below's raw conv output
          mbar_arrive_expect_tx(x_full, kStageTile);
          tma_load_4d(const_cast<uint8_t*>(xst), &tmX, x_full, n0 + c64 * 64, w0 + geo.hc[g], h0 + geo.hr[g], img);
        }
        chunk_to_stage(tmem_acc + c64 * 64, stage, row, valid, p.out_dt, nullptr);
        if (c64 == NCH - 1) {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) {
            if (PAIR) mbar_arrive_cluster(mapa_u32(smem_u32(&tmem_empty[as]), 0));
            else mbar_arrive(&tmem_empty[as]);
          }
        }
        fence_proxy_async_smem();
        named_bar_sync(bar_id, 128);
        if (elected && img < p.N && n0 + c64 * 64 < p.n_total && h0 + geo.hr[g] < p.H && w0 + geo.hc[g] < p.W) {
          tma_store_4d(&tmO, stage, n0 + c64 * 64, w0 + geo.hc[g], h0 + geo.hr[g], img);
          bulk_commit();
        }
        if (p.stats != nullptr) {
          if (p.out_dt == DT_F16) stats_rows32<DT_F16>(stage, gw * 32, lane, ssv[c64], sqv[c64]);
          else stats_rows32<DT_BF16>(stage, gw * 32, lane, ssv[c64], sqv[c64]);
        }
        if (bw) {
          const int ch = n0 + c64 * 64 + 2 * lane;
          float2 sc = make_float2(0.f, 0.f), sh = make_float2(0.f, 0.f);
          if (ch < p.n_total) { sc.x = __ldg(p.bw_scale + ch); sh.x = __ldg(p.bw_shift + ch); }
          if (ch + 1 < p.n_total) { sc.y = __ldg(p.bw_scale + ch + 1); sh.y = __ldg(p.bw_shift + ch + 1); }
          mbar_wait(x_full, xph);
          xph ^= 1;
          if (p.out_dt == DT_F16) bw_stats_rows32<DT_F16>(stage, xst, gw * 32, lane, sc, sh, ssv[c64], sqv[c64]);
          else bw_stats_rows32<DT_BF16>(stage, xst, gw * 32, lane, sc, sh, ssv[c64], sqv[c64]);
        }
      }
    }
    if (acc_on) {
      if (acc_n0 >= 0) {
#pragma unroll
        for (int j = 0; j < NCH; ++j) {
          if (det) flush_stats_ordered<8>(ssum, ssq, acc_n0 + j * 64 + 2 * lane, p.n_total, ssv[j], sqv[j], warp - 2, 3);
          else flush_stats(ssum, ssq, acc_n0 + j * 64 + 2 * lane, p.n_total, ssv[j], sqv[j]);
        }
      }
      named_bar_sync(3, kEpiThreads);
      for (int ch = threadIdx.x - 64; ch < p.n_total; ch += kEpiThreads) {
        const float a = ssum[ch], b = ssq[ch];
        if (det) {
          reinterpret_cast<float2*>(p.fin.partials)[static_cast<size_t>(blockIdx.x) * p.n_total + ch] = make_float2(a, b);
        } else if (a != 0.f || b != 0.f) {
          if (bw) {        // sums[c] = {sum dz, invstd * (sum dz*x - mean * sum dz), -} like bn_bwd_reduce_*_k
            const float mu = __ldg(p.bw_mean + ch), is = __ldg(p.bw_invstd + ch);
            atomicAdd(p.bw_sums + 3 * ch, static_cast<double>(a));
            atomicAdd(p.bw_sums + 3 * ch + 1,
                      static_cast<double>(is) * (static_cast<double>(b) - static_cast<double>(mu) * static_cast<double>(a)));
          } else {
            atomicAdd(p.stats + 2 * ch, static_cast<double>(a));
            atomicAdd(p.stats + 2 * ch + 1, static_cast<double>(b));
          }
        }
      }
      if (p.has_fin) bn_finalize_tail(p, threadIdx.x - 64, reinterpret_cast<int*>(tmem_ptr + 1));
    }
    if (elected) bulk_wait_read0();
  }
  __syncwarp();
  tc_fence_before();
  __syncthreads();
  if (PAIR) cluster_sync_all();           // no remote arrive / multicast commit may target a CTA that has exited
  if (warp == 1) {
    tc_fence_after();
    if (PAIR) tmem_dealloc2(tmem_base, 4 * BLOCK_N);
    else tmem_dealloc(tmem_base, 4 * BLOCK_N);
  }
}

// =====================================================================================
// 3x3 weight gradient with halo reuse (the two highest resolutions, Cout = 64 / 128).
//
//   dW[n][(tap, ci)] += sum_pixels X[pixel + tap][ci] * dY[pixel][n]           (MN-major operands, K = pixels)
//
// The generic MODE_WGRAD kernel re-loads the X tile once per tap and the dY tile once per pair of channel chunks:
// 48 KB of L2->SM traffic per 8 MMAs at N = 64, which pins those launches at 36-38 % tensor-pipe activity
// (profiles/ncu_full_r1e_summary.csv).  Here a k-block is one 16 x 8 pixel tile: the (16+2) x (8+2) halo block of X
// (per 64-channel chunk) and the dY tile are loaded ONCE and every tap is a descriptor start offset of
// (r*10 + s)*128 B into the halo block; eight consecutive pixels of a tile row are one 8-row swizzle group and tile
// rows are 10*128 B apart (SBO).  An M = 128 tile is two 64-channel groups LBO bytes apart: the two chunks of a
// chunk pair at one tap (LBO = halo block stride), or, when the layer has a single chunk (Cin = 64), two taps of the
// same block (LBO = their offset difference).  A CTA keeps the accumulators of up to 512 / BLOCK_N such tiles in
// TMEM (single-buffered: a work item is a long pixel range, its epilogue runs once) and walks its share of the
// pixel tiles (split-K), so one k-block of loads feeds 8 K-steps x up to 8 tiles of MMAs.
// Work item = (chunk pair, tap group, split).  fp32 red.add into the packed gradient, as the generic kernel.
// =====================================================================================
struct WgHaloArgs {
  int N, H, W, tiles_h, tiles_w;      // 16 x 8 pixel tiles
  int kchunks, nblk;                  // 64-channel chunks of X; halo blocks per work item (1 or 2)
  int units, n_tg;                    // M-tile units per chunk pair (9 taps, or 5 tap pairs when kchunks == 1); tap groups
  int npairs, splits;
  int n_total;                        // Cout
  float* dw;
  int dw_ld;
  int a_stride, stages;               // halo block bytes rounded to 1024; pipeline depth
  int a_dt, b_dt;
};
constexpr int kWgMaxStages = 6;
constexpr int kWgBlockBytes = 18 * 10 * 128;

template <int BLOCK_N>
__global__ void __launch_bounds__(kThreads, 1)
wgrad3x3_halo_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmDY,
                     const WgHaloArgs p) {
  constexpr int B_BYTES = (BLOCK_N / 64) * 16384;     // dY tile: 128 pixels x BLOCK_N channels
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem);           // [6] full, [6] empty, tmem_full, tmem_empty
  uint64_t* empty_bar = full_bar + kWgMaxStages;
  uint64_t* tmem_full = empty_bar + kWgMaxStages;
  uint64_t* tmem_empty = tmem_full + 1;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tmem_empty + 1);
  uint8_t* pipe = smem + 1024;
  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
  const int lane = threadIdx.x & 31;
  const int STG = p.stages;
  const uint32_t stage_bytes = static_cast<uint32_t>(p.nblk * p.a_stride + B_BYTES);
  const int tiles_per_img = p.tiles_h * p.tiles_w;
  const long long T = static_cast<long long>(p.N) * tiles_per_img;
  const int items_per_split = p.npairs * p.n_tg;
  const long long total_items = static_cast<long long>(items_per_split) * p.splits;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmX);
    tma_prefetch_desc(&tmDY);
    for (int s = 0; s < STG; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    mbar_init(tmem_full, 1);
    mbar_init(tmem_empty, 8);
    fence_mbar_init();
    fence_proxy_async_smem();
  }
  if (warp == 1) {
    tmem_alloc(tmem_ptr, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  grid_dep_launch();
  grid_dep_wait();            // everything above overlapped the previous kernel's tail; its outputs are visible from here

  // work item -> (pair, first unit, unit count, k-block range)
  auto decode = [&](long long item, int& pair, int& u0, int& nu, int& kb_begin, int& kb_end) {
    const int split = static_cast<int>(item / items_per_split);
    const int r = static_cast<int>(item - static_cast<long long>(split) * items_per_split);
    pair = r / p.n_tg;
    const int tg = r - pair * p.n_tg;
    u0 = (p.units * tg) / p.n_tg;
    nu = (p.units * (tg + 1)) / p.n_tg - u0;
    kb_begin = static_cast<int>(T * split / p.splits);
    kb_end = static_cast<int>(T * (split + 1) / p.splits);
  };

  if (warp == 0) {
    // =========================== TMA producer (whole warp converged) ===========================
    const uint32_t pipe_u = uniform_u32(smem_u32(pipe));
    const uint32_t full_u = uniform_u32(smem_u32(full_bar)), empty_u = full_u + kWgMaxStages * 8;
    uint32_t s = 0, ph = 0;
    for (long long item = blockIdx.x; item < total_items; item += gridDim.x) {
      int pair, u0, nu, kb_begin, kb_end;
      decode(item, pair, u0, nu, kb_begin, kb_end);
      for (int kb = kb_begin; kb < kb_end; ++kb) {
        const int im = kb / tiles_per_img;
        const int r = kb - im * tiles_per_img;
        const int h0 = (r / p.tiles_w) * 16, w0 = (r % p.tiles_w) * 8;
        mbar_wait_w(empty_u + s * 8, ph ^ 1);
        const uint32_t fb = full_u + s * 8;
        mbar_arrive_expect_tx_w(fb, static_cast<uint32_t>(p.nblk * kWgBlockBytes + B_BYTES));
        const uint32_t st = pipe_u + s * stage_bytes;
        for (int b = 0; b < p.nblk; ++b)
          tma_load_4d_w(st + b * p.a_stride, &tmX, fb, (2 * pair + b) * 64, w0 - 1, h0 - 1, im);
#pragma unroll
        for (int j = 0; j < BLOCK_N / 64; ++j)
          tma_load_4d_w(st + p.nblk * p.a_stride + j * 16384, &tmDY, fb, j * 64, w0, h0, im);
        if (++s == static_cast<uint32_t>(STG)) { s = 0; ph ^= 1; }
      }
    }
  } else if (warp == 1) {
    // =========================== MMA issuer (whole warp converged) ===========================
    const uint32_t idesc = make_idesc_16(128, BLOCK_N, 1, 1, p.a_dt, p.b_dt);
    const uint32_t pipe_u = uniform_u32(smem_u32(pipe));
    const uint32_t full_u = uniform_u32(smem_u32(full_bar)), empty_u = full_u + kWgMaxStages * 8;
    const uint32_t tfull_u = empty_u + kWgMaxStages * 8, tempty_u = tfull_u + 8;
    const uint32_t tmem_u = uniform_u32(tmem_base);
    const uint64_t dA = make_smem_desc_sw128(0, 0, 1280), dB = make_smem_desc_sw128(0, 16384, 1024);
    const uint32_t a_hi = static_cast<uint32_t>(dA >> 32), b_hi = static_cast<uint32_t>(dB >> 32);
    uint32_t s = 0, ph = 0, lt = 0;
    for (long long item = blockIdx.x; item < total_items; item += gridDim.x) {
      int pair, u0, nu, kb_begin, kb_end;
      decode(item, pair, u0, nu, kb_begin, kb_end);
      if (kb_end <= kb_begin) continue;
      mbar_wait_w(tempty_u, (lt & 1) ^ 1);          // the epilogue has drained the accumulators of the previous item
      ++lt;
      tc_fence_after();
      for (int kb = kb_begin; kb < kb_end; ++kb) {
        mbar_wait_w(full_u + s * 8, ph);
        tc_fence_after();
        const uint32_t st = pipe_u + s * stage_bytes;
        const uint32_t b_lo = static_cast<uint32_t>(dB) | ((st + p.nblk * p.a_stride) >> 4);
        const uint32_t acc0 = kb > kb_begin ? 1u : 0u;
#pragma unroll 1
        for (int i = 0; i < nu; ++i) {
          // start offset (rows of 128 B) of the first 64-row group and byte distance to the second one
          int off0, lbo;
          if (p.kchunks == 1) {
            const int t0 = 2 * (u0 + i), t1 = t0 + 1;
            off0 = (t0 / 3) * 10 + t0 % 3;
            lbo = t1 < 9 ? ((t1 / 3) * 10 + t1 % 3 - off0) * 128 : 0;
          } else {
            const int t0 = u0 + i;
            off0 = (t0 / 3) * 10 + t0 % 3;
            lbo = p.a_stride;
          }
          const uint32_t a_lo = ((st + off0 * 128) >> 4) | (static_cast<uint32_t>(lbo >> 4) << 16);
          const uint32_t td = tmem_u + i * BLOCK_N;
          // K = 16 pixels per MMA = two tile rows: 2 * 1280 B of the halo block, 2 * 1024 B of the dY tile
          umma_f16_off_w<0 * 160, 0 * 128>(td, a_lo, a_hi, b_lo, b_hi, idesc, acc0);
          umma_f16_off_w<1 * 160, 1 * 128>(td, a_lo, a_hi, b_lo, b_hi, idesc, 1u);
          umma_f16_off_w<2 * 160, 2 * 128>(td, a_lo, a_hi, b_lo, b_hi, idesc, 1u);
          umma_f16_off_w<3 * 160, 3 * 128>(td, a_lo, a_hi, b_lo, b_hi, idesc, 1u);
          umma_f16_off_w<4 * 160, 4 * 128>(td, a_lo, a_hi, b_lo, b_hi, idesc, 1u);
          umma_f16_off_w<5 * 160, 5 * 128>(td, a_lo, a_hi, b_lo, b_hi, idesc, 1u);
          umma_f16_off_w<6 * 160, 6 * 128>(td, a_lo, a_hi, b_lo, b_hi, idesc, 1u);
          umma_f16_off_w<7 * 160, 7 * 128>(td, a_lo, a_hi, b_lo, b_hi, idesc, 1u);
        }
        umma_commit_w(empty_u + s * 8);
        if (++s == static_cast<uint32_t>(STG)) { s = 0; ph ^= 1; }
      }
      umma_commit_w(tfull_u);
    }
  } else {
    // =========================== epilogue (warps 2..9): the groups split the 32-column slices ===========
    const int g = (warp - 2) >> 2;
    const int q = warp & 3;
    const int row = q * 32 + lane;
    uint32_t lt = 0;
    for (long long item = blockIdx.x; item < total_items; item += gridDim.x) {
      int pair, u0, nu, kb_begin, kb_end;
      decode(item, pair, u0, nu, kb_begin, kb_end);
      if (kb_end <= kb_begin) continue;
      mbar_wait(tmem_full, lt & 1);
      ++lt;
      tc_fence_after();
      for (int i = 0; i < nu; ++i) {
        // packed-gradient column of this accumulator row: group (tap, chunk) * 64 + channel
        int tap, cc;
        if (p.kchunks == 1) { tap = 2 * (u0 + i) + (row >> 6); cc = 0; }
        else { tap = u0 + i; cc = 2 * pair + (row >> 6); }
        const bool row_ok = tap < 9 && cc < p.kchunks;
        const long long kidx = static_cast<long long>(tap * p.kchunks + cc) * 64 + (row & 63);
        const uint32_t tmem_acc = tmem_base + i * BLOCK_N + (static_cast<uint32_t>(q * 32) << 16);
#pragma unroll 1
        for (int c = g; c < BLOCK_N / 32; c += 2) {
          uint32_t v[32];
          tmem_ld32(tmem_acc + c * 32, v);
          tmem_ld_wait();
          if (row_ok) {
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              const int n = c * 32 + j;
              if (n < p.n_total) red_add_f32(p.dw + static_cast<long long>(n) * p.dw_ld + kidx, __uint_as_float(v[j]));
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tmem_empty);
    }
  }
  __syncwarp();
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// =====================================================================================
// host side
// =====================================================================================
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* sym = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(sym);
  });
  return fn;
}

// bf16 tensor map, 128B swizzle, zero OOB fill.  dims/strides innermost first; strides in bytes for dims 1..rank-1.
static int make_map(CUtensorMap* m, const void* base, int rank, const uint64_t* dims, const uint64_t* strides,
                    const uint32_t* box, int dt) {
  EncodeTiledFn enc = get_encode();
  if (!enc) return HPRI_ERR_DRIVER;
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  cuuint64_t d[5], s[4];
  cuuint32_t b[5];
  for (int i = 0; i < rank; ++i) { d[i] = dims[i]; b[i] = box[i]; }
  for (int i = 0; i + 1 < rank; ++i) s[i] = strides[i];
  CUresult r = enc(m, dt == DT_F16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, rank, const_cast<void*>(base), d, s, b, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? HPRI_OK : HPRI_ERR_TENSORMAP;
}

static int check_view(const hpri_view_t& v) {
  if (!v.ptr || v.n <= 0 || v.h <= 0 || v.w <= 0 || v.c <= 0) return HPRI_ERR_ARG;
  if ((reinterpret_cast<uintptr_t>(v.ptr) & 15) || (v.pix_stride & 7) || (v.row_stride & 7) || (v.img_stride & 7))
    return HPRI_ERR_ALIGN;
  if (v.dtype != DT_BF16 && v.dtype != DT_F16) return HPRI_ERR_ARG;
  return HPRI_OK;
}

// NHWC view -> 4-D map (c, w, h, n), box {64, tw, th, 1}
static int map_nhwc(CUtensorMap* m, const hpri_view_t& v, int th, int tw) {
  uint64_t dims[4] = {(uint64_t)v.c, (uint64_t)v.w, (uint64_t)v.h, (uint64_t)v.n};
  uint64_t str[3] = {(uint64_t)v.pix_stride * 2, (uint64_t)v.row_stride * 2, (uint64_t)v.img_stride * 2};
  uint32_t box[4] = {64, (uint32_t)tw, (uint32_t)th, 1};
  return make_map(m, v.ptr, 4, dims, str, box, v.dtype);
}
// 2x2 stride-2 gather view of a high-res tensor: (c, w, a, h, n) for a fixed column parity b.
// v describes the HIGH-res tensor; (hl, wl) is the low-res pixel grid.
static int map_up2(CUtensorMap* m, const hpri_view_t& v, int hl, int wl, int b, int th, int tw) {
  if (v.h < 2 * hl || v.w < 2 * wl) return HPRI_ERR_ARG;
  uint64_t dims[5] = {(uint64_t)v.c, (uint64_t)wl, 2, (uint64_t)hl, (uint64_t)v.n};
  uint64_t str[4] = {(uint64_t)v.pix_stride * 4, (uint64_t)v.row_stride * 2, (uint64_t)v.row_stride * 4,
                     (uint64_t)v.img_stride * 2};
  uint32_t box[5] = {64, (uint32_t)tw, 1, (uint32_t)th, 1};
  const uint16_t* base = static_cast<const uint16_t*>(v.ptr) + (long long)b * v.pix_stride;
  return make_map(m, base, 5, dims, str, box, v.dtype);
}
static int map_weights(CUtensorMap* m, const void* w, int rows, int kpad, int block_n, int dt) {
  uint64_t dims[2] = {(uint64_t)kpad, (uint64_t)rows};
  uint64_t str[1] = {(uint64_t)kpad * 2};
  uint32_t box[2] = {64, (uint32_t)block_n};
  return make_map(m, w, 2, dims, str, box, dt);
}

// output map of a (channel-sliced) NHWC view: dims (n_store, w, h, n), box {64, bw, bh, 1}; the TMA unit clips
// whatever part of a staged tile falls outside
static int map_out(CUtensorMap* m, const hpri_view_t& v, int n_store, int bh, int bw) {
  uint64_t dims[4] = {(uint64_t)n_store, (uint64_t)v.w, (uint64_t)v.h, (uint64_t)v.n};
  uint64_t str[3] = {(uint64_t)v.pix_stride * 2, (uint64_t)v.row_stride * 2, (uint64_t)v.img_stride * 2};
  uint32_t box[4] = {64, (uint32_t)bw, (uint32_t)bh, 1};
  return make_map(m, v.ptr, 4, dims, str, box, v.dtype);
}
// ConvTranspose2d(k2, s2) destination for one (a, b): the pixels (2h+a, 2w+b) of the high-res view v, indexed by
// the low-res (w, h) of an hl x wl grid
static int map_out_up2(CUtensorMap* m, const hpri_view_t& v, int cout, int hl, int wl, int a, int b, int bh, int bw) {
  int wv = (v.w - b + 1) / 2, hv = (v.h - a + 1) / 2;
  if (wv > wl) wv = wl;
  if (hv > hl) hv = hl;
  if (wv <= 0 || hv <= 0) return HPRI_ERR_ARG;
  uint64_t dims[4] = {(uint64_t)cout, (uint64_t)wv, (uint64_t)hv, (uint64_t)v.n};
  uint64_t str[3] = {(uint64_t)v.pix_stride * 4, (uint64_t)v.row_stride * 4, (uint64_t)v.img_stride * 2};
  uint32_t box[4] = {64, (uint32_t)bw, (uint32_t)bh, 1};
  const uint16_t* base = static_cast<const uint16_t*>(v.ptr) + (long long)a * v.row_stride + (long long)b * v.pix_stride;
  return make_map(m, base, 4, dims, str, box, v.dtype);
}

static int ilog2(int v) {
  int l = 0;
  while ((1 << l) < v) ++l;
  return l;
}

// pick a th x tw = P pixel tile (powers of two) minimising the tile count; ties -> wider rows
static void pick_tile(int H, int W, int P, int* th, int* tw) {
  long long best = -1;
  for (int t = 1; t <= P; t *= 2) {
    const int w = t, h = P / t;
    if (w > 256 || h > 256) continue;
    const long long cnt = (long long)((H + h - 1) / h) * ((W + w - 1) / w);
    if (best < 0 || cnt <= best) { best = cnt; *th = h; *tw = w; }
  }
}
static void set_tile(IgemmArgs& a, int th, int tw) {
  a.th = th; a.tw = tw; a.ltw = ilog2(tw);
  a.tiles_h = (a.H + th - 1) / th; a.tiles_w = (a.W + tw - 1) / tw;
}

// SMs the persistent tcgen05 grids may fill.  hpri_set_sm_reserve(k) leaves k SMs free for kernels that must run
// CONCURRENTLY with them -- NCCL's all-reduce CTAs during the data-parallel backward: a persistent grid sized to every
// SM turns one wave into two as soon as a single CTA is displaced.  Seeded by the environment variable HPRI_SM_RESERVE.
static int g_sm_reserve = -1;
static int sm_count() {
  static int num_sms = 0;
  if (num_sms == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    if (cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || num_sms <= 0) num_sms = 148;
  }
  if (g_sm_reserve < 0) {
    const char* e = getenv("HPRI_SM_RESERVE");
    g_sm_reserve = e ? atoi(e) : 0;
    if (g_sm_reserve < 0) g_sm_reserve = 0;
  }
  const int n = num_sms - g_sm_reserve;
  return n < 2 ? 2 : n;
}

struct OutMaps { CUtensorMap m[4]; };

// Launch with programmatic stream serialisation (the kernels call grid_dep_wait() after their prologue) and, for the
// CTA-pair kernels, a cluster of two.  HPRI_PDL=0 falls back to plain stream order.
static bool pdl_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("HPRI_PDL");
    v = (e && atoi(e) == 0) ? 0 : 1;
  }
  return v == 1;
}
template <typename K, typename... A>
static cudaError_t launch_ex(K kern, long long grid, size_t smem, cudaStream_t stream, bool cluster2, A... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((unsigned)grid); cfg.blockDim = dim3(kThreads); cfg.dynamicSmemBytes = smem; cfg.stream = stream;
  cudaLaunchAttribute at[2];
  unsigned n = 0;
  if (cluster2) {
    at[n].id = cudaLaunchAttributeClusterDimension;
    at[n].val.clusterDim.x = 2; at[n].val.clusterDim.y = 1; at[n].val.clusterDim.z = 1;
    ++n;
  }
  if (pdl_enabled()) {
    at[n].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[n].val.programmaticStreamSerializationAllowed = 1;
    ++n;
  }
  cfg.attrs = at; cfg.numAttrs = n;
  ++g_launch_count;
  return cudaLaunchKernelEx(&cfg, kern, args...);
}

template <int BLOCK_N, int STAGES, int MODE>
static int launch_t(const CUtensorMap& a0, const CUtensorMap& a1, const CUtensorMap& b0, const CUtensorMap& b1,
                    const OutMaps& o, const IgemmArgs& args, long long grid, cudaStream_t stream) {
  using L = SmemLayout<BLOCK_N, STAGES, MODE>;
  auto kern = igemm_kernel<BLOCK_N, STAGES, MODE>;
  static std::once_flag once;
  static cudaError_t attr_err = cudaSuccess;
  std::call_once(once, [&] {
    attr_err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, L::ALLOC);
  });
  if (attr_err != cudaSuccess) return HPRI_ERR_CUDA;
  if (grid <= 0 || grid > 0x7FFFFFFFLL) return HPRI_ERR_ARG;
  if (grid > sm_count()) grid = sm_count();          // persistent: one CTA per SM walks the tile list
  return launch_ex(kern, grid, L::ALLOC, stream, false, a0, a1, b0, b1, o.m[0], o.m[1], o.m[2], o.m[3], args) == cudaSuccess
             ? HPRI_OK : HPRI_ERR_CUDA;
}

template <int MODE>
static int launch(int block_n, const CUtensorMap& a0, const CUtensorMap& a1, const CUtensorMap& b0,
                  const CUtensorMap& b1, const OutMaps& o, const IgemmArgs& args, long long grid,
                  cudaStream_t stream) {
  switch (block_n) {
    case 64: return launch_t<64, MODE == MODE_FWD ? 7 : 4, MODE>(a0, a1, b0, b1, o, args, grid, stream);
    case 128: return launch_t<128, MODE == MODE_FWD ? 5 : 3, MODE>(a0, a1, b0, b1, o, args, grid, stream);
    case 256: return launch_t<256, MODE == MODE_FWD ? 3 : 2, MODE>(a0, a1, b0, b1, o, args, grid, stream);
  }
  return HPRI_ERR_ARG;
}

static int pick_block_n(int n_total, int forced) {
  if (forced == 64 || forced == 128 || forced == 256) return forced;
  if (n_total <= 64) return 64;
  if (n_total <= 128) return 128;
  if (n_total % 256 == 0) return 256;
  return 128;
}


// halo kernel launcher -------------------------------------------------------------------
template <int BLOCK_N, bool PAIR>
static int launch_halo_t(const CUtensorMap& ma, const CUtensorMap& mb, const CUtensorMap& mo, const CUtensorMap& mx,
                         const IgemmArgs& args, const HaloGeom& geo, long long tiles, cudaStream_t stream) {
  using L = HaloSmem<BLOCK_N>;
  auto kern = conv3x3_halo_kernel<BLOCK_N, PAIR>;
  static std::once_flag once;
  static cudaError_t attr_err = cudaSuccess;
  std::call_once(once, [&] {
    attr_err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  });
  if (attr_err != cudaSuccess) return HPRI_ERR_CUDA;
  // `tiles` counts work items: pixel tiles x channel tiles, or (PAIR) pixel-tile pairs x channel tiles walked by a 2-CTA cluster
  const long long streams = PAIR ? sm_count() / 2 : sm_count();
  long long grid = (tiles < streams ? tiles : streams) * (PAIR ? 2 : 1);
  if (grid <= 0) return HPRI_ERR_ARG;
  const size_t smem = (size_t)L::PIPE_OFF + (size_t)args.a_stages * args.a_bytes +
                      (size_t)args.stages * (PAIR ? L::B_STAGE / 2 : L::B_STAGE) + (args.bw_sums ? 2 * kStageTile : 0) + 1024;
  return launch_ex(kern, grid, smem, stream, PAIR, ma, mb, mo, mx, args, geo) == cudaSuccess ? HPRI_OK : HPRI_ERR_CUDA;
}

// 256-pixel tile = two 16 x 8 halves, stacked (32 x 8) or side by side (16 x 16): minimise padded pixels;
// returns the waste fraction
static double pick_halo_tile(int H, int W, int* th, int* tw) {
  double best = 1e30;
  const int cand[2][2] = {{32, 8}, {16, 16}};
  for (auto& c : cand) {
    const int h = c[0], w = c[1];
    const double padded = (double)((H + h - 1) / h) * h * ((W + w - 1) / w) * w;
    if (padded < best) { best = padded; *th = h; *tw = w; }
  }
  return best / ((double)H * W) - 1.0;
}

static int g_conv_algo = -2;               // -1 heuristic, 0 generic kernel, 1 halo-reuse kernel on single CTAs, 2 on CTA pairs
static int conv_algo_override() {          // env HPRI_CONV_ALGO seeds it; hpri_set_conv_algo overrides
  if (g_conv_algo == -2) {
    const char* e = getenv("HPRI_CONV_ALGO");
    g_conv_algo = e ? atoi(e) : -1;
  }
  return g_conv_algo;
}

static int g_halo_a_stages = -1;           // depth of the halo-block ring: 2 (default) or 3; env HPRI_HALO_A_STAGES
static int halo_a_stages_override() {
  if (g_halo_a_stages < 0) {
    const char* e = getenv("HPRI_HALO_A_STAGES");
    g_halo_a_stages = (e && atoi(e) == 3) ? 3 : 2;
  }
  return g_halo_a_stages;
}

static int g_wgrad_algo = -2;              // -1 heuristic, 0 generic kernel only
static int wgrad_algo_override() {
  if (g_wgrad_algo == -2) {
    const char* e = getenv("HPRI_WGRAD_ALGO");
    g_wgrad_algo = e ? atoi(e) : -1;
  }
  return g_wgrad_algo;
}

}  // namespace hpri

using namespace hpri;

extern "C" int hpri_conv3x3_halo_ok(int h, int w, int w_rows) {
  if (h <= 0 || w <= 0 || w_rows <= 0) return 0;
  int hth = 0, htw = 0;
  const double waste = pick_halo_tile(h, w, &hth, &htw);
  const int ov = conv_algo_override();
  const int a_bytes = ((hth + 2) * (htw + 2) * 128 + 1023) / 1024 * 1024;
  const int stages = w_rows <= 64 ? HaloSmem<64>::b_stages_for(a_bytes, 2 * kStageTile)
                                  : HaloSmem<128>::b_stages_for(a_bytes, 2 * kStageTile);
  return (w_rows <= kHaloStatCh && stages >= 2 && (ov >= 1 || (ov < 0 && waste <= 1.0))) ? 1 : 0;
}

extern "C" int hpri_set_conv_algo(int algo) {
  if (algo < -1 || algo > 2) return HPRI_ERR_ARG;
  g_conv_algo = algo;
  return HPRI_OK;
}

extern "C" int hpri_set_halo_a_stages(int stages) {
  if (stages != 2 && stages != 3) return HPRI_ERR_ARG;
  g_halo_a_stages = stages;
  return HPRI_OK;
}

extern "C" int hpri_set_sm_reserve(int sms) {
  if (sms < 0 || sms > 64) return HPRI_ERR_ARG;
  g_sm_reserve = sms;
  return HPRI_OK;
}

extern "C" int hpri_set_wgrad_algo(int algo) {
  if (algo < -1 || algo > 0) return HPRI_ERR_ARG;
  g_wgrad_algo = algo;
  return HPRI_OK;
}

// -------------------------------------------------------------------------------------
// C ABI
// -------------------------------------------------------------------------------------
extern "C" int hpri_igemm_fwd(const hpri_view_t* x, const void* wpack, int w_dtype, int w_rows, int kpad, int taps,
                              const hpri_view_t* y, int n_store, const float* bias, double* stats, int accumulate,
                              int block_n, const hpri_bn_fin_t* fin, const hpri_bn_bwd_t* bw, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (!x || !y || !wpack) return HPRI_ERR_ARG;
  int rc;
  if ((rc = check_view(*x)) != HPRI_OK || (rc = check_view(*y)) != HPRI_OK) return rc;
  if (taps != 1 && taps != 9) return HPRI_ERR_ARG;
  if (x->n != y->n || x->h != y->h || x->w != y->w) return HPRI_ERR_ARG;
  if ((reinterpret_cast<uintptr_t>(wpack) & 15) || (kpad & 63) || (n_store & 7) || n_store > y->c) return HPRI_ERR_ALIGN;
  if (bias && (reinterpret_cast<uintptr_t>(bias) & 15)) return HPRI_ERR_ALIGN;
  const int kchunks = (x->c + 63) / 64;
  if (kpad != taps * kchunks * 64) return HPRI_ERR_ARG;
  IgemmArgs a{};
  a.N = x->n; a.H = x->h; a.W = x->w;
  a.taps = taps; a.tap_mode = taps == 9 ? TAP_3X3 : TAP_NONE; a.kchunks = kchunks;
  a.n_total = w_rows;
  a.up2 = 0; a.cout = w_rows;
  a.a_dt = x->dtype; a.b_dt = w_dtype; a.out_dt = y->dtype;
  if (x->dtype != w_dtype) return HPRI_ERR_ARG;     // kind::f16 takes A and B in one format
  a.bias = bias; a.stats = stats; a.accum = accumulate ? 1 : 0;
  if (accumulate && stats) return HPRI_ERR_ARG;
  if (stats && w_rows > kMaxStatCh) return HPRI_ERR_ARG;
  if (fin) {
    if (!stats || !fin->scale || !fin->shift || !fin->counter || fin->count <= 0) return HPRI_ERR_ARG;
    a.has_fin = 1;
    a.fin = *fin;
  }
  if (bias && (w_rows & 63)) return HPRI_ERR_ARG;   // the epilogue reads the bias in 64-float runs
  if (n_store <= 0) return HPRI_ERR_ARG;
  if (taps == 9 && bias == nullptr && !accumulate) {
    // halo-reuse kernel (256-pixel tiles)
    int hth = 0, htw = 0;
    const double waste = pick_halo_tile(a.H, a.W, &hth, &htw);
    const int ov = conv_algo_override();
    const bool stats_ok = !stats || w_rows <= kHaloStatCh;
    const int hbn = w_rows <= 64 ? 64 : 128;
    const int a_bytes = ((hth + 2) * (htw + 2) * 128 + 1023) / 1024 * 1024;
    const int extra = bw ? 2 * kStageTile : 0;           // x staging tiles of the fused BN-backward reduction
    // CTA pairs (cta_group::2: each CTA stages and reads half of the weight tile) unless HPRI_CONV_ALGO=1 asks for single
    // CTAs.  The 64-channel dgrads with the fused BN-backward reduction are epilogue-bound and measured 8 % slower in
    // pairs (the MMA warp then waits for the slower of two epilogues), so they stay on single CTAs.
    const bool pair = ov == 2 || (ov != 1 && !(bw && hbn == 64));
    const int b_slot = hbn * 128 / (pair ? 2 : 1);
    const int pipe_budget = hbn == 64 ? HaloSmem<64>::BUDGET - HaloSmem<64>::PIPE_OFF : HaloSmem<128>::BUDGET - HaloSmem<128>::PIPE_OFF;
    int a_stages = kHaloAStages;
    int stages = (pipe_budget - a_stages * a_bytes - extra) / b_slot;
    if (halo_a_stages_override() == 3 && (pipe_budget - 3 * a_bytes - extra) / b_slot >= 4) {   // a third halo block if >= 4 weight slots remain
      a_stages = 3;
      stages = (pipe_budget - 3 * a_bytes - extra) / b_slot;
    }
    if (stages > kHaloMaxBStages) stages = kHaloMaxBStages;
    if (stats_ok && stages >= 2 && (ov >= 1 || (ov < 0 && waste <= 1.0))) {
      set_tile(a, hth, htw);
      a.stages = stages; a.a_bytes = a_bytes; a.a_stages = a_stages;
      CUtensorMap mx{};
      if (bw) {
        if (stats || !bw->x || !bw->scale || !bw->shift || !bw->save_mean || !bw->save_invstd || !bw->sums) return HPRI_ERR_ARG;
        if ((rc = check_view(*bw->x)) != HPRI_OK) return rc;
        if (bw->x->n != y->n || bw->x->h != y->h || bw->x->w != y->w || bw->x->c < w_rows || bw->x->dtype != y->dtype)
          return HPRI_ERR_ARG;
        if (w_rows > kHaloStatCh) return HPRI_ERR_ARG;
        hpri_view_t xv = *bw->x;
        xv.c = w_rows;
        if ((rc = map_nhwc(&mx, xv, 16, 8)) != HPRI_OK) return rc;
        a.bw_scale = bw->scale; a.bw_shift = bw->shift; a.bw_mean = bw->save_mean; a.bw_invstd = bw->save_invstd;
        a.bw_sums = bw->sums;
        a.xstage_off = (hbn == 64 ? HaloSmem<64>::PIPE_OFF : HaloSmem<128>::PIPE_OFF) + a_stages * a_bytes + stages * b_slot;
      }
      HaloGeom geo{};
      geo.wb = htw + 2;
      geo.hr[0] = 0; geo.hc[0] = 0;
      geo.hr[1] = htw == 8 ? 16 : 0; geo.hc[1] = htw == 8 ? 0 : 8;
      CUtensorMap ma, mb, mo;
      uint64_t dims[4] = {(uint64_t)x->c, (uint64_t)x->w, (uint64_t)x->h, (uint64_t)x->n};
      uint64_t str[3] = {(uint64_t)x->pix_stride * 2, (uint64_t)x->row_stride * 2, (uint64_t)x->img_stride * 2};
      uint32_t box[4] = {64, (uint32_t)(htw + 2), (uint32_t)(hth + 2), 1};
      if ((rc = make_map(&ma, x->ptr, 4, dims, str, box, x->dtype)) != HPRI_OK) return rc;
      if ((rc = map_weights(&mb, wpack, w_rows, kpad, pair ? hbn / 2 : hbn, w_dtype)) != HPRI_OK) return rc;
      if ((rc = map_out(&mo, *y, n_store, 16, 8)) != HPRI_OK) return rc;
      const long long m_tiles = (long long)a.N * a.tiles_h * a.tiles_w;
      const long long tiles = (pair ? (m_tiles + 1) / 2 : m_tiles) * ((w_rows + hbn - 1) / hbn);
      if (!bw) mx = mo;
      if (pair) return hbn == 64 ? launch_halo_t<64, true>(ma, mb, mo, mx, a, geo, tiles, stream)
                                 : launch_halo_t<128, true>(ma, mb, mo, mx, a, geo, tiles, stream);
      return hbn == 64 ? launch_halo_t<64, false>(ma, mb, mo, mx, a, geo, tiles, stream)
                       : launch_halo_t<128, false>(ma, mb, mo, mx, a, geo, tiles, stream);
    }
  }
  if (bw) return HPRI_ERR_ARG;        // the fused reduction exists on the halo kernel only (hpri_conv3x3_halo_ok)
  int th, tw;
  pick_tile(a.H, a.W, 128, &th, &tw);
  set_tile(a, th, tw);
  const int bn = pick_block_n(w_rows, block_n);
  CUtensorMap ma, mb;
  OutMaps o;
  if ((rc = map_nhwc(&ma, *x, a.th, a.tw)) != HPRI_OK) return rc;
  if ((rc = map_weights(&mb, wpack, w_rows, kpad, bn, w_dtype)) != HPRI_OK) return rc;
  if ((rc = map_out(&o.m[0], *y, n_store, a.th, a.tw)) != HPRI_OK) return rc;
  o.m[1] = o.m[2] = o.m[3] = o.m[0];
  const long long grid = (long long)a.N * a.tiles_h * a.tiles_w * ((w_rows + bn - 1) / bn);
  return launch<MODE_FWD>(bn, ma, ma, mb, mb, o, a, grid, stream);
}

// ConvTranspose2d(k=2, s=2) forward: y[n, 2h+a, 2w+b, co] = bias[co] + sum_ci x[n,h,w,ci] W[ci,co,a,b]
// wpack rows are (a*2+b)*cout + co, K = ci.  y is the high-res destination view (channel offset baked in ptr).
extern "C" int hpri_convT2x2_fwd(const hpri_view_t* x, const void* wpack, int w_dtype, int cout, int kpad,
                                 const hpri_view_t* y, const float* bias, int block_n, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (!x || !y || !wpack) return HPRI_ERR_ARG;
  int rc;
  if ((rc = check_view(*x)) != HPRI_OK || (rc = check_view(*y)) != HPRI_OK) return rc;
  const int kchunks = (x->c + 63) / 64;
  if (kpad != kchunks * 64 || (cout & 63) || y->c < cout) return HPRI_ERR_ARG;
  if (bias && (reinterpret_cast<uintptr_t>(bias) & 15)) return HPRI_ERR_ALIGN;
  IgemmArgs a{};
  a.N = x->n; a.H = x->h; a.W = x->w;
  int th, tw;
  pick_tile(a.H, a.W, 128, &th, &tw);
  set_tile(a, th, tw);
  a.taps = 1; a.tap_mode = TAP_NONE; a.kchunks = kchunks; a.n_total = 4 * cout;
  a.up2 = 1; a.cout = cout;
  a.a_dt = x->dtype; a.b_dt = w_dtype; a.out_dt = y->dtype;
  if (x->dtype != w_dtype) return HPRI_ERR_ARG;
  a.bias = bias; a.stats = nullptr;
  // every 64-column chunk of a tile picks its own (a, b) output map, so wide tiles may span several of them
  const int bn = block_n == 64 || block_n == 128 || block_n == 256 ? block_n : 256;
  CUtensorMap ma, mb;
  OutMaps o;
  if ((rc = map_nhwc(&ma, *x, a.th, a.tw)) != HPRI_OK) return rc;
  if ((rc = map_weights(&mb, wpack, 4 * cout, kpad, bn, w_dtype)) != HPRI_OK) return rc;
  for (int ab = 0; ab < 4; ++ab)
    if ((rc = map_out_up2(&o.m[ab], *y, cout, a.H, a.W, ab >> 1, ab & 1, a.th, a.tw)) != HPRI_OK) return rc;
  const long long grid = (long long)a.N * a.tiles_h * a.tiles_w * ((4 * cout + bn - 1) / bn);
  return launch<MODE_FWD>(bn, ma, ma, mb, mb, o, a, grid, stream);
}

// ConvTranspose2d dgrad: dx[n,h,w,ci] = sum_{a,b,co} dy[n,2h+a,2w+b,co] W[ci,co,a,b]
// wpack rows = ci, K = (a*2+b)*kc*64 + co  (kc = ceil(cout/64)).
extern "C" int hpri_convT2x2_dgrad(const hpri_view_t* dy, const void* wpack, int w_dtype, int cin, int kpad,
                                   const hpri_view_t* dx, int block_n, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (!dy || !dx || !wpack) return HPRI_ERR_ARG;
  int rc;
  if ((rc = check_view(*dy)) != HPRI_OK || (rc = check_view(*dx)) != HPRI_OK) return rc;
  const int kchunks = (dy->c + 63) / 64;
  if (kpad != 4 * kchunks * 64) return HPRI_ERR_ARG;
  IgemmArgs a{};
  a.N = dx->n; a.H = dx->h; a.W = dx->w;
  int th, tw;
  pick_tile(a.H, a.W, 128, &th, &tw);
  set_tile(a, th, tw);
  a.taps = 4; a.tap_mode = TAP_UP2; a.kchunks = kchunks; a.n_total = cin;
  a.up2 = 0; a.cout = cin;
  a.a_dt = dy->dtype; a.b_dt = w_dtype; a.out_dt = dx->dtype;
  if (dy->dtype != w_dtype) return HPRI_ERR_ARG;
  const int bn = pick_block_n(cin, block_n);
  int n_store = (cin + 7) & ~7;
  if (n_store > dx->c) return HPRI_ERR_ARG;
  CUtensorMap m0, m1, mb;
  OutMaps o;
  if ((rc = map_up2(&m0, *dy, a.H, a.W, 0, a.th, a.tw)) != HPRI_OK) return rc;
  if ((rc = map_up2(&m1, *dy, a.H, a.W, 1, a.th, a.tw)) != HPRI_OK) return rc;
  if ((rc = map_weights(&mb, wpack, cin, kpad, bn, w_dtype)) != HPRI_OK) return rc;
  if ((rc = map_out(&o.m[0], *dx, n_store, a.th, a.tw)) != HPRI_OK) return rc;
  o.m[1] = o.m[2] = o.m[3] = o.m[0];
  const long long grid = (long long)a.N * a.tiles_h * a.tiles_w * ((cin + bn - 1) / bn);
  return launch<MODE_FWD>(bn, m0, m1, mb, mb, o, a, grid, stream);
}

// Weight gradient.  mode 0: 1x1 / Linear, 1: 3x3 pad 1, 2: ConvTranspose2d 2x2 (x low-res, dy high-res).
// dw is fp32 [n_total][dw_ld] in the forward pack layout and is ACCUMULATED into (caller zeroes it).
extern "C" int hpri_igemm_wgrad(const hpri_view_t* x, const hpri_view_t* dy, int mode, int n_total, float* dw,
                                int dw_ld, int block_n, int splits, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (!x || !dy || !dw) return HPRI_ERR_ARG;
  int rc;
  if ((rc = check_view(*x)) != HPRI_OK || (rc = check_view(*dy)) != HPRI_OK) return rc;
  if (mode < 0 || mode > 2) return HPRI_ERR_ARG;
  IgemmArgs a{};
  a.N = x->n; a.H = x->h; a.W = x->w;
  int th, tw;
  pick_tile(a.H, a.W, 128, &th, &tw);
  set_tile(a, th, tw);
  a.taps = mode == 1 ? 9 : 1; a.tap_mode = mode == 1 ? TAP_3X3 : (mode == 2 ? TAP_UP2 : TAP_NONE);
  a.kchunks = (x->c + 63) / 64;
  a.total_chunks = a.taps * a.kchunks;
  if (dw_ld != a.total_chunks * 64) return HPRI_ERR_ARG;
  a.n_total = n_total; a.dw = dw; a.dw_ld = dw_ld;
  a.a_dt = x->dtype; a.b_dt = dy->dtype; a.out_dt = DT_BF16;
  if (x->dtype != dy->dtype) return HPRI_ERR_ARG;   // kind::f16 takes A and B in one format
  if (mode == 1 && (n_total == 64 || n_total == 128) && (a.kchunks == 1 || a.kchunks % 2 == 0) && splits <= 0 &&
      block_n == 0 && wgrad_algo_override() != 0 && dy->c >= n_total) {
    // halo-reuse weight gradient (one X halo block + one dY tile per 128-pixel k-block feed every tap)
    if (dy->n != x->n || dy->h != x->h || dy->w != x->w) return HPRI_ERR_ARG;
    WgHaloArgs g{};
    g.N = x->n; g.H = x->h; g.W = x->w;
    g.tiles_h = (g.H + 15) / 16; g.tiles_w = (g.W + 7) / 8;
    g.kchunks = a.kchunks; g.nblk = a.kchunks == 1 ? 1 : 2;
    g.npairs = a.kchunks == 1 ? 1 : a.kchunks / 2;
    g.units = a.kchunks == 1 ? 5 : 9;
    const int tmax = 512 / n_total;
    g.n_tg = (g.units + tmax - 1) / tmax;
    g.n_total = n_total; g.dw = dw; g.dw_ld = dw_ld;
    g.a_stride = (kWgBlockBytes + 1023) / 1024 * 1024;
    g.a_dt = x->dtype; g.b_dt = dy->dtype;
    const int stage_bytes = g.nblk * g.a_stride + (n_total / 64) * 16384;
    g.stages = (227 * 1024 - 2048) / stage_bytes;
    if (g.stages > kWgMaxStages) g.stages = kWgMaxStages;
    const long long Tp = (long long)g.N * g.tiles_h * g.tiles_w;
    const long long mn = (long long)g.npairs * g.n_tg;
    const int sms = sm_count();
    double best = 1e30;
    int sp = 1;
    for (long long s = 1; s <= Tp && s * mn <= 2LL * sms + mn; ++s) {     // the epilogue is not overlapped: few, long items
      const long long waves = (s * mn + sms - 1) / sms;
      const double cost = (double)waves * ((double)((Tp + s - 1) / s) + 8.0);
      if (cost < best) { best = cost; sp = (int)s; }
    }
    g.splits = sp;
    if (g.stages >= 2) {
      CUtensorMap mx, mdy;
      uint64_t dims[4] = {(uint64_t)x->c, (uint64_t)x->w, (uint64_t)x->h, (uint64_t)x->n};
      uint64_t str[3] = {(uint64_t)x->pix_stride * 2, (uint64_t)x->row_stride * 2, (uint64_t)x->img_stride * 2};
      uint32_t box[4] = {64, 10, 18, 1};
      if ((rc = make_map(&mx, x->ptr, 4, dims, str, box, x->dtype)) != HPRI_OK) return rc;
      hpri_view_t dyv = *dy;
      dyv.c = n_total;
      if ((rc = map_nhwc(&mdy, dyv, 16, 8)) != HPRI_OK) return rc;
      const long long items = mn * g.splits;
      const long long grid = items < sms ? items : sms;
      const size_t smem = 2048 + (size_t)g.stages * stage_bytes;
      auto kern64 = wgrad3x3_halo_kernel<64>;
      auto kern128 = wgrad3x3_halo_kernel<128>;
      static std::once_flag once;
      static cudaError_t attr_err = cudaSuccess;
      std::call_once(once, [&] {
        attr_err = cudaFuncSetAttribute(kern64, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        if (attr_err == cudaSuccess)
          attr_err = cudaFuncSetAttribute(kern128, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
      });
      if (attr_err != cudaSuccess) return HPRI_ERR_CUDA;
      const cudaError_t le = n_total == 64 ? launch_ex(kern64, grid, smem, stream, false, mx, mdy, g)
                                           : launch_ex(kern128, grid, smem, stream, false, mx, mdy, g);
      return le == cudaSuccess ? HPRI_OK : HPRI_ERR_CUDA;
    }
  }
  int bn;
  if (mode == 2) {
    a.cout = n_total / 4;
    if (n_total % 4 || a.cout % 64 || dy->c != a.cout) return HPRI_ERR_ARG;
    bn = pick_block_n(a.cout, block_n);
    while (a.cout % bn) bn >>= 1;
  } else {
    if (dy->n != x->n || dy->h != x->h || dy->w != x->w) return HPRI_ERR_ARG;
    a.cout = n_total;
    bn = pick_block_n(n_total, block_n);
  }
  const long long T = (long long)a.N * a.tiles_h * a.tiles_w;
  const int m_tiles = (a.total_chunks + 1) / 2, n_tiles = (n_total + bn - 1) / bn;
  if (splits <= 0) {
    // persistent CTAs walk (m, n, split) tiles in waves of #SMs: pick the split count that minimises
    // waves x (k-blocks per tile + epilogue cost), i.e. never leave a wave with a single straggler tile.
    const long long mn = (long long)m_tiles * n_tiles;
    const double epi = 2.0;                             // tile epilogue (mostly overlapped) in k-block units
    const int sms = sm_count();
    double best = 1e30;
    splits = 1;
    for (long long s = 1; s <= T && s * mn <= 4LL * sms + mn; ++s) {
      const long long waves = (s * mn + sms - 1) / sms;
      const double cost = (double)waves * ((double)((T + s - 1) / s) + epi);
      if (cost < best) { best = cost; splits = (int)s; }
    }
  }
  if (splits > T) splits = (int)T;
  a.splits = splits;
  CUtensorMap ma, mb0, mb1;
  if ((rc = map_nhwc(&ma, *x, a.th, a.tw)) != HPRI_OK) return rc;
  if (mode == 2) {
    if ((rc = map_up2(&mb0, *dy, a.H, a.W, 0, a.th, a.tw)) != HPRI_OK) return rc;
    if ((rc = map_up2(&mb1, *dy, a.H, a.W, 1, a.th, a.tw)) != HPRI_OK) return rc;
  } else {
    if ((rc = map_nhwc(&mb0, *dy, a.th, a.tw)) != HPRI_OK) return rc;
    mb1 = mb0;
  }
  OutMaps o;
  o.m[0] = o.m[1] = o.m[2] = o.m[3] = ma;             // unused by the wgrad epilogue
  const long long grid = (long long)m_tiles * n_tiles * splits;
  return launch<MODE_WGRAD>(bn, ma, ma, mb0, mb1, o, a, grid, stream);
}
