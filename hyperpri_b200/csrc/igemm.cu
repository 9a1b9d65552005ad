// Implicit-GEMM convolution / GEMM family on tcgen05 (sm_100a).
//
// One warp-specialised kernel template covers every dense contraction on the hot path:
//   MODE_FWD   D[pixels, n] = sum_k A[pixels (+tap shift), k] * B[n, k]        (K-major operands)
//              3x3 pad-1 conv fprop and dgrad (model_parts.py:22,25), 1x1 / Linear (models.py:108),
//              ConvTranspose2d k2 s2 fprop (pixel-shuffle epilogue) and dgrad (2x2 stride-2 gather)
//              (model_parts.py:63-64).
//   MODE_WGRAD dW[n, (tap, c)] += sum_pixels X[pixel (+tap shift), c] * dY[pixel, n]   (MN-major operands)
//
// A operand tiles arrive by TMA straight from NHWC bf16 activations: a 4-D box {64 ch, tw, th, 1}
// at (c0, w0+dw, h0+dh, n).  Out-of-bounds box elements are zero-filled by the TMA unit, which
// implements the conv's zero padding, ragged channel counts (238 -> 240 -> 4 chunks of 64) and
// partial tiles at the image border for free.  Rows of 64 bf16 = 128 B land in the 128B-swizzled
// layout tcgen05.mma consumes.  Accumulators live in TMEM (128 lanes x BLOCK_N fp32 columns).
//
// Warp roles (192 threads): warp 0 = TMA producer, warp 1 = TMEM allocator + MMA issuer (one
// elected thread), warps 2..5 = epilogue (TMEM -> registers -> swizzled smem staging ->
// coalesced 16 B global stores, per-channel BatchNorm statistics from the staged tile).
#include "ptx.cuh"
#include "hyperpri_b200.h"

#include <mutex>

namespace hpri {

enum { MODE_FWD = 0, MODE_WGRAD = 1 };
enum { TAP_NONE = 0, TAP_3X3 = 1, TAP_UP2 = 2 };

struct IgemmArgs {
  int N, H, W;            // pixel grid walked by M (fwd) or K (wgrad)
  int th, tw, tiles_h, tiles_w;
  int taps, tap_mode, kchunks;
  int n_total;            // logical extent of the GEMM N dimension
  // ---- fwd epilogue
  uint16_t* out;          // 16-bit elements, format out_dt
  long long out_pix_stride, out_row_stride, out_img_stride;   // elements
  int out_h, out_w;       // store bounds (for up2: the upsampled extent)
  int n_store;            // channels to store per pixel (multiple of 8)
  int up2, cout;          // ConvT pixel shuffle: n = (a*2+b)*cout + co
  const float* bias;      // [n_total] (up2: [cout]) or null
  int accum;              // 1: y += result (read-modify-write of the bf16 destination)
  double* stats;          // [n_total][2] sum / sumsq or null
  // ---- wgrad epilogue
  float* dw;              // [n_total][dw_ld] fp32, accumulated with red.add
  int dw_ld, splits, total_chunks;
  int a_dt, b_dt, out_dt; // element formats (DT_BF16 / DT_F16) of the A, B operands and the fwd output
};

constexpr int kThreads = 192;
constexpr int A_BYTES = 16384;  // 128 rows x 128 B (fwd) or 2 chunks x 64 rows x 128 B (wgrad)

template <int BLOCK_N, int STAGES>
struct SmemLayout {
  static constexpr int B_BYTES = BLOCK_N * 128;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int PIPE_BYTES = STAGES * STAGE_BYTES;
  static constexpr int BAR_OFF = PIPE_BYTES;                 // full[STAGES], empty[STAGES], tmem_full
  static constexpr int TMEMPTR_OFF = BAR_OFF + (2 * STAGES + 1) * 8;
  static constexpr int VALID_OFF = TMEMPTR_OFF + 8;          // 128 row-valid bytes
  static constexpr int SSUM_OFF = VALID_OFF + 128;           // float[BLOCK_N] x 2
  static constexpr int TOTAL = SSUM_OFF + 2 * BLOCK_N * 4;
  static constexpr int ALLOC = TOTAL + 1024;                 // slack for manual 1024 B alignment
  static_assert(128 * BLOCK_N * 2 <= PIPE_BYTES, "epilogue staging must fit in the pipeline buffers");
};

template <int BLOCK_N, int STAGES, int MODE>
__global__ void __launch_bounds__(kThreads)
igemm_kernel(const __grid_constant__ CUtensorMap tmA0, const __grid_constant__ CUtensorMap tmA1,
             const __grid_constant__ CUtensorMap tmB0, const __grid_constant__ CUtensorMap tmB1,
             const IgemmArgs p) {
  using L = SmemLayout<BLOCK_N, STAGES>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + L::BAR_OFF);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full = empty_bar + STAGES;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(smem + L::TMEMPTR_OFF);
  uint8_t* row_valid = smem + L::VALID_OFF;
  float* ssum = reinterpret_cast<float*>(smem + L::SSUM_OFF);
  float* ssq = ssum + BLOCK_N;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  // ---------------- tile coordinates
  const int n_tiles = (p.n_total + BLOCK_N - 1) / BLOCK_N;
  int n_tile, m_tile, split = 0;
  {
    int bid = blockIdx.x;
    if (MODE == MODE_FWD) {
      n_tile = bid % n_tiles;
      m_tile = bid / n_tiles;
    } else {
      const int m_tiles = (p.total_chunks + 1) / 2;
      const int per_split = m_tiles * n_tiles;
      split = bid / per_split;
      bid -= split * per_split;
      n_tile = bid % n_tiles;
      m_tile = bid / n_tiles;
    }
  }
  const int n0 = n_tile * BLOCK_N;
  const int tiles_per_img = p.tiles_h * p.tiles_w;

  // K-loop extent
  int kb_begin, kb_end;   // fwd: k-blocks (tap, chunk); wgrad: pixel tiles
  if (MODE == MODE_FWD) {
    kb_begin = 0;
    kb_end = p.taps * p.kchunks;
  } else {
    const long long T = static_cast<long long>(p.N) * tiles_per_img;
    kb_begin = static_cast<int>(T * split / p.splits);
    kb_end = static_cast<int>(T * (split + 1) / p.splits);
  }
  const int num_kb = kb_end - kb_begin;

  // fwd: pixel tile of this CTA
  int img = 0, h0 = 0, w0 = 0;
  if (MODE == MODE_FWD) {
    img = m_tile / tiles_per_img;
    const int r = m_tile - img * tiles_per_img;
    h0 = (r / p.tiles_w) * p.th;
    w0 = (r % p.tiles_w) * p.tw;
  }

  // ---------------- one-time setup
  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmA0);
    tma_prefetch_desc(&tmB0);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(tmem_full, 1);
    fence_mbar_init();
    fence_proxy_async_smem();
  }
  if (warp == 1) {
    tmem_alloc(tmem_ptr, BLOCK_N);
    tmem_relinquish();
  }
  for (int i = threadIdx.x; i < 2 * BLOCK_N; i += kThreads) ssum[i] = 0.f;
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp == 0) {
    // =========================== TMA producer ===========================
    if (lane == 0) {
      for (int i = 0; i < num_kb; ++i) {
        const int s = i % STAGES;
        const uint32_t ph = (i / STAGES) & 1;
        mbar_wait(&empty_bar[s], ph ^ 1);
        mbar_arrive_expect_tx(&full_bar[s], L::STAGE_BYTES);
        uint8_t* sA = smem + s * L::STAGE_BYTES;
        uint8_t* sB = sA + A_BYTES;
        if (MODE == MODE_FWD) {
          const int kb = i;
          const int tap = kb / p.kchunks;
          const int cc = kb - tap * p.kchunks;
          if (p.tap_mode == TAP_UP2) {
            // ConvT dgrad: gather dy[2h+a, 2w+b]; 5-D view (c, w, a, h, n), one map per b
            tma_load_5d(sA, (tap & 1) ? &tmA1 : &tmA0, &full_bar[s], cc * 64, w0, tap >> 1, h0, img);
          } else {
            int dh = 0, dw = 0;
            if (p.tap_mode == TAP_3X3) { dh = tap / 3 - 1; dw = tap % 3 - 1; }
            tma_load_4d(sA, &tmA0, &full_bar[s], cc * 64, w0 + dw, h0 + dh, img);
          }
          tma_load_2d(sB, &tmB0, &full_bar[s], kb * 64, n0);
        } else {
          // wgrad: k-block = one pixel tile of 64 pixels
          const int t = kb_begin + i;
          const int im = t / tiles_per_img;
          const int r = t - im * tiles_per_img;
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
            tma_load_4d(sA + half * 8192, &tmA0, &full_bar[s], cc * 64, pw0 + dw, ph0 + dh, im);
          }
#pragma unroll
          for (int j = 0; j < BLOCK_N / 64; ++j) {
            if (p.tap_mode == TAP_UP2) {
              const int ab = n0 / p.cout;
              const int co0 = n0 - ab * p.cout;
              tma_load_5d(sB + j * 8192, (ab & 1) ? &tmB1 : &tmB0, &full_bar[s], co0 + j * 64, pw0, ab >> 1, ph0, im);
            } else {
              tma_load_4d(sB + j * 8192, &tmB0, &full_bar[s], n0 + j * 64, pw0, ph0, im);
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // =========================== MMA issuer ===========================
    if (lane == 0) {
      const uint32_t idesc = make_idesc_16(128, BLOCK_N, MODE == MODE_WGRAD, MODE == MODE_WGRAD, p.a_dt, p.b_dt);
      for (int i = 0; i < num_kb; ++i) {
        const int s = i % STAGES;
        const uint32_t ph = (i / STAGES) & 1;
        mbar_wait(&full_bar[s], ph);
        tc_fence_after();
        const uint32_t a_addr = smem_u32(smem + s * L::STAGE_BYTES);
        const uint32_t b_addr = a_addr + A_BYTES;
        uint64_t da, db;
        uint32_t kstep;   // descriptor start-address advance (in 16 B units) per UMMA_K = 16
        if (MODE == MODE_FWD) {
          da = make_smem_desc_sw128(a_addr, 16, 1024);
          db = make_smem_desc_sw128(b_addr, 16, 1024);
          kstep = 32 >> 4;          // 16 bf16 along the 128 B swizzle row
        } else {
          da = make_smem_desc_sw128(a_addr, 8192, 1024);   // LBO: next 64-channel group, SBO: next 8 pixel rows
          db = make_smem_desc_sw128(b_addr, 8192, 1024);
          kstep = 2048 >> 4;        // 16 pixel rows x 128 B
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          umma_bf16(tmem_base, da + static_cast<uint64_t>(k * kstep), db + static_cast<uint64_t>(k * kstep), idesc,
                    (i | k) != 0);
        }
        umma_commit(&empty_bar[s]);      // frees the smem slot once these MMAs have read it
      }
      umma_commit(tmem_full);            // accumulator complete
    }
  } else {
    // =========================== epilogue (warps 2..5) ===========================
    const int q = warp & 3;                 // TMEM lane quarter this warp may read
    const int row = q * 32 + lane;
    const int et = threadIdx.x - 64;        // 0..127
    if (MODE == MODE_FWD) {
      const int hl = row / p.tw, wl = row - hl * p.tw;
      const bool valid = (h0 + hl < p.H) && (w0 + wl < p.W);
      row_valid[row] = valid ? 1 : 0;
      mbar_wait(tmem_full, 0);
      tc_fence_after();
      uint8_t* stage = smem;                // pipeline buffers are idle once tmem_full fired
      constexpr int ROWB = BLOCK_N * 2;
      const int co_base = p.up2 ? (n0 % p.cout) : n0;
#pragma unroll 1
      for (int c = 0; c < BLOCK_N / 32; ++c) {
        uint32_t v[32];
        tmem_ld32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + c * 32, v);
        tmem_ld_wait();
        if (p.bias != nullptr) {
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const int ch = co_base + c * 32 + j;
            const float b = (ch < (p.up2 ? p.cout : p.n_total)) ? __ldg(p.bias + ch) : 0.f;
            v[j] = __float_as_uint(__uint_as_float(v[j]) + b);
          }
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          uint4 o;
          o.x = pack2(__uint_as_float(v[8 * i + 0]), __uint_as_float(v[8 * i + 1]), p.out_dt);
          o.y = pack2(__uint_as_float(v[8 * i + 2]), __uint_as_float(v[8 * i + 3]), p.out_dt);
          o.z = pack2(__uint_as_float(v[8 * i + 4]), __uint_as_float(v[8 * i + 5]), p.out_dt);
          o.w = pack2(__uint_as_float(v[8 * i + 6]), __uint_as_float(v[8 * i + 7]), p.out_dt);
          const int chunk = c * 4 + i;
          *reinterpret_cast<uint4*>(stage + row * ROWB + (((chunk & ~7) | ((chunk ^ row) & 7)) << 4)) = o;
        }
      }
      tc_fence_before();
      named_bar_sync(1, 128);
      // ---- per-channel statistics over the valid rows of the staged (bf16-rounded) tile
      if (p.stats != nullptr) {
        const int ew = warp - 2;
        for (int cp = lane; cp < BLOCK_N / 2; cp += 32) {
          float s0 = 0.f, s1 = 0.f, q0 = 0.f, q1 = 0.f;
          const int chunk = cp >> 2, word = cp & 3;
          for (int r = ew * 32; r < ew * 32 + 32; ++r) {
            if (!row_valid[r]) continue;
            const uint32_t u = *reinterpret_cast<const uint32_t*>(
                stage + r * ROWB + (((chunk & ~7) | ((chunk ^ r) & 7)) << 4) + word * 4);
            const float2 ab = unpack2(u, p.out_dt);
            s0 += ab.x; s1 += ab.y; q0 += ab.x * ab.x; q1 += ab.y * ab.y;
          }
          atomicAdd(&ssum[2 * cp], s0);
          atomicAdd(&ssum[2 * cp + 1], s1);
          atomicAdd(&ssq[2 * cp], q0);
          atomicAdd(&ssq[2 * cp + 1], q1);
        }
        named_bar_sync(1, 128);
        for (int ch = et; ch < BLOCK_N; ch += 128) {
          if (n0 + ch < p.n_total) {
            atomicAdd(p.stats + 2 * (n0 + ch), static_cast<double>(ssum[ch]));
            atomicAdd(p.stats + 2 * (n0 + ch) + 1, static_cast<double>(ssq[ch]));
          }
        }
      }
      // ---- coalesced store: consecutive threads take consecutive 16 B chunks of a pixel row
      constexpr int CPR = BLOCK_N / 8;
      int a_off = 0, b_off = 0;
      if (p.up2) { const int ab = n0 / p.cout; a_off = ab >> 1; b_off = ab & 1; }
#pragma unroll 4
      for (int i = 0; i < CPR; ++i) {
        const int idx = et + 128 * i;
        const int r = idx / CPR, chunk = idx % CPR;
        if (!row_valid[r]) continue;
        if (co_base + chunk * 8 >= p.n_store) continue;
        const int rh = r / p.tw, rw = r - rh * p.tw;
        int oh = h0 + rh, ow = w0 + rw;
        if (p.up2) { oh = 2 * oh + a_off; ow = 2 * ow + b_off; }
        if (oh >= p.out_h || ow >= p.out_w) continue;
        uint4 val = *reinterpret_cast<const uint4*>(stage + r * ROWB + (((chunk & ~7) | ((chunk ^ r) & 7)) << 4));
        uint16_t* dst = p.out + img * p.out_img_stride + oh * p.out_row_stride + ow * p.out_pix_stride +
                        co_base + chunk * 8;
        if (p.accum) {
          const uint4 old = *reinterpret_cast<const uint4*>(dst);
          const int dt = p.out_dt;
          float2 a, b;
          a = unpack2(val.x, dt); b = unpack2(old.x, dt); val.x = pack2(a.x + b.x, a.y + b.y, dt);
          a = unpack2(val.y, dt); b = unpack2(old.y, dt); val.y = pack2(a.x + b.x, a.y + b.y, dt);
          a = unpack2(val.z, dt); b = unpack2(old.z, dt); val.z = pack2(a.x + b.x, a.y + b.y, dt);
          a = unpack2(val.w, dt); b = unpack2(old.w, dt); val.w = pack2(a.x + b.x, a.y + b.y, dt);
        }
        *reinterpret_cast<uint4*>(dst) = val;
      }
    } else {
      // wgrad: rows = (chunk half, channel j); columns = n.  fp32 red.add into dW[n][k]
      if (num_kb > 0) {
        mbar_wait(tmem_full, 0);
        tc_fence_after();
        const int qc = 2 * m_tile + (row >> 6);
        const bool row_ok = qc < p.total_chunks;
        const long long kidx = static_cast<long long>(qc) * 64 + (row & 63);
#pragma unroll 1
        for (int c = 0; c < BLOCK_N / 32; ++c) {
          uint32_t v[32];
          tmem_ld32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + c * 32, v);
          tmem_ld_wait();
          if (row_ok) {
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              const int n = n0 + c * 32 + j;
              if (n < p.n_total) red_add_f32(p.dw + static_cast<long long>(n) * p.dw_ld + kidx, __uint_as_float(v[j]));
            }
          }
        }
        tc_fence_before();
      }
    }
  }
  __syncwarp();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, BLOCK_N);
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

template <int BLOCK_N, int STAGES, int MODE>
static int launch_t(const CUtensorMap& a0, const CUtensorMap& a1, const CUtensorMap& b0, const CUtensorMap& b1,
                    const IgemmArgs& args, long long grid, cudaStream_t stream) {
  using L = SmemLayout<BLOCK_N, STAGES>;
  auto kern = igemm_kernel<BLOCK_N, STAGES, MODE>;
  static std::once_flag once;
  static cudaError_t attr_err = cudaSuccess;
  std::call_once(once, [&] {
    attr_err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, L::ALLOC);
  });
  if (attr_err != cudaSuccess) return HPRI_ERR_CUDA;
  if (grid <= 0 || grid > 0x7FFFFFFFLL) return HPRI_ERR_ARG;
  kern<<<(unsigned)grid, kThreads, L::ALLOC, stream>>>(a0, a1, b0, b1, args);
  ++g_launch_count;
  return cudaGetLastError() == cudaSuccess ? HPRI_OK : HPRI_ERR_CUDA;
}

template <int MODE>
static int launch(int block_n, const CUtensorMap& a0, const CUtensorMap& a1, const CUtensorMap& b0,
                  const CUtensorMap& b1, const IgemmArgs& args, long long grid, cudaStream_t stream) {
  switch (block_n) {
    case 64: return launch_t<64, 4, MODE>(a0, a1, b0, b1, args, grid, stream);
    case 128: return launch_t<128, 3, MODE>(a0, a1, b0, b1, args, grid, stream);
    case 256: return launch_t<256, 4, MODE>(a0, a1, b0, b1, args, grid, stream);
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

}  // namespace hpri

using namespace hpri;

// -------------------------------------------------------------------------------------
// C ABI
// -------------------------------------------------------------------------------------
extern "C" int hpri_igemm_fwd(const hpri_view_t* x, const void* wpack, int w_dtype, int w_rows, int kpad, int taps,
                              const hpri_view_t* y, int n_store, const float* bias, double* stats, int accumulate,
                              int block_n, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (!x || !y || !wpack) return HPRI_ERR_ARG;
  int rc;
  if ((rc = check_view(*x)) != HPRI_OK || (rc = check_view(*y)) != HPRI_OK) return rc;
  if (taps != 1 && taps != 9) return HPRI_ERR_ARG;
  if (x->n != y->n || x->h != y->h || x->w != y->w) return HPRI_ERR_ARG;
  if ((reinterpret_cast<uintptr_t>(wpack) & 15) || (kpad & 63) || (n_store & 7) || n_store > y->c) return HPRI_ERR_ALIGN;
  const int kchunks = (x->c + 63) / 64;
  if (kpad != taps * kchunks * 64) return HPRI_ERR_ARG;
  IgemmArgs a{};
  a.N = x->n; a.H = x->h; a.W = x->w;
  pick_tile(a.H, a.W, 128, &a.th, &a.tw);
  a.tiles_h = (a.H + a.th - 1) / a.th; a.tiles_w = (a.W + a.tw - 1) / a.tw;
  a.taps = taps; a.tap_mode = taps == 9 ? TAP_3X3 : TAP_NONE; a.kchunks = kchunks;
  a.n_total = w_rows;
  a.out = static_cast<uint16_t*>(y->ptr);
  a.out_pix_stride = y->pix_stride; a.out_row_stride = y->row_stride; a.out_img_stride = y->img_stride;
  a.out_h = y->h; a.out_w = y->w; a.n_store = n_store; a.up2 = 0; a.cout = w_rows;
  a.a_dt = x->dtype; a.b_dt = w_dtype; a.out_dt = y->dtype;
  if (x->dtype != w_dtype) return HPRI_ERR_ARG;     // kind::f16 takes A and B in one format
  a.bias = bias; a.stats = stats; a.accum = accumulate ? 1 : 0;
  if (accumulate && stats) return HPRI_ERR_ARG;
  const int bn = pick_block_n(w_rows, block_n);
  CUtensorMap ma, mb;
  if ((rc = map_nhwc(&ma, *x, a.th, a.tw)) != HPRI_OK) return rc;
  if ((rc = map_weights(&mb, wpack, w_rows, kpad, bn, w_dtype)) != HPRI_OK) return rc;
  const long long grid = (long long)a.N * a.tiles_h * a.tiles_w * ((w_rows + bn - 1) / bn);
  return launch<MODE_FWD>(bn, ma, ma, mb, mb, a, grid, stream);
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
  if (kpad != kchunks * 64 || (cout & 63)) return HPRI_ERR_ARG;
  IgemmArgs a{};
  a.N = x->n; a.H = x->h; a.W = x->w;
  pick_tile(a.H, a.W, 128, &a.th, &a.tw);
  a.tiles_h = (a.H + a.th - 1) / a.th; a.tiles_w = (a.W + a.tw - 1) / a.tw;
  a.taps = 1; a.tap_mode = TAP_NONE; a.kchunks = kchunks; a.n_total = 4 * cout;
  a.out = static_cast<uint16_t*>(y->ptr);
  a.out_pix_stride = y->pix_stride; a.out_row_stride = y->row_stride; a.out_img_stride = y->img_stride;
  a.out_h = y->h; a.out_w = y->w; a.n_store = cout; a.up2 = 1; a.cout = cout;
  a.a_dt = x->dtype; a.b_dt = w_dtype; a.out_dt = y->dtype;
  if (x->dtype != w_dtype) return HPRI_ERR_ARG;
  a.bias = bias; a.stats = nullptr;
  int bn = pick_block_n(cout, block_n);
  while (cout % bn) bn >>= 1;
  CUtensorMap ma, mb;
  if ((rc = map_nhwc(&ma, *x, a.th, a.tw)) != HPRI_OK) return rc;
  if ((rc = map_weights(&mb, wpack, 4 * cout, kpad, bn, w_dtype)) != HPRI_OK) return rc;
  const long long grid = (long long)a.N * a.tiles_h * a.tiles_w * (4 * cout / bn);
  return launch<MODE_FWD>(bn, ma, ma, mb, mb, a, grid, stream);
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
  pick_tile(a.H, a.W, 128, &a.th, &a.tw);
  a.tiles_h = (a.H + a.th - 1) / a.th; a.tiles_w = (a.W + a.tw - 1) / a.tw;
  a.taps = 4; a.tap_mode = TAP_UP2; a.kchunks = kchunks; a.n_total = cin;
  a.out = static_cast<uint16_t*>(dx->ptr);
  a.out_pix_stride = dx->pix_stride; a.out_row_stride = dx->row_stride; a.out_img_stride = dx->img_stride;
  a.out_h = dx->h; a.out_w = dx->w; a.n_store = (cin + 7) & ~7; a.up2 = 0; a.cout = cin;
  a.a_dt = dy->dtype; a.b_dt = w_dtype; a.out_dt = dx->dtype;
  if (dy->dtype != w_dtype) return HPRI_ERR_ARG;
  const int bn = pick_block_n(cin, block_n);
  CUtensorMap m0, m1, mb;
  if ((rc = map_up2(&m0, *dy, a.H, a.W, 0, a.th, a.tw)) != HPRI_OK) return rc;
  if ((rc = map_up2(&m1, *dy, a.H, a.W, 1, a.th, a.tw)) != HPRI_OK) return rc;
  if ((rc = map_weights(&mb, wpack, cin, kpad, bn, w_dtype)) != HPRI_OK) return rc;
  const long long grid = (long long)a.N * a.tiles_h * a.tiles_w * ((cin + bn - 1) / bn);
  return launch<MODE_FWD>(bn, m0, m1, mb, mb, a, grid, stream);
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
  pick_tile(a.H, a.W, 64, &a.th, &a.tw);
  a.tiles_h = (a.H + a.th - 1) / a.th; a.tiles_w = (a.W + a.tw - 1) / a.tw;
  a.taps = mode == 1 ? 9 : 1; a.tap_mode = mode == 1 ? TAP_3X3 : (mode == 2 ? TAP_UP2 : TAP_NONE);
  a.kchunks = (x->c + 63) / 64;
  a.total_chunks = a.taps * a.kchunks;
  if (dw_ld != a.total_chunks * 64) return HPRI_ERR_ARG;
  a.n_total = n_total; a.dw = dw; a.dw_ld = dw_ld;
  a.a_dt = x->dtype; a.b_dt = dy->dtype; a.out_dt = DT_BF16;
  if (x->dtype != dy->dtype) return HPRI_ERR_ARG;   // kind::f16 takes A and B in one format
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
    // enough CTAs for ~4 waves of 148 SMs x 2 resident, at least 8 pixel tiles per CTA
    long long want = (4LL * 296 + m_tiles * n_tiles - 1) / (m_tiles * n_tiles);
    long long cap = T / 8 > 0 ? T / 8 : 1;
    splits = (int)(want < cap ? want : cap);
    if (splits < 1) splits = 1;
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
  const long long grid = (long long)m_tiles * n_tiles * splits;
  return launch<MODE_WGRAD>(bn, ma, ma, mb0, mb1, a, grid, stream);
}
