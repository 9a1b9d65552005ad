/* hyperpri_b200 -- C ABI of the B200-native HyperPRI segmentation hot path.
 *
 * The reference (GatorSense/HyperPRI) has no FFI: every FLOP on its hot path is a stock
 * torch.nn call (cuDNN / cuBLAS / ATen).  Each entry point below replaces one of those
 * library call sites (cited file:line, relative to the reference root).  All functions
 *   - take raw DEVICE pointers, plain sizes and a cudaStream_t passed as void*,
 *   - enqueue work on that stream and return immediately (no allocation, no host sync),
 *   - return HPRI_OK (0) or a negative HPRI_ERR_* code; they never fall back to a CPU path.
 *
 * Activations are NHWC 16-bit (fp16 in the engine; bf16 is accepted), described by hpri_view_t so that channel
 * sub-ranges of a concat buffer and cropped windows are expressed with strides instead of copies.
 */
#ifndef HYPERPRI_B200_H
#define HYPERPRI_B200_H

#ifdef __cplusplus
extern "C" {
#endif

#define HPRI_OK 0
#define HPRI_ERR_ARG (-1)       /* inconsistent shapes / null pointer / unsupported option */
#define HPRI_ERR_ALIGN (-2)     /* pointer or stride not 16-byte aligned */
#define HPRI_ERR_DRIVER (-3)    /* cuTensorMapEncodeTiled not obtainable from the driver */
#define HPRI_ERR_TENSORMAP (-4) /* the driver rejected a tensor map */
#define HPRI_ERR_CUDA (-5)      /* launch failed; see cudaGetLastError */

#define HPRI_BF16 0
#define HPRI_F16 1

/* NHWC view of a 16-bit tensor (dtype HPRI_BF16 or HPRI_F16): element strides; c = logical channels
 * visible through the view.  Activations and gradients are fp16 (BatchNorm keeps activations O(1); the loss
 * gradient is scaled by a power of two so gradients stay in fp16's normal range); all accumulation is fp32. */
typedef struct {
  void* ptr;
  int n, h, w, c;
  long long pix_stride, row_stride, img_stride;
  int dtype;
} hpri_view_t;

int hpri_abi_version(void);
/* number of CUDA kernels this library has launched in this process (bench.py's gpu_launches) */
long long hpri_launch_count(void);

/* ---- tensor-core contractions (tcgen05 implicit GEMM, csrc/igemm.cu) ------------------- */

/* y[n,h,w,:] = sum_{tap,c} x[n,h+dh,w+dw,c] * wpack[:, tap*kc*64+c] (+bias); taps = 9 (3x3, pad 1)
 * or 1 (1x1 / Linear).  Replaces nn.Conv2d 3x3 fprop AND dgrad (model_parts.py:22,25 -- dgrad uses a
 * transposed/flipped pack), Conv3d(1,64,(D,3,3)) (models.py:169), nn.Linear (models.py:108,102).
 * stats (nullable): double[w_rows][2] accumulating per-channel sum / sum-of-squares of the bf16
 * outputs for train-mode BatchNorm (model_parts.py:23,26; models.py:113,172,178).
 * accumulate=1: y += result (skip-gradient accumulation), not combinable with stats.
 * fin (nullable, needs stats): train-mode BatchNorm finalisation fused into the launch -- the last CTA to flush its
 * statistics (ticket counter) does what hpri_bn_finalize(training=1) does and zeroes stats and the counter. */
typedef struct {
  const float* gamma;          /* [C] or null (1) */
  const float* beta;           /* [C] or null (0) */
  const float* conv_bias;      /* [C] or null: re-added to running_mean (it cancels in the normalised output) */
  float* running_mean;         /* nullable */
  float* running_var;          /* nullable */
  long long* num_batches_tracked; /* nullable */
  float* scale;                /* out [C]: gamma * invstd */
  float* shift;                /* out [C]: beta - mean * scale */
  float* save_mean;            /* out [C], nullable */
  float* save_invstd;          /* out [C], nullable */
  unsigned int* counter;       /* one zero-initialised word owned by the layer */
  long long count;             /* elements per channel */
  float momentum, eps;
  float* partials;             /* nullable.  Deterministic statistics: scratch of >= 148 * w_rows * 2 floats.  Every CTA
                                * stores its per-channel partial (sum, sumsq) in its own slot -- its warps' contributions
                                * added in a fixed order -- and the last CTA adds the slots in CTA order (fp64), instead of
                                * fp64 atomics in arrival order: two runs on the same input give bit-identical outputs. */
} hpri_bn_fin_t;
/* bw (nullable; 3x3 dgrad launches on the halo kernel only, see hpri_conv3x3_halo_ok): the output y is the gradient dy
 * of the BatchNorm+ReLU layer below; the epilogue also accumulates pass 1 of its backward (what
 * hpri_bn_relu_bwd_reduce computes from x and dy) into sums, so only hpri_bn_relu_bwd_apply remains.  The caller zeroes
 * sums before the launch. */
typedef struct {
  const hpri_view_t* x;        /* raw conv output of the layer below: same n, h, w, channels and dtype as y */
  const float* scale;          /* its BatchNorm scale / shift (ReLU mask = x*scale+shift > 0) */
  const float* shift;
  const float* save_mean;
  const float* save_invstd;
  double* sums;                /* [C][3]: sum dz, invstd * (sum dz*x - mean * sum dz), unused */
} hpri_bn_bwd_t;
int hpri_igemm_fwd(const hpri_view_t* x, const void* wpack, int w_dtype, int w_rows, int kpad, int taps, const hpri_view_t* y,
                   int n_store, const float* bias, double* stats, int accumulate, int block_n, const hpri_bn_fin_t* fin,
                   const hpri_bn_bwd_t* bw, void* stream);
/* 1 when hpri_igemm_fwd(taps = 9) for an h x w image and w_rows output channels runs on the halo-reuse kernel
 * (which is what the fused `bw` reduction needs), 0 when it falls to the generic per-tap kernel. */
int hpri_conv3x3_halo_ok(int h, int w, int w_rows);

/* 3x3 kernel selection: -1 heuristic (default: halo-reuse kernel on CTA pairs -- tcgen05 cta_group::2, M = 256 --
 * except the 64-channel dgrads carrying the fused `bw` reduction, which stay on single CTAs), 0 generic per-tap kernel,
 * 1 halo-reuse kernel on single CTAs, 2 CTA pairs everywhere.  Seeded by the environment variable HPRI_CONV_ALGO. */
int hpri_set_conv_algo(int algo);
/* 3x3 weight-gradient kernel selection: -1 heuristic (halo-reuse weight-gradient kernel at Cout 64 / 128), 0 generic
 * per-tap kernel only.  Seeded by the environment variable HPRI_WGRAD_ALGO. */
int hpri_set_wgrad_algo(int algo);
/* Leave `sms` streaming multiprocessors out of the persistent tcgen05 grids (0 = use every SM): room for kernels that
 * run concurrently with them, i.e. NCCL's all-reduce CTAs in the data-parallel backward.  Seeded by HPRI_SM_RESERVE. */
int hpri_set_sm_reserve(int sms);
/* Depth of the halo-block ring of the 3x3 halo kernel: 2 (default) or 3 (taken when at least four weight-tap slots still
 * fit in shared memory).  Seeded by HPRI_HALO_A_STAGES. */
int hpri_set_halo_a_stages(int stages);

/* nn.ConvTranspose2d(k=2,s=2) fprop writing straight into the concat buffer (model_parts.py:63-64,
 * 74-87: pad + cat are absorbed by the destination view) and its dgrad. */
int hpri_convT2x2_fwd(const hpri_view_t* x, const void* wpack, int w_dtype, int cout, int kpad, const hpri_view_t* y,
                      const float* bias, int block_n, void* stream);
int hpri_convT2x2_dgrad(const hpri_view_t* dy, const void* wpack, int w_dtype, int cin, int kpad, const hpri_view_t* dx,
                        int block_n, void* stream);

/* Weight gradients (autograd of the three layer kinds above). mode 0 Linear/1x1, 1 conv3x3, 2 convT2x2.
 * dw: fp32 [n_total][dw_ld] in forward-pack layout, accumulated into. */
int hpri_igemm_wgrad(const hpri_view_t* x, const hpri_view_t* dy, int mode, int n_total, float* dw, int dw_ld,
                     int block_n, int splits, void* stream);

/* ---- weight layout conversion (csrc/elementwise.cu) -------------------------------------
 * dst[(g*R + r)][t*kc64 + c] = c < C ? src[g*sg + r*sr + tm(t)*st + c*sc] : 0,  tm(t) = flip ? T-1-t : t.
 * pack: fp32 torch-layout parameter -> 16-bit operand (dst_dtype).  unpack: fp32 packed gradient -> fp32 torch layout. */
int hpri_pack_weights(const float* src, void* dst, int dst_dtype, int G, int R, int T, int C, int kc64, long long sg,
                      long long sr, long long st, long long sc, int flip, void* stream);
/* unpack: dst = beta * dst + scale * packed (scale removes the power-of-two loss scale of the fp16 gradient path);
 * a non-finite result sets *flag (nullable) -- the overflow signal hpri_adam_step reads on the device. */
int hpri_unpack_grads(float* packed, float* dst, int G, int R, int T, int C, int kc64, long long sg,
                      long long sr, long long st, long long sc, int flip, float beta, int zero_src, float scale,
                      int* flag, void* stream);

/* Tiled conv3x3 specialisations: W[co][ci][3][3] -> forward operand and (optional) transposed, tap-flipped dgrad
 * operand in one pass (destination buffers must be zero-initialised once: padding is never written); and the
 * packed fp32 weight gradient back to W layout.  unpack (zero_src != 0 for the generic one, always for the conv3x3
 * one) resets the packed buffer to zero behind the read, so the next split-K hpri_igemm_wgrad launch can accumulate
 * into it without a memset. */
int hpri_pack_conv3x3(const float* w, int cout, int cin, void* dst_fwd, int fwd_dtype, void* dst_dgrad,
                      int dgrad_dtype, void* stream);
int hpri_unpack_conv3x3(float* packed, int cout, int cin, float* dst, void* stream);
/* Table-driven variants: ONE launch refreshes the 16-bit operands of every 3x3 layer after an optimizer step / unpacks
 * the gradients of a whole bucket (18 launches of ~14 us each were latency-, not bandwidth-bound).  `jobs` is a DEVICE
 * array; tile0 = number of 32x32 (co, ci) tiles of all previous jobs; total_tiles = tiles of all jobs. */
typedef struct {
  const float* w;       /* kind 0: [cout][cin][3][3]; kind 1: [cin][cout][2][2] */
  void* dst_fwd;        /* kind 0: [cout][9*kpad(cin)]; kind 1: [4*cout][kpad(cin)]; or null */
  void* dst_dgrad;      /* kind 0: [cin][9*kpad(cout)]; kind 1: [cin][4*kpad(cout)]; or null */
  float* grad_packed;   /* unpack: fp32 in the dst_fwd layout (zeroed behind the read) */
  float* grad_dst;      /* unpack: fp32 in the layout of w */
  int cout, cin, fwd_dtype, dgrad_dtype;
  int tile0;            /* number of 32 x 32 (co, ci) tiles of all previous jobs */
  int kind;             /* 0 conv3x3, 1 ConvTranspose2d(k=2,s=2) */
} hpri_conv3x3_job_t;
int hpri_pack_conv3x3_batch(const hpri_conv3x3_job_t* jobs, int njobs, int total_tiles, void* stream);
/* grad_dst = scale * grad_packed for every job; a non-finite value sets *flag (nullable). */
int hpri_unpack_conv3x3_batch(const hpri_conv3x3_job_t* jobs, int njobs, int total_tiles, float scale, int* flag,
                              void* stream);

/* ---- validation maths (src/PLTrainer.py:538-583: binned PR curve with thresholds=500, best-Dice threshold, counts) ----
 * One pass over a batch of logits accumulates (+=, caller zeroes once per sweep):
 *   hist_pos / hist_neg [n_thr]   : pixels whose p = sigmoid(logit) falls in bin i = max{i : thr[i] <= p}, by mask class
 *                                   (torchmetrics' binned curve: tps[i] = sum_{j>=i} hist_pos[j])
 *   cut_pos / cut_neg [n_cut + 1] : pixels with exactly k of the cuts below p (cut[k-1] < p <= cut[k]), so the
 *                                   confusion counts at ANY threshold cut[k] (p > cut[k]) follow without a second pass
 *   bce_sum                       : sum of the stable BCE-with-logits terms (double)
 * thr and cut are ascending device arrays. */
int hpri_pr_hist(const float* logits, const float* target, long long numel, const float* thr, int n_thr,
                 const float* cut, int n_cut, unsigned long long* hist_pos, unsigned long long* hist_neg,
                 unsigned long long* cut_pos, unsigned long long* cut_neg, double* bce_sum, void* stream);

/* ---- optimizer (src/PLTrainer.py:171-174: optim.Adam(lr, weight_decay)) --------------------------------
 * One launch updates every parameter tensor (torch.optim.Adam semantics: L2 weight decay added to the gradient,
 * bias-corrected moments, denom = sqrt(v)/sqrt(1-beta2^t) + eps).  `jobs` is a DEVICE array; block0 = number of
 * 1024-element blocks of all previous jobs. */
typedef struct {
  float* param;
  const float* grad;
  float* exp_avg;
  float* exp_avg_sq;
  long long numel;
  int block0, pad_;
} hpri_adam_job_t;
/* found_inf (nullable, device): when *found_inf != 0 the launch leaves parameters and moments untouched -- the
 * gradients of this step overflowed the loss-scaled fp16 range (set by the unpack / BatchNorm-backward kernels). */
int hpri_adam_step(const hpri_adam_job_t* jobs, int njobs, int total_blocks, double lr, double beta1, double beta2,
                   double eps, double weight_decay, int step, const int* found_inf, void* stream);

/* ConvTranspose2d(k=2,s=2): W[ci][co][2][2] <-> forward operand [(a*2+b)*cout + co][kpad(ci)] (a (ci, co) transpose,
 * tiled through shared memory); unpack zeroes the packed gradient behind the read.  Padding columns are not written. */
int hpri_pack_convT2x2(const float* w, int cin, int cout, void* dst_fwd, int fwd_dtype, void* stream);
int hpri_unpack_convT2x2(float* packed, int cin, int cout, float* dst, void* stream);

/* ---- ingest (src/dataset.py:266-270, 284-289) -------------------------------------------
 * src: fp32 [n][bands_total][H][W]; keeps bands [lo,hi), crops the (i0,j0,h,w) window, optional
 * horizontal / vertical flip, optional scalar rescale (the '/255 if max>10' rule is decided by the caller
 * with hpri_absmax), optional per-band (x-mean)/std; writes NHWC 16-bit (dst_dtype) with c_pad channels (zero filled). */
int hpri_hsi_ingest(const float* src, int n, int bands_total, int H, int W, int lo, int hi, int i0, int j0, int h,
                    int w, int flip_h, int flip_w, float scale, const float* band_mean, const float* band_std,
                    void* dst, int dst_dtype, int c_pad, void* stream);
/* Same for a cube that the host already holds in IEEE half precision (src: __half [n][bands_total][H][W]): the data
 * loader converts fp32 -> fp16 before the PCIe copy, which halves the bytes of the only host->device transfer of the
 * step; with scale == 1 and no normalisation the network input is bit-identical to the fp32 route (same rounding). */
int hpri_hsi_ingest_f16(const void* src, int n, int bands_total, int H, int W, int lo, int hi, int i0, int j0, int h,
                        int w, int flip_h, int flip_w, float scale, const float* band_mean, const float* band_std,
                        void* dst, int dst_dtype, int c_pad, void* stream);
int hpri_absmax(const float* src, long long numel, float* out_max, void* stream);

/* y = x converted between HPRI_F16 and HPRI_BF16 (tcgen05 kind::f16 requires both operands of one
 * MMA in the same format: wgrad pairs a bf16 copy of the fp16 activations with the bf16 gradients). */
int hpri_convert16(const hpri_view_t* x, const hpri_view_t* y, void* stream);

/* y = a * b elementwise over equal-shaped views (fp32 product, rounded once).  Replaces `x = x2 * x1` of Up with
 * use_attention=True (reference src/Experiments/model_parts.py:84-85) and the two products of its backward. */
int hpri_mul16(const hpri_view_t* a, const hpri_view_t* b, const hpri_view_t* y, void* stream);

/* nn.Upsample(scale_factor=2, mode='bilinear', align_corners=True) (reference src/Experiments/model_parts.py:57) of
 * x (n, h, w, c) into the top-left 2h x 2w of y (n, >=2h, >=2w, c); the rest of y -- the zero padding Up.forward adds
 * to reach the skip's size (model_parts.py:77-80) -- is written as zeros.  _bwd: dx (n, h, w, c) from the gradient view
 * dy of y's shape (gather form, no atomics). */
int hpri_upsample2_fwd(const hpri_view_t* x, const hpri_view_t* y, void* stream);
int hpri_upsample2_bwd(const hpri_view_t* dy, const hpri_view_t* dx, void* stream);

/* ---- BatchNorm / ReLU / MaxPool family ---------------------------------------------------- */
/* Turn accumulated (sum, sumsq) into scale/shift, saved mean/invstd, and the running-stat update
 * (biased var to normalise, unbiased for running_var, momentum 0.1, conv bias re-added to the mean);
 * zeroes stats for the next step.  training=0: coefficients from the running stats. */
int hpri_bn_finalize(double* stats, long long count, const float* gamma, const float* beta, const float* conv_bias,
                     float* running_mean, float* running_var, long long* num_batches_tracked, float momentum,
                     float eps, int training, float* scale, float* shift, float* save_mean, float* save_invstd,
                     int C, void* stream);
/* Traversal order of the HBM-bound BatchNorm kernels: 1 = last chunk first, so that a tensor a tcgen05 kernel has just
 * written in ascending tile order is read starting with the part still in L2; 0 (default) = ascending -- measured, the
 * reversed order gains nothing on B200.  Seeded by the environment variable HPRI_REVERSE_ELEMENTWISE.  Values are
 * unchanged (reductions differ in summation order only). */
int hpri_set_reverse_elementwise(int on);
/* Run-to-run reproducible reductions (the reference builds its Trainer with deterministic='warn', PLTrainer.py:430,439,447):
 * 1 = hpri_colsum runs one CTA per eight channels and hpri_sum_f32 one CTA (their CTAs otherwise meet in fp32 atomics in
 * arrival order).  The other
 * order-dependent sums are selected per call: BatchNorm statistics by hpri_bn_fin_t::partials, split-K weight gradients
 * by splits = 1, the fused dgrad + BatchNorm-backward reduction by bw = NULL; hpri_bn_relu_bwd_reduce is reproducible as
 * it is (fixed-order per-CTA partials, fp32-valued addends combined exactly in fp64). */
int hpri_set_deterministic(int on);
/* y = relu(x*scale+shift) (model_parts.py:23-24); optional fused MaxPool2d(2) output (model_parts.py:40). */
int hpri_bn_relu_apply(const hpri_view_t* x, const float* scale, const float* shift, const hpri_view_t* y,
                       const hpri_view_t* pooled, void* stream);
/* Backward of relu(bn(x)): pass 1 reduces sum(dz), sum(dz*xhat); pass 2 writes dx.
 * dy (nullable) direct gradient; dpool (nullable) gradient of the pooled output, routed to the arg-max;
 * head_w/dlogit (nullable): dy[p,c] += dlogit[p]*head_w[c] (OutConv backward, model_parts.py:96). */
/* _apply also writes the parameter gradients from the reduced sums: d{gamma,beta,head_w} = out_beta * (old) +
 * out_scale * sum (out_scale removes the loss scale; out_beta = 1 accumulates over per-image launches); a non-finite
 * value sets *flag (nullable). */
int hpri_bn_relu_bwd_reduce(const hpri_view_t* x, const float* scale, const float* shift, const float* save_mean,
                            const float* save_invstd, const hpri_view_t* dy, const hpri_view_t* dpool,
                            const float* head_w, const float* dlogit, double* sums /*[C][3]*/, void* stream);
int hpri_bn_relu_bwd_apply(const hpri_view_t* x, const float* scale, const float* shift, const float* save_mean,
                           const float* save_invstd, const float* gamma, const hpri_view_t* dy,
                           const hpri_view_t* dpool, const float* head_w, const float* dlogit, double* sums,
                           long long count, const hpri_view_t* dx, float* dgamma, float* dbeta, float* dhead_w,
                           float out_scale, float out_beta, int* flag, void* stream);

/* ---- head + loss (model_parts.py:96; PLTrainer.py:86) ------------------------------------ */
/* logits[n,0,h,w] = b + sum_c relu(x*scale+shift)[c]*w[c]; fp32 NCHW output. */
int hpri_head_fwd(const hpri_view_t* x, const float* scale, const float* shift, const float* w, const float* b,
                  float* logits, void* stream);
/* mean BCE-with-logits and dlogit = grad_scale*(sigmoid(x)-t)/numel; counts = TP,FP,FN,TN at thr. */
int hpri_bce_fwd_bwd(const float* logits, const float* target, long long numel, float grad_scale, float thr,
                     double* loss_sum, float* dlogit, unsigned long long* counts, void* stream);
/* out[c] = beta * out[c] + scale * sum over pixels of the view (ConvT bias grad; Linear bias grad);
 * out[0] = scale * sum(x). */
int hpri_colsum(const hpri_view_t* x, float* out, float beta, float scale, void* stream);
int hpri_sum_f32(const float* x, long long numel, float* out, float scale, void* stream);
/* x *= scale in place; *flag |= 1 if any result is non-finite (unscaling of loss-scaled fp16 gradients). */
int hpri_scale_check(float* x, long long numel, float scale, int* flag, void* stream);

#ifdef __cplusplus
}
#endif
#endif
