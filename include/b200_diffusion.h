/*
 * b200_diffusion.h — C-ABI of libb200diff.so: the sm_100a kernels behind the diffusers
 * denoising hot path.
 *
 * The reference (huggingface/diffusers) has no FFI of its own: every numeric op on this path
 * is a Python call into PyTorch/ATen.  Each entry point below therefore cites the reference
 * call site whose ATen op sequence it replaces (paths relative to
 * /root/reference/src/diffusers/).
 *
 * Conventions
 *   - plain pointers and sizes only; every pointer is a DEVICE pointer unless stated otherwise;
 *   - `stream` is a CUstream/cudaStream_t handle passed as void* (0 = legacy default stream);
 *     all work is enqueued on it, nothing synchronises, nothing allocates device memory, so
 *     calls are CUDA-graph capturable;
 *   - 16-bit tensors are bf16 (dtype 0) or fp16 (dtype 1); statistics / master copies are fp32;
 *   - image activations are NHWC ("pixel-major": [batch, H, W, C], C contiguous); token
 *     activations [rows, C] are the same memory;
 *   - return 0 on success, negative B200_ERR_* otherwise; b200_last_error() returns a
 *     thread-local message for the last failure.
 */
#ifndef B200_DIFFUSION_H_
#define B200_DIFFUSION_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B200_OK 0
#define B200_ERR_INVALID (-1)
#define B200_ERR_CUDA (-2)
#define B200_ERR_UNSUPPORTED (-3)

#define B200_DTYPE_BF16 0
#define B200_DTYPE_FP16 1

#define B200_ACT_NONE 0
#define B200_ACT_SILU 1
#define B200_ACT_GELU_ERF 2
#define B200_ACT_GELU_TANH 3
#define B200_ACT_QUICK_GELU 4 /* x * sigmoid(1.702 x): CLIP text encoders */

/* Library / device -------------------------------------------------------------------------- */
int b200_version(void);
const char* b200_last_error(void);
/* Validates the device (compute capability 10.x), resolves cuTensorMapEncodeTiled and raises the
 * dynamic shared-memory limits of every kernel.  Idempotent. */
int b200_init(int device);
int b200_num_sms(void);

/* -------------------------------------------------------------------------------------------
 * b200_conv_gemm — the one tensor-core contraction of the path (tcgen05 + TMEM + TMA).
 *
 *   y[p, n] = epilogue( sum_{tap, src, c} x_src[pix(p, tap), c] * w[n, k(tap, src, c)] )
 *
 * It is an implicit GEMM over NHWC pixels: M = batch*Ho*Wo output pixels, N output channels,
 * K = kh*kw*(C0+C1).  nn.Linear is the kh=kw=1, H=1, W=rows case.  Replaces
 *   nn.Conv2d 3x3/1x1, stride 1|2, pad 1|0    models/resnet.py:268,284,310; downsampling.py:114;
 *                                              upsampling.py:132; unets/unet_2d_condition.py:272,492
 *   torch.cat([h, skip], dim=1) feeding a conv unets/unet_2d_blocks.py:2444,2561  (two K sources)
 *   nn.Linear (+bias)                          models/attention_processor.py:2742-2750,2780;
 *                                              transformers/transformer_2d.py:475,504
 *   GEGLU                                      models/activations.py:113-123   (geglu = 1)
 *   Linear -> GELU(tanh)                       models/attention.py:1682 (Flux FeedForward)
 *   gate * Linear(x) + residual                transformers/transformer_flux.py:394-409,470-493
 *   conv + temb[:, :, None, None]              models/resnet.py:343-349 (rowvec)
 *   (x + h) / output_scale_factor              models/resnet.py:375 (residual)
 *
 * Execution: persistent CTAs, normally pairs of CTAs (cta_group::2) over two consecutive 128-row tiles, each staging
 * half of the weight tile; 32-column output slabs are staged in shared memory and written with TMA stores, a residual
 * operand is TMA-loaded into the slab ahead of time (tile width / pairing are picked by a cost model fitted to
 * measured K-chunk rates).
 * Epilogue order (fp32):  v = acc + bias[n];  v = act(v);  v *= gate[g, n];  v += rowvec[g, n];
 *                         v += residual[p, n];  y = round16(v)      with g = p / rows_per_group.
 * geglu: w rows are packed per BN-tile as [BN/2 value rows | BN/2 gate rows] (see
 * diffusers_b200/packing.py); y has N/2 columns, y = (a + b_a) * gelu_erf(g + b_g).
 * ------------------------------------------------------------------------------------------- */
typedef struct {
  const void* x[2];      /* NHWC sources concatenated along channels; x[1] may be NULL          */
  int32_t c[2];          /* channels of each source (c[1] = 0 when unused); multiples of 8      */
  int32_t ldx[2];        /* pixel stride of each source in elements (>= c[i], multiple of 8)    */
  int32_t batch, H, W;   /* input spatial shape (Linear: batch = 1, H = 1, W = rows)            */
  int32_t ksize;         /* 1 or 3 (square kernel; pad = ksize / 2)                             */
  int32_t stride;        /* 1 or 2 (2 requires even H and W)                                    */
  const void* w;         /* packed weights [N][Kp], Kp = ksize^2 * (rup64(c0) + rup64(c1)),     */
                         /* k = ((tap * nsrc + src) * rup64(c_src)) + channel, zero padded      */
  int32_t N;             /* rows of w (output channels; 2x the y columns when geglu)            */
  const void* bias;      /* [N] or NULL                                                         */
  int32_t act;           /* B200_ACT_*                                                          */
  int32_t geglu;         /* 0 | 1                                                               */
  const void* gate;      /* [groups, ld_gate] or NULL                                           */
  const void* rowvec;    /* [groups, ld_rowvec] or NULL                                         */
  int32_t ld_gate, ld_rowvec;
  int32_t rows_per_group; /* output pixels per group (conv: Ho*Wo); ignored without gate/rowvec */
  const void* residual;  /* [M, ldr] or NULL                                                    */
  int32_t ldr;
  void* y;               /* [batch, Ho, Wo, ldy]                                                */
  int32_t ldy;
  int32_t dtype;         /* B200_DTYPE_*                                                        */
  int32_t tile_n;        /* 0 = auto, else force BN in {32, 64, 96, 128, 160, 192, 256}         */
  int32_t out_fp32;      /* 1: y is float32 [.., ldy] (attention scores of the head_dim-512 path) */
  int32_t cluster_m;     /* 0 = auto, 1 = one CTA per tile, 2 = CTA pairs (cta_group::2) (tuning / tests)   */
  void* debug_timestamps; /* NULL, or int64 [grid][32] device buffer receiving per-CTA clock64 marks (tuning) */
  const void* prefetch;  /* NULL, or device memory (16-byte aligned) to pull into L2 while this launch runs:      */
  int64_t prefetch_bytes; /* the packed weights of the NEXT launch, which would otherwise start DRAM-latency-bound */
  /* ---- nn.LayerNorm folded into the Linear that consumes it (models/attention.py:986,1030,1056 -> attn1.to_q/k/v,
   * attn2.to_q, ff.net.0.proj).  With W' = W*gamma minus its row means (so that every row of W' sums to ~0),
   *     LN(x) W^T + b  =  rstd[m] * (x W'^T)[m,n] + (b + W beta)[n] :
   * the GEMM reads the RAW residual stream x, w holds W' (rounded to 16 bit once, at load time), bias holds b + W beta and the
   * epilogue scales each row by its rstd.  The row statistics come from the epilogue of the GEMM that PRODUCED x.  No
   * LayerNorm kernel, no normalised copy of x in HBM.  (The rounding of W' leaves row sums of ~sqrt(K) 2^-9 |w| instead of 0:
   * a relative output error of ~1e-3 * |mean(x)| / std(x), below the 16-bit rounding of LN(x) in the reference.)
   *   producer: row_stats_out != NULL -> for every output row m the epilogue stores b200_conv_gemm_row_stats_parts()
   *             pairs (sum, sum of squares) of disjoint subsets of the row's outputs AS ROUNDED to 16 bit, fp32
   *             [M][parts][2].  Needs the TMA-store epilogue (aligned y, N %% 32 == 0), no act / gate / rowvec / geglu / out_fp32.
   *   consumer: ln_stats != NULL -> those pairs (ln_parts of them per row, even); variance = E[x^2] - E[x]^2 over the K = c[0]
   *             values of the row.  ksize 1, one source, act none, no gate / rowvec / residual; works with geglu. */
  float* row_stats_out;
  const float* ln_stats;
  int32_t ln_parts;
  float ln_eps;
  /* ---- nearest-2x upsample folded into the 3x3 convolution that follows it (models/upsampling.py:175-190: F.interpolate then
   * self.conv).  Output pixel (2i + ph, 2j + pw) of conv3x3(nearest2x(x)) only ever sees the 2 x 2 input pixels
   * x[i + ph - 1 .. i + ph, j + pw - 1 .. j + pw]; the 3 x 3 taps that land on the same input pixel are summed into one weight
   * at load time (packing.pack_upsample_conv).  One launch per parity class with ksize = 2 and up2x_parity = 1 + 2 ph + pw:
   * H, W are the INPUT size, w is that class's packed [N, 4 * rup64(c0)] weight, y is the full [batch, 2H, 2W, ldy] output
   * of which the launch writes its pixels.  The upsampled tensor never exists and the convolution does 4/9 of the FLOPs.
   * bias / act only; needs the TMA-store epilogue. */
  int32_t up2x_parity;
  /* ---- stride-2 3x3 convolution with the padding on the bottom / right only: Downsample2D(padding=0) of the VAE encoder
   * (models/downsampling.py:141-143: F.pad(x, (0, 1, 0, 1)) then Conv2d(stride=2, padding=0)); output (ho, wo) reads input rows
   * 2 ho .. 2 ho + 2.  0 = the symmetric padding of 1 every other 3x3 convolution of the path uses. */
  int32_t pad_after_only;
  /* ---- per-head RMSNorm + rotary embedding of q and k folded into the fused QKV projection (FluxAttnProcessor:
   * transformers/transformer_flux.py:84-136 = to_q/to_k/to_v -> norm_q / norm_k (torch.nn.RMSNorm(head_dim, eps 1e-6)) ->
   * apply_rotary_emb (models/embeddings.py:1187-1231) -> attention).  With qk_cols > 0 the output columns [0, qk_cols) are
   * q heads then k heads of qk_head_dim (64 | 128) columns; each (row, head) is RMS-normalised over its columns, multiplied by
   * qk_norm_w ([2][qk_head_dim] 16-bit: the q weights, then the k weights) and rotated by the angle of position
   * rope_row0 + row: rope_cos / rope_sin are fp32 [qk_head_dim / 2][rope_ld] tables (entry [i][pos] = cos / sin of pair i at
   * position pos - position-minor so that the 32 rows of a warp read consecutive words).  Columns >= qk_cols (v) are plain.
   * Rounding points are those of the reference's eager ops.  Linear (H == 1) launches with bias / folded LayerNorm only. */
  int32_t qk_cols;
  int32_t qk_head_dim;
  const void* qk_norm_w;
  const float* rope_cos;
  const float* rope_sin;
  int32_t rope_ld;
  int32_t rope_row0;
  float qk_eps;
  /* ---- context parallelism (see the peer-memory section): with y_block_cols > 0 the N output columns of a plain linear are
   * N / y_block_cols (<= 8) blocks; block j is stored to y_peers[j] ([rows, ldy] buffers, columns [0, y_block_cols)) instead of
   * `y` - the joint QKV buffers of the ranks that own those heads, peer mappings over NVLink for the other ranks.  This is the
   * first all-to-all of Ulysses attention (models/attention_dispatch.py:2528-2540) done by the GEMM's own TMA stores.  With
   * qk_cols > 0 the [q | k | v] column pattern repeats in every block (qk_cols <= y_block_cols). */
  void* y_peers[8];
  int32_t y_block_cols;
} b200_conv_gemm_args;

int b200_conv_gemm(const b200_conv_gemm_args* args, void* stream);
/* (sum, sum of squares) pairs per output row that a launch with these arguments writes to row_stats_out (host helper). */
int32_t b200_conv_gemm_row_stats_parts(const b200_conv_gemm_args* args);
/* Kp for a given geometry (host helper used by the weight packer). */
int64_t b200_conv_gemm_packed_k(int32_t ksize, int32_t c0, int32_t c1);
/* BN the auto heuristic picks (the GEGLU packer must interleave with the same BN). */
int32_t b200_conv_gemm_pick_tile_n(int64_t M, int32_t N, int32_t geglu);

/* -------------------------------------------------------------------------------------------
 * b200_group_norm — nn.GroupNorm (+ fused SiLU) over NHWC, optionally over the channel concat of
 * two tensors (the skip-connection torch.cat is never materialised).  One launch when the slab of a (sample, group block)
 * fits the shared memory of a thread-block cluster (all UNet shapes but one): global -> shared once, exact two-pass
 * statistics reduced across the cluster through distributed shared memory, normalise/affine/activation from shared memory.
 * Otherwise two launches: statistics (shifted sums + Chan merge, deterministic) and normalise/affine/activation.
 * Replaces  models/resnet.py:326-327,349-362 (norm1/norm2 + nonlinearity),
 *           transformers/transformer_2d.py:466 (Transformer2DModel.norm, eps 1e-6),
 *           unets/unet_2d_condition.py:1227-1229 (conv_norm_out + conv_act),
 *           autoencoders/vae.py:303-305, models/attention_processor.py:2740 (VAE attention group_norm).
 * workspace: b200_group_norm_workspace_bytes() bytes, ZERO-INITIALISED once by the caller (the
 * kernels leave its counters zeroed); fp32 statistics.
 * ------------------------------------------------------------------------------------------- */
typedef struct {
  const void* x[2];   /* NHWC sources [batch, hw, c[i]] concatenated along channels            */
  int32_t c[2];       /* multiples of 8                                                         */
  int32_t ldx[2];
  int32_t batch, hw, groups;
  float eps;
  const void* gamma;  /* [C] or NULL, 16-byte aligned                                           */
  const void* beta;   /* [C] or NULL                                                            */
  int32_t act;        /* B200_ACT_NONE | B200_ACT_SILU                                          */
  void* y;            /* [batch, hw, ldy]                                                       */
  int32_t ldy;
  void* workspace;
  int64_t workspace_bytes;
  int32_t dtype;
} b200_group_norm_args;

int b200_group_norm(const b200_group_norm_args* args, void* stream);
/* Kernels the call above launches for a shape: 1 = single-pass cluster kernel (one read + one write of the activation; taken
 * whenever the slab of one (sample, block of groups) fits the shared memory of <= 8 clustered CTAs), 2 = statistics + apply. */
int32_t b200_group_norm_launches(int32_t hw, int32_t C, int32_t groups, int32_t c0, int32_t two_sources);
int64_t b200_group_norm_workspace_bytes(int32_t batch, int32_t hw, int32_t groups);

/* -------------------------------------------------------------------------------------------
 * b200_layer_norm — nn.LayerNorm over token rows (one warp per row, two-pass statistics in
 * registers), optional affine, optional AdaLN modulation  y = LN(x) * (1 + scale[g]) + shift[g]
 * with g = row / rows_per_group.
 * Replaces  models/attention.py:986,1030,1056 (BasicTransformerBlock norm1/2/3),
 *           models/normalization.py:167-170,199-202,348-351 (AdaLayerNormZero/-Single/-Continuous).
 * ------------------------------------------------------------------------------------------- */
typedef struct {
  const void* x;
  int32_t ldx, rows, cols; /* cols: multiple of 8, <= 4096 */
  float eps;
  const void* gamma;  /* [cols] or NULL */
  const void* beta;   /* [cols] or NULL */
  const void* scale;  /* [groups, ld_mod] or NULL */
  const void* shift;  /* [groups, ld_mod] or NULL */
  int32_t ld_mod, rows_per_group;
  void* y;
  int32_t ldy;
  int32_t dtype;
  int32_t rms;        /* 1 = RMSNorm: no mean subtraction, y = x * rsqrt(mean(x^2) + eps) * gamma  (T5LayerNorm,
                       * transformers models/t5/modeling_t5.py T5LayerNorm.forward); 0 = LayerNorm                */
} b200_layer_norm_args;

int b200_layer_norm(const b200_layer_norm_args* args, void* stream);

/* -------------------------------------------------------------------------------------------
 * b200_small_linear — nn.Linear for M <= 8 rows (time / text-time / AdaLN-modulation MLPs):
 *   y[m, n] = act_out(sum_k act_in(x[m, k]) * w[n, k] + bias[n]) (+ addend[m, n])
 * One warp per output column, weights streamed once with 16-byte loads (weight-bandwidth bound).
 * w is the plain nn.Linear weight [N, K] (K multiple of 8).
 * Replaces  models/embeddings.py:1262 (TimestepEmbedding), unets/unet_2d_condition.py:917-922
 *           (add_embedding), models/resnet.py:343-347 (time_emb_proj(nonlinearity(temb))),
 *           models/normalization.py:167,199,348 (AdaLN linear(silu(emb))).
 * ------------------------------------------------------------------------------------------- */
typedef struct {
  const void* x;
  int32_t ldx, M, K;
  const void* w;
  int32_t N;
  const void* bias;
  int32_t act_in, act_out;
  const void* addend; /* [M, ld_add] or NULL; added after rounding y to 16 bit (emb + aug_emb) */
  int32_t ld_add;
  void* y;
  int32_t ldy;
  int32_t dtype;
} b200_small_linear_args;

int b200_small_linear(const b200_small_linear_args* args, void* stream);

/* Boundary layout changes (model inputs/outputs are NCHW in the reference API) and nearest-2x
 * upsample (models/upsampling.py:175 F.interpolate(scale_factor=2.0, mode="nearest")). */
int b200_nchw_to_nhwc(const void* src, void* dst, int32_t batch, int32_t C, int32_t HW, int32_t ld_dst, int32_t dtype,
                      void* stream);
int b200_nhwc_to_nchw(const void* src, int32_t ld_src, void* dst, int32_t batch, int32_t C, int32_t HW, int32_t dtype,
                      void* stream);
int b200_upsample_nearest2x(const void* x, int32_t ldx, void* y, int32_t ldy, int32_t batch, int32_t H, int32_t W,
                            int32_t C, int32_t dtype, void* stream);

/* get_timestep_embedding (models/embeddings.py:27): t fp32 [n] -> out [n, dim] 16-bit. */
int b200_timestep_embedding(const float* t, int32_t n, void* out, int32_t ld_out, int32_t dim, int32_t flip_sin_to_cos,
                            float downscale_freq_shift, float scale, float max_period, int32_t dtype, void* stream);

/* Scheduler steps; sigma values are the scheduler's fp32 table entries passed by value.
 *   b200_euler_step       EulerDiscreteScheduler.step, epsilon prediction, s_churn = 0
 *                         (schedulers/scheduling_euler_discrete.py:751-789)
 *   b200_scale            EulerDiscreteScheduler.scale_model_input (:345): y = x / divisor
 *   b200_cfg_euler_step   pipeline_stable_diffusion_xl.py:1202-1203,1224-1225,1233 in one launch:
 *                         CFG combine, Euler update of the NCHW latents (in place), and the next
 *                         step's scaled, CFG-duplicated NHWC model input
 *   b200_flow_match_step  FlowMatchEulerDiscreteScheduler.step (scheduling_flow_match_euler_discrete.py:484-517) */
/*   b200_linear_step      prev = a*sample + b*m0 + c*m1 + s*noise (m0 / m1 / noise may be NULL), fp32 math, one rounding: the
 *                         update of DDIMScheduler.step with eta = 0 (schedulers/scheduling_ddim.py:384-520),
 *                         EulerAncestralDiscreteScheduler.step (scheduling_euler_ancestral_discrete.py:330-440) and
 *                         DPMSolverMultistepScheduler dpmsolver++ first / second order updates
 *                         (scheduling_dpmsolver_multistep.py:620-800); the data prediction x0 = (x - sigma_t eps) / alpha_t is
 *                         the same call */
int b200_linear_step(const void* sample, const void* m0, const void* m1, const void* noise, void* prev_sample, int64_t n, float a,
                     float b, float c, float s, int32_t dtype, void* stream);
int b200_euler_step(const void* model_output, const void* sample, void* prev_sample, int64_t n, float sigma,
                    float sigma_next, int32_t dtype, void* stream);
int b200_scale(const void* x, void* y, int64_t n, float divisor, int32_t dtype, void* stream);
int b200_cfg_euler_step(const void* eps_nhwc, int32_t ld_eps, void* latents_nchw, void* next_in_nhwc, int32_t ld_in,
                        int32_t batch, int32_t C, int32_t HW, float guidance_scale, int32_t do_cfg, float sigma,
                        float sigma_next, int32_t dtype, void* stream);
int b200_flow_match_step(const void* model_output, const void* sample, void* prev_sample, int64_t n, float sigma,
                         float sigma_next, int32_t dtype, void* stream);

/* -------------------------------------------------------------------------------------------
 * b200_attention — softmax(Q K^T * scale) V, no mask, no dropout (the only form the path uses),
 * as one FlashAttention-style tcgen05/TMEM kernel fed by TMA.  q/k/v/o are [batch, seq, heads,
 * head_dim] views given by row (token) and batch strides in elements; head h starts at column
 * h*head_dim, so fused-QKV GEMM outputs are consumed in place.
 * Replaces  F.scaled_dot_product_attention  models/attention_processor.py:2767 (AttnProcessor2_0,
 *           self S_k = S_q and cross S_k = 77) and models/attention_dispatch.py:3678-3717
 *           (_native_attention, the NATIVE backend FluxAttnProcessor dispatches to).
 * head_dim 64 (SDXL) or 128 (Flux).  scale <= 0 selects 1/sqrt(head_dim).
 * ------------------------------------------------------------------------------------------- */
typedef struct {
  const void* q;
  const void* k;
  const void* v;
  void* o;
  int32_t batch, heads, sq, sk, head_dim;
  int64_t q_row_stride, q_batch_stride;
  int64_t k_row_stride, k_batch_stride;
  int64_t v_row_stride, v_batch_stride;
  int64_t o_row_stride, o_batch_stride;
  float scale;
  int32_t dtype;
  int32_t nq_override; /* 0 = auto; 1 | 2 = query tiles (of 128 rows) per CTA */
  /* Optional scratch of >= b200_attention_workspace_bytes(...) bytes, 16-byte aligned, ZERO before its first use and private to
   * one stream.  With it, head_dim-64 launches whose query tiles do not fill the last wave of CTAs split the tiles of that wave
   * along the keys and merge the parts in a fixed order (bitwise reproducible; the kernel leaves the ticket words zero again).
   * NULL / too small: every tile is computed by one CTA.  Same results either way up to fp32 summation order of the merge. */
  void* workspace;
  int64_t workspace_bytes;
  /* Context parallelism (see the peer-memory section): with o_seg_rows > 0 (batch must be 1) output row r of head h goes to
   * o_seg[r / o_seg_rows] + (r % o_seg_rows) * o_row_stride + h * head_dim instead of `o` - the segments are the output
   * buffers of the ranks that own those rows (peer mappings; the local one for this rank's own rows).  head_dim 64 / 128
   * through the default kernel only. */
  int32_t o_seg_rows;
  void* o_seg[8];
} b200_attention_args;

int b200_attention(const b200_attention_args* args, void* stream);
/* Scratch the shape can use (0 = none; depends on the device's SM count: call after b200_init on the current device). */
int64_t b200_attention_workspace_bytes(int32_t batch, int32_t heads, int32_t sq, int32_t sk, int32_t head_dim);

/* Unfused attention helpers for head_dim 512 (AutoencoderKL mid-block attention, one head:
 * models/attention_processor.py:2725-2789 via unets/unet_2d_blocks.py:684-698).  TMEM cannot hold a
 * 128 x 512 fp32 output tile next to the scores, so that single layer runs as
 * b200_conv_gemm(out_fp32) -> b200_softmax_rows -> b200_conv_gemm against b200_transpose_16(V).
 *   b200_softmax_rows : p[r, :] = softmax(s[r, :] * scale), s fp32 [rows, ld_s], p 16-bit [rows, ld_p]
 *   b200_transpose_16 : dst[c, r] = src[r, c] for 16-bit elements */
int b200_softmax_rows(const float* s, int64_t ld_s, void* p, int64_t ld_p, int32_t rows, int32_t cols, float scale,
                      int32_t dtype, void* stream);
int b200_transpose_16(const void* src, int64_t ld_src, void* dst, int64_t ld_dst, int32_t rows, int32_t cols,
                      void* stream);

/* -------------------------------------------------------------------------------------------
 * b200_qk_norm_rope — per-head RMSNorm (with weight) of q and k followed by the rotary embedding, in
 * place on a fused QKV buffer [rows, ld] (q at column 0, k at column k_off, head h at h*head_dim).
 * Rows [0, txt_rows) of every `seq`-row sequence use the text-stream norm weights (joint attention puts
 * the text tokens first).  cos/sin are the fp32 [seq, head_dim] tables of FluxPosEmbed (repeat-interleaved).
 * Replaces  transformers/transformer_flux.py:102-119 (norm_q/norm_k/norm_added_q/norm_added_k =
 *           torch.nn.RMSNorm(head_dim, eps 1e-6); apply_rotary_emb models/embeddings.py:1187-1231).
 * ------------------------------------------------------------------------------------------- */
typedef struct {
  void* qkv;
  int64_t ld;
  int32_t rows, heads, head_dim, k_off;
  int32_t txt_rows, seq;
  const void* wq;      /* [head_dim] 16-bit, or NULL (no weight) */
  const void* wk;
  const void* wq_txt;  /* NULL = same as wq */
  const void* wk_txt;
  const float* cos_table; /* [seq, head_dim] or NULL (no rope) */
  const float* sin_table;
  float eps;
  int32_t dtype;
  int32_t txt_period; /* 0 = seq; else the text rows are [0, txt_rows) of every txt_period rows (rank-major joint order of context parallelism) */
} b200_qk_norm_rope_args;

int b200_qk_norm_rope(const b200_qk_norm_rope_args* args, void* stream);

/* DDPMScheduler.step (schedulers/scheduling_ddpm.py:513-565; epsilon prediction, fixed_small variance, optional
 * clip_sample): x0 = (x - sqrt(1-abar_t) eps)/sqrt(abar_t); prev = c0*x0 + c1*x (+ sigma*noise when noise != NULL).
 * The scalar coefficients come from the scheduler's fp32 tables on the host. */
int b200_ddpm_step(const void* model_output, const void* sample, const void* noise, void* prev_sample, int64_t n,
                   float sqrt_beta_prod_t, float sqrt_alpha_prod_t, float pred_original_coeff, float current_sample_coeff,
                   float sigma, int32_t clip_sample, float clip_range, int32_t dtype, void* stream);

/* -------------------------------------------------------------------------------------------
 * b200_text_attention — attention of the text encoders either side of the denoiser: head_dim 64, at most 512
 * keys, optional causal mask and additive bias, fp32 arithmetic.  q/k/v/o are [batch, seq, heads, 64] views given by row and
 * batch strides in elements, like b200_attention.
 * Replaces  the attention of transformers' CLIPTextModel / CLIPTextModelWithProjection (models/clip/modeling_clip.py
 *           CLIPAttention: scale head_dim^-0.5, causal mask) and T5EncoderModel (models/t5/modeling_t5.py T5Attention: no
 *           scaling, position_bias = relative_attention_bias[bucket(j - i)] added to the scores) as called from
 *           pipelines/stable_diffusion_xl/pipeline_stable_diffusion_xl.py:283 (encode_prompt) and
 *           pipelines/flux/pipeline_flux.py:217-387 (_get_t5_prompt_embeds / _get_clip_prompt_embeds).
 * out[i] = softmax_j(scale * q_i . k_j + bias[h, i, j], j <= i when causal) v_j
 * ------------------------------------------------------------------------------------------- */
typedef struct {
  const void* q;
  const void* k;
  const void* v;
  void* o;
  int32_t batch, heads, sq, sk; /* head_dim is 64; sk <= 512 */
  int64_t q_row_stride, q_batch_stride;
  int64_t k_row_stride, k_batch_stride;
  int64_t v_row_stride, v_batch_stride;
  int64_t o_row_stride, o_batch_stride;
  float scale;
  int32_t causal;
  const float* bias; /* fp32 [heads, sq, sk] or NULL */
  int32_t dtype;
} b200_text_attention_args;

int b200_text_attention(const b200_text_attention_args* args, void* stream);

/* -------------------------------------------------------------------------------------------
 * Peer memory over NVLink / NVSwitch — context parallelism (Ulysses) for one image on several GPUs of a node.
 * Replaces the all-to-all collectives of  models/attention_dispatch.py:2504-2580 (TemplatedUlyssesAttention:
 * _all_to_all_single on q/k/v before and on the output after the attention, funcol over NCCL) driven by
 * hooks/context_parallel.py:129,220 and transformers/transformer_flux.py:573-581 (_cp_plan).
 * No collective library sits on the data path: b200_conv_gemm stores the QKV tiles of the heads a peer owns straight
 * into that peer's buffer (`y` = a mapping returned by b200_peer_open), b200_attention stores every output row into
 * the peer that owns the row (o_seg below), and b200_peer_barrier separates the phases.
 *   b200_peer_alloc   : cudaMalloc + zero + export; `handle64` receives the 64-byte cudaIpcMemHandle_t (HOST pointer)
 *   b200_peer_open    : maps another process's buffer (same node) into this process; peer access is enabled lazily
 *   b200_peer_close / b200_peer_free : undo the two above
 *   b200_peer_barrier : one small kernel on `stream`.  flags is a HOST array of nranks device pointers: flags[r] =
 *                       this process's mapping of rank r's flag words (>= B200_MAX_PEERS uint32, zero at start);
 *                       epoch = one local uint32 in device memory (zero at start; advanced by the kernel, so the
 *                       call is CUDA-graph capturable).  Every rank must issue the same sequence of barriers.
 *                       A rank that waits longer than 20 s traps (a CUDA error on that rank, never a hung GPU).
 * ------------------------------------------------------------------------------------------- */
#define B200_MAX_PEERS 8
int b200_peer_alloc(int64_t bytes, void** ptr, unsigned char* handle64);
int b200_peer_open(const unsigned char* handle64, void** ptr);
int b200_peer_close(void* ptr);
int b200_peer_free(void* ptr);
int b200_peer_barrier(void* const* flags, void* epoch, int32_t rank, int32_t nranks, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* B200_DIFFUSION_H_ */
