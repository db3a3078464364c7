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
  int32_t tile_n;        /* 0 = auto, else force BN in {32, 64, 128, 256}                       */
} b200_conv_gemm_args;

int b200_conv_gemm(const b200_conv_gemm_args* args, void* stream);
/* Kp for a given geometry (host helper used by the weight packer). */
int64_t b200_conv_gemm_packed_k(int32_t ksize, int32_t c0, int32_t c1);
/* BN the auto heuristic picks (the GEGLU packer must interleave with the same BN). */
int32_t b200_conv_gemm_pick_tile_n(int64_t M, int32_t N, int32_t geglu);

#ifdef __cplusplus
}
#endif
#endif /* B200_DIFFUSION_H_ */
