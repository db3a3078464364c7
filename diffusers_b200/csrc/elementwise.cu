// Small HBM-/latency-bound kernels of the denoising loop: layout changes at the model boundary,
// nearest-2x upsample, sinusoidal timestep embedding, skinny (M <= 8) linear layers, and the
// scheduler steps (Euler / CFG+Euler fused / FlowMatch).
#include <string.h>

#include "common.cuh"
#include "host_common.h"

namespace b200 {

// ------------------------------------------------------------------------------------------------
// NCHW <-> NHWC (boundary tensors only: latents in, noise prediction / image out)
// ------------------------------------------------------------------------------------------------
template <typename T>
__global__ void nchw_to_nhwc_kernel(const T* __restrict__ src, T* __restrict__ dst, int C, int HW, int ld_dst,
                                    long long total) {
  pdl_trigger();
  pdl_wait();
  long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int c = static_cast<int>(i % ld_dst);
  const long long np = i / ld_dst;
  const int pix = static_cast<int>(np % HW);
  const long long n = np / HW;
  T v = T(0.f);
  if (c < C) v = src[(n * C + c) * HW + pix];
  dst[i] = v;
}

template <typename T>
__global__ void nhwc_to_nchw_kernel(const T* __restrict__ src, int ld_src, T* __restrict__ dst, int C, int HW,
                                    long long total) {
  pdl_trigger();
  pdl_wait();
  long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;  // index into dst (n, c, pix)
  if (i >= total) return;
  const int pix = static_cast<int>(i % HW);
  const long long nc = i / HW;
  const int c = static_cast<int>(nc % C);
  const long long n = nc / C;
  dst[i] = src[(n * HW + pix) * ld_src + c];
}

// ------------------------------------------------------------------------------------------------
// nearest-neighbour 2x upsample, NHWC, 8 channels (16 B) per thread
// ------------------------------------------------------------------------------------------------
__global__ void upsample2x_kernel(const uint4* __restrict__ x, int ldx8, uint4* __restrict__ y, int ldy8, int H, int W,
                                  int V, long long total) {
  pdl_trigger();
  pdl_wait();
  long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int v = static_cast<int>(i % V);
  long long r = i / V;
  const int ow = static_cast<int>(r % (2 * W));
  r /= (2 * W);
  const int oh = static_cast<int>(r % (2 * H));
  const long long n = r / (2 * H);
  const long long ip = (n * H + (oh >> 1)) * W + (ow >> 1);
  const long long op = (n * 2 * H + oh) * (2 * W) + ow;
  y[op * ldy8 + v] = x[ip * ldx8 + v];
}

// ------------------------------------------------------------------------------------------------
// sinusoidal timestep embedding (reference models/embeddings.py:27 get_timestep_embedding)
// ------------------------------------------------------------------------------------------------
template <bool FP16>
__global__ void timestep_embedding_kernel(const float* __restrict__ t, int n, void* out_, int ld_out, int dim,
                                          int flip_sin_to_cos, float downscale_freq_shift, float scale,
                                          float max_period) {
  pdl_trigger();
  pdl_wait();
  using H = Half16<FP16>;
  typename H::T* out = static_cast<typename H::T*>(out_);
  const int half = dim / 2;
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n * half) return;
  const int row = idx / half, i = idx - row * half;
  float exponent = __fmul_rn(-logf(max_period), static_cast<float>(i));
  exponent = __fdiv_rn(exponent, static_cast<float>(half) - downscale_freq_shift);
  const float freq = expf(exponent);
  float arg = __fmul_rn(t[row], freq);
  arg = __fmul_rn(scale, arg);
  const float s = sinf(arg), c = cosf(arg);
  typename H::T* o = out + static_cast<size_t>(row) * ld_out;
  if (flip_sin_to_cos) {
    o[i] = H::from_float(c);
    o[half + i] = H::from_float(s);
  } else {
    o[i] = H::from_float(s);
    o[half + i] = H::from_float(c);
  }
  if ((dim & 1) && i == 0) o[dim - 1] = H::from_float(0.f);
}

// ------------------------------------------------------------------------------------------------
// skinny linear: y[m, n] = act_out(sum_k act_in(x[m, k]) * w[n, k] + b[n]) (+ addend[m, n]), M <= 8.
// One warp per output column; weight rows streamed with 16-byte loads (weight-bandwidth bound).
// ------------------------------------------------------------------------------------------------
struct SmallLinearParams {
  const void* x;
  int ldx, M, K;
  const void* w;
  int N;
  const void* bias;
  int act_in, act_out;
  const void* addend;
  int ld_add;
  void* y;
  int ldy;
  int cpw;  // output columns per warp
};

template <bool FP16>
__global__ void __launch_bounds__(256) small_linear_kernel(const SmallLinearParams p) {
  pdl_trigger();
  pdl_wait();
  using H = Half16<FP16>;
  extern __shared__ float s_x[];  // [M][K] fp32 (rounded through the 16-bit type after act_in, like the reference)
  const typename H::T* x = static_cast<const typename H::T*>(p.x);
  for (int i = threadIdx.x; i < p.M * p.K; i += blockDim.x) {
    const int m = i / p.K, k = i - m * p.K;
    float v = H::to_float(x[static_cast<size_t>(m) * p.ldx + k]);
    if (p.act_in != ACT_NONE) v = H::to_float(H::from_float(apply_act(v, p.act_in)));
    s_x[i] = v;
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // cpw consecutive output columns per warp: staging x (and its activation: 3072 SiLUs for the Flux modulations) is paid per
  // CTA, and with one column per warp it cost more than the CTA's 8 dot products (ncu: XU pipe 31 %, 3.1 TB/s of weights)
  for (int ci = 0; ci < p.cpw; ++ci) {
  const int n = (blockIdx.x * (blockDim.x >> 5) + warp) * p.cpw + ci;
  if (n >= p.N) return;
  const typename H::T* wrow = static_cast<const typename H::T*>(p.w) + static_cast<size_t>(n) * p.K;
  float acc[8];
#pragma unroll
  for (int m = 0; m < 8; ++m) acc[m] = 0.f;
  const int nvec = p.K >> 3;
  for (int v = lane; v < nvec; v += 32) {
    uint4 u = *reinterpret_cast<const uint4*>(wrow + v * 8);
    float2 a = H::unpack(u.x), b = H::unpack(u.y), c = H::unpack(u.z), d = H::unpack(u.w);
    const float wv[8] = {a.x, a.y, b.x, b.y, c.x, c.y, d.x, d.y};
#pragma unroll
    for (int m = 0; m < 8; ++m) {
      if (m < p.M) {
        const float* xr = s_x + m * p.K + v * 8;
        float4 x0 = *reinterpret_cast<const float4*>(xr);
        float4 x1 = *reinterpret_cast<const float4*>(xr + 4);
        acc[m] += wv[0] * x0.x + wv[1] * x0.y + wv[2] * x0.z + wv[3] * x0.w + wv[4] * x1.x + wv[5] * x1.y +
                  wv[6] * x1.z + wv[7] * x1.w;
      }
    }
  }
#pragma unroll
  for (int m = 0; m < 8; ++m) acc[m] = warp_sum(acc[m]);
  if (lane == 0) {
    const float b = p.bias ? H::to_float(static_cast<const typename H::T*>(p.bias)[n]) : 0.f;
    typename H::T* y = static_cast<typename H::T*>(p.y);
    const typename H::T* add = static_cast<const typename H::T*>(p.addend);
#pragma unroll
    for (int m = 0; m < 8; ++m) {
      if (m < p.M) {
        float v = apply_act(acc[m] + b, p.act_out);
        if (add) v = H::to_float(H::from_float(v)) + H::to_float(add[static_cast<size_t>(m) * p.ld_add + n]);
        y[static_cast<size_t>(m) * p.ldy + n] = H::from_float(v);
      }
    }
  }
  }
}

// ------------------------------------------------------------------------------------------------
// scheduler steps.  The op order and the 16-bit rounding points mirror the reference eager code so a
// sampling loop stays on the reference trajectory (schedulers/scheduling_euler_discrete.py:751-789,
// scheduling_flow_match_euler_discrete.py:484-517, pipeline_stable_diffusion_xl.py:1202-1225).
// ------------------------------------------------------------------------------------------------
template <bool FP16>
__device__ __forceinline__ float r16(float v) {
  using H = Half16<FP16>;
  return H::to_float(H::from_float(v));
}

// prev = (x + ((x - (x - r16(sigma*eps))) / sigma) * (sigma_next - sigma)), x = float(sample)
template <bool FP16>
__device__ __forceinline__ float euler_update(float x, float eps, float sigma, float dt) {
  const float m = r16<FP16>(__fmul_rn(sigma, eps));
  const float x0 = __fsub_rn(x, m);
  const float d = __fdiv_rn(__fsub_rn(x, x0), sigma);
  return __fadd_rn(x, __fmul_rn(d, dt));
}

template <bool FP16>
__global__ void euler_step_kernel(const void* eps_, const void* sample_, void* prev_, long long n, float sigma,
                                  float dt) {
  pdl_trigger();
  pdl_wait();
  using H = Half16<FP16>;
  const typename H::T* eps = static_cast<const typename H::T*>(eps_);
  const typename H::T* sample = static_cast<const typename H::T*>(sample_);
  typename H::T* prev = static_cast<typename H::T*>(prev_);
  long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n) return;
  prev[i] = H::from_float(euler_update<FP16>(H::to_float(sample[i]), H::to_float(eps[i]), sigma, dt));
}

// One launch per SDXL step: classifier-free guidance + Euler update + the next step's model input.
//   eps_nhwc [2*B, HW, ld_eps]  model output, uncond rows first (reference concat order :1149-1151)
//   latents  [B, C, HW] (NCHW)  updated in place (16-bit, like the reference's `latents`)
//   next_in  [2*B, HW, ld_in]   scale_model_input(cat([latents]*2), t_next) in NHWC, zero padded channels
template <bool FP16>
__global__ void cfg_euler_step_kernel(const void* eps_, int ld_eps, void* latents_, void* next_in_, int ld_in, int B,
                                      int C, int HW, float guidance, float sigma, float dt, float next_div,
                                      int do_cfg) {
  pdl_trigger();
  pdl_wait();
  using H = Half16<FP16>;
  const typename H::T* eps = static_cast<const typename H::T*>(eps_);
  typename H::T* lat = static_cast<typename H::T*>(latents_);
  typename H::T* nin = static_cast<typename H::T*>(next_in_);
  long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;  // (b, pix, c_in) over ld_in
  const long long total = static_cast<long long>(B) * HW * ld_in;
  if (i >= total) return;
  const int c = static_cast<int>(i % ld_in);
  const long long bp = i / ld_in;
  const int pix = static_cast<int>(bp % HW);
  const int b = static_cast<int>(bp / HW);
  float out = 0.f;
  if (c < C) {
    float e;
    if (do_cfg) {
      const float eu = H::to_float(eps[(static_cast<long long>(b) * HW + pix) * ld_eps + c]);
      const float ec = H::to_float(eps[(static_cast<long long>(B + b) * HW + pix) * ld_eps + c]);
      // noise_pred_uncond + guidance_scale * (noise_pred_text - noise_pred_uncond), each op rounded to 16 bit
      const float diff = r16<FP16>(__fsub_rn(ec, eu));
      const float sc = r16<FP16>(__fmul_rn(guidance, diff));
      e = r16<FP16>(__fadd_rn(eu, sc));
    } else {
      e = H::to_float(eps[(static_cast<long long>(b) * HW + pix) * ld_eps + c]);
    }
    const long long li = (static_cast<long long>(b) * C + c) * HW + pix;
    const float x = H::to_float(lat[li]);
    const float nx = r16<FP16>(euler_update<FP16>(x, e, sigma, dt));
    lat[li] = H::from_float(nx);
    out = __fdiv_rn(nx, next_div);  // latent_model_input / ((sigma_next**2 + 1) ** 0.5), rounded on store
  }
  const typename H::T o = H::from_float(out);
  nin[i] = o;
  if (do_cfg) nin[i + total] = o;
}

// prev = r16(float(x) + r16(r16(dt) * v)).  dt = sigmas[i+1] - sigmas[i] is a 0-dim DEVICE tensor in the
// reference, so eager CUDA casts it to the 16-bit common dtype before the multiply.
template <bool FP16>
__global__ void flow_match_step_kernel(const void* v_, const void* sample_, void* prev_, long long n, float dt) {
  pdl_trigger();
  pdl_wait();
  using H = Half16<FP16>;
  const typename H::T* v = static_cast<const typename H::T*>(v_);
  const typename H::T* sample = static_cast<const typename H::T*>(sample_);
  typename H::T* prev = static_cast<typename H::T*>(prev_);
  long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float m = r16<FP16>(__fmul_rn(r16<FP16>(dt), H::to_float(v[i])));
  prev[i] = H::from_float(__fadd_rn(H::to_float(sample[i]), m));
}

// Generic first / second order stepper update: prev = a * sample + b * m0 + c * m1 + s * noise, fp32 math on the 16-bit
// operands, ONE rounding.  DDIM (eta = 0), Euler-ancestral and DPM-Solver++ (2M) steps are all of this form (their data
// predictions x0 = (x - sigma_t eps) / alpha_t are linear in (x, eps) too); the reference evaluates them as chains of 16-bit
// tensor ops, so this result is closer to the fp32 value of the same formula than the reference's own (tests compare both).
template <bool FP16>
__global__ void linear_step_kernel(const void* sample_, const void* m0_, const void* m1_, const void* noise_, void* prev_, long long n,
                                   float a, float b, float c, float s) {
  pdl_trigger();
  pdl_wait();
  using H = Half16<FP16>;
  const typename H::T* sample = static_cast<const typename H::T*>(sample_);
  const typename H::T* m0 = static_cast<const typename H::T*>(m0_);
  const typename H::T* m1 = static_cast<const typename H::T*>(m1_);
  const typename H::T* noise = static_cast<const typename H::T*>(noise_);
  typename H::T* prev = static_cast<typename H::T*>(prev_);
  long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float v = a * H::to_float(sample[i]);
  if (m0) v = fmaf(b, H::to_float(m0[i]), v);
  if (m1) v = fmaf(c, H::to_float(m1[i]), v);
  if (noise) v = fmaf(s, H::to_float(noise[i]), v);
  prev[i] = H::from_float(v);
}

// y = r16(x / div)  (EulerDiscreteScheduler.scale_model_input, scheduling_euler_discrete.py:345)
template <bool FP16>
__global__ void scale_kernel(const void* x_, void* y_, long long n, float div) {
  pdl_trigger();
  pdl_wait();
  using H = Half16<FP16>;
  const typename H::T* x = static_cast<const typename H::T*>(x_);
  typename H::T* y = static_cast<typename H::T*>(y_);
  long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n) return;
  y[i] = H::from_float(__fdiv_rn(H::to_float(x[i]), div));
}

// ------------------------------------------------------------------------------------------------
// row softmax fp32 -> 16 bit (one CTA per row) and 16-bit transpose (32x32 tiles through smem)
// ------------------------------------------------------------------------------------------------
template <bool FP16>
__global__ void __launch_bounds__(256) softmax_rows_kernel(const float* __restrict__ s, long long ld_s, void* p_,
                                                           long long ld_p, int cols, float scale_log2) {
  pdl_trigger();
  pdl_wait();
  using H = Half16<FP16>;
  __shared__ float red[8];
  const float* row = s + static_cast<long long>(blockIdx.x) * ld_s;
  typename H::T* prow = static_cast<typename H::T*>(p_) + static_cast<long long>(blockIdx.x) * ld_p;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float mx = -INFINITY;
  for (int c = threadIdx.x * 4; c < cols; c += blockDim.x * 4) {
    if (c + 4 <= cols) {
      float4 v = *reinterpret_cast<const float4*>(row + c);
      mx = fmaxf(mx, fmaxf(fmaxf(v.x, v.y), fmaxf(v.z, v.w)));
    } else {
      for (int j = c; j < cols; ++j) mx = fmaxf(mx, row[j]);
    }
  }
  mx = warp_max(mx);
  if (lane == 0) red[warp] = mx;
  __syncthreads();
  mx = red[0];
  for (int w = 1; w < (blockDim.x >> 5); ++w) mx = fmaxf(mx, red[w]);
  __syncthreads();
  const float m2 = mx * scale_log2;
  float sum = 0.f;
  for (int c = threadIdx.x * 4; c < cols; c += blockDim.x * 4) {
    if (c + 4 <= cols) {
      float4 v = *reinterpret_cast<const float4*>(row + c);
      sum += exp2f(fmaf(v.x, scale_log2, -m2)) + exp2f(fmaf(v.y, scale_log2, -m2)) +
             exp2f(fmaf(v.z, scale_log2, -m2)) + exp2f(fmaf(v.w, scale_log2, -m2));
    } else {
      for (int j = c; j < cols; ++j) sum += exp2f(fmaf(row[j], scale_log2, -m2));
    }
  }
  sum = warp_sum(sum);
  if (lane == 0) red[warp] = sum;
  __syncthreads();
  sum = 0.f;
  for (int w = 0; w < (blockDim.x >> 5); ++w) sum += red[w];
  const float inv = 1.0f / sum;
  for (int c = threadIdx.x * 4; c < cols; c += blockDim.x * 4) {
    if (c + 4 <= cols) {
      float4 v = *reinterpret_cast<const float4*>(row + c);
      uint2 o;
      o.x = H::pack(exp2f(fmaf(v.x, scale_log2, -m2)) * inv, exp2f(fmaf(v.y, scale_log2, -m2)) * inv);
      o.y = H::pack(exp2f(fmaf(v.z, scale_log2, -m2)) * inv, exp2f(fmaf(v.w, scale_log2, -m2)) * inv);
      *reinterpret_cast<uint2*>(prow + c) = o;
    } else {
      for (int j = c; j < cols; ++j) prow[j] = H::from_float(exp2f(fmaf(row[j], scale_log2, -m2)) * inv);
    }
  }
}

__global__ void transpose16_kernel(const uint16_t* __restrict__ src, long long ld_src, uint16_t* __restrict__ dst,
                                   long long ld_dst, int rows, int cols) {
  pdl_trigger();
  pdl_wait();
  __shared__ uint16_t tile[32][34];
  const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int r = r0 + i, c = c0 + threadIdx.x;
    if (r < rows && c < cols) tile[i][threadIdx.x] = src[static_cast<long long>(r) * ld_src + c];
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int c = c0 + i, r = r0 + threadIdx.x;
    if (r < rows && c < cols) dst[static_cast<long long>(c) * ld_dst + r] = tile[threadIdx.x][i];
  }
}

// DDPMScheduler.step (schedulers/scheduling_ddpm.py:513-565), epsilon prediction, fixed_small variance:
//   x0 = (x - sqrt(1-abar_t) * eps) / sqrt(abar_t); clip; prev = c0*x0 + c1*x + sigma*noise
// computed in fp32 from 16-bit tensors (the reference runs this path in the model dtype; config 0 is fp32 there).
template <bool FP16>
__global__ void ddpm_step_kernel(const void* eps_, const void* sample_, const void* noise_, void* prev_, long long n,
                                 float sqrt_beta_prod, float sqrt_alpha_prod, float c0, float c1, float sigma, int clip,
                                 float clip_range) {
  pdl_trigger();
  pdl_wait();
  using H = Half16<FP16>;
  const typename H::T* eps = static_cast<const typename H::T*>(eps_);
  const typename H::T* sample = static_cast<const typename H::T*>(sample_);
  const typename H::T* noise = static_cast<const typename H::T*>(noise_);
  typename H::T* prev = static_cast<typename H::T*>(prev_);
  long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float x = H::to_float(sample[i]);
  float x0 = __fdiv_rn(__fsub_rn(x, __fmul_rn(sqrt_beta_prod, H::to_float(eps[i]))), sqrt_alpha_prod);
  if (clip) x0 = fminf(fmaxf(x0, -clip_range), clip_range);
  float p = __fadd_rn(__fmul_rn(c0, x0), __fmul_rn(c1, x));
  if (noise) p = __fadd_rn(p, __fmul_rn(sigma, H::to_float(noise[i])));
  prev[i] = H::from_float(p);
}

static inline unsigned int blocks_for(long long n, int threads) {
  return static_cast<unsigned int>((n + threads - 1) / threads);
}

}  // namespace b200

extern "C" {

int b200_nchw_to_nhwc(const void* src, void* dst, int32_t batch, int32_t C, int32_t HW, int32_t ld_dst, int32_t dtype,
                      void* stream) {
  using namespace b200;
  B200_CHECK_ARG(src && dst && batch > 0 && C > 0 && HW > 0 && ld_dst >= C, "nchw_to_nhwc: bad args");
  (void)dtype;  // both 16-bit types move as raw 16-bit words; zero is all-bits-zero in both
  const long long total = static_cast<long long>(batch) * HW * ld_dst;
  launch_pdl(nchw_to_nhwc_kernel<__nv_bfloat16>, dim3(blocks_for(total, 256)), dim3(256), 0, static_cast<cudaStream_t>(stream), 
      static_cast<const __nv_bfloat16*>(src), static_cast<__nv_bfloat16*>(dst), C, HW, ld_dst, total);
  return check_launch("nchw_to_nhwc_kernel");
}

int b200_nhwc_to_nchw(const void* src, int32_t ld_src, void* dst, int32_t batch, int32_t C, int32_t HW, int32_t dtype,
                      void* stream) {
  using namespace b200;
  B200_CHECK_ARG(src && dst && batch > 0 && C > 0 && HW > 0 && ld_src >= C, "nhwc_to_nchw: bad args");
  (void)dtype;
  const long long total = static_cast<long long>(batch) * HW * C;
  launch_pdl(nhwc_to_nchw_kernel<__nv_bfloat16>, dim3(blocks_for(total, 256)), dim3(256), 0, static_cast<cudaStream_t>(stream), 
      static_cast<const __nv_bfloat16*>(src), ld_src, static_cast<__nv_bfloat16*>(dst), C, HW, total);
  return check_launch("nhwc_to_nchw_kernel");
}

int b200_upsample_nearest2x(const void* x, int32_t ldx, void* y, int32_t ldy, int32_t batch, int32_t H, int32_t W,
                            int32_t C, int32_t dtype, void* stream) {
  using namespace b200;
  B200_CHECK_ARG(x && y && batch > 0 && H > 0 && W > 0, "upsample: bad args");
  B200_CHECK_ARG(C % 8 == 0 && ldx % 8 == 0 && ldy % 8 == 0 && aligned16(x) && aligned16(y),
                 "upsample: channels / strides must be multiples of 8 and pointers 16-byte aligned");
  (void)dtype;
  const int V = C / 8;
  const long long total = static_cast<long long>(batch) * 4 * H * W * V;
  launch_pdl(upsample2x_kernel, dim3(blocks_for(total, 256)), dim3(256), 0, static_cast<cudaStream_t>(stream), 
      static_cast<const uint4*>(x), ldx / 8, static_cast<uint4*>(y), ldy / 8, H, W, V, total);
  return check_launch("upsample2x_kernel");
}

int b200_timestep_embedding(const float* t, int32_t n, void* out, int32_t ld_out, int32_t dim, int32_t flip_sin_to_cos,
                            float downscale_freq_shift, float scale, float max_period, int32_t dtype, void* stream) {
  using namespace b200;
  B200_CHECK_ARG(t && out && n > 0 && dim >= 2 && ld_out >= dim, "timestep_embedding: bad args");
  const int total = n * (dim / 2);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (dtype == B200_DTYPE_FP16)
    launch_pdl(timestep_embedding_kernel<true>, dim3(blocks_for(total, 128)), dim3(128), 0, st, t, n, out, ld_out, dim, flip_sin_to_cos,
                                                                           downscale_freq_shift, scale, max_period);
  else
    launch_pdl(timestep_embedding_kernel<false>, dim3(blocks_for(total, 128)), dim3(128), 0, st, t, n, out, ld_out, dim, flip_sin_to_cos,
                                                                            downscale_freq_shift, scale, max_period);
  return check_launch("timestep_embedding_kernel");
}

int b200_small_linear(const b200_small_linear_args* a, void* stream) {
  using namespace b200;
  B200_CHECK_ARG(a && a->x && a->w && a->y, "small_linear: null pointer");
  B200_CHECK_ARG(a->M >= 1 && a->M <= 8, "small_linear: M=%d (need 1..8)", a->M);
  B200_CHECK_ARG(a->K > 0 && a->K % 8 == 0 && aligned16(a->w), "small_linear: K=%d must be a multiple of 8", a->K);
  SmallLinearParams p;
  p.x = a->x; p.ldx = a->ldx; p.M = a->M; p.K = a->K; p.w = a->w; p.N = a->N; p.bias = a->bias;
  p.act_in = a->act_in; p.act_out = a->act_out; p.addend = a->addend; p.ld_add = a->ld_add; p.y = a->y; p.ldy = a->ldy;
  const size_t smem = static_cast<size_t>(a->M) * a->K * sizeof(float);
  B200_CHECK_ARG(smem <= 200 * 1024, "small_linear: M*K too large for shared memory");
  const bool fp16 = a->dtype == B200_DTYPE_FP16;
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(fp16 ? small_linear_kernel<true> : small_linear_kernel<false>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    if (e != cudaSuccess) return set_error(B200_ERR_CUDA, "small_linear smem attr: %s", cudaGetErrorString(e));
  }
  // enough columns per warp to amortise the per-CTA staging of x, but keep >= ~4 CTAs per SM
  int cpw = static_cast<int>(a->N / (8LL * 4 * num_sms()));
  cpw = cpw < 1 ? 1 : (cpw > 8 ? 8 : cpw);
  p.cpw = cpw;
  const int grid = (a->N + 8 * cpw - 1) / (8 * cpw);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (fp16)
    launch_pdl(small_linear_kernel<true>, dim3(grid), dim3(256), smem, st, p);
  else
    launch_pdl(small_linear_kernel<false>, dim3(grid), dim3(256), smem, st, p);
  return check_launch("small_linear_kernel");
}

int b200_euler_step(const void* model_output, const void* sample, void* prev_sample, int64_t n, float sigma,
                    float sigma_next, int32_t dtype, void* stream) {
  using namespace b200;
  B200_CHECK_ARG(model_output && sample && prev_sample && n > 0 && sigma > 0.f, "euler_step: bad args");
  const float dt = sigma_next - sigma;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (dtype == B200_DTYPE_FP16)
    launch_pdl(euler_step_kernel<true>, dim3(blocks_for(n, 256)), dim3(256), 0, st, model_output, sample, prev_sample, n, sigma, dt);
  else
    launch_pdl(euler_step_kernel<false>, dim3(blocks_for(n, 256)), dim3(256), 0, st, model_output, sample, prev_sample, n, sigma, dt);
  return check_launch("euler_step_kernel");
}

int b200_cfg_euler_step(const void* eps_nhwc, int32_t ld_eps, void* latents_nchw, void* next_in_nhwc, int32_t ld_in,
                        int32_t batch, int32_t C, int32_t HW, float guidance_scale, int32_t do_cfg, float sigma,
                        float sigma_next, int32_t dtype, void* stream) {
  using namespace b200;
  B200_CHECK_ARG(eps_nhwc && latents_nchw && next_in_nhwc && batch > 0 && C > 0 && HW > 0 && ld_in >= C && ld_eps >= C,
                 "cfg_euler_step: bad args");
  B200_CHECK_ARG(sigma > 0.f, "cfg_euler_step: sigma must be > 0");
  const float dt = sigma_next - sigma;
  // (sigma**2 + 1) ** 0.5 on a 0-dim fp32 tensor: fp32 square, add, sqrt
  const float next_div = sqrtf(sigma_next * sigma_next + 1.0f);
  const long long total = static_cast<long long>(batch) * HW * ld_in;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (dtype == B200_DTYPE_FP16)
    launch_pdl(cfg_euler_step_kernel<true>, dim3(blocks_for(total, 256)), dim3(256), 0, st, eps_nhwc, ld_eps, latents_nchw, next_in_nhwc,
                                                                        ld_in, batch, C, HW, guidance_scale, sigma, dt,
                                                                        next_div, do_cfg);
  else
    launch_pdl(cfg_euler_step_kernel<false>, dim3(blocks_for(total, 256)), dim3(256), 0, st, eps_nhwc, ld_eps, latents_nchw, next_in_nhwc,
                                                                         ld_in, batch, C, HW, guidance_scale, sigma, dt,
                                                                         next_div, do_cfg);
  return check_launch("cfg_euler_step_kernel");
}

int b200_flow_match_step(const void* model_output, const void* sample, void* prev_sample, int64_t n, float sigma,
                         float sigma_next, int32_t dtype, void* stream) {
  using namespace b200;
  B200_CHECK_ARG(model_output && sample && prev_sample && n > 0, "flow_match_step: bad args");
  const float dt = sigma_next - sigma;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (dtype == B200_DTYPE_FP16)
    launch_pdl(flow_match_step_kernel<true>, dim3(blocks_for(n, 256)), dim3(256), 0, st, model_output, sample, prev_sample, n, dt);
  else
    launch_pdl(flow_match_step_kernel<false>, dim3(blocks_for(n, 256)), dim3(256), 0, st, model_output, sample, prev_sample, n, dt);
  return check_launch("flow_match_step_kernel");
}

int b200_scale(const void* x, void* y, int64_t n, float divisor, int32_t dtype, void* stream) {
  using namespace b200;
  B200_CHECK_ARG(x && y && n > 0 && divisor != 0.f, "scale: bad args");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (dtype == B200_DTYPE_FP16)
    launch_pdl(scale_kernel<true>, dim3(blocks_for(n, 256)), dim3(256), 0, st, x, y, n, divisor);
  else
    launch_pdl(scale_kernel<false>, dim3(blocks_for(n, 256)), dim3(256), 0, st, x, y, n, divisor);
  return check_launch("scale_kernel");
}

int b200_softmax_rows(const float* s, int64_t ld_s, void* p, int64_t ld_p, int32_t rows, int32_t cols, float scale,
                      int32_t dtype, void* stream) {
  using namespace b200;
  B200_CHECK_ARG(s && p && rows > 0 && cols > 0, "softmax_rows: bad args");
  B200_CHECK_ARG(ld_s % 4 == 0 && ld_p % 4 == 0 && aligned16(s) && (reinterpret_cast<uintptr_t>(p) & 7u) == 0,
                 "softmax_rows: strides must be multiples of 4 elements, s 16-byte / p 8-byte aligned");
  const float sl2 = scale * 1.4426950408889634f;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (dtype == B200_DTYPE_FP16)
    launch_pdl(softmax_rows_kernel<true>, dim3(rows), dim3(256), 0, st, s, ld_s, p, ld_p, cols, sl2);
  else
    launch_pdl(softmax_rows_kernel<false>, dim3(rows), dim3(256), 0, st, s, ld_s, p, ld_p, cols, sl2);
  return check_launch("softmax_rows_kernel");
}

int b200_transpose_16(const void* src, int64_t ld_src, void* dst, int64_t ld_dst, int32_t rows, int32_t cols,
                      void* stream) {
  using namespace b200;
  B200_CHECK_ARG(src && dst && rows > 0 && cols > 0, "transpose_16: bad args");
  dim3 grid((cols + 31) / 32, (rows + 31) / 32), block(32, 8);
  launch_pdl(transpose16_kernel, dim3(grid), dim3(block), 0, static_cast<cudaStream_t>(stream), 
      static_cast<const uint16_t*>(src), ld_src, static_cast<uint16_t*>(dst), ld_dst, rows, cols);
  return check_launch("transpose16_kernel");
}

int b200_linear_step(const void* sample, const void* m0, const void* m1, const void* noise, void* prev_sample, int64_t n, float a,
                     float b, float c, float s, int32_t dtype, void* stream) {
  using namespace b200;
  B200_CHECK_ARG(sample && prev_sample && n > 0, "linear_step: bad args");
  B200_CHECK_ARG(dtype == B200_DTYPE_BF16 || dtype == B200_DTYPE_FP16, "linear_step: dtype %d", dtype);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (dtype == B200_DTYPE_FP16)
    launch_pdl(linear_step_kernel<true>, dim3(blocks_for(n, 256)), dim3(256), 0, st, sample, m0, m1, noise, prev_sample, n, a, b, c, s);
  else
    launch_pdl(linear_step_kernel<false>, dim3(blocks_for(n, 256)), dim3(256), 0, st, sample, m0, m1, noise, prev_sample, n, a, b, c, s);
  return check_launch("linear_step_kernel");
}

int b200_ddpm_step(const void* model_output, const void* sample, const void* noise, void* prev_sample, int64_t n,
                   float sqrt_beta_prod_t, float sqrt_alpha_prod_t, float pred_original_coeff, float current_sample_coeff,
                   float sigma, int32_t clip_sample, float clip_range, int32_t dtype, void* stream) {
  using namespace b200;
  B200_CHECK_ARG(model_output && sample && prev_sample && n > 0 && sqrt_alpha_prod_t > 0.f, "ddpm_step: bad args");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (dtype == B200_DTYPE_FP16)
    launch_pdl(ddpm_step_kernel<true>, dim3(blocks_for(n, 256)), dim3(256), 0, st, model_output, sample, noise, prev_sample, n, sqrt_beta_prod_t,
                                                              sqrt_alpha_prod_t, pred_original_coeff, current_sample_coeff,
                                                              sigma, clip_sample, clip_range);
  else
    launch_pdl(ddpm_step_kernel<false>, dim3(blocks_for(n, 256)), dim3(256), 0, st, model_output, sample, noise, prev_sample, n, sqrt_beta_prod_t,
                                                               sqrt_alpha_prod_t, pred_original_coeff, current_sample_coeff,
                                                               sigma, clip_sample, clip_range);
  return check_launch("ddpm_step_kernel");
}

}  // extern "C"
