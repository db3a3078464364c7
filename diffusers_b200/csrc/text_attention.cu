// b200_text_attention: attention of the text encoders (CLIP-L / OpenCLIP-bigG: 77 tokens, causal; T5-XXL: 512 tokens, additive
// relative-position bias, no scaling) - head_dim 64, at most 512 keys, run ONCE per prompt (1 GFLOP for CLIP, 0.1 TFLOP for T5-XXL).
//
// These shapes do not fit the tcgen05 kernel of the denoisers (no mask / bias there, 128-row query tiles against 77 keys), and at this
// size the work is latency, not throughput: a CTA stages the whole K (transposed) and V of one (batch, head) in shared memory and each
// warp computes four query rows in fp32 - scores with a lane per pair of keys (conflict-free 32-bit words of the transposed K),
// softmax in registers, P V with a lane per pair of output columns.
#include "common.cuh"
#include "host_common.h"

namespace b200 {

struct TextAttnParams {
  const void* q;
  const void* k;
  const void* v;
  void* o;
  int batch, heads, sq, sk, skp;  // skp = sk rounded up to 64 (padded keys are masked)
  long long q_rs, q_bs, k_rs, k_bs, v_rs, v_bs, o_rs, o_bs;
  float scale;
  int causal;
  const float* bias;  // [heads][sq][sk] or nullptr
};

constexpr int kTextQB = 32;      // query rows per CTA (8 warps x 4)
constexpr int kTextMaxSk = 512;

template <bool FP16>
__global__ void __launch_bounds__(256) text_attention_kernel(const TextAttnParams p) {
  using H = Half16<FP16>;
  extern __shared__ __align__(16) uint8_t smem[];
  const int skp = p.skp;
  uint32_t* kt = reinterpret_cast<uint32_t*>(smem);                     // [64][skp / 2] words: kt[d][j / 2] = (K[j][d], K[j + 1][d])
  uint32_t* vs = kt + 64 * (skp >> 1);                                  // [skp][32] words: V rows
  float* ps = reinterpret_cast<float*>(vs + skp * 32);                  // [8 warps][skp]
  float* qs = ps + 8 * skp;                                             // [8 warps][64]
  pdl_trigger();
  pdl_wait();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.z, head = blockIdx.y;
  const typename H::T* kbase = static_cast<const typename H::T*>(p.k) + b * p.k_bs + head * 64;
  const typename H::T* vbase = static_cast<const typename H::T*>(p.v) + b * p.v_bs + head * 64;
  // ---- stage K^T and V (rows >= sk are zero)
  for (int idx = threadIdx.x; idx < skp * 32; idx += 256) {
    const int j = idx >> 5, w = idx & 31;  // key j, word w (columns 2w, 2w + 1)
    uint32_t kw = 0, vw = 0;
    if (j < p.sk) {
      kw = *reinterpret_cast<const uint32_t*>(kbase + j * p.k_rs + 2 * w);
      vw = *reinterpret_cast<const uint32_t*>(vbase + j * p.v_rs + 2 * w);
    }
    vs[j * 32 + w] = vw;
    // transposed: 16-bit element (d, j) of kt
    uint16_t* kt16 = reinterpret_cast<uint16_t*>(kt);
    kt16[(2 * w) * skp + j] = static_cast<uint16_t>(kw & 0xffffu);
    kt16[(2 * w + 1) * skp + j] = static_cast<uint16_t>(kw >> 16);
  }
  __syncthreads();
  const int npair = skp >> 6;  // pairs of keys per lane (each lane owns keys 64 t + 2 lane, + 1)
  float* myp = ps + warp * skp;
  float* myq = qs + warp * 64;
#pragma unroll 1
  for (int qi = 0; qi < kTextQB / 8; ++qi) {
    const int row = blockIdx.x * kTextQB + warp * (kTextQB / 8) + qi;
    if (row >= p.sq) break;  // warp-uniform
    const typename H::T* qrow = static_cast<const typename H::T*>(p.q) + b * p.q_bs + static_cast<long long>(row) * p.q_rs + head * 64;
    {
      const float2 f = H::unpack(*reinterpret_cast<const uint32_t*>(qrow + 2 * lane));
      myq[2 * lane] = f.x * p.scale;
      myq[2 * lane + 1] = f.y * p.scale;
    }
    __syncwarp();
    float s[kTextMaxSk / 32];  // 16 scores per lane: pairs t = 0..7
#pragma unroll
    for (int t = 0; t < kTextMaxSk / 64; ++t) s[2 * t] = s[2 * t + 1] = 0.f;
#pragma unroll 4
    for (int d = 0; d < 64; ++d) {
      const float qd = myq[d];
      const uint32_t* krow = kt + d * (skp >> 1) + lane;
#pragma unroll
      for (int t = 0; t < kTextMaxSk / 64; ++t) {
        if (t < npair) {
          const float2 kk = H::unpack(krow[t * 32]);
          s[2 * t] = fmaf(qd, kk.x, s[2 * t]);
          s[2 * t + 1] = fmaf(qd, kk.y, s[2 * t + 1]);
        }
      }
    }
    const float* brow = p.bias ? p.bias + (static_cast<long long>(head) * p.sq + row) * p.sk : nullptr;
    float m = -INFINITY;
#pragma unroll
    for (int t = 0; t < kTextMaxSk / 64; ++t) {
      if (t < npair) {
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const int j = 64 * t + 2 * lane + e;
          float x = s[2 * t + e];
          if (brow && j < p.sk) x += brow[j];
          if (j >= p.sk || (p.causal && j > row)) x = -INFINITY;
          s[2 * t + e] = x;
          m = fmaxf(m, x);
        }
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    float l = 0.f;
#pragma unroll
    for (int t = 0; t < kTextMaxSk / 64; ++t) {
      if (t < npair) {
        const float e0 = __expf(s[2 * t] - m), e1 = __expf(s[2 * t + 1] - m);  // exp(-inf) = 0 for masked keys
        l += e0 + e1;
        *reinterpret_cast<float2*>(myp + 64 * t + 2 * lane) = make_float2(e0, e1);
      }
    }
    l = warp_sum(l);
    __syncwarp();
    float o0 = 0.f, o1 = 0.f;
    const int jmax = p.causal ? min(p.sk, row + 1) : p.sk;
#pragma unroll 4
    for (int j = 0; j < jmax; ++j) {
      const float pj = myp[j];
      const float2 vv = H::unpack(vs[j * 32 + lane]);
      o0 = fmaf(pj, vv.x, o0);
      o1 = fmaf(pj, vv.y, o1);
    }
    const float inv = 1.0f / l;
    typename H::T* orow = static_cast<typename H::T*>(p.o) + b * p.o_bs + static_cast<long long>(row) * p.o_rs + head * 64;
    *reinterpret_cast<uint32_t*>(orow + 2 * lane) = H::pack(o0 * inv, o1 * inv);
    __syncwarp();
  }
}

static size_t text_attention_smem(int skp) { return static_cast<size_t>(64) * skp * 2 + static_cast<size_t>(skp) * 128 + 8 * skp * 4 + 8 * 64 * 4; }

int init_text_attention() {
  const int mx = static_cast<int>(text_attention_smem(kTextMaxSk));
  cudaError_t e = cudaFuncSetAttribute(text_attention_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, mx);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(text_attention_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, mx);
  if (e != cudaSuccess) return set_error(B200_ERR_CUDA, "text_attention smem attr: %s", cudaGetErrorString(e));
  return 0;
}

}  // namespace b200

extern "C" {

int b200_text_attention(const b200_text_attention_args* a, void* stream) {
  using namespace b200;
  B200_CHECK_ARG(a && a->q && a->k && a->v && a->o, "text_attention: null pointer");
  B200_CHECK_ARG(a->batch > 0 && a->heads > 0 && a->sq > 0 && a->sk > 0 && a->sk <= kTextMaxSk, "text_attention: bad shape (at most %d keys)", kTextMaxSk);
  B200_CHECK_ARG(a->batch <= 65535 && a->heads <= 65535, "text_attention: batch / heads too large");
  B200_CHECK_ARG(a->dtype == B200_DTYPE_BF16 || a->dtype == B200_DTYPE_FP16, "text_attention: dtype %d", a->dtype);
  const long long strides[8] = {a->q_row_stride, a->q_batch_stride, a->k_row_stride, a->k_batch_stride, a->v_row_stride, a->v_batch_stride,
                                a->o_row_stride, a->o_batch_stride};
  for (long long st : strides) B200_CHECK_ARG(st % 2 == 0, "text_attention: strides must be even (32-bit accesses)");
  B200_CHECK_ARG(((reinterpret_cast<uintptr_t>(a->q) | reinterpret_cast<uintptr_t>(a->k) | reinterpret_cast<uintptr_t>(a->v) | reinterpret_cast<uintptr_t>(a->o)) & 3u) == 0,
                 "text_attention: pointers must be 4-byte aligned");
  TextAttnParams p;
  p.q = a->q; p.k = a->k; p.v = a->v; p.o = a->o;
  p.batch = a->batch; p.heads = a->heads; p.sq = a->sq; p.sk = a->sk;
  p.skp = rup(a->sk, 64);
  p.q_rs = a->q_row_stride; p.q_bs = a->q_batch_stride; p.k_rs = a->k_row_stride; p.k_bs = a->k_batch_stride;
  p.v_rs = a->v_row_stride; p.v_bs = a->v_batch_stride; p.o_rs = a->o_row_stride; p.o_bs = a->o_batch_stride;
  p.scale = a->scale;
  p.causal = a->causal;
  p.bias = a->bias;
  const dim3 grid(cdiv(a->sq, kTextQB), a->heads, a->batch);
  const size_t smem = text_attention_smem(p.skp);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  cudaError_t e = a->dtype == B200_DTYPE_FP16 ? launch_pdl(text_attention_kernel<true>, grid, dim3(256), smem, st, p)
                                              : launch_pdl(text_attention_kernel<false>, grid, dim3(256), smem, st, p);
  if (e != cudaSuccess) return set_error(B200_ERR_CUDA, "text_attention launch: %s", cudaGetErrorString(e));
  return 0;
}

}  // extern "C"
