// EXPERIMENTAL - off by default (B200_ATTN_SPLIT=1), NOT yet validated on hardware; kept apart from attention_pipe.cu so
// that the validated kernel's source is untouched.
//
// The pipelined head_dim-64 attention of attention_pipe.cu with TWO softmax warps per 32 query rows, each taking 32 of the
// 64 keys of a half (12 warps per CTA, 80 registers, still two CTAs per SM).  The row maximum is exchanged through shared
// memory once per half (named barrier among the 256 softmax threads), the row sums are kept per warp and added at the end;
// the rescale of O and the final normalisation are split by columns.  Intent: halve the softmax latency per half and
// double the warps the schedulers can interleave (attention_pipe.cu: tensor pipe still only ~25 % active).
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include <type_traits>

#include "common.cuh"
#include "host_common.h"

namespace b200 {

struct AttnSplitParams {
  CUtensorMap q_map;  // box {64, 128, 1, 1}
  CUtensorMap k_map;  // box {64, 64, 1, 1}
  CUtensorMap v_map;  // box {64, 64, 1, 1}
  void* o;
  long long o_row_stride, o_batch_stride;
  int batch, heads, sq, sk;
  int q_tiles;    // ceil(sq / 128)
  int kv_halves;  // ceil(sk / 64)
  float scale_log2;
};

struct AttnSplitCfg {
  static constexpr int HD = 64;
  static constexpr int Q_BYTES = 128 * 64 * 2;   // 16 KB
  static constexpr int KV_BYTES = 64 * 64 * 2;   // 8 KB: one 64-key half of K or V
  static constexpr int KS = 4, VS = 4;
  static constexpr int TMEM_COLS = 256;
  static constexpr int S_COL = 0, P_COL = 128, O_COL = 192;
  static constexpr int XCHG_BYTES = 2 * 2 * 128 * 4;  // [half parity][warp group][row] partial row maxima / sums
  static constexpr int SMEM_BYTES = Q_BYTES + (KS + VS) * KV_BYTES + 1024 + 256 + XCHG_BYTES;
};

__device__ __forceinline__ float split_ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float split_ex2_poly(float x) {  // see ex2_poly in attention.cu
  x = fmaxf(x, -126.0f);
  const float t = x + 12582912.0f;
  const float f = x - (t - 12582912.0f);
  float p = fmaf(f, 0.05583828f, 0.24263948f);
  p = fmaf(p, f, 0.69313675f);
  p = fmaf(p, f, 0.99992454f);
  return __int_as_float(__float_as_int(p) + (__float_as_int(t) << 23));
}

template <bool FP16>
__global__ void __launch_bounds__(384, 2) attention64_split_kernel(const __grid_constant__ AttnSplitParams p) {
  using Cfg = AttnSplitCfg;
  using H = Half16<FP16>;
  constexpr int KS = Cfg::KS, VS = Cfg::VS;

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* s_q = smem;                        // [128][64]
  uint8_t* s_k = s_q + Cfg::Q_BYTES;          // [KS][64][64]
  uint8_t* s_v = s_k + KS * Cfg::KV_BYTES;    // [VS][64][64]
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_v + VS * Cfg::KV_BYTES);
  uint64_t* q_full = bars;               // [1]
  uint64_t* k_full = q_full + 1;         // [KS]
  uint64_t* k_empty = k_full + KS;       // [KS]
  uint64_t* v_full = k_empty + KS;       // [VS]
  uint64_t* v_empty = v_full + VS;       // [VS]
  uint64_t* s_full = v_empty + VS;       // [2]  S(h) landed in TMEM
  uint64_t* p_full = s_full + 2;         // [2]  softmax wrote P(h) and is done with S(h): 128 arrivals
  uint64_t* pv_done = p_full + 2;        // [2]  P(h) V(h) accumulated into O (P buffer reusable, O consistent)
  uint64_t* o_full = pv_done + 2;        // [1]
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(o_full + 1);
  static_assert((1 + 2 * KS + 2 * VS + 2 + 2 + 2 + 1) * 8 + 4 <= 256, "barrier area");
  float* xchg = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(bars) + 256);  // [2][2][128]

  pdl_trigger();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int qt = blockIdx.x % p.q_tiles;
  const int bh = blockIdx.x / p.q_tiles;
  const int head = bh % p.heads;
  const int b = bh / p.heads;
  const int q_row0 = qt * 128;
  const int n_half = p.kv_halves;

  if (warp == 0 && lane == 0) {
    prefetch_tensormap(&p.q_map);
    prefetch_tensormap(&p.k_map);
    prefetch_tensormap(&p.v_map);
    mbar_init(q_full, 1);
    mbar_init(o_full, 1);
    for (int i = 0; i < KS; ++i) {
      mbar_init(&k_full[i], 1);
      mbar_init(&k_empty[i], 1);
    }
    for (int i = 0; i < VS; ++i) {
      mbar_init(&v_full[i], 1);
      mbar_init(&v_empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&s_full[i], 1);
      mbar_init(&p_full[i], 256);
      mbar_init(&pv_done[i], 1);
    }
    fence_barrier_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_ptr, Cfg::TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  pdl_wait();

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (elect_one()) {
      mbar_expect_tx(q_full, Cfg::Q_BYTES);
      tma_load_4d(s_q, &p.q_map, q_full, 0, q_row0, head, b);
    }
    int ks = 0, vs = 0;
    uint32_t kph = 0, vph = 0;
    for (int h = 0; h < n_half; ++h) {
      mbar_wait(&k_empty[ks], kph ^ 1u);
      if (elect_one()) {
        mbar_expect_tx(&k_full[ks], Cfg::KV_BYTES);
        tma_load_4d(s_k + ks * Cfg::KV_BYTES, &p.k_map, &k_full[ks], 0, h * 64, head, b);
      }
      if (++ks == KS) { ks = 0; kph ^= 1u; }
      mbar_wait(&v_empty[vs], vph ^ 1u);
      if (elect_one()) {
        mbar_expect_tx(&v_full[vs], Cfg::KV_BYTES);
        tma_load_4d(s_v + vs * Cfg::KV_BYTES, &p.v_map, &v_full[vs], 0, h * 64, head, b);
      }
      if (++vs == VS) { vs = 0; vph ^= 1u; }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    constexpr uint32_t idesc_qk = make_idesc(128, 64, FP16, false, false);
    constexpr uint32_t idesc_pv = make_idesc(128, 64, FP16, false, true);
    int ks = 0, vs = 0;
    uint32_t kph = 0, vph = 0;
    const uint32_t qa = smem_u32(s_q);

    auto issue_s = [&](int buf, int kstage) {  // S_buf = Q K_half^T: 128 x 64 x 64, four K steps of 16
      const uint64_t qd = make_smem_desc_sw128(qa, 16, 1024);
      const uint64_t kd = make_smem_desc_sw128(smem_u32(s_k + kstage * Cfg::KV_BYTES), 16, 1024);
#pragma unroll
      for (int k = 0; k < 4; ++k)
        umma_ss(tmem_base + Cfg::S_COL + buf * 64, qd + 2u * k, kd + 2u * k, idesc_qk, k != 0 ? 1u : 0u);
    };
    auto issue_pv = [&](int buf, int vstage, bool accumulate) {  // O += P_buf V_half: K = 64 keys, four steps of 16
      const uint32_t va = smem_u32(s_v + vstage * Cfg::KV_BYTES);
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const uint64_t vd = make_smem_desc_sw128(va + k * 2048, Cfg::KV_BYTES, 1024);  // MN-major: 16 keys = 2 x 8-row groups
        umma_ts(tmem_base + Cfg::O_COL, tmem_base + Cfg::P_COL + buf * 32 + k * 8, vd, idesc_pv, (accumulate || k != 0) ? 1u : 0u);
      }
    };

    mbar_wait(q_full, 0);
    for (int h0 = 0; h0 < 2 && h0 < n_half; ++h0) {  // prologue: both score buffers filled
      mbar_wait(&k_full[ks], kph);
      tc_fence_after();
      if (elect_one()) {
        issue_s(h0, ks);
        umma_commit(&s_full[h0]);
        umma_commit(&k_empty[ks]);
      }
      if (++ks == KS) { ks = 0; kph ^= 1u; }
    }
    for (int h = 0; h < n_half; ++h) {
      const int buf = h & 1;
      mbar_wait(&p_full[buf], (h >> 1) & 1u);
      mbar_wait(&v_full[vs], vph);
      tc_fence_after();
      if (elect_one()) {
        issue_pv(buf, vs, h > 0);
        umma_commit(&v_empty[vs]);
        umma_commit(&pv_done[buf]);
      }
      if (++vs == VS) { vs = 0; vph ^= 1u; }
      if (h + 2 < n_half) {  // the softmax has released S_buf: refill it two halves ahead
        mbar_wait(&k_full[ks], kph);
        tc_fence_after();
        if (elect_one()) {
          issue_s(buf, ks);
          umma_commit(&s_full[buf]);
          umma_commit(&k_empty[ks]);
        }
        if (++ks == KS) { ks = 0; kph ^= 1u; }
      }
    }
    if (elect_one()) umma_commit(o_full);
  } else if (warp >= 4) {
    // ===================== softmax warps: two threads per query row, 32 keys of each half apiece =====================
    const int q = warp & 3;            // TMEM lane quarter
    const int grp = (warp - 4) >> 2;   // which 32 keys of a half / which 32 columns of O this warp owns
    const int row = q * 32 + lane;
    const uint32_t lane_off = static_cast<uint32_t>(q * 32) << 16;
    const uint32_t o_t = tmem_base + lane_off + Cfg::O_COL;
    const float sc = p.scale_log2;
    float m = 0.f, l = 0.f;
    using Masked = std::integral_constant<bool, true>;
    using Full = std::integral_constant<bool, false>;

    for (int h = 0; h < n_half; ++h) {
      const int buf = h & 1;
      const uint32_t s_t = tmem_base + lane_off + Cfg::S_COL + buf * 64;
      const uint32_t p_t = tmem_base + lane_off + Cfg::P_COL + buf * 32;
      mbar_wait_warp(&s_full[buf], (h >> 1) & 1u);
      tc_fence_after();
      const int kv_left = p.sk - h * 64;
      const bool partial = kv_left < 64;
      uint32_t a0[32];
      tmem_ld32(s_t + grp * 32, a0);
      tmem_wait_ld();
      const int col0 = grp * 32;  // key index (within the half) of a0[0]
      float mx = -INFINITY;
      if (!partial) {
#pragma unroll
        for (int k = 0; k < 32; ++k) mx = fmaxf(mx, __uint_as_float(a0[k]));
      } else {
#pragma unroll
        for (int k = 0; k < 32; ++k) {
          if (col0 + k < kv_left) mx = fmaxf(mx, __uint_as_float(a0[k]));
        }
      }
      {
        // both threads of a row must use the same maximum: exchange the partial maxima (double-buffered by half parity:
        // the barrier of half h+1 orders every read of half h before the writes of half h+2)
        float* xm = xchg + (h & 1) * 256;
        xm[grp * 128 + row] = mx;
        named_bar_sync(1, 256);
        mx = fmaxf(mx, xm[(grp ^ 1) * 128 + row]);
      }
      mx *= sc;  // sc > 0
      if (h == 0) {
        m = (mx == -INFINITY) ? 0.f : mx;
      } else {
        const bool need = mx > m + 8.0f;
        if (__any_sync(0xffffffffu, need)) {
          // every P V issued so far (halves <= h-1, in order on the tensor pipe) must have landed before O is rescaled
          mbar_wait_warp(&pv_done[(h - 1) & 1], ((h - 1) >> 1) & 1u);
          tc_fence_after();
          const float alpha = need ? split_ex2_approx(m - mx) : 1.0f;
          if (need) {
            m = mx;
            l *= alpha;
          }
          uint32_t t0[32];
          tmem_ld32(o_t + grp * 32, t0);  // each warp of the pair rescales its 32 columns of O
          tmem_wait_ld();
#pragma unroll
          for (int k = 0; k < 32; ++k) t0[k] = __float_as_uint(__uint_as_float(t0[k]) * alpha);
          tmem_st32(o_t + grp * 32, t0);
          tmem_wait_st();
        }
      }
      // P_buf still feeds P V of half h-2 until that MMA completes
      if (h >= 2) mbar_wait_warp(&pv_done[buf], ((h - 2) >> 1) & 1u);

      auto exps32 = [&](auto masked, const uint32_t (&src)[32], uint32_t (&dst)[16], int c0) {
#pragma unroll
        for (int k = 0; k < 16; ++k) {
          float p0 = split_ex2_approx(fmaf(__uint_as_float(src[2 * k]), sc, -m));
          float p1 = (k & 1) ? split_ex2_poly(fmaf(__uint_as_float(src[2 * k + 1]), sc, -m))
                             : split_ex2_approx(fmaf(__uint_as_float(src[2 * k + 1]), sc, -m));
          if constexpr (decltype(masked)::value) {
            if (c0 + 2 * k >= kv_left) p0 = 0.f;
            if (c0 + 2 * k + 1 >= kv_left) p1 = 0.f;
          }
          l += p0 + p1;
          dst[k] = H::pack(p0, p1);
        }
      };
      uint32_t pk[16];
      if (partial) exps32(Masked{}, a0, pk, col0); else exps32(Full{}, a0, pk, col0);
      tmem_st16(p_t + grp * 16, pk);  // 32 keys -> 16 P columns
      tmem_wait_st();
      tc_fence_before();
      mbar_arrive(&p_full[buf]);
    }

    {  // the two partial row sums (same running max) add up
      float* xl = xchg + (n_half & 1) * 256;
      xl[grp * 128 + row] = l;
      named_bar_sync(1, 256);
      l += xl[(grp ^ 1) * 128 + row];
    }

    // ---- epilogue: O / l -> global
    mbar_wait_warp(o_full, 0);
    tc_fence_after();
    const int qrow = q_row0 + row;
    const bool valid = qrow < p.sq;
    const float inv_l = 1.0f / l;
    typename H::T* orow = static_cast<typename H::T*>(p.o) + static_cast<long long>(b) * p.o_batch_stride +
                          static_cast<long long>(qrow) * p.o_row_stride + head * 64;
    {
      const int c = grp;  // this warp's 32 columns of O
      uint32_t v[32];
      tmem_ld32(o_t + c * 32, v);
      tmem_wait_ld();
      if (valid) {
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          uint4 o;
          o.x = H::pack(__uint_as_float(v[g * 8 + 0]) * inv_l, __uint_as_float(v[g * 8 + 1]) * inv_l);
          o.y = H::pack(__uint_as_float(v[g * 8 + 2]) * inv_l, __uint_as_float(v[g * 8 + 3]) * inv_l);
          o.z = H::pack(__uint_as_float(v[g * 8 + 4]) * inv_l, __uint_as_float(v[g * 8 + 5]) * inv_l);
          o.w = H::pack(__uint_as_float(v[g * 8 + 6]) * inv_l, __uint_as_float(v[g * 8 + 7]) * inv_l);
          *reinterpret_cast<uint4*>(orow + c * 32 + g * 8) = o;
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
  }
}

static bool g_split_ready = false;

// Never fatal for b200_init: the experimental kernel must not be able to take the validated paths down with it.
int init_attention_split() {
  cudaError_t e = cudaFuncSetAttribute(attention64_split_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, AttnSplitCfg::SMEM_BYTES);
  if (e == cudaSuccess)
    e = cudaFuncSetAttribute(attention64_split_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, AttnSplitCfg::SMEM_BYTES);
  if (e != cudaSuccess) cudaGetLastError();  // clear the sticky-less error state
  g_split_ready = (e == cudaSuccess);
  return 0;
}

bool attention_split_enabled() {
  static const bool on = getenv("B200_ATTN_SPLIT") && atoi(getenv("B200_ATTN_SPLIT")) != 0;
  return on;
}

// head_dim 64, one query tile per CTA; arguments already validated by b200_attention
int launch_attention64_split(const b200_attention_args* a, cudaStream_t st) {
  if (!g_split_ready) return set_error(B200_ERR_UNSUPPORTED, "attention (split): kernel attributes could not be set at init");
  AttnSplitParams prm;
  memset(&prm, 0, sizeof(prm));
  auto mk = [&](CUtensorMap* m, const void* base, int rows, long long row_stride, long long batch_stride, uint32_t box_rows,
                const char* what) {
    const uint32_t box[4] = {64u, box_rows, 1u, 1u};
    const uint64_t dims[4] = {64u, static_cast<uint64_t>(rows), static_cast<uint64_t>(a->heads), static_cast<uint64_t>(a->batch)};
    const uint64_t str[3] = {static_cast<uint64_t>(row_stride) * 2, 64u * 2,
                             static_cast<uint64_t>(batch_stride > 0 ? batch_stride : row_stride * rows) * 2};
    return make_tensor_map_16b(m, base, 4, dims, str, box, what);
  };
  int r;
  if ((r = mk(&prm.q_map, a->q, a->sq, a->q_row_stride, a->q_batch_stride, 128u, "attention Q"))) return r;
  if ((r = mk(&prm.k_map, a->k, a->sk, a->k_row_stride, a->k_batch_stride, 64u, "attention K (halves)"))) return r;
  if ((r = mk(&prm.v_map, a->v, a->sk, a->v_row_stride, a->v_batch_stride, 64u, "attention V (halves)"))) return r;
  prm.o = a->o;
  prm.o_row_stride = a->o_row_stride;
  prm.o_batch_stride = a->o_batch_stride;
  prm.batch = a->batch;
  prm.heads = a->heads;
  prm.sq = a->sq;
  prm.sk = a->sk;
  prm.q_tiles = (a->sq + 127) / 128;
  prm.kv_halves = (a->sk + 63) / 64;
  const float scale = a->scale > 0.f ? a->scale : 0.125f;
  prm.scale_log2 = scale * 1.4426950408889634f;
  const long long grid_ll = static_cast<long long>(a->batch) * a->heads * prm.q_tiles;
  B200_CHECK_ARG(grid_ll < (1ll << 31), "attention: grid too large");
  const dim3 grid(static_cast<unsigned>(grid_ll));
  cudaError_t e = a->dtype == B200_DTYPE_FP16
                      ? launch_pdl(attention64_split_kernel<true>, grid, dim3(384), AttnSplitCfg::SMEM_BYTES, st, prm)
                      : launch_pdl(attention64_split_kernel<false>, grid, dim3(384), AttnSplitCfg::SMEM_BYTES, st, prm);
  if (e != cudaSuccess) return set_error(B200_ERR_CUDA, "attention (split) launch: %s", cudaGetErrorString(e));
  return 0;
}

}  // namespace b200
