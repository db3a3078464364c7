// b200_attention: FlashAttention-style fused softmax(Q K^T * scale) V on tcgen05 tensor cores.
//
//   CTA = NQ query tiles of 128 rows of one (batch, head); K/V streamed in 128-row blocks by TMA.
//   warp 0        TMA producer (Q once; K and V rings)
//   warp 1        MMA issuer   S_i = Q_i K_j^T   (SS, both operands K-major, 128B-swizzled smem)
//                              O_i += P_i V_j    (TS: P from TMEM, V MN-major from smem)
//   warp 2        TMEM allocator
//   warps 4..7    softmax warpgroup for tile 0 (thread = row: no shuffles), 8..11 for tile 1
//
//   TMEM columns: S_i (128 fp32) at i*128, P_i (16-bit, 64 columns) aliases the upper half of S_i,
//   O_i (HD fp32) at NQ*128 + i*HD.  Online softmax with lazy rescaling: the running max only
//   moves (and O is only rescaled) when it grew by more than 2^8, which keeps exp2 arguments
//   bounded and is exact up to rounding.
#include <math.h>
#include <string.h>

#include <type_traits>

#include "common.cuh"
#include "host_common.h"

namespace b200 {

struct AttnParams {
  CUtensorMap q_map, k_map, v_map;
  void* o;
  long long o_row_stride, o_batch_stride;
  int batch, heads, sq, sk;
  int q_tiles;       // ceil(sq / (128*NQ))
  int kv_blocks;     // ceil(sk / 128)
  float scale_log2;  // scale * log2(e)
};

template <int HD, int NQ>
struct AttnCfg {
  static constexpr int SLABS = HD / 64;               // 64-wide (128 B) column slabs per row
  static constexpr int TILE_BYTES = 128 * HD * 2;      // one 128-row tile of Q, K or V
  static constexpr int SLAB_BYTES = 128 * 64 * 2;      // 16 KB
  static constexpr int K_STAGES = (HD == 64) ? 3 : 2;
  static constexpr int V_STAGES = 2;
  static constexpr int TMEM_COLS = (NQ == 2) ? 512 : 256;
  static constexpr int SMEM_BYTES = (NQ + K_STAGES + V_STAGES) * TILE_BYTES + 1024 + 256;
  static constexpr int THREADS = 128 + 128 * NQ;
};

__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// POLY: score pairs of every 8 whose exponentials are evaluated by the FMA-pipe polynomial instead of MUFU (1..3)
template <int HD, int NQ, bool FP16, int POLY = 1>
__global__ void __launch_bounds__(AttnCfg<HD, NQ>::THREADS, (HD == 64 && NQ == 1) ? 2 : 1) attention_kernel(const __grid_constant__ AttnParams p) {
  using Cfg = AttnCfg<HD, NQ>;
  using H = Half16<FP16>;
  constexpr int SLABS = Cfg::SLABS;
  constexpr int KS = Cfg::K_STAGES, VS = Cfg::V_STAGES;

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* s_q = smem;                                   // [NQ][SLABS][128][64]
  uint8_t* s_k = s_q + NQ * Cfg::TILE_BYTES;             // [KS][SLABS][128][64]
  uint8_t* s_v = s_k + KS * Cfg::TILE_BYTES;             // [VS][SLABS][128][64]
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_v + VS * Cfg::TILE_BYTES);
  uint64_t* q_full = bars;              // [NQ]
  uint64_t* k_full = q_full + 2;        // [KS]
  uint64_t* k_empty = k_full + 3;       // [KS]
  uint64_t* v_full = k_empty + 3;       // [VS]
  uint64_t* v_empty = v_full + 2;       // [VS]
  uint64_t* s_full = v_empty + 2;       // [NQ]
  uint64_t* p_full = s_full + 2;        // [NQ]
  uint64_t* o_full = p_full + 2;        // [NQ]
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(o_full + 2);

  pdl_trigger();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  // work decomposition: blockIdx.x -> (q tile group, head, batch)
  const int qt = blockIdx.x % p.q_tiles;
  const int bh = blockIdx.x / p.q_tiles;
  const int head = bh % p.heads;
  const int b = bh / p.heads;
  const int q_row0 = qt * 128 * NQ;

  if (warp == 0 && lane == 0) {
    prefetch_tensormap(&p.q_map);
    prefetch_tensormap(&p.k_map);
    prefetch_tensormap(&p.v_map);
    for (int i = 0; i < NQ; ++i) {
      mbar_init(&q_full[i], 1);
      mbar_init(&s_full[i], 1);
      mbar_init(&p_full[i], 128);
      mbar_init(&o_full[i], 1);
    }
    for (int i = 0; i < KS; ++i) {
      mbar_init(&k_full[i], 1);
      mbar_init(&k_empty[i], 1);
    }
    for (int i = 0; i < VS; ++i) {
      mbar_init(&v_full[i], 1);
      mbar_init(&v_empty[i], 1);
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
  const int n_blocks = p.kv_blocks;
  pdl_wait();

  if (warp == 0) {
    // ===================== TMA producer (whole warp, one elected lane issues: see elect_one) =====================
    if (elect_one()) {
      for (int i = 0; i < NQ; ++i) {
        mbar_expect_tx(&q_full[i], Cfg::TILE_BYTES);
        for (int s = 0; s < SLABS; ++s)
          tma_load_4d(s_q + i * Cfg::TILE_BYTES + s * Cfg::SLAB_BYTES, &p.q_map, &q_full[i], s * 64, q_row0 + i * 128,
                      head, b);
      }
    }
    int ks = 0, vs = 0;
    uint32_t kph = 0, vph = 0;
    for (int j = 0; j < n_blocks; ++j) {
      mbar_wait(&k_empty[ks], kph ^ 1u);
      if (elect_one()) {
        mbar_expect_tx(&k_full[ks], Cfg::TILE_BYTES);
        for (int s = 0; s < SLABS; ++s)
          tma_load_4d(s_k + ks * Cfg::TILE_BYTES + s * Cfg::SLAB_BYTES, &p.k_map, &k_full[ks], s * 64, j * 128, head, b);
      }
      if (++ks == KS) { ks = 0; kph ^= 1u; }
      mbar_wait(&v_empty[vs], vph ^ 1u);
      if (elect_one()) {
        mbar_expect_tx(&v_full[vs], Cfg::TILE_BYTES);
        for (int s = 0; s < SLABS; ++s)
          tma_load_4d(s_v + vs * Cfg::TILE_BYTES + s * Cfg::SLAB_BYTES, &p.v_map, &v_full[vs], s * 64, j * 128, head, b);
      }
      if (++vs == VS) { vs = 0; vph ^= 1u; }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (whole warp, one elected lane issues) =====================
    {
      constexpr uint32_t idesc_qk = make_idesc(128, 128, FP16, false, false);
      constexpr uint32_t idesc_pv = make_idesc(128, 64, FP16, false, true);
      int ks = 0, vs = 0;
      uint32_t kph = 0, vph = 0;

      auto issue_s = [&](int i, int kstage) {
        // S_i = Q_i K^T : K-dim = HD, 16 per instruction, 4 per 64-wide slab
        const uint32_t qa = smem_u32(s_q + i * Cfg::TILE_BYTES);
        const uint32_t ka = smem_u32(s_k + kstage * Cfg::TILE_BYTES);
#pragma unroll
        for (int s = 0; s < SLABS; ++s) {
          const uint64_t qd = make_smem_desc_sw128(qa + s * Cfg::SLAB_BYTES, 16, 1024);
          const uint64_t kd = make_smem_desc_sw128(ka + s * Cfg::SLAB_BYTES, 16, 1024);
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_ss(tmem_base + i * 128, qd + 2u * k, kd + 2u * k, idesc_qk, (s | k) != 0 ? 1u : 0u);
        }
      };
      auto issue_pv = [&](int i, int vstage, bool accumulate) {
        // O_i[:, slab] += P_i V[:, slab] : K-dim = 128 kv rows, 16 per instruction (2 x 8-row groups = 2048 B)
        const uint32_t va = smem_u32(s_v + vstage * Cfg::TILE_BYTES);
        const uint32_t p_t = tmem_base + i * 128 + 64;  // P lives in the upper half of the S row
        const uint32_t o_t = tmem_base + NQ * 128 + i * HD;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
#pragma unroll
          for (int s = 0; s < SLABS; ++s) {
            const uint64_t vd = make_smem_desc_sw128(va + s * Cfg::SLAB_BYTES + k * 2048, Cfg::SLAB_BYTES, 1024);
            umma_ts(o_t + s * 64, p_t + k * 8, vd, idesc_pv, (accumulate || k != 0) ? 1u : 0u);
          }
        }
      };

      // prologue: S_i(0)
      mbar_wait(&k_full[0], 0);
      for (int i = 0; i < NQ; ++i) {
        mbar_wait(&q_full[i], 0);
        tc_fence_after();
        if (elect_one()) {
          issue_s(i, 0);
          umma_commit(&s_full[i]);
        }
      }
      if (elect_one()) umma_commit(&k_empty[0]);
      ks = (KS > 1) ? 1 : 0;
      if (KS == 1) kph ^= 1u;

      for (int j = 0; j < n_blocks; ++j) {
        mbar_wait(&v_full[vs], vph);
        const bool has_next = (j + 1 < n_blocks);
        for (int i = 0; i < NQ; ++i) {
          mbar_wait(&p_full[i], j & 1);
          tc_fence_after();
          if (elect_one()) {
            issue_pv(i, vs, j > 0);
            if (i == NQ - 1) umma_commit(&v_empty[vs]);
          }
          if (has_next) {
            if (i == 0) mbar_wait(&k_full[ks], kph);
            tc_fence_after();
            if (elect_one()) {
              issue_s(i, ks);
              umma_commit(&s_full[i]);
              if (i == NQ - 1) umma_commit(&k_empty[ks]);
            }
          } else if (elect_one()) {
            umma_commit(&o_full[i]);
          }
        }
        if (has_next) {
          if (++ks == KS) { ks = 0; kph ^= 1u; }
        }
        if (++vs == VS) { vs = 0; vph ^= 1u; }
      }
    }
  } else if (warp >= 4) {
    // ===================== softmax warpgroups =====================
    const int i = (warp - 4) >> 2;   // query tile handled by this warpgroup
    const int q = (warp - 4) & 3;    // TMEM lane quarter
    const int row = q * 32 + lane;
    const uint32_t lane_off = static_cast<uint32_t>(q * 32) << 16;
    const uint32_t s_t = tmem_base + lane_off + i * 128;
    const uint32_t o_t = tmem_base + lane_off + NQ * 128 + i * HD;
    const float sc = p.scale_log2;
    float m = 0.f;
    float l0 = 0.f, l1 = 0.f, l2 = 0.f, l3 = 0.f;  // four partial row sums: no serial FADD chain

    for (int j = 0; j < n_blocks; ++j) {
      mbar_wait_warp(&s_full[i], j & 1);
      tc_fence_after();
      const int kv_left = p.sk - j * 128;  // valid columns in this block
      const bool partial = kv_left < 128;

      // Columns 64..127 of the score row (H1) are read from TMEM once and stay in registers for both the max and the
      // exp pass; columns 0..63 (H0) are read twice.  3 TMEM waits per block instead of 8.  P (16-bit, 64 columns)
      // is written over the UPPER half of the S row: H1 is consumed from registers first, so nothing live is clobbered.
      uint32_t a1[32], b1[32], t0[32];
      tmem_ld32(s_t + 64, a1);
      tmem_ld32(s_t + 96, b1);
      tmem_ld32(s_t + 0, t0);
      tmem_wait_ld();
      float mx = -INFINITY;
      if (!partial) {
        // 3-input max (FMNMX3) into four independent accumulators: 48 ALU instructions for 96 scores, chains of 12
        float x0 = -INFINITY, x1 = -INFINITY, x2 = -INFINITY, x3 = -INFINITY;
#pragma unroll
        for (int k = 0; k < 32; k += 8) {
          x0 = fmax3(x0, __uint_as_float(t0[k + 0]), __uint_as_float(t0[k + 1]));
          x1 = fmax3(x1, __uint_as_float(t0[k + 2]), __uint_as_float(t0[k + 3]));
          x2 = fmax3(x2, __uint_as_float(t0[k + 4]), __uint_as_float(t0[k + 5]));
          x3 = fmax3(x3, __uint_as_float(t0[k + 6]), __uint_as_float(t0[k + 7]));
          x0 = fmax3(x0, __uint_as_float(a1[k + 0]), __uint_as_float(a1[k + 1]));
          x1 = fmax3(x1, __uint_as_float(a1[k + 2]), __uint_as_float(a1[k + 3]));
          x2 = fmax3(x2, __uint_as_float(a1[k + 4]), __uint_as_float(a1[k + 5]));
          x3 = fmax3(x3, __uint_as_float(a1[k + 6]), __uint_as_float(a1[k + 7]));
          x0 = fmax3(x0, __uint_as_float(b1[k + 0]), __uint_as_float(b1[k + 1]));
          x1 = fmax3(x1, __uint_as_float(b1[k + 2]), __uint_as_float(b1[k + 3]));
          x2 = fmax3(x2, __uint_as_float(b1[k + 4]), __uint_as_float(b1[k + 5]));
          x3 = fmax3(x3, __uint_as_float(b1[k + 6]), __uint_as_float(b1[k + 7]));
        }
        mx = fmax3(x0, x1, fmaxf(x2, x3));
      } else {
#pragma unroll
        for (int k = 0; k < 32; ++k) {
          if (k < kv_left) mx = fmaxf(mx, __uint_as_float(t0[k]));
          if (64 + k < kv_left) mx = fmaxf(mx, __uint_as_float(a1[k]));
          if (96 + k < kv_left) mx = fmaxf(mx, __uint_as_float(b1[k]));
        }
      }
      tmem_ld32(s_t + 32, t0);
      tmem_wait_ld();
      if (!partial) {
        float x0 = mx, x1 = -INFINITY, x2 = -INFINITY, x3 = -INFINITY;
#pragma unroll
        for (int k = 0; k < 32; k += 8) {
          x0 = fmax3(x0, __uint_as_float(t0[k + 0]), __uint_as_float(t0[k + 1]));
          x1 = fmax3(x1, __uint_as_float(t0[k + 2]), __uint_as_float(t0[k + 3]));
          x2 = fmax3(x2, __uint_as_float(t0[k + 4]), __uint_as_float(t0[k + 5]));
          x3 = fmax3(x3, __uint_as_float(t0[k + 6]), __uint_as_float(t0[k + 7]));
        }
        mx = fmax3(x0, x1, fmaxf(x2, x3));
      } else {
#pragma unroll
        for (int k = 0; k < 32; ++k)
          if (32 + k < kv_left) mx = fmaxf(mx, __uint_as_float(t0[k]));
      }
      mx *= sc;  // sc > 0
      if (j == 0) {
        m = (mx == -INFINITY) ? 0.f : mx;
      } else {
        const bool need = mx > m + 8.0f;
        if (__any_sync(0xffffffffu, need)) {
          const float alpha = need ? ex2_approx(m - mx) : 1.0f;
          if (need) {
            m = mx;
            l0 *= alpha; l1 *= alpha; l2 *= alpha; l3 *= alpha;
          }
#pragma unroll 1
          for (int c = 0; c < HD / 32; ++c) {
            tmem_ld32(o_t + c * 32, t0);
            tmem_wait_ld();
#pragma unroll
            for (int k = 0; k < 32; ++k) t0[k] = __float_as_uint(__uint_as_float(t0[k]) * alpha);
            tmem_st32(o_t + c * 32, t0);
          }
          tmem_wait_st();
        }
      }

      // exp2(S*sc - m): three of four on MUFU, one on the FMA pipes (ex2_poly); pairs packed to 16 bit
      // (the ragged-tail masking is compiled out of the full-block path: as predicated compares/selects it cost two
      // instructions per score in EVERY block, a third of the loop)
      auto exps32 = [&](auto masked, const uint32_t (&src)[32], uint32_t* dst, int col0) {
        if constexpr (decltype(masked)::value) {
#pragma unroll
          for (int k = 0; k < 16; ++k) {
            float p0 = ex2_approx(fmaf(__uint_as_float(src[2 * k]), sc, -m));
            float p1 = ex2_approx(fmaf(__uint_as_float(src[2 * k + 1]), sc, -m));
            if (col0 + 2 * k >= kv_left) p0 = 0.f;
            if (col0 + 2 * k + 1 >= kv_left) p1 = 0.f;
            l0 += p0;
            l1 += p1;
            dst[k] = H::pack(p0, p1);
          }
        } else {
          // packed fp32x2 scale-and-shift; of every 8 score pairs POLY go to the FMA-pipe polynomial and the rest to MUFU;
          // pairwise packed row sums into 4 accumulators
          const float2 sc2 = make_float2(sc, sc), nm2 = make_float2(-m, -m);
#pragma unroll
          for (int k = 0; k < 16; k += 8) {
            float2 x[8], e[8];
#pragma unroll
            for (int j = 0; j < 8; ++j)
              x[j] = __ffma2_rn(make_float2(__uint_as_float(src[2 * (k + j)]), __uint_as_float(src[2 * (k + j) + 1])), sc2, nm2);
#pragma unroll
            for (int j = 0; j < 8; ++j)
              e[j] = ((POLY >= 1 && j == 5) || (POLY >= 2 && j == 2) || (POLY >= 3 && j == 7)) ? ex2_poly_pair(x[j])
                                                                                              : make_float2(fast_ex2(x[j].x), fast_ex2(x[j].y));
            const float2 s01 = __fadd2_rn(e[0], e[1]), s23 = __fadd2_rn(e[2], e[3]), s45 = __fadd2_rn(e[4], e[5]), s67 = __fadd2_rn(e[6], e[7]);
            const float2 sa = __fadd2_rn(s01, s23), sb = __fadd2_rn(s45, s67);
            l0 += sa.x; l1 += sa.y; l2 += sb.x; l3 += sb.y;
#pragma unroll
            for (int j = 0; j < 8; ++j) dst[k + j] = H::pack(e[j].x, e[j].y);
          }
        }
      };
      using Masked = std::integral_constant<bool, true>;
      using Full = std::integral_constant<bool, false>;
      {
        uint32_t pk[32];
        if (partial) {
          exps32(Masked{}, a1, pk, 64);
          exps32(Masked{}, b1, pk + 16, 96);
        } else {
          exps32(Full{}, a1, pk, 64);
          exps32(Full{}, b1, pk + 16, 96);
        }
        tmem_st32(s_t + 64 + 32, pk);  // P columns of kv 64..127
      }
      tmem_ld32(s_t + 0, a1);
      tmem_ld32(s_t + 32, b1);
      tmem_wait_ld();
      {
        uint32_t pk[32];
        if (partial) {
          exps32(Masked{}, a1, pk, 0);
          exps32(Masked{}, b1, pk + 16, 32);
        } else {
          exps32(Full{}, a1, pk, 0);
          exps32(Full{}, b1, pk + 16, 32);
        }
        tmem_st32(s_t + 64, pk);  // P columns of kv 0..63
      }
      tmem_wait_st();
      tc_fence_before();
      mbar_arrive(&p_full[i]);
    }

    // ---- epilogue: O / l -> global
    mbar_wait_warp(&o_full[i], 0);
    tc_fence_after();
    const int qrow = q_row0 + i * 128 + row;
    const bool valid = qrow < p.sq;
    const float inv_l = 1.0f / ((l0 + l1) + (l2 + l3));
    typename H::T* orow = static_cast<typename H::T*>(p.o) + static_cast<long long>(b) * p.o_batch_stride +
                          static_cast<long long>(qrow) * p.o_row_stride + head * HD;
#pragma unroll 1
    for (int c = 0; c < HD / 32; ++c) {
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

// attention_pipe.cu: the head_dim-64, one-query-tile variant with the score tile pipelined in 64-key halves
int init_attention_pipe();
bool attention_pipe_enabled();
int launch_attention64_pipe(const b200_attention_args* a, cudaStream_t st);
// attention64.cu: the head_dim-64 default (software-pipelined softmax, register reallocation, KV split with in-kernel combine)
int init_attention64();
bool attention64_enabled();
bool attention128_v2_enabled();
int launch_attention64(const b200_attention_args* a, cudaStream_t st);
long long attention64_workspace_bytes(long long tiles, int kv_halves);

template <int HD, int NQ, bool FP16, int POLY = 1>
static int attn_set_attr() {
  cudaError_t e = cudaFuncSetAttribute(attention_kernel<HD, NQ, FP16, POLY>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       AttnCfg<HD, NQ>::SMEM_BYTES);
  if (e != cudaSuccess) return set_error(B200_ERR_CUDA, "attention smem attr: %s", cudaGetErrorString(e));
  return 0;
}

int init_attention() {
  int r;
  if ((r = attn_set_attr<64, 1, false>())) return r;
  if ((r = attn_set_attr<64, 1, true>())) return r;
  if ((r = attn_set_attr<64, 2, false>())) return r;
  if ((r = attn_set_attr<64, 2, true>())) return r;
  if ((r = attn_set_attr<128, 1, false>())) return r;
  if ((r = attn_set_attr<128, 1, true>())) return r;
  if ((r = attn_set_attr<128, 2, false>())) return r;
  if ((r = attn_set_attr<128, 2, true>())) return r;
  if ((r = attn_set_attr<128, 2, false, 2>())) return r;
  if ((r = attn_set_attr<128, 2, true, 2>())) return r;
  if ((r = attn_set_attr<128, 2, false, 3>())) return r;
  if ((r = attn_set_attr<128, 2, true, 3>())) return r;
  if ((r = init_attention_pipe())) return r;
  if ((r = init_attention64())) return r;
  return 0;
}

template <int HD, int NQ, bool FP16, int POLY = 1>
static int attn_launch(const AttnParams& prm, int grid, cudaStream_t st) {
  cudaError_t e = launch_pdl(attention_kernel<HD, NQ, FP16, POLY>, dim3(grid), dim3(AttnCfg<HD, NQ>::THREADS), AttnCfg<HD, NQ>::SMEM_BYTES, st, prm);
  if (e != cudaSuccess) return set_error(B200_ERR_CUDA, "attention launch: %s", cudaGetErrorString(e));
  return 0;
}

}  // namespace b200

extern "C" {

int64_t b200_attention_workspace_bytes(int32_t batch, int32_t heads, int32_t sq, int32_t sk, int32_t head_dim) {
  using namespace b200;
  if (head_dim != 64 || batch <= 0 || heads <= 0 || sq <= 0 || sk <= 0 || !attention64_enabled()) return 0;
  return attention64_workspace_bytes(static_cast<long long>(batch) * heads * ((sq + 127) / 128), (sk + 63) / 64);
}

int b200_attention(const b200_attention_args* a, void* stream) {
  using namespace b200;
  B200_CHECK_ARG(a && a->q && a->k && a->v && (a->o || a->o_seg_rows > 0), "attention: null pointer");
  B200_CHECK_ARG(a->head_dim == 64 || a->head_dim == 128, "attention: head_dim %d (64 or 128 supported)", a->head_dim);
  B200_CHECK_ARG(a->batch > 0 && a->heads > 0 && a->sq > 0 && a->sk > 0, "attention: bad shape");
  B200_CHECK_ARG(aligned16(a->q) && aligned16(a->k) && aligned16(a->v) && aligned16(a->o), "attention: alignment");
  B200_CHECK_ARG(a->q_row_stride % 8 == 0 && a->k_row_stride % 8 == 0 && a->v_row_stride % 8 == 0 &&
                     a->o_row_stride % 8 == 0 && a->q_batch_stride % 8 == 0 && a->k_batch_stride % 8 == 0 &&
                     a->v_batch_stride % 8 == 0 && a->o_batch_stride % 8 == 0,
                 "attention: strides must be multiples of 8 elements");
  const int HD = a->head_dim;
  if (HD == 64 && a->nq_override != 2 && attention64_enabled())
    return launch_attention64(a, static_cast<cudaStream_t>(stream));
  if (HD == 128 && a->nq_override == 0 && attention128_v2_enabled())
    return launch_attention64(a, static_cast<cudaStream_t>(stream));
  B200_CHECK_ARG(a->o_seg_rows <= 0, "attention: o_seg (context parallelism) needs the default kernel (attention64.cu)");
  if (HD == 64 && a->nq_override != 2 && attention_pipe_enabled())
    return launch_attention64_pipe(a, static_cast<cudaStream_t>(stream));

  AttnParams prm;
  memset(&prm, 0, sizeof(prm));
  const uint32_t box[4] = {64u, 128u, 1u, 1u};
  auto mk = [&](CUtensorMap* m, const void* base, int rows, long long row_stride, long long batch_stride,
                const char* what) {
    const uint64_t dims[4] = {static_cast<uint64_t>(HD), static_cast<uint64_t>(rows), static_cast<uint64_t>(a->heads),
                              static_cast<uint64_t>(a->batch)};
    const uint64_t str[3] = {static_cast<uint64_t>(row_stride) * 2, static_cast<uint64_t>(HD) * 2,
                             static_cast<uint64_t>(batch_stride > 0 ? batch_stride : row_stride * rows) * 2};
    return make_tensor_map_16b(m, base, 4, dims, str, box, what);
  };
  int r;
  if ((r = mk(&prm.q_map, a->q, a->sq, a->q_row_stride, a->q_batch_stride, "attention Q"))) return r;
  if ((r = mk(&prm.k_map, a->k, a->sk, a->k_row_stride, a->k_batch_stride, "attention K"))) return r;
  if ((r = mk(&prm.v_map, a->v, a->sk, a->v_row_stride, a->v_batch_stride, "attention V"))) return r;
  prm.o = a->o;
  prm.o_row_stride = a->o_row_stride;
  prm.o_batch_stride = a->o_batch_stride;
  prm.batch = a->batch;
  prm.heads = a->heads;
  prm.sq = a->sq;
  prm.sk = a->sk;
  prm.kv_blocks = (a->sk + 127) / 128;
  const float scale = a->scale > 0.f ? a->scale : 1.0f / sqrtf(static_cast<float>(HD));
  prm.scale_log2 = scale * 1.4426950408889634f;

  // head_dim 64: one 128-row query tile per CTA, two CTAs co-resident per SM (96 KB smem, 256 TMEM columns each) so
  // one CTA's MMAs overlap the other's softmax and the tail wave is finer grained; head_dim 128: two query tiles
  // ping-pong inside one CTA (the K/V stages of two CTAs would not fit)
  const long long ctas2 = static_cast<long long>(a->batch) * a->heads * ((a->sq + 255) / 256);
  int nq = (HD == 128 && a->sq > 128 && 2 * ctas2 >= num_sms()) ? 2 : 1;
  if (a->nq_override == 1 || a->nq_override == 2) nq = a->nq_override;
  prm.q_tiles = (a->sq + 128 * nq - 1) / (128 * nq);
  const long long grid_ll = static_cast<long long>(a->batch) * a->heads * prm.q_tiles;
  B200_CHECK_ARG(grid_ll < (1ll << 31), "attention: grid too large");
  const int grid = static_cast<int>(grid_ll);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const bool fp16 = a->dtype == B200_DTYPE_FP16;
#define B200_ATTN(HDv, NQv)                                                     \
  return fp16 ? attn_launch<HDv, NQv, true>(prm, grid, st) : attn_launch<HDv, NQv, false>(prm, grid, st)
  if (HD == 64) {
    if (nq == 2) { B200_ATTN(64, 2); } else { B200_ATTN(64, 1); }
  } else {
    if (nq == 2) {
      static const int poly = getenv("B200_ATTN_POLY128") ? atoi(getenv("B200_ATTN_POLY128")) : 1;  // tuning knob: 1, 2 or 3
      if (poly >= 3) return fp16 ? attn_launch<128, 2, true, 3>(prm, grid, st) : attn_launch<128, 2, false, 3>(prm, grid, st);
      if (poly == 2) return fp16 ? attn_launch<128, 2, true, 2>(prm, grid, st) : attn_launch<128, 2, false, 2>(prm, grid, st);
      B200_ATTN(128, 2);
    } else {
      B200_ATTN(128, 1);
    }
  }
#undef B200_ATTN
}

}  // extern "C"
