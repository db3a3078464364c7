// b200_conv_gemm: implicit-GEMM convolution / linear layer on tcgen05 tensor cores.
//
//   persistent CTAs (one per SM), 12 warps; two CTAs of a cluster normally work as a PAIR on two consecutive m tiles
//   (cta_group::2: one 256 x BN MMA issued by the leader, each CTA stages its own A rows and HALF of the weight tile):
//     warp 0       TMA producer    : per 64-wide K chunk, one 4-D box of NHWC pixels (the im2col row block, staged
//                                    128B-swizzled in shared memory; halo pixels are zero-filled by TMA out-of-bounds
//                                    handling) + one 2-D box of packed weights
//     warp 1       MMA issuer      : tcgen05.mma 128(256) x BN x 16, fp32 accumulators in TMEM, double-buffered
//                                    across tiles
//     warp 2       TMEM allocator, then L2 prefetch of the NEXT launch's weights
//     warp 3       slab manager    : hands 32-column output slabs to the epilogue, TMA-loading the residual tile
//                                    into them ahead of time
//     warps 4..11  epilogue        : two warps per TMEM lane quarter on alternate 32-column chunks: tcgen05.ld ->
//                                    bias/act/gate/rowvec/residual (or GEGLU) -> swizzled smem slab -> TMA store
//   Warps 0, 1 and 3 run warp-convergent and predicate only the issue on elect.sync (see elect_one in common.cuh).
//
// A-operand tile = 128 output pixels arranged as a bw x bh rectangle (bw*bh = 128) so that one
// TMA box per filter tap fetches exactly the shifted input pixels.  nn.Linear is the 1x1, H = 1
// case (bw = 128, bh = 1).
#include <stdlib.h>
#include <string.h>

#include "common.cuh"
#include "host_common.h"

namespace b200 {

// compact activations for the epilogue (code size is performance-critical there):
//   SiLU      x * sigmoid(x)
//   GELU-tanh 0.5 x (1 + tanh(u)) == x * sigmoid(2u),  u = sqrt(2/pi) (x + 0.044715 x^3)
//   GELU-erf  0.5 x (1 + erf(x / sqrt 2)) with Abramowitz-Stegun 7.1.26 (|erf error| < 1.5e-7, far below 16-bit rounding)
// (ex2 / rcp in their .ftz forms: __expf / __fdividef without -ftz carry range fix-ups around the MUFU instruction)
__device__ __forceinline__ float rcp_fast(float x) {
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ float exp_fast(float x) { return fast_ex2(x * 1.4426950408889634f); }
__device__ __forceinline__ float sigmoid_fast(float t) { return rcp_fast(1.0f + exp_fast(-t)); }
__device__ __forceinline__ float gelu_erf_fast(float x) {
  const float z = fabsf(x) * 0.70710678118654752f;
  const float t = rcp_fast(fmaf(0.3275911f, z, 1.0f));
  float poly = fmaf(t, 1.061405429f, -1.453152027f);
  poly = fmaf(poly, t, 1.421413741f);
  poly = fmaf(poly, t, -0.284496736f);
  poly = fmaf(poly, t, 0.254829592f);
  const float erf_abs = 1.0f - poly * t * exp_fast(-z * z);
  return 0.5f * x * (1.0f + copysignf(erf_abs, x));
}
__device__ __forceinline__ float act_fast(float x, int act) {
  if (act == ACT_GELU_ERF) return gelu_erf_fast(x);
  // QuickGELU (CLIP text encoders, transformers activations.py QuickGELUActivation): x * sigmoid(1.702 x)
  const float a = (act == ACT_SILU) ? 1.0f : (act == ACT_QUICK_GELU ? 1.702f : 1.5957691216057308f);  // 2*sqrt(2/pi)
  const float b = (act == ACT_GELU_TANH) ? 0.07135481627f : 0.0f;                                      // 2*sqrt(2/pi)*0.044715
  return x * sigmoid_fast(x * fmaf(b, x * x, a));
}

struct ConvGemmParams {
  CUtensorMap a_map[8];  // [src][parity]; stride-1 convs only use parity 0
  CUtensorMap w_map;     // box {64, BN / cm}: a CTA of a pair stages its half of the weight tile
  CUtensorMap y_map;     // output, box {32, bw, bh, 1}, SWIZZLE_64B (TMA-store epilogue)
  CUtensorMap r_map;     // residual, same box: TMA-loaded into the output slab ahead of the epilogue (tma_res)
  int batch, Ho, Wo;
  int bw, bh, bw_shift;
  int tiles_w, tiles_h, m_tiles, n_tiles;
  int cm;                // 1, or 2 = CTA pair (cta_group::2 MMA over two consecutive m tiles)
  int m_groups;          // ceil(m_tiles / cm)
  int total_groups;      // n_tiles * m_groups
  // n / d as (n * M) >> 40 with M = 2^40 / d + 1 (exact while n * d < 2^40: tile counts are far below that): the three divisions of
  // decode_tile sit on the start-up path of every role of every launch
  unsigned long long div_m_groups, div_per_img, div_tiles_w;
  int N;  // rows of w
  int num_taps, nsrc;
  int chunks[2];
  int k_chunks;  // num_taps * (chunks[0] + chunks[1])
  int8_t tap_map[9], tap_dh[9], tap_dw[9];
  const void* bias;
  const void* gate;
  const void* rowvec;
  const void* residual;
  void* y;
  int ld_gate, ld_rowvec, ldr, ldy, rows_per_group, act;
  int vec_ok;    // y / residual / gate / rowvec / bias allow 16-byte accesses
  int out_fp32;  // y is float (attention scores of the unfused head_dim-512 path)
  int tma_store; // epilogue stages 32-column slabs in smem and writes them with TMA (needs vec_ok, 16-bit y)
  int tma_res;   // residual tiles are staged by the loader warp (needs tma_store, 16-bit residual, no GEGLU)
  const uint8_t* pf_ptr;  // next launch's weights: pulled into L2 while this launch runs (nullptr = none)
  long long pf_bytes;
  int dbg_mode;  // tuning: 1 = producer stops loading after the first pipeline round, 2 = MMA thread issues no MMAs
  long long* dbg; // optional [grid][32] clock64 timestamps (tuning aid)
  // LayerNorm folded into the consuming GEMM (see b200_conv_gemm_args)
  float2* stats_out;       // producer: [M][stats_parts] (sum, sumsq) of the rounded outputs, one pair per (n tile, epilogue half)
  int stats_parts;         // 2 * n_tiles
  const float2* ln_stats;  // consumer: [M][ln_parts]
  int ln_parts;
  float ln_eps, ln_inv_k;
  // per-head RMSNorm + rotary embedding of the q / k columns of a fused QKV projection (QKR instantiations; see b200_conv_gemm_args)
  int qk_cols, qk_hd;      // columns [0, qk_cols) are q then k heads of qk_hd (64 | 128) columns each
  const void* qk_w;        // 16-bit [2][qk_hd]: RMSNorm weights of q, then of k
  const float* rope_cos;   // fp32 [qk_hd / 2][rope_ld], position-minor (lanes = consecutive rows read consecutive words)
  const float* rope_sin;
  int rope_ld, rope_row0;  // position of output row r = rope_row0 + r
  float qk_eps;
  // context parallelism: output column block j (y_block_cols columns) is stored through y_map (j == 0) / y_map_peer[j - 1]: the
  // buffers of the ranks that own those heads (peer mappings over NVLink); the q / k / v pattern of QKR repeats per block
  CUtensorMap y_map_peer[7];
  int y_block_cols;  // 0 = one output tensor
};

template <int BN, bool PAIR>
struct ConvGemmCfg {
  static constexpr int BM = 128, BK = 64;
  static constexpr int A_BYTES = BM * BK * 2;
  static constexpr int B_BYTES = (PAIR ? BN / 2 : BN) * BK * 2;  // a CTA of a pair holds half of the weight tile
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int SLAB_BYTES = 128 * 32 * 2;  // one 32-column output slab (64 B rows)
  // 6 output slabs = 3 per epilogue half: a half stores chunk i, works on chunk i+1 and the residual of chunk i+2 streams
  // into the third (with 2 per half every chunk waited for the TMA store issued just before it to finish reading its slab:
  // the full store latency on the critical path of every 32 columns)
  static constexpr int NSLAB = 6;
  static constexpr int RAW_STAGES = (232448 - NSLAB * SLAB_BYTES - 1024 - 512) / STAGE_BYTES;
  static constexpr int STAGES = RAW_STAGES > 8 ? 8 : RAW_STAGES;
  // accumulator buffers sit at power-of-two column offsets
  static constexpr int ACC_STRIDE = (BN <= 32) ? 32 : (BN <= 64 ? 64 : (BN <= 128 ? 128 : 256));
  static constexpr int TMEM_COLS = 2 * ACC_STRIDE;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + NSLAB * SLAB_BYTES + 1024 /*align slack*/ + 512 /*barriers*/;
  static_assert(SMEM_BYTES <= 232448 && STAGES >= 3, "shared memory budget");
};

struct TileCoord {
  int img, h0, w0, n0;
};

// group g -> (n block, m group); CTA `rank` of the cluster takes m tile m_group*cm + rank (may be a phantom tile
// past the end: its loads are all out of bounds -> zeros, and it stores nothing)
__device__ __forceinline__ TileCoord decode_tile(const ConvGemmParams& p, int g, int rank, int BN) {
  TileCoord c;
  int n_blk = static_cast<int>((static_cast<unsigned long long>(g) * p.div_m_groups) >> 40);
  int m = (g - n_blk * p.m_groups) * p.cm + rank;
  int per_img = p.tiles_w * p.tiles_h;
  c.img = static_cast<int>((static_cast<unsigned long long>(m) * p.div_per_img) >> 40);
  int r = m - c.img * per_img;
  int th = static_cast<int>((static_cast<unsigned long long>(r) * p.div_tiles_w) >> 40);
  int tw = r - th * p.tiles_w;
  c.h0 = th * p.bh;
  c.w0 = tw * p.bw;
  c.n0 = n_blk * BN;
  return c;
}

// Slow path of the epilogue for shapes the vector path cannot take (kept out of line: it must not bloat the hot loop).
template <bool GEGLU, bool FP16>
__device__ __noinline__ void epilogue_scalar(const uint32_t* v, const uint32_t* gv, const ConvGemmParams& p,
                                             const typename Half16<FP16>::T* bias, const typename Half16<FP16>::T* gate_row,
                                             const typename Half16<FP16>::T* rv_row, const typename Half16<FP16>::T* res_row,
                                             typename Half16<FP16>::T* y_row, long long pix, int yc0, int wcol0, int bn,
                                             int n_limit) {
  using H = Half16<FP16>;
#pragma unroll 1
  for (int j = 0; j < 32; ++j) {
    const int n = yc0 + j;
    if (n >= n_limit) break;
    float x = __uint_as_float(v[j]);
    if (GEGLU) {
      float gg = __uint_as_float(gv[j]);
      if (bias) {
        x += H::to_float(bias[wcol0 + j]);
        gg += H::to_float(bias[wcol0 + bn / 2 + j]);
      }
      x *= (p.act == ACT_GELU_TANH) ? act_fast(gg, ACT_GELU_TANH) : gelu_erf_fast(gg);
    } else {
      if (bias) x += H::to_float(bias[n]);
      if (p.act != ACT_NONE) x = act_fast(x, p.act);
      if (gate_row) x *= H::to_float(gate_row[n]);
      if (rv_row) x += H::to_float(rv_row[n]);
      if (res_row) x += H::to_float(res_row[n]);
    }
    if (p.out_fp32)
      reinterpret_cast<float*>(p.y)[pix * p.ldy + n] = x;
    else
      y_row[n] = H::from_float(x);
  }
}

template <int BN, bool GEGLU, bool FP16, bool PAIR, bool QKR = false>
__global__ void __launch_bounds__(384, 1) conv_gemm_kernel(const __grid_constant__ ConvGemmParams p) {
  using Cfg = ConvGemmCfg<BN, PAIR>;
  using H = Half16<FP16>;
  constexpr int STAGES = Cfg::STAGES;

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* slabs = smem + STAGES * Cfg::STAGE_BYTES;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(slabs + Cfg::NSLAB * Cfg::SLAB_BYTES);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tfull_bar = empty_bar + STAGES;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint64_t* rfull_bar = tempty_bar + 2;          // residual slab landed (loader warp -> epilogue)
  uint64_t* sfree_bar = rfull_bar + Cfg::NSLAB;  // the TMA store that read the slab has drained it (epilogue -> loader)
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(sfree_bar + Cfg::NSLAB);
  static_assert((2 * STAGES + 4 + 2 * Cfg::NSLAB) * 8 + 4 <= 512, "barrier area");

  const long long t_entry = clock64();
  pdl_trigger();
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int cm = p.cm;
  const int rank = cm > 1 ? static_cast<int>(cluster_ctarank()) : 0;
  const int cluster = cm > 1 ? static_cast<int>(cluster_id_x()) : static_cast<int>(blockIdx.x);
  const int n_clusters = cm > 1 ? static_cast<int>(num_clusters_x()) : static_cast<int>(gridDim.x);

  if (warp == 0) {
    // one barrier per lane (the 2 * STAGES + 4 + 2 * NSLAB <= 32 barriers are contiguous): a single lane initialising all of
    // them costs one serialised shared-memory operation per barrier at the head of every launch
    constexpr int NBAR = 2 * STAGES + 4 + 2 * Cfg::NSLAB;
    static_assert(NBAR <= 32, "one barrier per lane");
    if (lane < NBAR) {
      uint32_t count = 1;
      if (lane < STAGES) count = PAIR ? 2 : 1;  // full: pair = one expect_tx arrive from each CTA's producer (the leader's copy is used)
      if (lane >= 2 * STAGES + 2 && lane < 2 * STAGES + 4) count = PAIR ? 16 : 8;  // tempty: 8 epilogue warps; pair: those of both CTAs
      mbar_init(&full_bar[lane], count);
    }
    if (lane == 0) {
      for (int s = 0; s < p.nsrc; ++s) prefetch_tensormap(&p.a_map[s * 4]);
      prefetch_tensormap(&p.w_map);
      if (p.tma_store) prefetch_tensormap(&p.y_map);
      if (p.tma_res) prefetch_tensormap(&p.r_map);
    }
    fence_barrier_init();
  }
  if (warp == 2) {
    if (PAIR) {
      tmem_alloc2(tmem_ptr, Cfg::TMEM_COLS);
      tmem_relinquish2();
    } else {
      tmem_alloc(tmem_ptr, Cfg::TMEM_COLS);
      tmem_relinquish();
    }
  }
  tc_fence_before();
  __syncthreads();
  if (cm > 1) cluster_sync_all();  // peers' barriers are initialised before any remote arrive / multicast write
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  pdl_wait();  // everything above overlapped the previous kernel's tail; no global access before this point
  // (the marks of thread 0 are kept in registers and stored at the very end: a store here makes the compiler peel lane 0 of
  // the producer warp onto its own path, and the timeline then measures that divergence)
  long long* dbg = p.dbg ? p.dbg + static_cast<size_t>(blockIdx.x) * 32 : nullptr;
  const long long t_prologue = clock64();
  long long t_decoded = 0, t_stage_free = 0;

  const int dbg_mode = p.dbg_mode;
  if (warp == 0) {
    // ===================== TMA producer (whole warp, one elected lane issues) =====================
    int stage = 0;
    uint32_t phase = 0;
    const int b_rows = BN / cm;  // rows of the weight tile this CTA fetches (pair: its half)
    for (int g = cluster; g < p.total_groups; g += n_clusters) {
      TileCoord tc = decode_tile(p, g, rank, BN);
      if (g == cluster) t_decoded = clock64();
      int kc = 0;
      for (int tap = 0; tap < p.num_taps; ++tap) {
        const int mp = p.tap_map[tap];
        const int cw = tc.w0 + p.tap_dw[tap];
        const int ch = tc.h0 + p.tap_dh[tap];
        for (int s = 0; s < p.nsrc; ++s) {
          const CUtensorMap* am = &p.a_map[s * 4 + mp];
          for (int cc = 0; cc < p.chunks[s]; ++cc) {
            mbar_wait(&empty_bar[stage], phase ^ 1u);  // the MMAs that read this stage (in both CTAs) have retired
            if (kc == 0 && g == cluster) t_stage_free = clock64();
            uint8_t* sa = smem + stage * Cfg::STAGE_BYTES;
            const bool skip_loads = dbg_mode == 1 && (kc >= STAGES || g != cluster);  // tuning: MMA rate alone
            if (elect_one()) {
              if (skip_loads) {
                if (PAIR) {
                  mbar_arrive_cluster(leader_addr(&full_bar[stage]));
                } else {
                  mbar_arrive(&full_bar[stage]);
                }
              } else if (PAIR) {
                // own 128 rows of A + own half of the weight tile; the bytes are credited to the LEADER's barrier
                const uint32_t lf = leader_addr(&full_bar[stage]);
                mbar_expect_tx_cluster(lf, Cfg::STAGE_BYTES);
                tma_load_4d_2sm(sa, am, lf, cc * 64, cw, ch, tc.img);
                tma_load_2d_2sm(sa + Cfg::A_BYTES, &p.w_map, lf, kc * 64, tc.n0 + rank * b_rows);
              } else {
                mbar_expect_tx(&full_bar[stage], Cfg::STAGE_BYTES);
                tma_load_4d(sa, am, &full_bar[stage], cc * 64, cw, ch, tc.img);
                tma_load_2d(sa + Cfg::A_BYTES, &p.w_map, &full_bar[stage], kc * 64, tc.n0);
              }
              if (dbg && kc == 0 && g == cluster) dbg[2] = clock64();
            }
            ++kc;
            if (++stage == STAGES) {
              stage = 0;
              phase ^= 1u;
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (whole warp, one elected lane issues) =====================
    if (!PAIR || rank == 0) {
      constexpr uint32_t idesc = make_idesc(PAIR ? 256 : 128, BN, FP16, false, false);
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int g = cluster; g < p.total_groups; g += n_clusters, ++it) {
        const int acc = it & 1;
        const uint32_t acc_phase = (it >> 1) & 1u;
        mbar_wait(&tempty_bar[acc], acc_phase ^ 1u);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * Cfg::ACC_STRIDE;
        for (int kc = 0; kc < p.k_chunks; ++kc) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t a_addr = smem_u32(smem + stage * Cfg::STAGE_BYTES);
          const uint64_t a_desc = make_smem_desc_sw128(a_addr, 16, 1024);
          const uint64_t b_desc = make_smem_desc_sw128(a_addr + Cfg::A_BYTES, 16, 1024);
          if (elect_one()) {
            if (dbg && kc == 0 && it == 0) dbg[3] = clock64();
            if (dbg_mode != 2) {  // (2 = tuning: TMA fill rate without operand reads)
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                // +32 B per UMMA_K=16 step inside the 128 B swizzle atom -> +2 in the (addr>>4) field
                if (PAIR)
                  umma_ss2(d_tmem, a_desc + 2u * k, b_desc + 2u * k, idesc, (kc | k) != 0 ? 1u : 0u);
                else
                  umma_ss(d_tmem, a_desc + 2u * k, b_desc + 2u * k, idesc, (kc | k) != 0 ? 1u : 0u);
              }
            }
            if (PAIR)
              umma_commit2_mcast(&empty_bar[stage], 3);  // frees the stage in both CTAs
            else
              umma_commit(&empty_bar[stage]);
          }
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1u;
          }
        }
        if (elect_one()) {
          if (PAIR)
            umma_commit2_mcast(&tfull_bar[acc], 3);  // both CTAs' epilogues read their own 128 accumulator rows
          else
            umma_commit(&tfull_bar[acc]);
          if (dbg) dbg[4] = clock64();
        }
      }
    }
  } else if (warp == 2) {
    // ===================== L2 prefetch of the NEXT launch's weights =====================
    // Weights (5 GB for SDXL) never survive in L2 from one forward to the next, so every launch would start with
    // DRAM-latency-bound weight fetches; the previous launch pulls them into L2 in the background instead.
    if (p.pf_ptr) {
      constexpr long long PIECE = 4096;
      const long long pieces = (p.pf_bytes + PIECE - 1) / PIECE;
      for (long long i = static_cast<long long>(blockIdx.x) * 32 + lane; i < pieces; i += static_cast<long long>(gridDim.x) * 32) {
        const long long off = i * PIECE;
        const long long rem = p.pf_bytes - off;
        l2_prefetch_bulk(p.pf_ptr + off, static_cast<uint32_t>(rem < PIECE ? (rem & ~15ll) : PIECE));
      }
    }
  } else if (warp == 3) {
    // ===================== slab manager (whole warp, one elected lane issues) =====================
    // Hands the 32-column output slabs to the epilogue in launch order.  With a residual operand the residual tile of
    // the chunk is TMA-loaded into the very slab the chunk's output will be written to (same box, same swizzle), up
    // to NSLAB chunks ahead and overlapping the main loop, so the epilogue never issues a global load for it.
    if (p.tma_store) {
      uint32_t k = 0;
      constexpr int OUT_COLS = GEGLU ? BN / 2 : BN;
      const int n_limit = GEGLU ? (p.N >> 1) : p.N;
      const bool load_res = p.tma_res != 0;
      for (int g = cluster; g < p.total_groups; g += n_clusters) {
        TileCoord tc = decode_tile(p, g, rank, BN);
        const int ycol0 = GEGLU ? (tc.n0 >> 1) : tc.n0;
#pragma unroll 1
        for (int c = 0; c < OUT_COLS / 32; ++c, ++k) {
          const int yc0 = ycol0 + c * 32;
          if (yc0 >= n_limit) break;
          const uint32_t s = k % Cfg::NSLAB;
          mbar_wait(&sfree_bar[s], ((k / Cfg::NSLAB) & 1u) ^ 1u);
          if (elect_one()) {
            if (load_res) {
              mbar_expect_tx(&rfull_bar[s], Cfg::SLAB_BYTES);
              tma_load_4d(slabs + s * Cfg::SLAB_BYTES, &p.r_map, &rfull_bar[s], yc0, tc.w0, tc.h0, tc.img);  // phantom tile: zeros
            } else {
              mbar_arrive(&rfull_bar[s]);
            }
          }
        }
      }
    }
  } else if (warp >= 4) {
    // ===================== epilogue: 8 warps =====================
    // A warp may only read the TMEM lane quarter (warp % 4); the two warps of a quarter take alternate 32-column
    // chunks, so every SM sub-partition has two epilogue warps to interleave (one warp alone is issue-latency bound).
    // Code size matters: the chunk loop stays rolled and the common case (bias and/or residual only) has its own
    // lean body, so the hot path stays resident in the instruction cache.
    const int q = warp & 3;
    const int half = (warp - 4) >> 2;
    const int row = q * 32 + lane;
    const int rh = row >> p.bw_shift;
    const int rw = row & (p.bw - 1);
    const typename H::T* bias = static_cast<const typename H::T*>(p.bias);
    const typename H::T* gate = static_cast<const typename H::T*>(p.gate);
    const typename H::T* rowvec = static_cast<const typename H::T*>(p.rowvec);
    const typename H::T* residual = static_cast<const typename H::T*>(p.residual);
    typename H::T* y = static_cast<typename H::T*>(p.y);
    const int n_limit = GEGLU ? (p.N >> 1) : p.N;  // number of y columns
    const bool tma_store = p.tma_store != 0;
    const bool tma_res = !GEGLU && p.tma_res != 0;
    const bool issuer = (q == 0 && lane == 0);  // one per half: issues and tracks that half's TMA stores
    const uint32_t bar_id = 1 + half;
    const int sw = (row >> 1) & 3;  // SWIZZLE_64B: 16-byte unit index ^= bits [7,9) of the byte offset
    const int act = p.act;
    const bool lean = !GEGLU && tma_store && act == ACT_NONE && gate == nullptr && rowvec == nullptr;
    const float2* ln_stats = p.ln_stats;
    float2* stats_out = p.stats_out;
    uint32_t k = 0;           // running chunk count over all tiles of this CTA (slab k % NSLAB, half k & 1)
    uint32_t prev_k = 0, prev2_k = 0;  // issuer: chunks of this half's (up to two) stores still in flight, newest first
    bool have_prev = false, have_prev2 = false;
    int it = 0;
    for (int g = cluster; g < p.total_groups; g += n_clusters, ++it) {
      const int acc = it & 1;
      TileCoord tc = decode_tile(p, g, rank, BN);
      const int oh = tc.h0 + rh, ow = tc.w0 + rw;
      const bool real_tile = tc.img < p.batch;
      const bool valid = real_tile && (oh < p.Ho) && (ow < p.Wo);
      const long long pix = (static_cast<long long>(tc.img) * p.Ho + oh) * p.Wo + ow;
      const long long grp = (valid && (gate != nullptr || rowvec != nullptr)) ? (pix / p.rows_per_group) : 0;
      const typename H::T* gate_row = gate ? gate + grp * p.ld_gate : nullptr;
      const typename H::T* rv_row = rowvec ? rowvec + grp * p.ld_rowvec : nullptr;
      const typename H::T* res_row = (residual && valid && !tma_res) ? residual + pix * p.ldr : nullptr;
      typename H::T* y_row = y + pix * p.ldy;

      constexpr int OUT_COLS = GEGLU ? BN / 2 : BN;
      const int ycol0 = GEGLU ? (tc.n0 >> 1) : tc.n0;

      // folded LayerNorm: this row's rstd from the producer's partial sums (fixed summation order).  All loads of a batch are
      // issued before the first use (one L2 round trip per 16 parts); runs before the accumulator is needed.
      float ln_r = 1.0f;
      if (ln_stats != nullptr && valid) {
        const float4* sp = reinterpret_cast<const float4*>(ln_stats + pix * p.ln_parts);  // ln_parts is even: 16-byte rows
        const int n4 = p.ln_parts >> 1;
        float s0 = 0.f, s1 = 0.f;
        for (int i0 = 0; i0 < n4; i0 += 8) {
          float4 t[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) t[i] = (i0 + i < n4) ? sp[i0 + i] : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            s0 += t[i].x + t[i].z;
            s1 += t[i].y + t[i].w;
          }
        }
        const float mean = s0 * p.ln_inv_k;
        const float var = fmaxf(fmaf(-mean, mean, s1 * p.ln_inv_k), 0.0f);
        ln_r = rsqrtf(var + p.ln_eps);
      }
      float qk_rstd = 1.0f;          // QKR: 1 / rms of this row over the head being written
      float st_s = 0.f, st_q = 0.f;  // producer side of a folded LayerNorm: sums of this thread's ROUNDED outputs of the tile
      if (bias != nullptr && real_tile && lane * 64 < BN && tc.n0 + lane * 64 < p.N) prefetch_l1(bias + tc.n0 + lane * 64);  // the tile's bias: L1 hits in the chunk loop

      mbar_wait_warp(&tfull_bar[acc], (it >> 1) & 1u);
      tc_fence_after();
      if (dbg && issuer && half == 0) dbg[5] = clock64();
      const uint32_t t_row = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc * Cfg::ACC_STRIDE;

      uint32_t v[32];
      bool preloaded = false;
#pragma unroll 1
      for (int c = 0; c < OUT_COLS / 32; ++c, ++k) {
        const int yc0 = ycol0 + c * 32;
        if (yc0 >= n_limit) break;  // tile-uniform
        if ((k & 1u) != static_cast<uint32_t>(half)) continue;
        const bool chunk_vec = p.vec_ok && (yc0 + 32 <= n_limit);
        const int bcol = GEGLU ? (tc.n0 + c * 32) : yc0;
        if (!preloaded) tmem_ld32(t_row + c * 32, v);  // (else issued during the previous chunk's store phase)
        uint32_t gv[32];
        if (GEGLU) tmem_ld32(t_row + BN / 2 + c * 32, gv);
        uint4 bc[4];  // bias of this chunk (same for every row: L1 hits after the first warp)
#pragma unroll
        for (int j = 0; j < 4; ++j)
          bc[j] = (chunk_vec && bias) ? *reinterpret_cast<const uint4*>(bias + bcol + j * 8) : make_uint4(0, 0, 0, 0);
        uint8_t* slab = slabs + (k % Cfg::NSLAB) * Cfg::SLAB_BYTES;
        if (tma_store) {
          if (q == 0 && have_prev2) {
            // the store BEFORE the previous one has long been read out: hand its slab back to the slab manager (the
            // previous store may still be reading its slab: nobody waits for it here)
            if (elect_one()) {
              bulk_wait_group_read<1>();
              mbar_arrive(&sfree_bar[prev2_k % Cfg::NSLAB]);
            }
            have_prev2 = false;
          }
          mbar_wait_warp(&rfull_bar[k % Cfg::NSLAB], (k / Cfg::NSLAB) & 1u);  // slab is ours (and holds the residual if any)
        }
        tmem_wait_ld();
        const bool stamp = dbg && issuer && it == 0 && c < 2;
        if (stamp) dbg[8 + c * 4] = clock64();
        const int ycb = (QKR && p.y_block_cols > 0) ? yc0 % p.y_block_cols : yc0;  // column inside its destination block ([q | k | v] per block)
        if (QKR && ycb < p.qk_cols) {
          // ---- q / k head columns of a fused QKV projection: per-head RMSNorm (weight) + rotary embedding before the store, so
          // that the attention kernel reads finished q / k (transformer_flux.py:102-119 without the extra pass over the buffer).
          // Rounding points are those of the reference's eager ops (and of qk_norm_rope_kernel): the projection is rounded to 16
          // bit, the normalised value twice (x * rstd, then * weight), the rotation is computed in fp32 and rounded once.
          const int cph = p.qk_hd >> 5;       // 32-column chunks per head (2 | 4)
          (void)0;
          const int cin = c & (cph - 1);      // this chunk's index inside its head (heads never straddle tiles: BN % 128 == 0)
          if (cin == half) {
            // first of this warp's chunks of the head: sum of squares over the WHOLE head (the other warp of the quarter owns the
            // alternate chunks and does the same; TMEM reads are cheap, an exchange would couple the two halves)
            float ss = 0.f;
#pragma unroll 1
            for (int cc = 0; cc < cph; ++cc) {
              uint32_t u[32];
              tmem_ld32(t_row + (c - cin + cc) * 32, u);
              const typename H::T* bp = bias + (yc0 - cin * 32 + cc * 32);
              uint4 bb[4];
#pragma unroll
              for (int j = 0; j < 4; ++j) bb[j] = bias ? *reinterpret_cast<const uint4*>(bp + j * 8) : make_uint4(0, 0, 0, 0);
              tmem_wait_ld();
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                const float2 t0 = H::unpack(bb[j].x), t1 = H::unpack(bb[j].y), t2 = H::unpack(bb[j].z), t3 = H::unpack(bb[j].w);
                const float bq[8] = {t0.x, t0.y, t1.x, t1.y, t2.x, t2.y, t3.x, t3.y};
#pragma unroll
                for (int e = 0; e < 8; ++e) {
                  const float q = H::to_float(H::from_float(fmaf(__uint_as_float(u[j * 8 + e]), ln_r, bq[e])));
                  ss = fmaf(q, q, ss);
                }
              }
            }
            qk_rstd = rsqrtf(ss * (p.qk_hd == 128 ? (1.0f / 128.0f) : (1.0f / 64.0f)) + p.qk_eps);
          }
          const int hcol = cin * 32;  // first column of the chunk inside its head
          const typename H::T* nw = static_cast<const typename H::T*>(p.qk_w) + ((ycb - hcol) >= (p.qk_cols >> 1) ? p.qk_hd : 0) + hcol;
          const long long pos = valid ? (p.rope_row0 + pix) : 0;
          const float* cs_p = p.rope_cos + static_cast<long long>(hcol >> 1) * p.rope_ld + pos;
          const float* sn_p = p.rope_sin + static_cast<long long>(hcol >> 1) * p.rope_ld + pos;
#pragma unroll
          for (int j8 = 0; j8 < 4; ++j8) {
            uint4* sp = reinterpret_cast<uint4*>(slab + row * 64 + ((j8 ^ sw) << 4));
            const uint4 w4 = *reinterpret_cast<const uint4*>(nw + j8 * 8);
            const float2 b0 = H::unpack(bc[j8].x), b1 = H::unpack(bc[j8].y), b2 = H::unpack(bc[j8].z), b3 = H::unpack(bc[j8].w);
            const float2 w0 = H::unpack(w4.x), w1 = H::unpack(w4.y), w2 = H::unpack(w4.z), w3 = H::unpack(w4.w);
            const float bq[8] = {b0.x, b0.y, b1.x, b1.y, b2.x, b2.y, b3.x, b3.y};
            const float wq[8] = {w0.x, w0.y, w1.x, w1.y, w2.x, w2.y, w3.x, w3.y};
            float f[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) {
              float x = H::to_float(H::from_float(fmaf(__uint_as_float(v[j8 * 8 + e]), ln_r, bq[e])));
              x = H::to_float(H::from_float(x * qk_rstd));
              f[e] = H::to_float(H::from_float(x * wq[e]));
            }
#pragma unroll
            for (int e = 0; e < 8; e += 2) {
              const float cs = cs_p[static_cast<long long>(j8 * 4 + (e >> 1)) * p.rope_ld];
              const float sn = sn_p[static_cast<long long>(j8 * 4 + (e >> 1)) * p.rope_ld];
              const float re = f[e], im = f[e + 1];
              f[e] = re * cs + (-im) * sn;
              f[e + 1] = im * cs + re * sn;
            }
            uint4 o;
            o.x = H::pack(f[0], f[1]);
            o.y = H::pack(f[2], f[3]);
            o.z = H::pack(f[4], f[5]);
            o.w = H::pack(f[6], f[7]);
            *sp = o;
          }
        } else if (lean) {
#pragma unroll
          for (int j8 = 0; j8 < 4; ++j8) {
            uint4* sp = reinterpret_cast<uint4*>(slab + row * 64 + ((j8 ^ sw) << 4));
            float f[8];
            {  // acc * rstd + bias (rstd = 1 without a folded LayerNorm)
              float2 t0 = H::unpack(bc[j8].x), t1 = H::unpack(bc[j8].y), t2 = H::unpack(bc[j8].z), t3 = H::unpack(bc[j8].w);
              f[0] = fmaf(__uint_as_float(v[j8 * 8 + 0]), ln_r, t0.x); f[1] = fmaf(__uint_as_float(v[j8 * 8 + 1]), ln_r, t0.y);
              f[2] = fmaf(__uint_as_float(v[j8 * 8 + 2]), ln_r, t1.x); f[3] = fmaf(__uint_as_float(v[j8 * 8 + 3]), ln_r, t1.y);
              f[4] = fmaf(__uint_as_float(v[j8 * 8 + 4]), ln_r, t2.x); f[5] = fmaf(__uint_as_float(v[j8 * 8 + 5]), ln_r, t2.y);
              f[6] = fmaf(__uint_as_float(v[j8 * 8 + 6]), ln_r, t3.x); f[7] = fmaf(__uint_as_float(v[j8 * 8 + 7]), ln_r, t3.y);
            }
            if (tma_res) {  // the residual sits where this thread is about to write its output
              const uint4 r4 = *sp;
              float2 t0 = H::unpack(r4.x), t1 = H::unpack(r4.y), t2 = H::unpack(r4.z), t3 = H::unpack(r4.w);
              f[0] += t0.x; f[1] += t0.y; f[2] += t1.x; f[3] += t1.y;
              f[4] += t2.x; f[5] += t2.y; f[6] += t3.x; f[7] += t3.y;
            }
            uint4 o;
            o.x = H::pack(f[0], f[1]);
            o.y = H::pack(f[2], f[3]);
            o.z = H::pack(f[4], f[5]);
            o.w = H::pack(f[6], f[7]);
            *sp = o;
            if (stats_out != nullptr) {
              const float2 r0 = H::unpack(o.x), r1 = H::unpack(o.y), r2 = H::unpack(o.z), r3 = H::unpack(o.w);
              st_s += ((r0.x + r0.y) + (r1.x + r1.y)) + ((r2.x + r2.y) + (r3.x + r3.y));
              st_q = fmaf(r0.x, r0.x, fmaf(r0.y, r0.y, fmaf(r1.x, r1.x, fmaf(r1.y, r1.y, st_q))));
              st_q = fmaf(r2.x, r2.x, fmaf(r2.y, r2.y, fmaf(r3.x, r3.x, fmaf(r3.y, r3.y, st_q))));
            }
          }
        } else if (chunk_vec) {
#pragma unroll
          for (int j8 = 0; j8 < 4; ++j8) {
            const int yc = yc0 + j8 * 8;
            uint4* sp = reinterpret_cast<uint4*>(slab + row * 64 + ((j8 ^ sw) << 4));
            float f[8];
            {  // acc * rstd + bias (rstd = 1 without a folded LayerNorm)
              float2 t0 = H::unpack(bc[j8].x), t1 = H::unpack(bc[j8].y), t2 = H::unpack(bc[j8].z), t3 = H::unpack(bc[j8].w);
              f[0] = fmaf(__uint_as_float(v[j8 * 8 + 0]), ln_r, t0.x); f[1] = fmaf(__uint_as_float(v[j8 * 8 + 1]), ln_r, t0.y);
              f[2] = fmaf(__uint_as_float(v[j8 * 8 + 2]), ln_r, t1.x); f[3] = fmaf(__uint_as_float(v[j8 * 8 + 3]), ln_r, t1.y);
              f[4] = fmaf(__uint_as_float(v[j8 * 8 + 4]), ln_r, t2.x); f[5] = fmaf(__uint_as_float(v[j8 * 8 + 5]), ln_r, t2.y);
              f[6] = fmaf(__uint_as_float(v[j8 * 8 + 6]), ln_r, t3.x); f[7] = fmaf(__uint_as_float(v[j8 * 8 + 7]), ln_r, t3.y);
            }
            if (GEGLU) {
              // packed rows: [n0, n0+BN/2) value, [n0+BN/2, n0+BN) gate
              uint4 bg = bias ? *reinterpret_cast<const uint4*>(bias + bcol + BN / 2 + j8 * 8) : make_uint4(0, 0, 0, 0);
              float2 t0 = H::unpack(bg.x), t1 = H::unpack(bg.y), t2 = H::unpack(bg.z), t3 = H::unpack(bg.w);
              const float gb[8] = {t0.x, t0.y, t1.x, t1.y, t2.x, t2.y, t3.x, t3.y};
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                const float gx = fmaf(__uint_as_float(gv[j8 * 8 + j]), ln_r, gb[j]);
                f[j] *= (act == ACT_GELU_TANH) ? act_fast(gx, ACT_GELU_TANH) : gelu_erf_fast(gx);  // tanh gate: T5 gated-gelu ("gelu_new")
              }
            } else {
              if (act != ACT_NONE) {
#pragma unroll
                for (int j = 0; j < 8; ++j) f[j] = act_fast(f[j], act);
              }
              if (gate_row) {
                uint4 b = *reinterpret_cast<const uint4*>(gate_row + yc);
                float2 t0 = H::unpack(b.x), t1 = H::unpack(b.y), t2 = H::unpack(b.z), t3 = H::unpack(b.w);
                f[0] *= t0.x; f[1] *= t0.y; f[2] *= t1.x; f[3] *= t1.y;
                f[4] *= t2.x; f[5] *= t2.y; f[6] *= t3.x; f[7] *= t3.y;
              }
              if (rv_row) {
                uint4 b = *reinterpret_cast<const uint4*>(rv_row + yc);
                float2 t0 = H::unpack(b.x), t1 = H::unpack(b.y), t2 = H::unpack(b.z), t3 = H::unpack(b.w);
                f[0] += t0.x; f[1] += t0.y; f[2] += t1.x; f[3] += t1.y;
                f[4] += t2.x; f[5] += t2.y; f[6] += t3.x; f[7] += t3.y;
              }
              if (tma_res || res_row) {
                const uint4 r4 = tma_res ? *sp : *reinterpret_cast<const uint4*>(res_row + yc);
                float2 t0 = H::unpack(r4.x), t1 = H::unpack(r4.y), t2 = H::unpack(r4.z), t3 = H::unpack(r4.w);
                f[0] += t0.x; f[1] += t0.y; f[2] += t1.x; f[3] += t1.y;
                f[4] += t2.x; f[5] += t2.y; f[6] += t3.x; f[7] += t3.y;
              }
            }
            if (p.out_fp32) {
              if (valid) {
                float* yf = reinterpret_cast<float*>(p.y) + pix * p.ldy + yc;
                *reinterpret_cast<float4*>(yf) = make_float4(f[0], f[1], f[2], f[3]);
                *reinterpret_cast<float4*>(yf + 4) = make_float4(f[4], f[5], f[6], f[7]);
              }
            } else {
              uint4 o;
              o.x = H::pack(f[0], f[1]);
              o.y = H::pack(f[2], f[3]);
              o.z = H::pack(f[4], f[5]);
              o.w = H::pack(f[6], f[7]);
              if (tma_store)
                *sp = o;
              else if (valid)
                *reinterpret_cast<uint4*>(y_row + yc) = o;
            }
          }
        } else if (valid) {
          // rare shapes (N not a multiple of 32 / unaligned buffers): scalar, element by element
          uint32_t tv[32], tg[32];  // stack copies: only this cold branch touches local memory
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            tv[j] = v[j];
            tg[j] = GEGLU ? gv[j] : 0u;
          }
          epilogue_scalar<GEGLU, FP16>(tv, tg, p, bias, gate_row, rv_row, res_row, y_row, pix, yc0, tc.n0 + c * 32, BN, n_limit);
        }
        if (stamp) dbg[9 + c * 4] = clock64();
        // this warp's next chunk (c + 2: the halves alternate) starts its TMEM read now, under the fence / barrier / store below
        preloaded = false;
        if (!GEGLU && c + 2 < OUT_COLS / 32 && ycol0 + (c + 2) * 32 < n_limit) {
          tmem_ld32(t_row + (c + 2) * 32, v);
          preloaded = true;
        }
        if (tma_store) {
          fence_proxy_async_smem();  // generic-proxy smem writes -> visible to the TMA engine
          named_bar_sync(bar_id, 128);
          if (q == 0) {
            if (elect_one()) {
              if (real_tile) {
                if (p.y_block_cols > 0) {
                  const int dest = yc0 / p.y_block_cols;  // the all-to-all of Ulysses: this chunk belongs to the heads of rank `dest`
                  tma_store_4d(dest == 0 ? &p.y_map : &p.y_map_peer[dest - 1], slab, yc0 - dest * p.y_block_cols, tc.w0, tc.h0, tc.img);
                } else {
                  tma_store_4d(&p.y_map, slab, yc0, tc.w0, tc.h0, tc.img);
                }
              }
              bulk_commit_group();
              if (stamp) dbg[11 + c * 4] = clock64();
            }
            have_prev2 = have_prev;
            prev2_k = prev_k;
            have_prev = true;
            prev_k = k;
          }
        }
      }
      if (stats_out != nullptr && valid) stats_out[pix * p.stats_parts + ((tc.n0 / BN) << 1) + half] = make_float2(st_s, st_q);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (PAIR)
          mbar_arrive_cluster(leader_addr(&tempty_bar[acc]));
        else
          mbar_arrive(&tempty_bar[acc]);
      }
      // (stores still in flight carry over into the next tile: with three slabs per half one is always free for the next
      // tile's first residual chunk, and nobody stalls on a store that has just been issued)
    }
    if (q == 0 && (have_prev || have_prev2)) {  // the stores must have read their slabs before the CTA gives up its shared memory
      if (elect_one()) bulk_wait_group_read<0>();
    }
    if (dbg && issuer && half == 0) dbg[6] = clock64();
  }

  tc_fence_before();
  __syncthreads();
  if (cm > 1) cluster_sync_all();  // nobody exits while a peer can still multicast into it / arrive on its barriers
  if (warp == 2) {
    tc_fence_after();
    if (PAIR)
      tmem_dealloc2(tmem_base, Cfg::TMEM_COLS);
    else
      tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
  }
  if (dbg && threadIdx.x == 0) {
    dbg[0] = t_entry;
    dbg[1] = t_prologue;
    dbg[24] = t_decoded;
    dbg[25] = t_stage_free;
    dbg[7] = clock64();
  }
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
template <int BN, bool GEGLU, bool FP16, bool QKR = false>
static int set_smem_attr() {
  cudaError_t e = cudaFuncSetAttribute(conv_gemm_kernel<BN, GEGLU, FP16, false, QKR>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       ConvGemmCfg<BN, false>::SMEM_BYTES);
  if (e == cudaSuccess)
    e = cudaFuncSetAttribute(conv_gemm_kernel<BN, GEGLU, FP16, true, QKR>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             ConvGemmCfg<BN, true>::SMEM_BYTES);
  if (e == cudaSuccess)
    e = cudaFuncSetAttribute(conv_gemm_kernel<BN, GEGLU, FP16, true, QKR>, cudaFuncAttributeNonPortableClusterSizeAllowed, 0);
  if (e != cudaSuccess) return set_error(B200_ERR_CUDA, "conv_gemm smem attr BN=%d: %s", BN, cudaGetErrorString(e));
  return 0;
}

int init_conv_gemm() {
  int r = 0;
#define B200_SET(BN)                                   \
  if ((r = set_smem_attr<BN, false, false>())) return r; \
  if ((r = set_smem_attr<BN, false, true>())) return r;
  B200_SET(32) B200_SET(64) B200_SET(96) B200_SET(128) B200_SET(160) B200_SET(192) B200_SET(256)
#undef B200_SET
#define B200_SETG(BN)                                 \
  if ((r = set_smem_attr<BN, true, false>())) return r; \
  if ((r = set_smem_attr<BN, true, true>())) return r;
  B200_SETG(64) B200_SETG(128) B200_SETG(256)
#undef B200_SETG
#define B200_SETQ(BN)                                        \
  if ((r = set_smem_attr<BN, false, false, true>())) return r; \
  if ((r = set_smem_attr<BN, false, true, true>())) return r;
  B200_SETQ(128) B200_SETQ(256)
#undef B200_SETQ
  return 0;
}

template <int BN, bool GEGLU, bool FP16, bool PAIR, bool QKR = false>
static int launch_one(const ConvGemmParams& prm, int grid, cudaStream_t st) {
  constexpr int cm = PAIR ? 2 : 1;
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(384);
  cfg.dynamicSmemBytes = ConvGemmCfg<BN, PAIR>::SMEM_BYTES;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = cm;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = pdl_enabled() ? 1 : 0;
  cfg.attrs = attr;
  cfg.numAttrs = 2;
  if (cm > 1) {
    // GPC boundaries can strand SMs for clusters: never launch more clusters than can be co-resident, the
    // persistent tile loop strides by the number of clusters actually launched
    static int max_active = 0;
    if (max_active == 0) {
      int n = 0;
      cudaLaunchConfig_t q = cfg;
      q.gridDim = dim3(num_sms() / cm * cm);
      if (cudaOccupancyMaxActiveClusters(&n, conv_gemm_kernel<BN, GEGLU, FP16, PAIR, QKR>, &q) != cudaSuccess || n <= 0) {
        cudaGetLastError();
        n = num_sms() / cm;
      }
      max_active = n;
    }
    if (grid > max_active * cm) cfg.gridDim = dim3(max_active * cm);
  }
  cudaError_t e = cudaLaunchKernelEx(&cfg, conv_gemm_kernel<BN, GEGLU, FP16, PAIR, QKR>, prm);
  if (e != cudaSuccess) return set_error(B200_ERR_CUDA, "conv_gemm launch (BN=%d cm=%d grid=%d): %s", BN, cm, grid, cudaGetErrorString(e));
  return 0;
}

template <bool GEGLU, bool FP16, bool PAIR>
static int launch_bn(int bn, const ConvGemmParams& prm, int grid, cudaStream_t st) {
  if (!GEGLU && prm.qk_cols > 0) {  // fused RMSNorm + rotary embedding of the q / k columns
    if (bn == 256) return launch_one<256, false, FP16, PAIR, true>(prm, grid, st);
    if (bn == 128) return launch_one<128, false, FP16, PAIR, true>(prm, grid, st);
    return set_error(B200_ERR_UNSUPPORTED, "conv_gemm: qk_rope needs tile_n 128 or 256 (got %d)", bn);
  }
  switch (bn) {
    case 256: return launch_one<256, GEGLU, FP16, PAIR>(prm, grid, st);
    case 128: return launch_one<128, GEGLU, FP16, PAIR>(prm, grid, st);
    case 64: return launch_one<64, GEGLU, FP16, PAIR>(prm, grid, st);
    default: break;
  }
  if (GEGLU) return set_error(B200_ERR_UNSUPPORTED, "geglu needs tile_n in {64, 128, 256} (got %d)", bn);
  switch (bn) {
    case 192: return launch_one<192, false, FP16, PAIR>(prm, grid, st);
    case 160: return launch_one<160, false, FP16, PAIR>(prm, grid, st);
    case 96: return launch_one<96, false, FP16, PAIR>(prm, grid, st);
    case 32: return launch_one<32, false, FP16, PAIR>(prm, grid, st);
    default: return set_error(B200_ERR_INVALID, "conv_gemm: unsupported tile_n %d", bn);
  }
}

// Cost model (SM cycles, measured on B200 with tools/gemm_chunk_rate.py / gemm_timeline.py) used to pick the tile
// width BN and cm (1 = one CTA per tile, 2 = a CTA pair driving one 256 x BN cta_group::2 MMA, each CTA staging half
// of the weight tile).  Per 64-wide K chunk the tensor pipe needs ~2*BN cycles, but the TMA/mbarrier round of a stage
// has a floor of ~400 (pair) / ~450-470 (single) cycles whatever the tile width; the 8-warp epilogue drains two
// 32-column chunks per ~1450 cycles and overlaps the next tile's main loop when a CTA owns several tiles.
static double chunk_cycles(int bn, int cm) {
  if (cm == 2) return bn <= 192 ? 400.0 : 1.96 * bn;
  if (bn <= 160) return 445.0 + 0.18 * bn;
  return bn <= 192 ? 543.0 : 2.31 * bn;
}
static double launch_cost(int bn, int cm, int k_chunks, bool geglu, long long waves) {
  const double main_loop = k_chunks * chunk_cycles(bn, cm) + 900.0;  // + pipeline fill / accumulator hand-over
  const int n_chunks = (geglu ? bn / 2 : bn) / 32;
  const double epi = 700.0 + 1450.0 * ((n_chunks + 1) / 2) * (geglu ? 1.6 : 1.0);
  const double hidden = epi > main_loop ? epi - main_loop : 0.0;  // part of an epilogue the next main loop cannot cover
  return 3000.0 + (cm == 2 ? 900.0 : 0.0) + waves * main_loop + (waves - 1) * hidden + epi;
}

static void pick_config(long long m_tiles, int N, int k_chunks, int geglu, int force_bn, int force_cm_arg, int* bn_out, int* cm_out, bool qkr = false) {
  const int sms = num_sms();
  static const int env_cm = getenv("B200_FORCE_CM") ? atoi(getenv("B200_FORCE_CM")) : 0;  // test knob
  const int force_cm = force_cm_arg ? force_cm_arg : env_cm;
  double best = 1e30;
  int best_bn = 0, best_cm = 1;
  static const int kBn[] = {32, 64, 96, 128, 160, 192, 256};  // ascending: ties go to the narrower tile
  for (int bn : kBn) {
    if (force_bn && bn != force_bn) continue;
    if (geglu && !(bn == 256 || bn == 128 || bn == 64)) continue;
    if (geglu && N % bn != 0) continue;
    if (qkr && !(bn == 256 || bn == 128)) continue;  // heads of 64 / 128 columns must not straddle tiles
    const long long n_tiles = (N + bn - 1) / bn;
    static const bool no_pair = getenv("B200_NO_PAIR") && atoi(getenv("B200_NO_PAIR")) != 0;  // tuning knob
    for (int cm : {1, 2}) {
      if (force_cm ? cm != force_cm : (cm == 2 && no_pair)) continue;
      if ((bn / cm) % 8 != 0) continue;
      if (cm > 1 && m_tiles < cm) continue;
      const long long groups = n_tiles * ((m_tiles + cm - 1) / cm);
      const long long clusters = sms / cm;
      const long long waves = (groups + clusters - 1) / clusters;
      const double cost = launch_cost(bn, cm, k_chunks, geglu != 0, waves);
      if (cost < best * 0.999) {
        best = cost;
        best_bn = bn;
        best_cm = cm;
      }
    }
  }
  *bn_out = best_bn;
  *cm_out = best_cm;
}

static int pick_tile_n(long long M, int N, int geglu) {
  if (geglu) {
    // the packer interleaves value/gate rows per tile, so BN must divide N; fixed per N so that weights packed
    // once serve every M
    for (int bn : {256, 128, 64})
      if (N % bn == 0) return bn;
    return 0;
  }
  int bn, cm;
  pick_config((M + 127) / 128, N, 20, 0, 0, 0, &bn, &cm);
  return bn;
}

}  // namespace b200

extern "C" {

int64_t b200_conv_gemm_packed_k(int32_t ksize, int32_t c0, int32_t c1) {
  return static_cast<int64_t>(ksize) * ksize * (b200::rup(c0, 64) + (c1 > 0 ? b200::rup(c1, 64) : 0));
}

int32_t b200_conv_gemm_pick_tile_n(int64_t M, int32_t N, int32_t geglu) { return b200::pick_tile_n(M, N, geglu); }

/* Partial (sum, sum of squares) pairs per output row that a launch with these arguments writes to row_stats_out:
 * 2 per N tile of the configuration the launch will pick. */
int32_t b200_conv_gemm_row_stats_parts(const b200_conv_gemm_args* a) {
  using namespace b200;
  if (!a || a->N <= 0 || a->batch <= 0 || a->H <= 0 || a->W <= 0 || (a->ksize != 1 && a->ksize != 2 && a->ksize != 3)) return 0;
  const int Ho = (a->stride == 2) ? a->H / 2 : a->H, Wo = (a->stride == 2) ? a->W / 2 : a->W;
  long long best_tiles = -1;
  for (int bw = 128; bw >= 1; bw >>= 1) {
    const long long tiles = static_cast<long long>(cdiv(Wo, bw)) * cdiv(Ho, 128 / bw);
    if (best_tiles < 0 || tiles < best_tiles) best_tiles = tiles;
  }
  const int nsrc = (a->x[1] != nullptr && a->c[1] > 0) ? 2 : 1;
  const int k_chunks = a->ksize * a->ksize * (cdiv(a->c[0], 64) + (nsrc == 2 ? cdiv(a->c[1], 64) : 0));
  int bn = 0, cm = 1;
  pick_config(best_tiles * a->batch, a->N, k_chunks, a->geglu, a->geglu && !a->tile_n ? pick_tile_n(0, a->N, 1) : a->tile_n, a->cluster_m, &bn, &cm);
  return bn ? 2 * cdiv(a->N, bn) : 0;
}

int b200_conv_gemm(const b200_conv_gemm_args* a, void* stream) {
  using namespace b200;
  B200_CHECK_ARG(a != nullptr, "conv_gemm: null args");
  B200_CHECK_ARG(a->x[0] && a->w && (a->y || a->y_block_cols > 0), "conv_gemm: null x/w/y");
  if (a->y_block_cols > 0) {
    B200_CHECK_ARG(a->y_block_cols % 32 == 0 && a->N % a->y_block_cols == 0 && a->N / a->y_block_cols <= 8 && !a->geglu && !a->out_fp32 && a->ksize == 1 &&
                       !a->residual && !a->row_stats_out,
                   "conv_gemm: y_peers splits the N = %d output columns of a plain linear into at most 8 blocks of y_block_cols = %d (a multiple of 32)", a->N,
                   a->y_block_cols);
    for (int j = 0; j < a->N / a->y_block_cols; ++j)
      B200_CHECK_ARG(a->y_peers[j] != nullptr && aligned16(a->y_peers[j]), "conv_gemm: y_peers[%d] null or not 16-byte aligned", j);
  }
  B200_CHECK_ARG(a->ksize == 1 || a->ksize == 3 || (a->ksize == 2 && a->up2x_parity >= 1 && a->up2x_parity <= 4),
                 "conv_gemm: ksize %d (need 1 or 3; 2 only as a parity class of the folded nearest-2x upsample)", a->ksize);
  const bool up2x = a->ksize == 2;
  if (up2x)
    B200_CHECK_ARG(a->stride == 1 && !a->geglu && !a->gate && !a->rowvec && !a->residual && !a->out_fp32 && !a->row_stats_out && !a->ln_stats,
                   "conv_gemm: the folded-upsample form takes bias / act only");
  else
    B200_CHECK_ARG(a->up2x_parity == 0, "conv_gemm: up2x_parity needs ksize 2");
  B200_CHECK_ARG(a->stride == 1 || a->stride == 2, "conv_gemm: stride %d (need 1 or 2)", a->stride);
  B200_CHECK_ARG(!a->pad_after_only || (a->stride == 2 && a->ksize == 3), "conv_gemm: pad_after_only is the stride-2 3x3 case");
  B200_CHECK_ARG(a->batch > 0 && a->H > 0 && a->W > 0 && a->N > 0, "conv_gemm: bad shape");
  B200_CHECK_ARG(a->dtype == B200_DTYPE_BF16 || a->dtype == B200_DTYPE_FP16, "conv_gemm: dtype %d", a->dtype);
  const int nsrc = (a->x[1] != nullptr && a->c[1] > 0) ? 2 : 1;
  for (int s = 0; s < nsrc; ++s) {
    B200_CHECK_ARG(a->c[s] > 0 && a->c[s] % 8 == 0, "conv_gemm: c[%d]=%d must be a positive multiple of 8", s, a->c[s]);
    B200_CHECK_ARG(a->ldx[s] >= a->c[s] && a->ldx[s] % 8 == 0, "conv_gemm: ldx[%d]=%d invalid", s, a->ldx[s]);
    B200_CHECK_ARG(aligned16(a->x[s]), "conv_gemm: x[%d] not 16-byte aligned", s);
  }
  B200_CHECK_ARG(aligned16(a->w), "conv_gemm: w not 16-byte aligned");
  if (a->stride == 2) B200_CHECK_ARG(a->H % 2 == 0 && a->W % 2 == 0, "conv_gemm: stride 2 needs even H, W");
  if (a->geglu) B200_CHECK_ARG(a->N % 64 == 0, "conv_gemm: geglu needs N %% 64 == 0 (N=%d)", a->N);
  if (a->gate || a->rowvec) B200_CHECK_ARG(a->rows_per_group > 0, "conv_gemm: rows_per_group must be > 0");
  B200_CHECK_ARG(a->cluster_m == 0 || a->cluster_m == 1 || a->cluster_m == 2, "conv_gemm: cluster_m %d (0, 1 or 2)", a->cluster_m);
  B200_CHECK_ARG(a->tile_n == 0 || a->tile_n == 32 || a->tile_n == 64 || a->tile_n == 96 || a->tile_n == 128 ||
                     a->tile_n == 160 || a->tile_n == 192 || a->tile_n == 256,
                 "conv_gemm: tile_n %d not in {32,64,96,128,160,192,256}", a->tile_n);

  const int Ho = (a->stride == 1) ? a->H : a->H / 2;
  const int Wo = (a->stride == 1) ? a->W : a->W / 2;

  ConvGemmParams prm;
  memset(&prm, 0, sizeof(prm));

  // ---- output-pixel rectangle (bw x bh = 128) minimising the tile count
  int best_bw = 128;
  long long best_tiles = -1;
  for (int bw = 128; bw >= 1; bw >>= 1) {
    int bh = 128 / bw;
    long long tiles = static_cast<long long>(cdiv(Wo, bw)) * cdiv(Ho, bh);
    if (best_tiles < 0 || tiles < best_tiles) {
      best_tiles = tiles;
      best_bw = bw;
    }
  }
  prm.bw = best_bw;
  prm.bh = 128 / best_bw;
  prm.bw_shift = 0;
  while ((1 << prm.bw_shift) < prm.bw) ++prm.bw_shift;
  prm.batch = a->batch;
  prm.Ho = Ho;
  prm.Wo = Wo;
  prm.tiles_w = cdiv(Wo, prm.bw);
  prm.tiles_h = cdiv(Ho, prm.bh);
  prm.m_tiles = prm.tiles_w * prm.tiles_h * a->batch;
  prm.N = a->N;
  prm.nsrc = nsrc;
  prm.num_taps = a->ksize * a->ksize;
  prm.chunks[0] = cdiv(a->c[0], 64);
  prm.chunks[1] = nsrc == 2 ? cdiv(a->c[1], 64) : 0;
  prm.k_chunks = prm.num_taps * (prm.chunks[0] + prm.chunks[1]);

  int bn = 0, cm = 1;
  const bool qkr = a->qk_cols > 0;
  if (qkr) {
    B200_CHECK_ARG(a->qk_head_dim == 64 || a->qk_head_dim == 128, "conv_gemm: qk_head_dim %d (64 or 128)", a->qk_head_dim);
    B200_CHECK_ARG(a->qk_cols % (2 * a->qk_head_dim) == 0 && a->qk_cols <= a->N && a->N % 64 == 0,
                   "conv_gemm: qk_cols %d must be q heads + k heads of %d columns inside N = %d (N %% 64 == 0)", a->qk_cols, a->qk_head_dim, a->N);
    B200_CHECK_ARG(a->ksize == 1 && a->stride == 1 && a->H == 1 && !a->geglu && !a->gate && !a->rowvec && !a->residual && !a->out_fp32 && !a->row_stats_out &&
                       a->act == B200_ACT_NONE,
                   "conv_gemm: qk_rope is the epilogue of a plain linear projection (bias and a folded LayerNorm only)");
    B200_CHECK_ARG(a->qk_norm_w && a->rope_cos && a->rope_sin && aligned16(a->qk_norm_w), "conv_gemm: qk_rope needs qk_norm_w (16-byte aligned), rope_cos and rope_sin");
    B200_CHECK_ARG(a->rope_ld > 0 && a->rope_row0 >= 0 && a->rope_row0 + a->W <= a->rope_ld, "conv_gemm: rope table of %d positions does not cover rows [%d, %d)",
                   a->rope_ld, a->rope_row0, a->rope_row0 + a->W);
    B200_CHECK_ARG(a->tile_n == 0 || a->tile_n == 128 || a->tile_n == 256, "conv_gemm: qk_rope needs tile_n 128 or 256");
    if (a->y_block_cols > 0)
      B200_CHECK_ARG(a->qk_cols <= a->y_block_cols && a->y_block_cols % 128 == 0, "conv_gemm: qk_cols %d must fit a block of y_block_cols = %d (a multiple of 128)",
                     a->qk_cols, a->y_block_cols);
  }
  pick_config(prm.m_tiles, a->N, prm.k_chunks, a->geglu, a->geglu && !a->tile_n ? pick_tile_n(0, a->N, 1) : a->tile_n, a->cluster_m, &bn, &cm, qkr);
  B200_CHECK_ARG(bn != 0, "conv_gemm: no valid tile configuration (N=%d geglu=%d tile_n=%d)", a->N, a->geglu, a->tile_n);
  if (a->geglu) B200_CHECK_ARG(a->N % bn == 0, "conv_gemm: geglu N=%d not a multiple of tile_n=%d", a->N, bn);
  prm.cm = cm;
  prm.n_tiles = cdiv(a->N, bn);
  prm.m_groups = cdiv(prm.m_tiles, cm);
  prm.total_groups = prm.n_tiles * prm.m_groups;
  B200_CHECK_ARG(static_cast<long long>(prm.total_groups) * prm.m_groups < (1ll << 38) && static_cast<long long>(prm.m_groups) * cm < (1ll << 19),
                 "conv_gemm: too many tiles (%d groups)", prm.total_groups);
  prm.div_m_groups = (1ull << 40) / static_cast<unsigned long long>(prm.m_groups) + 1;
  prm.div_per_img = (1ull << 40) / static_cast<unsigned long long>(prm.tiles_w * prm.tiles_h) + 1;
  prm.div_tiles_w = (1ull << 40) / static_cast<unsigned long long>(prm.tiles_w) + 1;

  // ---- filter taps: which (parity) tensor map and which box shift each tap uses
  int tap = 0;
  for (int r = 0; r < a->ksize; ++r) {
    for (int s = 0; s < a->ksize; ++s, ++tap) {
      if (a->ksize == 1) {
        prm.tap_map[tap] = 0;
        prm.tap_dh[tap] = 0;
        prm.tap_dw[tap] = 0;
      } else if (up2x) {
        // output pixel (2i + ph, 2j + pw) of conv3x3(nearest2x(x)) reads x rows {i + ph - 1, i + ph} and columns {j + pw - 1, j + pw}
        const int ph = (a->up2x_parity - 1) >> 1, pw = (a->up2x_parity - 1) & 1;
        prm.tap_map[tap] = 0;
        prm.tap_dh[tap] = static_cast<int8_t>(ph - 1 + r);
        prm.tap_dw[tap] = static_cast<int8_t>(pw - 1 + s);
      } else if (a->stride == 1) {
        prm.tap_map[tap] = 0;
        prm.tap_dh[tap] = static_cast<int8_t>(r - 1);
        prm.tap_dw[tap] = static_cast<int8_t>(s - 1);
      } else if (a->pad_after_only) {
        // input row 2*ho + r:  r=0 -> even rows of ho, r=1 -> odd rows of ho, r=2 -> even rows of ho+1 (out of bounds at the bottom: zeros)
        const int ph = (r == 1) ? 1 : 0, pw = (s == 1) ? 1 : 0;
        prm.tap_map[tap] = static_cast<int8_t>(ph * 2 + pw);
        prm.tap_dh[tap] = static_cast<int8_t>(r == 2 ? 1 : 0);
        prm.tap_dw[tap] = static_cast<int8_t>(s == 2 ? 1 : 0);
      } else {
        // input row 2*ho + r - 1:  r=0 -> odd rows of ho-1, r=1 -> even rows of ho, r=2 -> odd rows of ho
        const int ph = (r == 1) ? 0 : 1, pw = (s == 1) ? 0 : 1;
        prm.tap_map[tap] = static_cast<int8_t>(ph * 2 + pw);
        prm.tap_dh[tap] = static_cast<int8_t>(r == 0 ? -1 : 0);
        prm.tap_dw[tap] = static_cast<int8_t>(s == 0 ? -1 : 0);
      }
    }
  }

  // ---- tensor maps
  const uint32_t abox[4] = {64u, static_cast<uint32_t>(prm.bw), static_cast<uint32_t>(prm.bh), 1u};
  for (int s = 0; s < nsrc; ++s) {
    const uint64_t ld = static_cast<uint64_t>(a->ldx[s]);
    if (a->stride == 1) {
      const uint64_t dims[4] = {static_cast<uint64_t>(a->c[s]), static_cast<uint64_t>(a->W), static_cast<uint64_t>(a->H),
                                static_cast<uint64_t>(a->batch)};
      const uint64_t str[3] = {ld * 2, ld * 2 * a->W, ld * 2 * a->W * a->H};
      int r = make_tensor_map_16b(&prm.a_map[s * 4], a->x[s], 4, dims, str, abox, "conv_gemm A");
      if (r) return r;
    } else {
      const int npar = (a->ksize == 1) ? 1 : 4;
      for (int par = 0; par < npar; ++par) {
        const int ph = par >> 1, pw = par & 1;
        const uint8_t* base = static_cast<const uint8_t*>(a->x[s]) + (static_cast<uint64_t>(ph) * a->W + pw) * ld * 2;
        const uint64_t dims[4] = {static_cast<uint64_t>(a->c[s]), static_cast<uint64_t>(a->W / 2),
                                  static_cast<uint64_t>(a->H / 2), static_cast<uint64_t>(a->batch)};
        const uint64_t str[3] = {ld * 4, ld * 4 * a->W, ld * 2 * a->W * a->H};
        int r = make_tensor_map_16b(&prm.a_map[s * 4 + par], base, 4, dims, str, abox, "conv_gemm A (stride 2)");
        if (r) return r;
      }
    }
  }
  {
    const uint64_t Kp = static_cast<uint64_t>(b200_conv_gemm_packed_k(a->ksize, a->c[0], nsrc == 2 ? a->c[1] : 0));
    const uint64_t dims[2] = {Kp, static_cast<uint64_t>(a->N)};
    const uint64_t str[1] = {Kp * 2};
    const uint32_t box[2] = {64u, static_cast<uint32_t>(bn / cm)};
    int r = make_tensor_map_16b(&prm.w_map, a->w, 2, dims, str, box, "conv_gemm W");
    if (r) return r;
  }

  prm.bias = a->bias;
  prm.gate = a->gate;
  prm.rowvec = a->rowvec;
  prm.residual = a->residual;
  prm.y = a->y;
  prm.ld_gate = a->ld_gate;
  prm.ld_rowvec = a->ld_rowvec;
  prm.ldr = a->ldr;
  prm.ldy = a->ldy;
  prm.rows_per_group = a->rows_per_group > 0 ? a->rows_per_group : 1;
  prm.act = a->act;
  bool vec = aligned16(a->y) && (a->ldy % 8 == 0);
  if (a->bias) vec = vec && aligned16(a->bias);
  if (a->gate) vec = vec && aligned16(a->gate) && (a->ld_gate % 8 == 0);
  if (a->rowvec) vec = vec && aligned16(a->rowvec) && (a->ld_rowvec % 8 == 0);
  if (a->residual) vec = vec && aligned16(a->residual) && (a->ldr % 8 == 0);
  prm.vec_ok = vec ? 1 : 0;
  prm.out_fp32 = a->out_fp32 ? 1 : 0;
  prm.dbg = static_cast<long long*>(a->debug_timestamps);
  static const int dbg_mode = getenv("B200_GEMM_DEBUG_MODE") ? atoi(getenv("B200_GEMM_DEBUG_MODE")) : 0;  // wrong results!
  prm.dbg_mode = dbg_mode;
  const int n_out = a->geglu ? a->N / 2 : a->N;
  static const bool no_tma_store = getenv("B200_NO_TMA_STORE") && atoi(getenv("B200_NO_TMA_STORE")) != 0;  // tuning knob
  prm.tma_store = (vec && !a->out_fp32 && n_out % 32 == 0 && (!no_tma_store || up2x)) ? 1 : 0;
  if (up2x) B200_CHECK_ARG(prm.tma_store, "conv_gemm: the folded-upsample form needs the TMA-store epilogue (aligned y, N %% 32 == 0)");
  if (prm.tma_store) {
    const uint64_t ldy = static_cast<uint64_t>(a->ldy);
    const uint64_t dims[4] = {static_cast<uint64_t>(n_out), static_cast<uint64_t>(Wo), static_cast<uint64_t>(Ho),
                              static_cast<uint64_t>(a->batch)};
    const uint32_t ybox[4] = {32u, static_cast<uint32_t>(prm.bw), static_cast<uint32_t>(prm.bh), 1u};
    int r;
    if (up2x) {
      // y is the [batch, 2H, 2W, ldy] image; this launch owns the pixels (2i + ph, 2j + pw): a view with doubled pixel strides
      const int ph = (a->up2x_parity - 1) >> 1, pw = (a->up2x_parity - 1) & 1;
      const uint64_t str[3] = {ldy * 2 * 2, ldy * 2 * (2 * Wo) * 2, ldy * 2 * (2 * Wo) * (2 * Ho)};
      const uint8_t* base = static_cast<const uint8_t*>(a->y) + (static_cast<uint64_t>(ph) * (2 * Wo) + pw) * ldy * 2;
      r = make_tensor_map_16b(&prm.y_map, base, 4, dims, str, ybox, "conv_gemm Y (upsampled parity view)", 64);
    } else if (a->y_block_cols > 0) {
      const uint64_t bdims[4] = {static_cast<uint64_t>(a->y_block_cols), dims[1], dims[2], dims[3]};
      const uint64_t str[3] = {ldy * 2, ldy * 2 * Wo, ldy * 2 * Wo * Ho};
      const int nblk = n_out / a->y_block_cols;
      r = 0;
      for (int j = 0; j < nblk && !r; ++j)
        r = make_tensor_map_16b(j == 0 ? &prm.y_map : &prm.y_map_peer[j - 1], a->y_peers[j], 4, bdims, str, ybox, "conv_gemm Y (peer block)", 64);
      prm.y_block_cols = a->y_block_cols;
    } else {
      const uint64_t str[3] = {ldy * 2, ldy * 2 * Wo, ldy * 2 * Wo * Ho};
      r = make_tensor_map_16b(&prm.y_map, a->y, 4, dims, str, ybox, "conv_gemm Y", 64);
    }
    if (r) return r;
  }
  if (a->y_block_cols > 0)
    B200_CHECK_ARG(prm.tma_store, "conv_gemm: y_peers needs the TMA-store epilogue (16-byte aligned buffers, N %% 32 == 0, B200_NO_TMA_STORE unset)");

  prm.tma_res = (prm.tma_store && a->residual && !a->geglu) ? 1 : 0;
  if (prm.tma_res) {
    const uint64_t ldr = static_cast<uint64_t>(a->ldr);
    const uint64_t dims[4] = {static_cast<uint64_t>(n_out), static_cast<uint64_t>(Wo), static_cast<uint64_t>(Ho),
                              static_cast<uint64_t>(a->batch)};
    const uint64_t str[3] = {ldr * 2, ldr * 2 * Wo, ldr * 2 * Wo * Ho};
    const uint32_t rbox[4] = {32u, static_cast<uint32_t>(prm.bw), static_cast<uint32_t>(prm.bh), 1u};
    int r = make_tensor_map_16b(&prm.r_map, a->residual, 4, dims, str, rbox, "conv_gemm residual", 64);
    if (r) return r;
  }
  if (a->row_stats_out) {
    B200_CHECK_ARG(prm.tma_store && !a->geglu && a->act == B200_ACT_NONE && !a->gate && !a->rowvec && !a->out_fp32,
                   "conv_gemm: row_stats_out needs the plain TMA-store epilogue (aligned y, N %% 32 == 0, no act / gate / rowvec / geglu)");
    B200_CHECK_ARG(aligned16(a->row_stats_out), "conv_gemm: row_stats_out not 16-byte aligned");
    prm.stats_out = reinterpret_cast<float2*>(a->row_stats_out);
    prm.stats_parts = 2 * prm.n_tiles;
  }
  if (a->ln_stats) {
    B200_CHECK_ARG(a->ksize == 1 && a->stride == 1 && nsrc == 1, "conv_gemm: folded LayerNorm needs a 1x1 / linear GEMM with one source");
    B200_CHECK_ARG(a->ln_parts > 0 && a->ln_parts % 2 == 0 && a->ln_eps > 0.f, "conv_gemm: folded LayerNorm needs an even ln_parts > 0 and ln_eps > 0");
    B200_CHECK_ARG(a->act == B200_ACT_NONE && !a->gate && !a->rowvec && !a->residual && !a->out_fp32,
                   "conv_gemm: folded LayerNorm excludes act / gate / rowvec / residual / out_fp32");
    B200_CHECK_ARG(prm.tma_store && vec, "conv_gemm: folded LayerNorm needs the vector TMA-store epilogue (aligned y, y columns %% 32 == 0)");
    B200_CHECK_ARG(aligned16(a->ln_stats), "conv_gemm: ln_stats not 16-byte aligned");
    prm.ln_stats = reinterpret_cast<const float2*>(a->ln_stats);
    prm.ln_parts = a->ln_parts;
    prm.ln_eps = a->ln_eps;
    prm.ln_inv_k = 1.0f / static_cast<float>(a->c[0]);
  }
  if (qkr) {
    B200_CHECK_ARG(prm.tma_store && vec, "conv_gemm: qk_rope needs the vector TMA-store epilogue (aligned y / bias)");
    prm.qk_cols = a->qk_cols;
    prm.qk_hd = a->qk_head_dim;
    prm.qk_w = a->qk_norm_w;
    prm.rope_cos = a->rope_cos;
    prm.rope_sin = a->rope_sin;
    prm.rope_ld = a->rope_ld;
    prm.rope_row0 = a->rope_row0;
    prm.qk_eps = a->qk_eps;
  }
  static const bool no_pf = getenv("B200_NO_WEIGHT_PREFETCH") && atoi(getenv("B200_NO_WEIGHT_PREFETCH")) != 0;  // tuning knob
  if (a->prefetch && a->prefetch_bytes > 0 && !no_pf) {
    B200_CHECK_ARG(aligned16(a->prefetch), "conv_gemm: prefetch pointer not 16-byte aligned");
    prm.pf_ptr = static_cast<const uint8_t*>(a->prefetch);
    prm.pf_bytes = a->prefetch_bytes;
  }

  const int clusters_avail = num_sms() / cm;
  const int n_clusters = prm.total_groups < clusters_avail ? prm.total_groups : clusters_avail;
  const int grid = n_clusters * cm;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const bool fp16 = a->dtype == B200_DTYPE_FP16;
  if (cm == 2) {
    if (a->geglu) return fp16 ? launch_bn<true, true, true>(bn, prm, grid, st) : launch_bn<true, false, true>(bn, prm, grid, st);
    return fp16 ? launch_bn<false, true, true>(bn, prm, grid, st) : launch_bn<false, false, true>(bn, prm, grid, st);
  }
  if (a->geglu) return fp16 ? launch_bn<true, true, false>(bn, prm, grid, st) : launch_bn<true, false, false>(bn, prm, grid, st);
  return fp16 ? launch_bn<false, true, false>(bn, prm, grid, st) : launch_bn<false, false, false>(bn, prm, grid, st);
}

}  // extern "C"
