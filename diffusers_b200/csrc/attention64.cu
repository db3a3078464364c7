// head_dim-64 attention (SDXL self / cross attention): the default for head_dim 64 with one query tile per CTA
// (B200_ATTN_V2=0 falls back to the kernel in attention.cu).
//
// Structure (2 CTAs per SM, 256 threads each):
//   warp 0  TMA producer (Q once, K ring in 64-key halves); warp 3: TMA producer of the V ring
//   warp 1  tcgen05.mma issuer:  S(h) = Q K_h^T  (SS, 128 x 64 x 64)   and   O += P(h) V_h  (TS: P from TMEM, V MN-major)
//   warp 2  TMEM allocator (idle afterwards)                      -> warps 0-3 give their registers away (setmaxnreg.dec)
//   warps 4-7  softmax warpgroup, thread = query row               -> 200 registers per thread (setmaxnreg.inc)
//
//   TMEM (256 columns): S0 [0,64)  S1 [64,128)  P0 [128,160)  P1 [160,192)  O [192,256)
//   MMA warp:      S(0) S(1) | PV(0) S(2) | PV(1) S(3) | ...          (S(h+2) refills the buffer softmax(h) released)
//
// What round 1's pipelined kernel left on the table (ncu: tensor pipe 20-23 % active, the softmax warps - two per SM
// sub-partition - stalled on their own dependency chains: TMEM load -> serial row max -> exponentials -> TMEM store):
//  * SOFTWARE-PIPELINED softmax: while the exponentials of half h occupy the MUFU / FMA pipes, the same thread already holds
//    the scores of half h+1 in a second register set and reduces their row maximum on the ALU pipe.  The two halves are
//    independent instruction streams in one basic block, so the scheduler interleaves them and no pipe waits for a TMEM round
//    trip.  128 score registers + 16 packed P need the register reallocation above.
//  * cheaper arithmetic: 3-input max (FMNMX3), packed fp32x2 FMA / ADD for the scale-and-shift, the row sum and the
//    exponential polynomial, 4 independent partial sums instead of one serial chain; kPolyPairs of every 8 score pairs
//    are evaluated as a degree-3 polynomial on the FMA pipe instead of MUFU (16 ex2 / clk / SM).  The polynomial costs ~6
//    issue slots per element against 1 for MUFU: it only pays once the kernel is MUFU-bound (see kPolyPairs).
//  * S(h+2) is issued as soon as every softmax thread holds S(h) in registers (s_free), i.e. AHEAD of P(h) V(h): with the
//    scores read one step early, waiting for P(h) first left the softmax without scores at the start of every step
//    (measured: 248 us -> 153 us at 4096 tokens).
//  * TAIL SPLIT: with one query tile per CTA and 2 CTAs per SM the SDXL shapes quantise badly - 640 tiles on 296 slots are
//    2.16 waves (4096 tokens), 320 tiles 1.08 waves (1024 tokens) - and the last, nearly empty wave costs as much as a full
//    one.  The tiles of that last wave are therefore split along the keys over `split` CTAs each (so that they fill the
//    machine once more, for 1/split of a tile's time); every part parks its unnormalised O, m and l in a workspace and the
//    part that arrives last (ticket counter per tile) merges all parts IN INDEX ORDER - the result does not depend on which
//    part that is.  Splitting EVERY tile was measured earlier and does not pay (the merge costs ~3-5 us per CTA, more than a
//    fuller last wave returns); splitting the tail pays that once.
// Numerics are those of attention.cu: online softmax in the exp2 domain, lazy rescale (the running max only moves on > 2^8
// growth), P rounded to 16 bit for the P V GEMM, fp32 row sums of the unrounded P.
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include <type_traits>

#include "common.cuh"
#include "host_common.h"

namespace b200 {

struct Attn64Params {
  CUtensorMap q_map;  // box {64, 128, 1, 1}
  CUtensorMap k_map;  // box {64, 64, 1, 1}
  CUtensorMap v_map;  // box {64, 64, 1, 1}
  void* o;
  long long o_row_stride, o_batch_stride;
  int batch, heads, sq, sk;
  int q_tiles;           // ceil(sq / 128)
  int kv_halves;         // ceil(sk / 64)
  float scale_log2;
  // tail split: CTAs [0, n_whole) take whole tiles; CTA n_whole + u takes part u % split of tile n_whole + u / split
  int n_whole, split;
  float* ws_part;        // [tail tiles][split][64 * 128 O (column-major: [col][row]) + 128 m + 128 l] fp32
  int* ws_ticket;        // [tail tiles], zero between launches (the merging CTA resets it)
  // context parallelism: output row r lives in o_seg[r / o_seg_rows] (the buffer of the rank that owns it, a peer mapping
  // over NVLink for the other ranks) at row r % o_seg_rows; o_seg_rows == 0 -> plain `o`
  void* o_seg[8];
  int o_seg_rows;
};

// HD = 64 (SDXL) or 128 (Flux).  Operands live in shared memory as HD / 64 "slabs" of 64 columns (128-byte swizzled rows).
// head_dim 128 keeps ONE score buffer (S(h+1) is issued as soon as the softmax holds S(h) in registers, one step ahead instead
// of two) so that S + P0 + P1 + O still fit 256 TMEM columns and two CTAs share an SM.
template <int HD>
struct Attn64Cfg {
  static_assert(HD == 64 || HD == 128, "head_dim");
  static constexpr int SLABS = HD / 64;
  static constexpr int Q_SLAB = 128 * 64 * 2;    // 16 KB
  static constexpr int KV_SLAB = 64 * 64 * 2;    // 8 KB
  static constexpr int Q_BYTES = SLABS * Q_SLAB;
  static constexpr int KV_BYTES = SLABS * KV_SLAB;  // one 64-key half of K or V
  // head_dim 128: three K stages and two V stages are what two CTAs per SM leave room for (115,712 bytes each): K(h+2) is
  // needed half a step after its stage frees, V(h) a step and a half after
  static constexpr int KS = HD == 64 ? 4 : 3, VS = HD == 64 ? 4 : 2;
  static constexpr int LA = HD == 64 ? 2 : 1;    // score buffers = halves S is issued ahead of the softmax
  static constexpr int TMEM_COLS = 256;
  static constexpr int S_COL = 0, P_COL = LA * 64, O_COL = P_COL + 64;
  static_assert(O_COL + HD <= TMEM_COLS, "TMEM budget");
  // (head_dim 128 has no room for alignment slack: the dynamic shared memory is declared 1024-byte aligned and checked)
  static constexpr int ALIGN_SLACK = HD == 64 ? 1024 : 0;
  static constexpr int SMEM_BYTES = Q_BYTES + (KS + VS) * KV_BYTES + ALIGN_SLACK + 256;
  static_assert(2 * (SMEM_BYTES + 1024) <= 233472, "two CTAs per SM");
  static constexpr int THREADS = 256;
  static constexpr int REGS_LOW = 56, REGS_HIGH = 200;  // 128 * (56 + 200) = 32768 = half of the SM's register file
};

// POLY: score pairs (of every 8) whose exponentials go to the FMA-pipe polynomial instead of MUFU (0..3)
// SPLIT: the launch has tail parts (see TAIL SPLIT above); the plain instance keeps the key range compile-time [0, kv_halves)
template <int HD, bool FP16, int POLY, bool SPLIT>
__global__ void __launch_bounds__(Attn64Cfg<HD>::THREADS, 2) attention64_kernel(const __grid_constant__ Attn64Params p) {
  using Cfg = Attn64Cfg<HD>;
  static_assert(!SPLIT || HD == 64, "the tail split merges 64 output columns per thread");
  constexpr int SLABS = Cfg::SLABS, LA = Cfg::LA;
  using H = Half16<FP16>;
  constexpr int KS = Cfg::KS, VS = Cfg::VS;

  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  if (Cfg::ALIGN_SLACK == 0 && smem != smem_raw) __trap();  // the swizzled tiles need 1024-byte alignment
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
  uint64_t* s_free = pv_done + 2;        // [2]  every softmax thread holds S(h) in registers: the buffer can take S(h+2): 128 arrivals
  uint64_t* o_full = s_free + 2;         // [1]
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(o_full + 1);
  int* ticket_s = reinterpret_cast<int*>(tmem_ptr + 1);  // tail split: this part's arrival number, broadcast to the softmax warps
  static_assert((1 + 2 * KS + 2 * VS + 2 + 2 + 2 + 2 + 1) * 8 + 8 <= 256, "barrier area");

  pdl_trigger();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  int tile = blockIdx.x, part = 0;
  const bool is_part = SPLIT && tile >= p.n_whole;
  if (is_part) {
    const int u = tile - p.n_whole;
    tile = p.n_whole + u / p.split;
    part = u - (u / p.split) * p.split;
  }
  const int qt = tile % p.q_tiles;
  const int bh = tile / p.q_tiles;
  const int head = bh % p.heads;
  const int b = bh / p.heads;
  const int q_row0 = qt * 128;
  const int h_begin = (SPLIT && is_part) ? (part * p.kv_halves) / p.split : 0;
  const int h_end = (SPLIT && is_part) ? ((part + 1) * p.kv_halves) / p.split : p.kv_halves;
  const int n_half = h_end - h_begin;

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
      mbar_init(&p_full[i], 128);
      mbar_init(&pv_done[i], 1);
      mbar_init(&s_free[i], 128);
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

  if (warp < 4) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(Cfg::REGS_LOW));
    if (warp == 0) {
      // ===================== TMA producer =====================
      if (elect_one()) {
        mbar_expect_tx(q_full, Cfg::Q_BYTES);
#pragma unroll
        for (int s = 0; s < SLABS; ++s) tma_load_4d(s_q + s * Cfg::Q_SLAB, &p.q_map, q_full, s * 64, q_row0, head, b);
      }
      // K ring here, V ring in warp 3: one in-order loop over both made every K load wait for the V stage before it, i.e. for
      // a P V two halves back - with two stages per ring (head_dim 128) K(h+3) was then requested half a step before S(h+3)
      // needed it and the softmax waited for scores at the start of every step
      int ks = 0;
      uint32_t kph = 0;
      for (int h = h_begin; h < h_end; ++h) {
        mbar_wait(&k_empty[ks], kph ^ 1u);
        if (elect_one()) {
          mbar_expect_tx(&k_full[ks], Cfg::KV_BYTES);
#pragma unroll
          for (int s = 0; s < SLABS; ++s)
            tma_load_4d(s_k + ks * Cfg::KV_BYTES + s * Cfg::KV_SLAB, &p.k_map, &k_full[ks], s * 64, h * 64, head, b);
        }
        if (++ks == KS) { ks = 0; kph ^= 1u; }
      }
    } else if (warp == 3) {
      // ===================== TMA producer of V =====================
      int vs = 0;
      uint32_t vph = 0;
      for (int h = h_begin; h < h_end; ++h) {
        mbar_wait(&v_empty[vs], vph ^ 1u);
        if (elect_one()) {
          mbar_expect_tx(&v_full[vs], Cfg::KV_BYTES);
#pragma unroll
          for (int s = 0; s < SLABS; ++s)
            tma_load_4d(s_v + vs * Cfg::KV_BYTES + s * Cfg::KV_SLAB, &p.v_map, &v_full[vs], s * 64, h * 64, head, b);
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

      auto issue_s = [&](int buf, int kstage) {  // S_buf = Q K_half^T: 128 x 64 x HD, four K steps of 16 per 64-column slab
#pragma unroll
        for (int s = 0; s < SLABS; ++s) {
          const uint64_t qd = make_smem_desc_sw128(qa + s * Cfg::Q_SLAB, 16, 1024);
          const uint64_t kd = make_smem_desc_sw128(smem_u32(s_k + kstage * Cfg::KV_BYTES + s * Cfg::KV_SLAB), 16, 1024);
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_ss(tmem_base + Cfg::S_COL + buf * 64, qd + 2u * k, kd + 2u * k, idesc_qk, (s | k) != 0 ? 1u : 0u);
        }
      };
      auto issue_pv = [&](int buf, int vstage, bool accumulate) {  // O += P_buf V_half: K = 64 keys, four steps of 16
        const uint32_t va = smem_u32(s_v + vstage * Cfg::KV_BYTES);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
#pragma unroll
          for (int s = 0; s < SLABS; ++s) {  // O[:, slab s] += P V[:, slab s] (N = 64 per instruction)
            const uint64_t vd = make_smem_desc_sw128(va + s * Cfg::KV_SLAB + k * 2048, Cfg::KV_SLAB, 1024);  // MN-major: 16 keys = 2 x 8-row groups
            umma_ts(tmem_base + Cfg::O_COL + s * 64, tmem_base + Cfg::P_COL + buf * 32 + k * 8, vd, idesc_pv, (accumulate || k != 0) ? 1u : 0u);
          }
        }
      };

      mbar_wait(q_full, 0);
      // S is always issued TWO halves ahead of the P V it precedes; with one score buffer (head_dim 128) S(h+2) additionally waits
      // until the softmax has read S(h+1) into registers (s_free), which happens in the middle of softmax step h - well before
      // P(h) exists.  (Issuing S(h+1) only after P(h-1) V had been issued left the softmax waiting for the score MMA at the start
      // of every step: 362 us instead of 272 for the Flux shape.)
      auto issue_scores = [&](int hh) {  // S(hh) into buffer hh % LA
        if (hh >= LA) mbar_wait(&s_free[(hh - LA) % LA], ((hh - LA) / LA) & 1u);
        mbar_wait(&k_full[ks], kph);
        tc_fence_after();
        if (elect_one()) {
          issue_s(hh % LA, ks);
          umma_commit(&s_full[hh % LA]);
          umma_commit(&k_empty[ks]);
        }
        if (++ks == KS) { ks = 0; kph ^= 1u; }
      };
      for (int h0 = 0; h0 < 2 && h0 < n_half; ++h0) issue_scores(h0);
      for (int h = 0; h < n_half; ++h) {
        const int buf = h & 1;
        if (h + 2 < n_half) issue_scores(h + 2);
        mbar_wait(&p_full[buf], (h >> 1) & 1u);
        mbar_wait(&v_full[vs], vph);
        tc_fence_after();
        if (elect_one()) {
          issue_pv(buf, vs, h > 0);
          umma_commit(&v_empty[vs]);
          umma_commit(&pv_done[buf]);
        }
        if (++vs == VS) { vs = 0; vph ^= 1u; }
      }
      if (elect_one()) umma_commit(o_full);
    }
  } else {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(Cfg::REGS_HIGH));
    // ===================== softmax warpgroup (thread = query row) =====================
    const int q = warp - 4;
    const int row = q * 32 + lane;
    const uint32_t lane_off = static_cast<uint32_t>(q * 32) << 16;
    const uint32_t o_t = tmem_base + lane_off + Cfg::O_COL;
    const float sc = p.scale_log2;
    float m = 0.f;
    float l0 = 0.f, l1 = 0.f, l2 = 0.f, l3 = 0.f;

    // scores of the last half of the sequence beyond sk: -inf (their exponentials are 0, V rows there are TMA zero fill)
    auto mask_tail = [&](uint32_t (&s)[64], int hg) {
      const int kv_left = p.sk - hg * 64;
      if (kv_left < 64) {
#pragma unroll
        for (int k = 0; k < 64; ++k)
          if (k >= kv_left) s[k] = 0xff800000u;
      }
    };
    auto load_scores = [&](uint32_t (&s)[64], int h) {  // issue only; tmem_wait_ld() before use
      const uint32_t s_t = tmem_base + lane_off + Cfg::S_COL + (h % LA) * 64;
      tmem_ld32_at<0>(s_t, s);
      tmem_ld32_at<32>(s_t + 32, s);
    };
    auto row_max = [&](const uint32_t (&s)[64]) {
      float a0 = -INFINITY, a1 = -INFINITY, a2 = -INFINITY, a3 = -INFINITY;
#pragma unroll
      for (int k = 0; k < 64; k += 8) {
        a0 = fmax3(a0, __uint_as_float(s[k + 0]), __uint_as_float(s[k + 1]));
        a1 = fmax3(a1, __uint_as_float(s[k + 2]), __uint_as_float(s[k + 3]));
        a2 = fmax3(a2, __uint_as_float(s[k + 4]), __uint_as_float(s[k + 5]));
        a3 = fmax3(a3, __uint_as_float(s[k + 6]), __uint_as_float(s[k + 7]));
      }
      return fmax3(a0, a1, fmaxf(a2, a3));
    };
    // 32 scores s[OFF .. OFF+32) -> 16 packed P.  Of every 16 exponentials 10 go to MUFU and 6 (three packed pairs) to the
    // FMA-pipe polynomial: MUFU 40 x 8 cycles, FMA pipe (scale, polynomial, row sums) ~300, ALU (max, clamps, exponent
    // insertion, packing) ~270 cycles per warp and 64-key half - the three pipes finish together.
    auto exps32 = [&](const uint32_t (&s)[64], auto off_tag, uint32_t (&dst)[16]) {
      constexpr int OFF = decltype(off_tag)::value;
      const float2 sc2 = make_float2(sc, sc), nm2 = make_float2(-m, -m);
#pragma unroll
      for (int k = 0; k < 32; k += 16) {
        float2 x[8], e[8];
#pragma unroll
        for (int j = 0; j < 8; ++j)
          x[j] = __ffma2_rn(make_float2(__uint_as_float(s[OFF + k + 2 * j]), __uint_as_float(s[OFF + k + 2 * j + 1])), sc2, nm2);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          if ((POLY >= 1 && j == 5) || (POLY >= 2 && j == 2) || (POLY >= 3 && j == 7))
            e[j] = ex2_poly_pair(x[j]);
          else
            e[j] = make_float2(fast_ex2(x[j].x), fast_ex2(x[j].y));
        }
        const float2 s01 = __fadd2_rn(e[0], e[1]), s23 = __fadd2_rn(e[2], e[3]), s45 = __fadd2_rn(e[4], e[5]), s67 = __fadd2_rn(e[6], e[7]);
        const float2 sa_ = __fadd2_rn(s01, s23), sb_ = __fadd2_rn(s45, s67);
        l0 += sa_.x; l1 += sa_.y; l2 += sb_.x; l3 += sb_.y;
#pragma unroll
        for (int j = 0; j < 8; ++j) dst[k / 2 + j] = H::pack(e[j].x, e[j].y);
      }
    };
    using Off0 = std::integral_constant<int, 0>;
    using Off32 = std::integral_constant<int, 32>;

    // one softmax step: exponentials of `cur` (local half h) and, interleaved with them, the row max of `nxt` (half h+1).
    //   MODE 0  main loop: there is a next half and it is complete (branch-free: one basic block from the TMEM wait to the
    //           P store, so the scheduler overlaps the two independent streams)
    //   MODE 1  the next half is the last of the sequence and may be ragged (masked)
    //   MODE 2  no next half
    auto step = [&](uint32_t (&cur)[64], uint32_t (&nxt)[64], int h, auto mode_tag) {
      constexpr int MODE = decltype(mode_tag)::value;
      constexpr bool has_next = MODE != 2;
      const int buf = h & 1;
      const uint32_t p_t = tmem_base + lane_off + Cfg::P_COL + buf * 32;
      // With two score buffers S(h+1) was issued a whole step ago and is read first thing; with one buffer (head_dim 128) it was
      // issued in the middle of the previous step behind up to ~1000 cycles of queued MMAs of both CTAs of the SM, so its read
      // comes after the first block of exponentials (which only needs `cur`): the step never starts by waiting for the tensor pipe.
      if (has_next && LA > 1) {
        mbar_wait_warp(&s_full[(h + 1) % LA], ((h + 1) / LA) & 1u);
        tc_fence_after();
        load_scores(nxt, h + 1);
      }
      // P_buf still feeds P V of half h-2 until that MMA completes
      if (h >= 2) mbar_wait_warp(&pv_done[buf], ((h - 2) >> 1) & 1u);
      uint32_t pk[16];
      exps32(cur, Off0{}, pk);
      tmem_st16(p_t, pk);  // keys 0..31 of the half -> P columns 0..15
      float mx = -INFINITY;
      if (has_next) {
        if (LA == 1) {
          mbar_wait_warp(&s_full[0], (h + 1) & 1u);
          tc_fence_after();
          load_scores(nxt, h + 1);
        }
        tmem_wait_ld();
        tc_fence_before();
        mbar_arrive(&s_free[(h + 1) % LA]);  // S(h+1) is in registers: its TMEM buffer may take S(h+1+LA)
        if (MODE == 1) mask_tail(nxt, h_begin + h + 1);
        mx = row_max(nxt);  // independent of the exponentials below
      }
      exps32(cur, Off32{}, pk);
      tmem_st16(p_t + 16, pk);  // keys 32..63 -> P columns 16..31
      tmem_wait_st();
      tc_fence_before();
      mbar_arrive(&p_full[buf]);
      if (has_next) {
        mx *= sc;  // sc > 0
        const bool need = mx > m + 8.0f;
        if (__any_sync(0xffffffffu, need)) {
          // every P V issued so far (halves <= h, in order on the tensor pipe) must have landed before O is rescaled
          mbar_wait_warp(&pv_done[buf], (h >> 1) & 1u);
          tc_fence_after();
          const float alpha = need ? fast_ex2(m - mx) : 1.0f;
          if (need) {
            m = mx;
            l0 *= alpha; l1 *= alpha; l2 *= alpha; l3 *= alpha;
          }
          uint32_t t0[32];
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
    };
    using ModeMain = std::integral_constant<int, 0>;
    using ModeLastNext = std::integral_constant<int, 1>;
    using ModeFinal = std::integral_constant<int, 2>;

    uint32_t sa[64], sb[64];
    mbar_wait_warp(&s_full[0], 0);
    tc_fence_after();
    load_scores(sa, 0);
    tmem_wait_ld();
    tc_fence_before();
    mbar_arrive(&s_free[0]);
    mask_tail(sa, h_begin);
    {
      const float mx = row_max(sa) * sc;
      m = (mx == -INFINITY) ? 0.f : mx;
    }
    // halves 0 .. n_half-3 have a complete successor (main loop, two steps per trip so that the register sets swap roles
    // without moves); half n_half-2 precedes the possibly ragged last half; half n_half-1 has no successor
    int h = 0;
#pragma unroll 1
    for (; h + 3 < n_half; h += 2) {
      step(sa, sb, h, ModeMain{});
      step(sb, sa, h + 1, ModeMain{});
    }
    if (h + 2 < n_half) {  // one more main-loop step: afterwards the current half sits in sb -> move it to sa
      step(sa, sb, h, ModeMain{});
      ++h;
#pragma unroll
      for (int k = 0; k < 64; ++k) sa[k] = sb[k];
    }
    if (h + 1 < n_half) {
      step(sa, sb, h, ModeLastNext{});
      ++h;
    } else {
#pragma unroll
      for (int k = 0; k < 64; ++k) sb[k] = sa[k];
    }
    step(sb, sa, h, ModeFinal{});
    float l = (l0 + l1) + (l2 + l3);

    mbar_wait_warp(o_full, 0);
    tc_fence_after();
    const int qrow = q_row0 + row;
    typename H::T* orow = static_cast<typename H::T*>(p.o) + static_cast<long long>(b) * p.o_batch_stride +
                          static_cast<long long>(qrow) * p.o_row_stride + head * HD;
    if (p.o_seg_rows > 0) {  // the all-to-all back to the row owners, folded into this store
      const int seg = min(qrow / p.o_seg_rows, 7);
      orow = static_cast<typename H::T*>(p.o_seg[seg]) + static_cast<long long>(qrow - seg * p.o_seg_rows) * p.o_row_stride + head * HD;
    }
    bool parked = false;
    if constexpr (SPLIT) if (is_part) {
      parked = true;
      // ---- one part of a split tile: park (O, m, l), take a ticket; the last part to arrive merges all of them
      constexpr int PART_FLOATS = 64 * 128 + 256;
      float* const tile_ws = p.ws_part + static_cast<size_t>(tile - p.n_whole) * p.split * PART_FLOATS;
      float* const mine = tile_ws + static_cast<size_t>(part) * PART_FLOATS;
#pragma unroll 1
      for (int c = 0; c < 2; ++c) {
        uint32_t v[32];
        tmem_ld32(o_t + c * 32, v);
        tmem_wait_ld();
#pragma unroll
        for (int k = 0; k < 32; ++k) mine[(c * 32 + k) * 128 + row] = __uint_as_float(v[k]);  // [col][row]: coalesced over the warp
      }
      mine[64 * 128 + row] = m;
      mine[64 * 128 + 128 + row] = l;
      __threadfence();
      named_bar_sync(1, 128);  // the 128 softmax threads: every row of this part is fenced before the ticket is taken
      if (threadIdx.x == 128) *ticket_s = atomicAdd(&p.ws_ticket[tile - p.n_whole], 1);
      named_bar_sync(1, 128);
      if (*ticket_s == p.split - 1) {
        __threadfence();
        // merge in part order (NOT arrival order): out = sum_i 2^(m_i - M) O_i / sum_i 2^(m_i - M) l_i, M = max_i m_i
        float M = -INFINITY;
        for (int i = 0; i < p.split; ++i) M = fmaxf(M, __ldcg(tile_ws + static_cast<size_t>(i) * PART_FLOATS + 64 * 128 + row));
        float acc[64];
#pragma unroll
        for (int k = 0; k < 64; ++k) acc[k] = 0.f;
        float L = 0.f;
#pragma unroll 1
        for (int i = 0; i < p.split; ++i) {
          const float* src = tile_ws + static_cast<size_t>(i) * PART_FLOATS;
          float t[64];
#pragma unroll
          for (int k = 0; k < 64; ++k) t[k] = __ldcg(src + k * 128 + row);  // all 64 loads in flight: one L2 round trip per part
          const float wgt = fast_ex2(__ldcg(src + 64 * 128 + row) - M);
          L = fmaf(wgt, __ldcg(src + 64 * 128 + 128 + row), L);
#pragma unroll
          for (int k = 0; k < 64; ++k) acc[k] = fmaf(wgt, t[k], acc[k]);
        }
        if (qrow < p.sq) {
          const float inv_l = 1.0f / L;
#pragma unroll
          for (int g = 0; g < 8; ++g) {
            uint4 o;
            o.x = H::pack(acc[g * 8 + 0] * inv_l, acc[g * 8 + 1] * inv_l);
            o.y = H::pack(acc[g * 8 + 2] * inv_l, acc[g * 8 + 3] * inv_l);
            o.z = H::pack(acc[g * 8 + 4] * inv_l, acc[g * 8 + 5] * inv_l);
            o.w = H::pack(acc[g * 8 + 6] * inv_l, acc[g * 8 + 7] * inv_l);
            *reinterpret_cast<uint4*>(orow + g * 8) = o;
          }
        }
        if (threadIdx.x == 128) p.ws_ticket[tile - p.n_whole] = 0;  // ready for the next launch (stream order)
      }
    }
    if (!parked) {
      // ---- epilogue: O / l -> global
      const bool valid = qrow < p.sq;
      const float inv_l = 1.0f / l;
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
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
  }
}

template <int HD, bool FP16, int POLY, bool SPLIT>
static cudaError_t a64_attr() {
  return cudaFuncSetAttribute(attention64_kernel<HD, FP16, POLY, SPLIT>, cudaFuncAttributeMaxDynamicSharedMemorySize, Attn64Cfg<HD>::SMEM_BYTES);
}

int init_attention64() {
  cudaError_t e = a64_attr<64, false, 0, false>();
  if (e == cudaSuccess) e = a64_attr<64, true, 0, false>();
  if (e == cudaSuccess) e = a64_attr<64, false, 1, false>();
  if (e == cudaSuccess) e = a64_attr<64, true, 1, false>();
  if (e == cudaSuccess) e = a64_attr<64, false, 1, true>();
  if (e == cudaSuccess) e = a64_attr<64, true, 1, true>();
  if (e == cudaSuccess) e = a64_attr<128, false, 1, false>();
  if (e == cudaSuccess) e = a64_attr<128, true, 1, false>();
  if (e != cudaSuccess) return set_error(B200_ERR_CUDA, "attention (head_dim 64) smem attr: %s", cudaGetErrorString(e));
  return 0;
}

bool attention64_enabled() {
  static const bool on = !(getenv("B200_ATTN_V2") && atoi(getenv("B200_ATTN_V2")) == 0);
  return on;
}
bool attention128_v2_enabled() {  // head_dim 128 through this kernel (B200_ATTN128_V2=0: attention.cu, two query tiles per CTA)
  static const bool on = attention64_enabled() && !(getenv("B200_ATTN128_V2") && atoi(getenv("B200_ATTN128_V2")) == 0);
  return on;
}

// ---- tail split (see the header comment).  slots = CTAs resident at once (2 per SM).
struct Attn64Split {
  int tail;   // tiles of the last, partial wave
  int split;  // parts per tail tile (1 = no split)
};
constexpr int kAttn64TicketBytes = 4096;  // up to 1024 tail tiles (there are < 2 * SMs of them)
constexpr long long kAttn64PartBytes = (64 * 128 + 256) * 4;

static int attention64_slots() {
  static const int slots = [] {
    int dev = 0, sms = 148;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    return 2 * sms;
  }();
  return slots;
}

static Attn64Split attention64_split(long long tiles, int kv_halves) {
  static const bool off = getenv("B200_ATTN_TAIL_SPLIT") && atoi(getenv("B200_ATTN_TAIL_SPLIT")) == 0;  // tuning / test knob
  Attn64Split r{0, 1};
  const int slots = attention64_slots();
  const int tail = static_cast<int>(tiles % slots);
  // measured on B200 (tools/bench_attention.py): 4096 keys (64 halves, 48 tail tiles 6-way) 155.8 -> 141.1 us; 1024 keys (16 halves,
  // 24 tail tiles 8-way) 30.5 -> 31.5 us: a CTA that is alone on its SM runs ~1.7x faster than one of a pair, so the last wave is
  // far cheaper than a full one and only long tiles repay a part's ~4 us of start-up, parking and merging
  static const int min_halves = getenv("B200_ATTN_SPLIT_MIN_HALVES") ? atoi(getenv("B200_ATTN_SPLIT_MIN_HALVES")) : 32;
  if (off || tail == 0 || kv_halves < min_halves || tail * 4 > kAttn64TicketBytes) return r;
  int s = slots / tail;
  if (s > kv_halves / 2) s = kv_halves / 2;  // at least two 64-key halves per part
  if (s > 8) s = 8;
  // a part costs its share of the tile plus ~4 us of start-up, parking and merging (a 64-key step is ~0.75 us): split only when
  // the tail wave gets clearly shorter
  if (s < 2 || 0.75 * kv_halves * (1.0 - 1.0 / s) < 8.0) return r;
  r.tail = tail;
  r.split = s;
  return r;
}

long long attention64_workspace_bytes(long long tiles, int kv_halves) {
  const Attn64Split sp = attention64_split(tiles, kv_halves);
  return sp.split > 1 ? kAttn64TicketBytes + static_cast<long long>(sp.tail) * sp.split * kAttn64PartBytes : 0;
}

// head_dim 64 or 128, one query tile per CTA; arguments already validated by b200_attention
int launch_attention64(const b200_attention_args* a, cudaStream_t st) {
  const int HD = a->head_dim;
  Attn64Params prm;
  memset(&prm, 0, sizeof(prm));
  auto mk = [&](CUtensorMap* m, const void* base, int rows, long long row_stride, long long batch_stride, uint32_t box_rows,
                const char* what) {
    const uint32_t box[4] = {64u, box_rows, 1u, 1u};
    const uint64_t dims[4] = {static_cast<uint64_t>(HD), static_cast<uint64_t>(rows), static_cast<uint64_t>(a->heads), static_cast<uint64_t>(a->batch)};
    const uint64_t str[3] = {static_cast<uint64_t>(row_stride) * 2, static_cast<uint64_t>(HD) * 2,
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
  if (a->o_seg_rows > 0) {
    B200_CHECK_ARG(a->batch == 1, "attention: o_seg needs batch 1");
    const int nseg = (a->sq + a->o_seg_rows - 1) / a->o_seg_rows;
    B200_CHECK_ARG(nseg <= 8, "attention: %d output segments (at most 8)", nseg);
    for (int i = 0; i < nseg; ++i) {
      B200_CHECK_ARG(a->o_seg[i] != nullptr && (reinterpret_cast<uintptr_t>(a->o_seg[i]) & 15u) == 0, "attention: o_seg[%d] null or not 16-byte aligned", i);
      prm.o_seg[i] = a->o_seg[i];
    }
    for (int i = nseg; i < 8; ++i) prm.o_seg[i] = a->o_seg[nseg - 1];
    prm.o_seg_rows = a->o_seg_rows;
  }
  prm.batch = a->batch;
  prm.heads = a->heads;
  prm.sq = a->sq;
  prm.sk = a->sk;
  prm.q_tiles = (a->sq + 127) / 128;
  prm.kv_halves = (a->sk + 63) / 64;
  const float scale = a->scale > 0.f ? a->scale : (HD == 64 ? 0.125f : 0.08838834764831845f);
  prm.scale_log2 = scale * 1.4426950408889634f;
  const long long tiles = static_cast<long long>(a->batch) * a->heads * prm.q_tiles;
  B200_CHECK_ARG(tiles < (1ll << 30), "attention: grid too large");
  const Attn64Split sp = HD == 64 ? attention64_split(tiles, prm.kv_halves) : Attn64Split{0, 1};
  long long grid_ll = tiles;
  prm.n_whole = static_cast<int>(tiles);
  prm.split = 1;
  if (sp.split > 1 && a->workspace != nullptr && a->workspace_bytes >= attention64_workspace_bytes(tiles, prm.kv_halves)) {
    B200_CHECK_ARG((reinterpret_cast<uintptr_t>(a->workspace) & 15u) == 0, "attention: workspace not 16-byte aligned");
    prm.n_whole = static_cast<int>(tiles - sp.tail);
    prm.split = sp.split;
    prm.ws_ticket = static_cast<int*>(a->workspace);
    prm.ws_part = reinterpret_cast<float*>(static_cast<uint8_t*>(a->workspace) + kAttn64TicketBytes);
    grid_ll = prm.n_whole + static_cast<long long>(sp.tail) * sp.split;
  }
  const bool fp16 = a->dtype == B200_DTYPE_FP16;
  // measured on B200 (tools/bench_attention.py, 4096 / 1024 tokens): POLY 0 -> 154.8 / 30.0 us, 1 -> 146.5 / 28.7 us, 3 -> 154.7 / 30.1 us
  static const int poly = getenv("B200_ATTN_POLY") ? atoi(getenv("B200_ATTN_POLY")) : 1;  // tuning knob: 0 or 1
  const dim3 grid(static_cast<unsigned>(grid_ll)), block(Attn64Cfg<64>::THREADS);
  cudaError_t e;
#define B200_A64(D, P, S) (fp16 ? launch_pdl(attention64_kernel<D, true, P, S>, grid, block, Attn64Cfg<D>::SMEM_BYTES, st, prm) \
                                : launch_pdl(attention64_kernel<D, false, P, S>, grid, block, Attn64Cfg<D>::SMEM_BYTES, st, prm))
  if (HD == 128) e = B200_A64(128, 1, false);
  else if (prm.split > 1) e = B200_A64(64, 1, true);
  else if (poly >= 1) e = B200_A64(64, 1, false);
  else e = B200_A64(64, 0, false);
#undef B200_A64
  if (e != cudaSuccess) return set_error(B200_ERR_CUDA, "attention (head_dim %d) launch: %s", HD, cudaGetErrorString(e));
  return 0;
}

}  // namespace b200
