// HBM-bound normalisation kernels (coalesced 16-byte accesses, warp-shuffle / shared reductions):
//   b200_group_norm : GroupNorm(+SiLU) over NHWC, optionally over the channel concat of two tensors
//   b200_layer_norm : LayerNorm over token rows, optional affine and AdaLN modulation
#include <stdlib.h>
#include <string.h>

#include "common.cuh"
#include "host_common.h"

namespace b200 {

// ------------------------------------------------------------------------------------------------
// GroupNorm
// ------------------------------------------------------------------------------------------------
struct GroupNormParams {
  const void* x[2];
  int c[2], ldx[2];
  int C, V, V0;  // total channels, 8-wide vectors per pixel (total / in source 0)
  int batch, hw, groups, cg;
  float eps;
  const void* gamma;
  const void* beta;
  int act;
  void* y;
  int ldy;
  int chunks, ppc;   // stats: chunks per image, pixels per chunk
  int P;             // stats: pixels processed in parallel per CTA
  float* partial;    // [batch][chunks][groups][3]  (count, mean, M2)
  float* stats;      // [batch][groups][2]          (mean, rstd)
  unsigned int* counter;  // [batch], zero between launches
  int apply_ppb;     // apply: pixels per CTA
};

__device__ __forceinline__ void chan_combine(float& na, float& ma, float& M2a, float nb, float mb, float M2b) {
  if (nb == 0.f) return;
  float n = na + nb;
  float d = mb - ma;
  ma += d * (nb / n);
  M2a += M2b + d * d * (na * nb / n);
  na = n;
}

// grid (chunks, batch); block V*P threads.  Thread (pl, cv) owns 8 fixed channels and walks pixels
// chunk_start + pl, + P, ...  All sums of a CTA are taken relative to one representative value per group (the group's
// first channel at the chunk's first pixel), so the fp32 variance does not cancel and the in-CTA reduction is plain
// additions in a fixed order (deterministic).  Chunk partials (count, mean, M2) are merged with Chan's formula by
// the last CTA of the image to finish, 8 lanes per group.
template <bool FP16>
__global__ void group_norm_stats_kernel(const GroupNormParams p) {
  pdl_trigger();
  pdl_wait();
  using H = Half16<FP16>;
  extern __shared__ float2 s_mm[];  // [P][C] (sum, sum of squares) of shifted values per (pixel lane, channel)
  __shared__ float s_cnt[64];       // per pixel lane count (P <= 64)
  __shared__ unsigned int s_last;

  const int n = blockIdx.y, chunk = blockIdx.x;
  const int t = threadIdx.x;
  const int pl = t / p.V, cv = t - pl * p.V;
  const int pix0 = chunk * p.ppc;
  const int pix1 = min(p.hw, pix0 + p.ppc);

  const int src = cv < p.V0 ? 0 : 1;
  const int coff = (src == 0 ? cv : cv - p.V0) * 8;
  const typename H::T* xb = static_cast<const typename H::T*>(p.x[src]) +
                            (static_cast<size_t>(n) * p.hw) * p.ldx[src] + coff;
  const int ld = p.ldx[src];
  auto rep_value = [&](int g) -> float {  // group g's first channel at the chunk's first pixel
    const int c = g * p.cg;
    const int s2 = c < p.c[0] ? 0 : 1;
    const typename H::T* q = static_cast<const typename H::T*>(p.x[s2]) +
                             (static_cast<size_t>(n) * p.hw + pix0) * p.ldx[s2] + (s2 ? c - p.c[0] : c);
    return H::to_float(*q);
  };

  float sh[8], s[8], ss[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) s[j] = ss[j] = sh[j] = 0.f;
  float cnt = 0.f;
  if (pl < p.P) {
    {  // group of channel cv * 8 + j without eight integer divisions: one division, then count up
      int g = (cv * 8) / p.cg, r = cv * 8 - g * p.cg;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        sh[j] = rep_value(g);
        if (++r == p.cg) {
          r = 0;
          ++g;
        }
      }
    }
    const int step = p.P;
    for (int pix = pix0 + pl; pix < pix1; pix += 4 * step) {
      uint4 u[4];
#pragma unroll
      for (int k = 0; k < 4; ++k)  // 4 independent 16-byte loads in flight per thread
        if (pix + k * step < pix1) u[k] = *reinterpret_cast<const uint4*>(xb + static_cast<size_t>(pix + k * step) * ld);
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        if (pix + k * step < pix1) {
          float2 a = H::unpack(u[k].x), b = H::unpack(u[k].y), c = H::unpack(u[k].z), d = H::unpack(u[k].w);
          float v[8] = {a.x, a.y, b.x, b.y, c.x, c.y, d.x, d.y};
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            float e = v[j] - sh[j];
            s[j] += e;
            ss[j] += e * e;
          }
          cnt += 1.f;
        }
      }
    }
    const int cbase = cv * 8;
#pragma unroll
    for (int j = 0; j < 8; ++j) s_mm[pl * p.C + cbase + j] = make_float2(s[j], ss[j]);
    if (cv == 0) s_cnt[pl] = cnt;
  }
  __syncthreads();
  float* part = p.partial + (static_cast<size_t>(n) * p.chunks + chunk) * p.groups * 3;
  // 8 lanes per group add strided subsets of the (pixel lane, channel) sums, then a fixed shuffle tree: deterministic
  for (int g0 = 0; g0 < p.groups; g0 += blockDim.x / 8) {
    const int g = g0 + t / 8, l8 = t & 7;
    float S = 0.f, SS = 0.f;
    if (g < p.groups) {
      const int items = p.P * p.cg;
      for (int i = l8; i < items; i += 8) {
        const int pl2 = i / p.cg, c = g * p.cg + (i - pl2 * p.cg);
        const float2 v = s_mm[pl2 * p.C + c];
        S += v.x;
        SS += v.y;
      }
    }
#pragma unroll
    for (int o = 4; o > 0; o >>= 1) {
      S += __shfl_down_sync(0xffffffffu, S, o, 8);
      SS += __shfl_down_sync(0xffffffffu, SS, o, 8);
    }
    if (g < p.groups && l8 == 0) {
      float npix = 0.f;
      for (int pl2 = 0; pl2 < p.P; ++pl2) npix += s_cnt[pl2];
      const float cntg = npix * static_cast<float>(p.cg);
      const float K = rep_value(g);
      const float m = cntg > 0.f ? S / cntg : 0.f;
      part[g * 3 + 0] = cntg;
      part[g * 3 + 1] = K + m;
      part[g * 3 + 2] = fmaxf(SS - S * m, 0.f);
    }
  }
  __threadfence();
  __syncthreads();
  if (t == 0) {
    unsigned int ticket = atomicAdd(&p.counter[n], 1u);
    s_last = (ticket == static_cast<unsigned int>(p.chunks - 1)) ? 1u : 0u;
  }
  __syncthreads();
  if (s_last) {
    __threadfence();
    // merge the chunk partials of this image: 8 lanes per group take every 8th chunk (fixed order), then a fixed
    // shuffle tree merges the lanes
    const float* pp = p.partial + static_cast<size_t>(n) * p.chunks * p.groups * 3;
    for (int g0 = 0; g0 < p.groups; g0 += blockDim.x / 8) {
      const int g = g0 + t / 8, l8 = t & 7;
      float na = 0.f, ma = 0.f, M2a = 0.f;
      if (g < p.groups) {
        // 8 chunks per batch: 24 independent L2 loads in flight (ld.global.cg: the partials were written by other
        // SMs), then Chan merges in the same fixed chunk order as ever
        for (int ch0 = l8; ch0 < p.chunks; ch0 += 64) {
          float cb[8], mb[8], Mb[8];
#pragma unroll
          for (int u = 0; u < 8; ++u) {
            const int ch = ch0 + u * 8;
            cb[u] = 0.f;
            mb[u] = 0.f;
            Mb[u] = 0.f;
            if (ch < p.chunks) {
              const float* q = pp + (static_cast<size_t>(ch) * p.groups + g) * 3;
              cb[u] = __ldcg(q);
              mb[u] = __ldcg(q + 1);
              Mb[u] = __ldcg(q + 2);
            }
          }
#pragma unroll
          for (int u = 0; u < 8; ++u) chan_combine(na, ma, M2a, cb[u], mb[u], Mb[u]);
        }
      }
#pragma unroll
      for (int o = 4; o > 0; o >>= 1) {
        const float nb = __shfl_down_sync(0xffffffffu, na, o, 8);
        const float mb = __shfl_down_sync(0xffffffffu, ma, o, 8);
        const float Mb = __shfl_down_sync(0xffffffffu, M2a, o, 8);
        if (l8 < o) chan_combine(na, ma, M2a, nb, mb, Mb);
      }
      if (g < p.groups && l8 == 0) {
        const float var = M2a / na;
        p.stats[(static_cast<size_t>(n) * p.groups + g) * 2 + 0] = ma;
        p.stats[(static_cast<size_t>(n) * p.groups + g) * 2 + 1] = rsqrtf(var + p.eps);
      }
    }
    if (t == 0) p.counter[n] = 0u;
  }
}

// grid (apply_chunks, batch); block V*P threads like the statistics pass: thread (pl, cv) owns 8 fixed channels, keeps
// their scale = rstd*gamma and bias = beta - mean*rstd*gamma in registers and walks pixels pl, pl+P, ... of its chunk with
// four 16-byte loads in flight.  y = act(x * scale + bias).
template <bool FP16>
__global__ void __launch_bounds__(512) group_norm_apply_kernel(const GroupNormParams p) {
  pdl_trigger();
  pdl_wait();
  using H = Half16<FP16>;
  const int n = blockIdx.y;
  const int t = threadIdx.x;
  const int pl = t / p.V, cv = t - pl * p.V;
  if (pl >= p.P) return;
  const int pix0 = blockIdx.x * p.apply_ppb;
  const int pix1 = min(p.hw, pix0 + p.apply_ppb);
  const typename H::T* gamma = static_cast<const typename H::T*>(p.gamma);
  const typename H::T* beta = static_cast<const typename H::T*>(p.beta);
  float sc[8], bi[8];
  {
    // this thread's 8 channels: gamma / beta as one 16-byte load each, the group index by counting up from one division (eight
    // runtime divisions cost more than the thread's share of the pixel loop at the small shapes)
    uint4 g4 = make_uint4(0, 0, 0, 0), b4 = make_uint4(0, 0, 0, 0);
    if (gamma) g4 = *reinterpret_cast<const uint4*>(gamma + cv * 8);
    if (beta) b4 = *reinterpret_cast<const uint4*>(beta + cv * 8);
    const float2 ga01 = H::unpack(g4.x), ga23 = H::unpack(g4.y), ga45 = H::unpack(g4.z), ga67 = H::unpack(g4.w);
    const float2 be01 = H::unpack(b4.x), be23 = H::unpack(b4.y), be45 = H::unpack(b4.z), be67 = H::unpack(b4.w);
    const float gav[8] = {ga01.x, ga01.y, ga23.x, ga23.y, ga45.x, ga45.y, ga67.x, ga67.y};
    const float bev[8] = {be01.x, be01.y, be23.x, be23.y, be45.x, be45.y, be67.x, be67.y};
    int g = (cv * 8) / p.cg, r = cv * 8 - g * p.cg;
    const float2* st = reinterpret_cast<const float2*>(p.stats) + static_cast<size_t>(n) * p.groups;
    float2 mr = st[g];  // (mean, rstd)
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float ga = gamma ? gav[j] : 1.f;
      const float be = beta ? bev[j] : 0.f;
      sc[j] = mr.y * ga;
      bi[j] = be - mr.x * mr.y * ga;
      if (++r == p.cg) {
        r = 0;
        ++g;
        if (j < 7) mr = st[g < p.groups ? g : p.groups - 1];
      }
    }
  }
  const int src = cv < p.V0 ? 0 : 1;
  const int ld = p.ldx[src];
  const typename H::T* xb = static_cast<const typename H::T*>(p.x[src]) + (static_cast<size_t>(n) * p.hw) * ld +
                            (src == 0 ? cv : cv - p.V0) * 8;
  typename H::T* yb = static_cast<typename H::T*>(p.y) + (static_cast<size_t>(n) * p.hw) * p.ldy + cv * 8;
  const bool do_silu = p.act == ACT_SILU;
  const int step = p.P;
  for (int pix = pix0 + pl; pix < pix1; pix += 4 * step) {
    uint4 u[4];
#pragma unroll
    for (int k = 0; k < 4; ++k)
      if (pix + k * step < pix1) u[k] = *reinterpret_cast<const uint4*>(xb + static_cast<size_t>(pix + k * step) * ld);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      if (pix + k * step < pix1) {
        float2 a = H::unpack(u[k].x), b = H::unpack(u[k].y), c = H::unpack(u[k].z), d = H::unpack(u[k].w);
        float v[8] = {a.x, a.y, b.x, b.y, c.x, c.y, d.x, d.y};
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float o = fmaf(v[j], sc[j], bi[j]);
          v[j] = do_silu ? silu_fast(o) : o;
        }
        uint4 o;
        o.x = H::pack(v[0], v[1]);
        o.y = H::pack(v[2], v[3]);
        o.z = H::pack(v[4], v[5]);
        o.w = H::pack(v[6], v[7]);
        *reinterpret_cast<uint4*>(yb + static_cast<size_t>(pix + k * step) * p.ldy) = o;
      }
    }
  }
}

static void gn_plan(int batch, int hw, int C, int* chunks, int* ppc) {
  // ~2 CTAs per SM over the whole batch, at least 8 pixels per chunk, at most 256 chunks per image (workspace)
  long long total = static_cast<long long>(batch) * hw;
  int pp = static_cast<int>((total + 2LL * num_sms() - 1) / (2LL * num_sms()));
  if (pp < 8) pp = 8;
  if (pp * 256 < hw) pp = (hw + 255) / 256;
  if (pp > hw) pp = hw;
  *ppc = pp;
  *chunks = (hw + pp - 1) / pp;
  (void)C;
}

// ------------------------------------------------------------------------------------------------
// GroupNorm, single pass over HBM ("slab" kernel).  The statistics kernel above re-reads what the apply kernel reads again:
// three passes over the activation for two algorithmic ones, and two launches.  Here one thread-block CLUSTER owns one
// (sample, block of GB consecutive groups) and its CTAs split the pixels; each CTA stages its slab - pixels x (GB * C/groups)
// channels, <= 200 KB - in shared memory with 16-byte loads, the cluster reduces sum and sum of squared deviations through
// distributed shared memory (every CTA reads the partials of all ranks in rank order: same value everywhere, deterministic),
// and the slab is normalised (+ affine, + SiLU) from shared memory and written once.  Exact two-pass variance (the data sit
// in shared memory), one read + one write of the activation, one launch.
// GB is the smallest count of groups whose channel span is a multiple of 8 (16 bytes); thread t always works on vector
// t % vpp of a pixel, so its 8 channels (and their scale / shift) are fixed.
// ------------------------------------------------------------------------------------------------
struct GroupNormSlabParams {
  const void* x[2];
  int c0, ldx[2];
  int C, hw, groups, cg;
  int gb;        // groups per unit
  int span;      // channels per unit = gb * cg (multiple of 8)
  int vpp;       // 16-byte vectors per pixel of a unit = span / 8
  int units;     // groups / gb
  int ppc;       // pixels per CTA (hw split over the cluster)
  float eps;
  const void* gamma;
  const void* beta;
  int act;
  void* y;
  int ldy;
};

constexpr int kGnSlabThreads = 960;  // divisible by every vpp in use (1, 2, 3, 4, 5, 6, 8, 10, 12, 15, 16, 20)
constexpr int kGnSlabMaxBytes = 196608;

__device__ __forceinline__ uint32_t map_to_rank(uint32_t smem_addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_addr), "r"(rank));
  return r;
}
__device__ __forceinline__ float ld_dsmem_f32(uint32_t cluster_addr) {
  float v;
  asm volatile("ld.shared::cluster.f32 %0, [%1];" : "=f"(v) : "r"(cluster_addr) : "memory");
  return v;
}
__device__ __forceinline__ uint32_t cluster_nctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(r));
  return r;
}

template <bool FP16>
__global__ void __launch_bounds__(kGnSlabThreads, 1) group_norm_slab_kernel(const GroupNormSlabParams p) {
  pdl_trigger();
  using H = Half16<FP16>;
  extern __shared__ uint8_t gn_smem[];
  uint4* slab = reinterpret_cast<uint4*>(gn_smem);                                  // [ppc][vpp] vectors of 8 channels
  float* part = reinterpret_cast<float*>(gn_smem + static_cast<size_t>(p.ppc) * p.vpp * 16);  // [threads / vpp][span] thread partials
  __shared__ float part2[kGnSlabThreads];
  __shared__ float red[2][32];     // this CTA's per-group partial (sum | sum of squared deviations), read by the whole cluster
  __shared__ float g_mean[32], g_rstd[32];

  const uint32_t rank = cluster_ctarank(), csize = cluster_nctarank();
  const int unit = blockIdx.y % p.units, n = blockIdx.y / p.units;
  const int t = threadIdx.x;
  const int v = t % p.vpp, lane_p = t / p.vpp, p_step = kGnSlabThreads / p.vpp;
  const int pix0 = rank * p.ppc;
  const int npix = max(0, min(p.hw, pix0 + p.ppc) - pix0);
  const int ch0 = unit * p.span + v * 8;  // first of this thread's 8 channels
  const int src = ch0 < p.c0 ? 0 : 1;
  const typename H::T* xb = static_cast<const typename H::T*>(p.x[src]) + (static_cast<size_t>(n) * p.hw + pix0) * p.ldx[src] +
                            (src ? ch0 - p.c0 : ch0);
  const int ld = p.ldx[src];
  pdl_wait();

  // ---- pass 0: global -> shared, per-channel sums
  float s[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) s[j] = 0.f;
  for (int px = lane_p; px < npix; px += 4 * p_step) {
    uint4 u[4];
#pragma unroll
    for (int k = 0; k < 4; ++k)
      if (px + k * p_step < npix) u[k] = *reinterpret_cast<const uint4*>(xb + static_cast<size_t>(px + k * p_step) * ld);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      if (px + k * p_step < npix) {
        slab[static_cast<size_t>(px + k * p_step) * p.vpp + v] = u[k];
        const float2 a = H::unpack(u[k].x), b = H::unpack(u[k].y), c = H::unpack(u[k].z), d = H::unpack(u[k].w);
        s[0] += a.x; s[1] += a.y; s[2] += b.x; s[3] += b.y; s[4] += c.x; s[5] += c.y; s[6] += d.x; s[7] += d.y;
      }
    }
  }
  // CTA reduction in a fixed order, three levels: thread partials [p_step][span] -> R1 = 960 / span partial rows per channel
  // (every thread sums a strided set of rows) -> one warp per group sums its R1 x cg values (strided per lane, shuffle tree)
  auto cta_group_totals = [&](const float (&val)[8], int which) {
#pragma unroll
    for (int j = 0; j < 8; ++j) part[lane_p * p.span + v * 8 + j] = val[j];
    __syncthreads();
    const int R1 = kGnSlabThreads / p.span;
    {
      const int c = t % p.span, rr = t / p.span;
      float acc = 0.f;
      for (int r = rr; r < p_step; r += R1) acc += part[r * p.span + c];
      part2[rr * p.span + c] = acc;
    }
    __syncthreads();
    const int w = t >> 5, ln = t & 31;
    if (w < p.gb) {
      float acc = 0.f;
      const int cnt = R1 * p.cg;
      for (int i = ln; i < cnt; i += 32) {
        const int rr = i / p.cg, c = w * p.cg + (i - rr * p.cg);
        acc += part2[rr * p.span + c];
      }
      acc = warp_sum(acc);
      if (ln == 0) red[which][w] = acc;
    }
  };
  auto cluster_total = [&](int which, int g) {  // every CTA sums the ranks in the same order
    float acc = 0.f;
    const uint32_t a = smem_u32(&red[which][g]);
    for (uint32_t r = 0; r < csize; ++r) acc += ld_dsmem_f32(map_to_rank(a, r));
    return acc;
  };
  const float inv_cnt = 1.0f / (static_cast<float>(p.hw) * static_cast<float>(p.cg));
  cta_group_totals(s, 0);
  cluster_sync_all();
  if (t < p.gb) g_mean[t] = cluster_total(0, t) * inv_cnt;
  __syncthreads();

  // ---- pass 1 (shared memory only): squared deviations from the group mean
  float mean8[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) mean8[j] = g_mean[(v * 8 + j) / p.cg];
  float q[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) q[j] = 0.f;
  for (int px = lane_p; px < npix; px += p_step) {
    const uint4 u = slab[static_cast<size_t>(px) * p.vpp + v];
    const float2 a = H::unpack(u.x), b = H::unpack(u.y), c = H::unpack(u.z), d = H::unpack(u.w);
    const float e[8] = {a.x - mean8[0], a.y - mean8[1], b.x - mean8[2], b.y - mean8[3], c.x - mean8[4], c.y - mean8[5], d.x - mean8[6], d.y - mean8[7]};
#pragma unroll
    for (int j = 0; j < 8; ++j) q[j] = fmaf(e[j], e[j], q[j]);
  }
  cta_group_totals(q, 1);
  cluster_sync_all();
  if (t < p.gb) g_rstd[t] = rsqrtf(cluster_total(1, t) * inv_cnt + p.eps);
  __syncthreads();

  // ---- pass 2: normalise + affine (+ SiLU) from shared memory, one write
  float sc[8], sh[8];
  {
    const typename H::T* gam = static_cast<const typename H::T*>(p.gamma);
    const typename H::T* bet = static_cast<const typename H::T*>(p.beta);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int g = (v * 8 + j) / p.cg;
      const float ga = gam ? H::to_float(gam[ch0 + j]) : 1.0f;
      const float be = bet ? H::to_float(bet[ch0 + j]) : 0.0f;
      sc[j] = g_rstd[g] * ga;
      sh[j] = fmaf(-g_mean[g], sc[j], be);
    }
  }
  typename H::T* yb = static_cast<typename H::T*>(p.y) + (static_cast<size_t>(n) * p.hw + pix0) * p.ldy + ch0;
  const bool silu = p.act == ACT_SILU;
  for (int px = lane_p; px < npix; px += p_step) {
    const uint4 u = slab[static_cast<size_t>(px) * p.vpp + v];
    const float2 a = H::unpack(u.x), b = H::unpack(u.y), c = H::unpack(u.z), d = H::unpack(u.w);
    float f[8] = {a.x, a.y, b.x, b.y, c.x, c.y, d.x, d.y};
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      f[j] = fmaf(f[j], sc[j], sh[j]);
      if (silu) f[j] = silu_fast(f[j]);
    }
    uint4 o;
    o.x = H::pack(f[0], f[1]);
    o.y = H::pack(f[2], f[3]);
    o.z = H::pack(f[4], f[5]);
    o.w = H::pack(f[6], f[7]);
    *reinterpret_cast<uint4*>(yb + static_cast<size_t>(px) * p.ldy) = o;
  }
  cluster_sync_all();  // nobody leaves while a peer may still read its `red`
}

// Returns 0 when the slab kernel cannot take the shape (the two-kernel path does), else the cluster size.
static int gn_slab_plan(int batch, int hw, int C, int groups, int c0, int nsrc, GroupNormSlabParams* p) {
  if (groups > 32 * 32 || C % groups) return 0;
  const int cg = C / groups;
  int gb = 1;
  while (gb <= 32 && ((gb * cg) % 8 != 0 || groups % gb != 0)) ++gb;
  if (gb > 32) return 0;
  const int span = gb * cg, vpp = span / 8;
  if (kGnSlabThreads % span != 0 || gb > kGnSlabThreads / 32) return 0;  // (implies 960 % vpp == 0)
  if (nsrc == 2 && c0 % 8 != 0) return 0;
  // Clusters of up to 4 CTAs: measured on B200 (tools/library_bar.py --only norm, us per launch, slab vs two kernels):
  // 4096 x 640: 18.5 vs 28.8, 4096 x 1280: 27.4 vs 39.1, 1024 x 1280: 14.2 vs 22.0, 1024 x 2560: 18.6 vs 25.6 - but
  // 16384 x 320, which needs clusters of 8, 48.6 vs 39.7: those shapes keep the two-kernel path.
  static const int max_cl = getenv("B200_GN_SLAB_MAX_CLUSTER") ? atoi(getenv("B200_GN_SLAB_MAX_CLUSTER")) : 4;  // tuning knob (1..8)
  for (int cl = 1; cl <= max_cl; cl *= 2) {
    const int ppc = (hw + cl - 1) / cl;
    const long long slab = static_cast<long long>(ppc) * span * 2;
    const long long part = static_cast<long long>(kGnSlabThreads / vpp) * span * 4;
    if (slab + part <= kGnSlabMaxBytes + 32768 && slab <= kGnSlabMaxBytes) {
      // the smallest cluster that fits; wider clusters while fewer than ~100 CTAs would be at work (smaller slabs per SM)
      while (cl < max_cl && static_cast<long long>(cl) * (groups / gb) * batch < 96 && hw / (2 * cl) >= 64) cl *= 2;
      if (p) {
        p->cg = cg; p->gb = gb; p->span = span; p->vpp = vpp; p->units = groups / gb; p->ppc = (hw + cl - 1) / cl;
      }
      return cl;
    }
  }
  return 0;
}

// ------------------------------------------------------------------------------------------------
// LayerNorm: one warp per row, the row held in registers as packed 16-bit (<= 16 x 8 values per lane)
// ------------------------------------------------------------------------------------------------
struct LayerNormParams {
  const void* x;
  int ldx, rows, cols;
  float eps;
  const void* gamma;
  const void* beta;
  const void* scale;
  const void* shift;
  int ld_mod, rows_per_group;
  void* y;
  int ldy;
  int rms;  // 1: no mean subtraction (T5LayerNorm / RMSNorm): y = x * rsqrt(mean(x^2) + eps) * gamma
};

template <bool FP16, int NV>
__global__ void __launch_bounds__(256) layer_norm_kernel(const LayerNormParams p) {
  pdl_trigger();
  pdl_wait();
  using H = Half16<FP16>;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int row = blockIdx.x * (blockDim.x >> 5) + warp;
  if (row >= p.rows) return;
  const int nvec = p.cols >> 3;
  const typename H::T* x = static_cast<const typename H::T*>(p.x) + static_cast<size_t>(row) * p.ldx;
  uint4 r[NV];
  float sum = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int v = lane + i * 32;
    if (v < nvec) {
      r[i] = *reinterpret_cast<const uint4*>(x + v * 8);
      float2 a = H::unpack(r[i].x), b = H::unpack(r[i].y), c = H::unpack(r[i].z), d = H::unpack(r[i].w);
      sum += (a.x + a.y) + (b.x + b.y) + (c.x + c.y) + (d.x + d.y);
    }
  }
  const float mean = p.rms ? 0.f : warp_sum(sum) / static_cast<float>(p.cols);
  float sq = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int v = lane + i * 32;
    if (v < nvec) {
      float2 a = H::unpack(r[i].x), b = H::unpack(r[i].y), c = H::unpack(r[i].z), d = H::unpack(r[i].w);
      float e;
      e = a.x - mean; sq += e * e; e = a.y - mean; sq += e * e;
      e = b.x - mean; sq += e * e; e = b.y - mean; sq += e * e;
      e = c.x - mean; sq += e * e; e = c.y - mean; sq += e * e;
      e = d.x - mean; sq += e * e; e = d.y - mean; sq += e * e;
    }
  }
  const float rstd = rsqrtf(warp_sum(sq) / static_cast<float>(p.cols) + p.eps);
  const typename H::T* gamma = static_cast<const typename H::T*>(p.gamma);
  const typename H::T* beta = static_cast<const typename H::T*>(p.beta);
  const int grp = (p.scale || p.shift) ? row / p.rows_per_group : 0;
  const typename H::T* scale = p.scale ? static_cast<const typename H::T*>(p.scale) + static_cast<size_t>(grp) * p.ld_mod : nullptr;
  const typename H::T* shift = p.shift ? static_cast<const typename H::T*>(p.shift) + static_cast<size_t>(grp) * p.ld_mod : nullptr;
  typename H::T* y = static_cast<typename H::T*>(p.y) + static_cast<size_t>(row) * p.ldy;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int v = lane + i * 32;
    if (v < nvec) {
      float2 a = H::unpack(r[i].x), b = H::unpack(r[i].y), c = H::unpack(r[i].z), d = H::unpack(r[i].w);
      float f[8] = {a.x, a.y, b.x, b.y, c.x, c.y, d.x, d.y};
#pragma unroll
      for (int j = 0; j < 8; ++j) f[j] = (f[j] - mean) * rstd;
      if (gamma) {
        uint4 g = *reinterpret_cast<const uint4*>(gamma + v * 8);
        float2 g0 = H::unpack(g.x), g1 = H::unpack(g.y), g2 = H::unpack(g.z), g3 = H::unpack(g.w);
        f[0] *= g0.x; f[1] *= g0.y; f[2] *= g1.x; f[3] *= g1.y; f[4] *= g2.x; f[5] *= g2.y; f[6] *= g3.x; f[7] *= g3.y;
      }
      if (beta) {
        uint4 g = *reinterpret_cast<const uint4*>(beta + v * 8);
        float2 g0 = H::unpack(g.x), g1 = H::unpack(g.y), g2 = H::unpack(g.z), g3 = H::unpack(g.w);
        f[0] += g0.x; f[1] += g0.y; f[2] += g1.x; f[3] += g1.y; f[4] += g2.x; f[5] += g2.y; f[6] += g3.x; f[7] += g3.y;
      }
      if (scale) {
        uint4 g = *reinterpret_cast<const uint4*>(scale + v * 8);
        float2 g0 = H::unpack(g.x), g1 = H::unpack(g.y), g2 = H::unpack(g.z), g3 = H::unpack(g.w);
        f[0] *= 1.f + g0.x; f[1] *= 1.f + g0.y; f[2] *= 1.f + g1.x; f[3] *= 1.f + g1.y;
        f[4] *= 1.f + g2.x; f[5] *= 1.f + g2.y; f[6] *= 1.f + g3.x; f[7] *= 1.f + g3.y;
      }
      if (shift) {
        uint4 g = *reinterpret_cast<const uint4*>(shift + v * 8);
        float2 g0 = H::unpack(g.x), g1 = H::unpack(g.y), g2 = H::unpack(g.z), g3 = H::unpack(g.w);
        f[0] += g0.x; f[1] += g0.y; f[2] += g1.x; f[3] += g1.y; f[4] += g2.x; f[5] += g2.y; f[6] += g3.x; f[7] += g3.y;
      }
      uint4 o;
      o.x = H::pack(f[0], f[1]);
      o.y = H::pack(f[2], f[3]);
      o.z = H::pack(f[4], f[5]);
      o.w = H::pack(f[6], f[7]);
      *reinterpret_cast<uint4*>(y + v * 8) = o;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// per-head RMSNorm (weight) + rotary embedding on the q and k slices of a fused QKV buffer, in place.
// One warp per (token row, head).  Rounding points follow the reference eager ops: rms_norm output is
// rounded to 16 bit twice (normalised value, then * weight), rope is computed in fp32 and rounded once.
// ------------------------------------------------------------------------------------------------
struct QkNormRopeParams {
  void* qkv;
  long long ld;
  int rows, heads, k_off;
  int txt_rows;  // rows [0, txt_rows) of every sequence use the *_txt weights (joint attention: text first)
  int seq;       // rows per sequence (rope position = row % seq)
  int txt_period;  // the text rows are [0, txt_rows) of every txt_period positions (== seq unless context-parallel)
  const void* wq;
  const void* wk;
  const void* wq_txt;
  const void* wk_txt;
  const float* cos_t;  // [seq, HD] or nullptr
  const float* sin_t;
  float eps;
};

template <bool FP16, int HD>
__global__ void __launch_bounds__(256) qk_norm_rope_kernel(const QkNormRopeParams p) {
  pdl_trigger();
  pdl_wait();
  using H = Half16<FP16>;
  constexpr int EPL = HD / 32;  // elements per lane (2 or 4): rotation pairs stay inside a lane
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // grid (rows, ceil(heads / 8)): no 64-bit division per warp (three of them cost more than the warp's 512 bytes of work; ncu
  // before: XU pipe 53 %, 40 us for 66 MB)
  const int head = blockIdx.y * (blockDim.x >> 5) + warp;
  if (head >= p.heads) return;
  const long long row = blockIdx.x;
  const int pos = static_cast<int>(static_cast<unsigned>(blockIdx.x) % static_cast<unsigned>(p.seq));
  const bool is_txt = static_cast<int>(static_cast<unsigned>(pos) % static_cast<unsigned>(p.txt_period)) < p.txt_rows;
  typename H::T* base = static_cast<typename H::T*>(p.qkv) + row * p.ld + head * HD + lane * EPL;
  float cs[EPL], sn[EPL];
  if (p.cos_t) {
#pragma unroll
    for (int j = 0; j < EPL; ++j) {
      cs[j] = p.cos_t[static_cast<long long>(pos) * HD + lane * EPL + j];
      sn[j] = p.sin_t[static_cast<long long>(pos) * HD + lane * EPL + j];
    }
  }
#pragma unroll
  for (int which = 0; which < 2; ++which) {
    typename H::T* ptr = base + (which ? p.k_off : 0);
    const typename H::T* w = static_cast<const typename H::T*>(which ? (is_txt ? p.wk_txt : p.wk) : (is_txt ? p.wq_txt : p.wq));
    float x[EPL];
    if (EPL == 4) {
      uint2 u = *reinterpret_cast<const uint2*>(ptr);
      float2 a = H::unpack(u.x), b = H::unpack(u.y);
      x[0] = a.x; x[1] = a.y; x[EPL - 2] = b.x; x[EPL - 1] = b.y;
    } else {
      uint32_t u = *reinterpret_cast<const uint32_t*>(ptr);
      float2 a = H::unpack(u);
      x[0] = a.x; x[1] = a.y;
    }
    float ss = 0.f;
#pragma unroll
    for (int j = 0; j < EPL; ++j) ss += x[j] * x[j];
    ss = warp_sum(ss);
    const float rstd = rsqrtf(ss * (1.0f / static_cast<float>(HD)) + p.eps);  // HD is a power of two: same value as ss / HD
#pragma unroll
    for (int j = 0; j < EPL; ++j) {
      float v = H::to_float(H::from_float(x[j] * rstd));
      if (w) v = H::to_float(H::from_float(v * H::to_float(w[lane * EPL + j])));
      x[j] = v;
    }
    if (p.cos_t) {
#pragma unroll
      for (int j = 0; j < EPL; j += 2) {
        const float re = x[j], im = x[j + 1];
        x[j] = re * cs[j] + (-im) * sn[j];
        x[j + 1] = im * cs[j + 1] + re * sn[j + 1];
      }
    }
    if (EPL == 4) {
      uint2 o;
      o.x = H::pack(x[0], x[1]);
      o.y = H::pack(x[EPL - 2], x[EPL - 1]);
      *reinterpret_cast<uint2*>(ptr) = o;
    } else {
      *reinterpret_cast<uint32_t*>(ptr) = H::pack(x[0], x[1]);
    }
  }
}

}  // namespace b200

extern "C" {

int64_t b200_group_norm_workspace_bytes(int32_t batch, int32_t hw, int32_t groups) {
  int chunks, ppc;
  b200::gn_plan(batch, hw, 0, &chunks, &ppc);
  // counters [1024] + stats [batch][groups][2] + partial [batch][chunks][groups][3]; chunks <= 256
  int64_t floats = 1024 + static_cast<int64_t>(batch) * 256 * groups * 3 + static_cast<int64_t>(batch) * groups * 2;
  (void)chunks;
  return floats * 4 + 256;
}

/* Kernels b200_group_norm launches for this shape: 1 (single-pass slab kernel) or 2 (statistics + apply). */
int32_t b200_group_norm_launches(int32_t hw, int32_t C, int32_t groups, int32_t c0, int32_t two_sources) {
  static const bool no_slab = getenv("B200_GN_NO_SLAB") && atoi(getenv("B200_GN_NO_SLAB")) != 0;
  if (hw <= 0 || C <= 0 || groups <= 0 || C % groups) return 2;
  return (!no_slab && b200::gn_slab_plan(2, hw, C, groups, c0, two_sources ? 2 : 1, nullptr) > 0) ? 1 : 2;
}

int b200_group_norm(const b200_group_norm_args* a, void* stream) {
  using namespace b200;
  B200_CHECK_ARG(a && a->x[0] && a->y && a->workspace, "group_norm: null pointer");
  const int nsrc = (a->x[1] && a->c[1] > 0) ? 2 : 1;
  const int C = a->c[0] + (nsrc == 2 ? a->c[1] : 0);
  B200_CHECK_ARG(a->groups > 0 && a->groups <= 256 && C % a->groups == 0, "group_norm: C=%d groups=%d", C, a->groups);
  for (int s = 0; s < nsrc; ++s) {
    B200_CHECK_ARG(a->c[s] % 8 == 0 && a->ldx[s] % 8 == 0 && a->ldx[s] >= a->c[s] && aligned16(a->x[s]),
                   "group_norm: source %d needs channels/stride multiple of 8 and 16-byte alignment", s);
  }
  B200_CHECK_ARG(a->ldy % 8 == 0 && a->ldy >= C && aligned16(a->y), "group_norm: y stride/alignment");
  B200_CHECK_ARG(a->act == B200_ACT_NONE || a->act == B200_ACT_SILU, "group_norm: act must be NONE or SILU");
  B200_CHECK_ARG((!a->gamma || aligned16(a->gamma)) && (!a->beta || aligned16(a->beta)), "group_norm: gamma / beta must be 16-byte aligned");
  B200_CHECK_ARG(a->workspace_bytes >= b200_group_norm_workspace_bytes(a->batch, a->hw, a->groups),
                 "group_norm: workspace too small");
  B200_CHECK_ARG(a->batch > 0 && a->hw > 0, "group_norm: bad shape");

  static const bool no_slab = getenv("B200_GN_NO_SLAB") && atoi(getenv("B200_GN_NO_SLAB")) != 0;  // tuning / test knob
  {
    GroupNormSlabParams sp;
    memset(&sp, 0, sizeof(sp));
    const int cl = no_slab ? 0 : gn_slab_plan(a->batch, a->hw, C, a->groups, a->c[0], nsrc, &sp);
    if (cl > 0) {
      sp.x[0] = a->x[0];
      sp.x[1] = nsrc == 2 ? a->x[1] : nullptr;
      sp.c0 = nsrc == 2 ? a->c[0] : C;
      sp.ldx[0] = a->ldx[0];
      sp.ldx[1] = nsrc == 2 ? a->ldx[1] : 0;
      sp.C = C; sp.hw = a->hw; sp.groups = a->groups; sp.eps = a->eps;
      sp.gamma = a->gamma; sp.beta = a->beta; sp.act = a->act; sp.y = a->y; sp.ldy = a->ldy;
      const size_t smem = static_cast<size_t>(sp.ppc) * sp.span * 2 + static_cast<size_t>(kGnSlabThreads / sp.vpp) * sp.span * 4;
      const bool fp16s = a->dtype == B200_DTYPE_FP16;
      auto kern = fp16s ? group_norm_slab_kernel<true> : group_norm_slab_kernel<false>;
      cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
      if (e != cudaSuccess) return set_error(B200_ERR_CUDA, "group_norm (slab) smem attr: %s", cudaGetErrorString(e));
      cudaLaunchConfig_t cfg;
      memset(&cfg, 0, sizeof(cfg));
      cfg.gridDim = dim3(cl, sp.units * a->batch);
      cfg.blockDim = dim3(kGnSlabThreads);
      cfg.dynamicSmemBytes = smem;
      cfg.stream = static_cast<cudaStream_t>(stream);
      cudaLaunchAttribute attr[2];
      attr[0].id = cudaLaunchAttributeClusterDimension;
      attr[0].val.clusterDim.x = cl;
      attr[0].val.clusterDim.y = 1;
      attr[0].val.clusterDim.z = 1;
      attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
      attr[1].val.programmaticStreamSerializationAllowed = pdl_enabled() ? 1 : 0;
      cfg.attrs = attr;
      cfg.numAttrs = 2;
      e = cudaLaunchKernelEx(&cfg, kern, sp);
      if (e != cudaSuccess) return set_error(B200_ERR_CUDA, "group_norm (slab, cluster %d) launch: %s", cl, cudaGetErrorString(e));
      return 0;
    }
  }

  GroupNormParams p;
  memset(&p, 0, sizeof(p));
  p.x[0] = a->x[0];
  p.x[1] = nsrc == 2 ? a->x[1] : nullptr;
  p.c[0] = a->c[0];
  p.c[1] = nsrc == 2 ? a->c[1] : 0;
  p.ldx[0] = a->ldx[0];
  p.ldx[1] = nsrc == 2 ? a->ldx[1] : 0;
  p.C = C;
  p.V = C / 8;
  p.V0 = a->c[0] / 8;
  p.batch = a->batch;
  p.hw = a->hw;
  p.groups = a->groups;
  p.cg = C / a->groups;
  p.eps = a->eps;
  p.gamma = a->gamma;
  p.beta = a->beta;
  p.act = a->act;
  p.y = a->y;
  p.ldy = a->ldy;
  gn_plan(a->batch, a->hw, C, &p.chunks, &p.ppc);
  B200_CHECK_ARG(p.V <= 512, "group_norm: C=%d too large (<= 4096)", C);
  p.P = 512 / p.V;
  if (p.P < 1) p.P = 1;
  if (p.P > 64) p.P = 64;
  if (p.P > p.ppc) p.P = p.ppc;
  // workspace layout: [1024 completion counters | stats | partials].  The counters live at a FIXED offset so that
  // they stay zero between launches whatever batch / shape the previous call used.
  B200_CHECK_ARG(a->batch <= 1024, "group_norm: batch %d > 1024", a->batch);
  float* ws = static_cast<float*>(a->workspace);
  p.counter = reinterpret_cast<unsigned int*>(ws);
  p.stats = ws + 1024;
  p.partial = p.stats + static_cast<size_t>(a->batch) * a->groups * 2;
  {
    // ~4 CTAs per SM for the apply pass, every pixel lane of a CTA gets at least 4 pixels (the per-thread scale/bias
    // prologue amortises over them)
    long long total = static_cast<long long>(a->batch) * a->hw;
    long long ppb = (total + 4LL * num_sms() - 1) / (4LL * num_sms());
    if (ppb < 4LL * p.P) ppb = 4LL * p.P;
    if (ppb > a->hw) ppb = a->hw;
    p.apply_ppb = static_cast<int>(ppb);
  }

  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const bool fp16 = a->dtype == B200_DTYPE_FP16;
  {
    int threads = p.V * p.P;
    threads = (threads + 31) / 32 * 32;
    if (threads < 32 * ((a->groups + 31) / 32)) threads = 32 * ((a->groups + 31) / 32);
    size_t smem = static_cast<size_t>(p.P) * C * sizeof(float2);
    dim3 grid(p.chunks, a->batch);
    if (smem > 48 * 1024) {
      cudaError_t e = cudaFuncSetAttribute(fp16 ? group_norm_stats_kernel<true> : group_norm_stats_kernel<false>,
                                           cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
      if (e != cudaSuccess) return set_error(B200_ERR_CUDA, "group_norm smem attr: %s", cudaGetErrorString(e));
    }
    if (fp16)
      launch_pdl(group_norm_stats_kernel<true>, dim3(grid), dim3(threads), smem, st, p);
    else
      launch_pdl(group_norm_stats_kernel<false>, dim3(grid), dim3(threads), smem, st, p);
    int r = check_launch("group_norm_stats_kernel");
    if (r) return r;
  }
  {
    dim3 grid((a->hw + p.apply_ppb - 1) / p.apply_ppb, a->batch);
    int threads = (p.V * p.P + 31) / 32 * 32;
    if (fp16)
      launch_pdl(group_norm_apply_kernel<true>, dim3(grid), dim3(threads), 0, st, p);
    else
      launch_pdl(group_norm_apply_kernel<false>, dim3(grid), dim3(threads), 0, st, p);
    return check_launch("group_norm_apply_kernel");
  }
}

int b200_layer_norm(const b200_layer_norm_args* a, void* stream) {
  using namespace b200;
  B200_CHECK_ARG(a && a->x && a->y, "layer_norm: null pointer");
  B200_CHECK_ARG(a->cols > 0 && a->cols % 8 == 0 && a->cols <= 4096, "layer_norm: cols=%d (need multiple of 8, <= 4096)",
                 a->cols);
  B200_CHECK_ARG(a->ldx % 8 == 0 && a->ldy % 8 == 0 && aligned16(a->x) && aligned16(a->y), "layer_norm: alignment");
  if (a->scale || a->shift)
    B200_CHECK_ARG(a->rows_per_group > 0 && a->ld_mod % 8 == 0, "layer_norm: modulation needs rows_per_group, ld_mod %% 8");
  if (a->rows <= 0) return 0;
  LayerNormParams p;
  p.x = a->x;
  p.ldx = a->ldx;
  p.rows = a->rows;
  p.cols = a->cols;
  p.eps = a->eps;
  p.gamma = a->gamma;
  p.beta = a->beta;
  p.scale = a->scale;
  p.shift = a->shift;
  p.ld_mod = a->ld_mod;
  p.rows_per_group = a->rows_per_group > 0 ? a->rows_per_group : 1;
  p.rms = a->rms ? 1 : 0;
  p.y = a->y;
  p.ldy = a->ldy;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const bool fp16 = a->dtype == B200_DTYPE_FP16;
  const int warps = 8;
  const int grid = (a->rows + warps - 1) / warps;
  const int nvec = a->cols / 8;
#define B200_LN(NV)                                              \
  if (fp16)                                                      \
    launch_pdl(layer_norm_kernel<true, NV>, dim3(grid), dim3(warps * 32), 0, st, p); \
  else                                                           \
    launch_pdl(layer_norm_kernel<false, NV>, dim3(grid), dim3(warps * 32), 0, st, p);
  if (nvec <= 32 * 3) {
    B200_LN(3)
  } else if (nvec <= 32 * 5) {
    B200_LN(5)
  } else if (nvec <= 32 * 12) {
    B200_LN(12)
  } else {
    B200_LN(16)
  }
#undef B200_LN
  return check_launch("layer_norm_kernel");
}

int b200_qk_norm_rope(const b200_qk_norm_rope_args* a, void* stream) {
  using namespace b200;
  B200_CHECK_ARG(a && a->qkv, "qk_norm_rope: null pointer");
  B200_CHECK_ARG(a->head_dim == 64 || a->head_dim == 128, "qk_norm_rope: head_dim %d (64 or 128)", a->head_dim);
  B200_CHECK_ARG(a->rows > 0 && a->heads > 0 && a->seq > 0 && a->ld % 4 == 0 && a->k_off % 4 == 0,
                 "qk_norm_rope: bad shape / stride");
  B200_CHECK_ARG((reinterpret_cast<uintptr_t>(a->qkv) & 7u) == 0, "qk_norm_rope: qkv must be 8-byte aligned");
  B200_CHECK_ARG((a->cos_table == nullptr) == (a->sin_table == nullptr), "qk_norm_rope: need both cos and sin or neither");
  QkNormRopeParams p;
  p.qkv = a->qkv; p.ld = a->ld; p.rows = a->rows; p.heads = a->heads; p.k_off = a->k_off;
  p.txt_rows = a->txt_rows; p.seq = a->seq;
  p.txt_period = a->txt_period > 0 ? a->txt_period : a->seq;
  p.wq = a->wq; p.wk = a->wk;
  p.wq_txt = a->wq_txt ? a->wq_txt : a->wq;
  p.wk_txt = a->wk_txt ? a->wk_txt : a->wk;
  p.cos_t = a->cos_table; p.sin_t = a->sin_table; p.eps = a->eps;
  B200_CHECK_ARG(a->heads <= 8 * 65535, "qk_norm_rope: too many heads");
  const dim3 grid(static_cast<unsigned int>(a->rows), static_cast<unsigned int>((a->heads + 7) / 8));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const bool fp16 = a->dtype == B200_DTYPE_FP16;
  if (a->head_dim == 128) {
    if (fp16) launch_pdl(qk_norm_rope_kernel<true, 128>, grid, dim3(256), 0, st, p);
    else launch_pdl(qk_norm_rope_kernel<false, 128>, grid, dim3(256), 0, st, p);
  } else {
    if (fp16) launch_pdl(qk_norm_rope_kernel<true, 64>, grid, dim3(256), 0, st, p);
    else launch_pdl(qk_norm_rope_kernel<false, 64>, grid, dim3(256), 0, st, p);
  }
  return check_launch("qk_norm_rope_kernel");
}

}  // extern "C"
