// Microbenchmark: TMEM read (tcgen05.ld 32x32b.x32) throughput per SM with 4 or 8 reading warps, and the latency of one load.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/micro/_bin/tmem_bw tools/micro/tmem_bw.cu && tools/micro/_bin/tmem_bw
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

template <int X>
__device__ __forceinline__ void ld(uint32_t taddr, uint32_t (&v)[32]);
template <>
__device__ __forceinline__ void ld<32>(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
        "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
        "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
}

// mode 0: back-to-back loads, 4 in flight (throughput); mode 1: one load, wait, dependent next (latency)
template <int MODE>
__global__ void __launch_bounds__(256, 1) tmem_read(int iters, long long* cycles, uint32_t* sink) {
  __shared__ uint32_t tmem_ptr;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_ptr)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  const uint32_t base = tmem_ptr + (static_cast<uint32_t>((warp & 3) * 32) << 16) + (warp >> 2) * 256;
  uint32_t a[32], b[32], c[32], d[32];
  uint32_t acc = 0;
  __syncthreads();
  const long long t0 = clock64();
  if (MODE == 0) {
    for (int i = 0; i < iters; ++i) {
      ld<32>(base, a);
      ld<32>(base + 32, b);
      ld<32>(base + 64, c);
      ld<32>(base + 96, d);
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      acc ^= a[0] ^ b[1] ^ c[2] ^ d[3];
    }
  } else {
    uint32_t off = 0;
    for (int i = 0; i < iters * 4; ++i) {
      ld<32>(base + off, a);
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      off = (a[0] & 1u) * 32;  // next address depends on the data
      acc ^= a[5];
    }
  }
  const long long t1 = clock64();
  __syncthreads();
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
  sink[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_ptr), "r"(512));
}

int main() {
  long long* cyc;
  uint32_t* sink;
  cudaMalloc(&cyc, 148 * sizeof(long long));
  cudaMalloc(&sink, 148 * 256 * 4);
  const int iters = 2000;
  for (int threads : {32, 128, 256}) {
    for (int mode = 0; mode < 2; ++mode) {
      for (int rep = 0; rep < 2; ++rep) {
        if (mode == 0) tmem_read<0><<<148, threads>>>(iters, cyc, sink);
        else tmem_read<1><<<148, threads>>>(iters, cyc, sink);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("error: %s\n", cudaGetErrorString(e)); return 1; }
      }
      long long h[148];
      cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
      const double per_ld = static_cast<double>(h[0]) / (iters * 4.0);
      const double bytes = (threads / 32) * 4096.0;  // per round of one load per warp
      printf("%3d threads (%d warps) %s: %.1f cycles per x32 load per warp -> %.1f B/clk/SM\n", threads, threads / 32,
             mode == 0 ? "4 loads in flight" : "dependent loads  ", per_ld, bytes / per_ld);
    }
  }
  return 0;
}
