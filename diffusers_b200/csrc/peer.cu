// Peer memory over NVLink for context parallelism (Ulysses): exportable device buffers, peer mappings of the other
// ranks' buffers, and a device-side barrier between the GPUs of one node.
//
// The data path never calls a collective library: the QKV GEMM's epilogue stores its tiles straight into the peer that
// owns the head (b200_conv_gemm with `y` = a peer mapping), the attention kernel stores every output row into the peer
// that owns the row (b200_attention_args.o_seg), and b200_peer_barrier orders the two phases.
#include "common.cuh"
#include "host_common.h"

namespace b200 {

struct PeerBarrierParams {
  uint32_t* flags[B200_MAX_PEERS];  // flags[r]: rank r's flag words (this process's mapping of them), B200_MAX_PEERS words each
  uint32_t* epoch;                  // this rank's barrier count (device memory, local)
  int rank, nranks;
};

// One CTA, one thread per rank.  Thread t publishes this rank's new epoch into rank t's flag word [rank] (release at
// system scope: every write the preceding kernels of this stream made to peer memory is ordered before it), then waits
// until rank t's epoch has arrived in the local flag word [t].  The epoch lives in device memory and is advanced by the
// kernel itself, so a captured graph replays correctly.
__global__ void peer_barrier_kernel(const PeerBarrierParams p) {
  __shared__ uint32_t s_epoch;
  if (threadIdx.x == 0) s_epoch = *p.epoch + 1u;
  __syncthreads();
  const uint32_t e = s_epoch;
  const int t = threadIdx.x;
  if (t < p.nranks) {
    __threadfence_system();
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p.flags[t] + p.rank), "r"(e) : "memory");
    const uint32_t* mine = p.flags[p.rank] + t;
    uint32_t spins = 0;
    uint64_t t0 = 0;
    for (;;) {
      uint32_t v;
      asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(mine) : "memory");
      if (static_cast<int32_t>(v - e) >= 0) break;
      if ((++spins & 0xffu) == 0) {
        const uint64_t now = global_timer_ns();
        if (t0 == 0) t0 = now;
        else if (now - t0 > 20000000000ull) __trap();  // a peer that never arrives is a CUDA error, not a hung GPU
      }
    }
    __threadfence_system();
  }
  __syncthreads();
  if (threadIdx.x == 0) *p.epoch = e;
}

}  // namespace b200

extern "C" {

int b200_peer_alloc(int64_t bytes, void** ptr, unsigned char* handle64) {
  B200_CHECK_ARG(bytes > 0 && ptr != nullptr && handle64 != nullptr, "b200_peer_alloc: bytes > 0, ptr and handle must be given");
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "handle size");
  void* p = nullptr;
  B200_CHECK_CUDA(cudaMalloc(&p, static_cast<size_t>(bytes)));
  B200_CHECK_CUDA(cudaMemset(p, 0, static_cast<size_t>(bytes)));
  B200_CHECK_CUDA(cudaDeviceSynchronize());
  cudaIpcMemHandle_t h;
  cudaError_t e = cudaIpcGetMemHandle(&h, p);
  if (e != cudaSuccess) {
    cudaFree(p);
    return b200::set_error(B200_ERR_CUDA, "cudaIpcGetMemHandle failed: %s", cudaGetErrorString(e));
  }
  memcpy(handle64, &h, 64);
  *ptr = p;
  return 0;
}

int b200_peer_open(const unsigned char* handle64, void** ptr) {
  B200_CHECK_ARG(handle64 != nullptr && ptr != nullptr, "b200_peer_open: null argument");
  cudaIpcMemHandle_t h;
  memcpy(&h, handle64, 64);
  void* p = nullptr;
  B200_CHECK_CUDA(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
  *ptr = p;
  return 0;
}

int b200_peer_close(void* ptr) {
  B200_CHECK_ARG(ptr != nullptr, "b200_peer_close: null pointer");
  B200_CHECK_CUDA(cudaIpcCloseMemHandle(ptr));
  return 0;
}

int b200_peer_free(void* ptr) {
  B200_CHECK_ARG(ptr != nullptr, "b200_peer_free: null pointer");
  B200_CHECK_CUDA(cudaFree(ptr));
  return 0;
}

int b200_peer_barrier(void* const* flags, void* epoch, int32_t rank, int32_t nranks, void* stream) {
  B200_CHECK_ARG(flags != nullptr && epoch != nullptr, "b200_peer_barrier: null argument");
  B200_CHECK_ARG(nranks >= 1 && nranks <= B200_MAX_PEERS && rank >= 0 && rank < nranks, "b200_peer_barrier: rank %d of %d (at most %d ranks)",
                 rank, nranks, B200_MAX_PEERS);
  b200::PeerBarrierParams p;
  memset(&p, 0, sizeof(p));
  for (int r = 0; r < nranks; ++r) {
    B200_CHECK_ARG(flags[r] != nullptr, "b200_peer_barrier: flags[%d] is null", r);
    p.flags[r] = static_cast<uint32_t*>(flags[r]);
  }
  p.epoch = static_cast<uint32_t*>(epoch);
  p.rank = rank;
  p.nranks = nranks;
  // a plain launch (no programmatic dependent launch): it must start after the preceding kernels have completed
  b200::peer_barrier_kernel<<<1, 32, 0, static_cast<cudaStream_t>(stream)>>>(p);
  return b200::check_launch("peer_barrier_kernel");
}

}  // extern "C"
