// Peer exchange of the data-parallel sharded retrieval without NCCL: every rank owns one "symmetric" buffer that its
// peers map over CUDA IPC (NVLink / NVSwitch peer memory); a push kernel stores this rank's block straight into every
// peer's buffer and raises a per-source flag there, a wait kernel on the consumer side spins on its local flags.
//
//   all-gather of the query embeddings : push(q [B,D] -> every peer's q_all[rank])            + wait
//   all-to-all of the per-shard top-k  : push(lists of peer j's queries -> peer j's mine[rank]) + wait
//
// Flags carry an epoch: push raises flag[src] to (*epoch + 1), wait spins until every flag > *epoch and then increments
// *epoch.  Both kernels of a (slot, exchange) pair are stream-ordered on every rank and every rank runs the same
// sequence, so the device-resident epochs stay in step without any host value -- the whole step is CUDA-graph capturable.
// Buffer reuse needs no back-pressure: a rank can only push step s+depth of a slot after its own step s of that slot
// has completed, which required every peer to have consumed step s (see serving.py).
#include <string.h>

#include "common.cuh"

namespace t2p {

constexpr int PUSH_THREADS = 512;

struct PushArgs {
  void* base[T2P_MAX_PEERS];
  const uint8_t* src[2];
  size_t src_stride[2], dst_off[2], bytes[2];
  size_t flag_off;  // of flag[my_rank] inside every peer's buffer
  const unsigned long long* epoch;
};

__global__ void __launch_bounds__(PUSH_THREADS) peer_push_kernel(const PushArgs a) {
  const int j = blockIdx.x;  // destination peer
  uint8_t* dst_base = static_cast<uint8_t*>(a.base[j]);
#pragma unroll
  for (int seg = 0; seg < 2; ++seg) {
    const size_t n16 = a.bytes[seg] >> 4;
    const uint4* s = reinterpret_cast<const uint4*>(a.src[seg] + (size_t)j * a.src_stride[seg]);
    uint4* d = reinterpret_cast<uint4*>(dst_base + a.dst_off[seg]);
    for (size_t i = threadIdx.x; i < n16; i += PUSH_THREADS) d[i] = s[i];
  }
  __threadfence_system();  // this thread's peer stores are visible system-wide before the flag
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned long long e = *a.epoch + 1ull;
    volatile unsigned long long* flag = reinterpret_cast<volatile unsigned long long*>(dst_base + a.flag_off);
    *flag = e;
  }
}

__global__ void peer_wait_kernel(volatile unsigned long long* flags, int n_peers, unsigned long long* epoch) {
  const unsigned long long want = *epoch + 1ull;
  if ((int)threadIdx.x < n_peers) {
    const long long t0 = clock64();
    while (flags[threadIdx.x] < want) {
      if (clock64() - t0 > 60000000000LL) __trap();  // ~30 s: a peer died; surface it as a CUDA error, not a hang
    }
  }
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) *epoch = want;
}

}  // namespace t2p

using namespace t2p;

extern "C" {

int t2p_enable_peer_access(int peer_device) {
  int dev = 0;
  T2P_CUDA(cudaGetDevice(&dev));
  if (peer_device == dev) return T2P_OK;
  int can = 0;
  T2P_CUDA(cudaDeviceCanAccessPeer(&can, dev, peer_device));
  T2P_REQUIRE(can, T2P_ERR_UNSUPPORTED, "device %d cannot access device %d (no NVLink / PCIe peer path)", dev, peer_device);
  cudaError_t e = cudaDeviceEnablePeerAccess(peer_device, 0);
  if (e == cudaErrorPeerAccessAlreadyEnabled) {
    cudaGetLastError();
    return T2P_OK;
  }
  T2P_CUDA(e);
  return T2P_OK;
}

// Symmetric buffers are plain cudaMalloc allocations shared through legacy CUDA IPC handles: peer access
// (cudaDeviceEnablePeerAccess / lazily at open) covers them for KERNEL loads and stores, which is not the case for the
// VMM-backed blocks of a caching allocator (measured: kernels fault on torch-IPC tensors of another device).
int t2p_ipc_alloc(size_t bytes, void** d_ptr, void* handle64) {
  T2P_REQUIRE(d_ptr && handle64 && bytes > 0, T2P_ERR_INVALID, "ipc_alloc: bad argument");
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "CUDA IPC handles are 64 bytes");
  void* p = nullptr;
  T2P_CUDA(cudaMalloc(&p, bytes));
  cudaError_t e = cudaMemset(p, 0, bytes);
  cudaIpcMemHandle_t h;
  if (e == cudaSuccess) e = cudaIpcGetMemHandle(&h, p);
  if (e != cudaSuccess) {
    cudaFree(p);
    t2p::set_error("ipc_alloc: %s", cudaGetErrorString(e));
    return T2P_ERR_CUDA;
  }
  memcpy(handle64, &h, sizeof(h));
  *d_ptr = p;
  return T2P_OK;
}

int t2p_ipc_open(const void* handle64, void** d_ptr) {
  T2P_REQUIRE(handle64 && d_ptr, T2P_ERR_INVALID, "ipc_open: bad argument");
  cudaIpcMemHandle_t h;
  memcpy(&h, handle64, sizeof(h));
  T2P_CUDA(cudaIpcOpenMemHandle(d_ptr, h, cudaIpcMemLazyEnablePeerAccess));
  return T2P_OK;
}

int t2p_ipc_close(void* d_ptr) {
  if (d_ptr) T2P_CUDA(cudaIpcCloseMemHandle(d_ptr));
  return T2P_OK;
}

int t2p_ipc_free(void* d_ptr) {
  if (d_ptr) T2P_CUDA(cudaFree(d_ptr));
  return T2P_OK;
}

int t2p_peer_push(const t2p_peers* peers, const void* d_src0, size_t src_stride0, size_t dst_off0, size_t bytes0,
                  const void* d_src1, size_t src_stride1, size_t dst_off1, size_t bytes1, size_t flag_off,
                  const uint64_t* d_epoch, t2p_stream stream) {
  T2P_REQUIRE(peers && d_src0 && d_epoch, T2P_ERR_INVALID, "peer_push: null argument");
  T2P_REQUIRE(peers->n_peers >= 1 && peers->n_peers <= T2P_MAX_PEERS && peers->my_rank >= 0 && peers->my_rank < peers->n_peers,
              T2P_ERR_INVALID, "peer_push: n_peers=%d my_rank=%d", peers->n_peers, peers->my_rank);
  T2P_REQUIRE(bytes0 % 16 == 0 && bytes1 % 16 == 0 && dst_off0 % 16 == 0 && dst_off1 % 16 == 0 && src_stride0 % 16 == 0 &&
                  src_stride1 % 16 == 0 && flag_off % 8 == 0 && (reinterpret_cast<uintptr_t>(d_src0) & 15) == 0 &&
                  (reinterpret_cast<uintptr_t>(d_src1) & 15) == 0,
              T2P_ERR_INVALID, "peer_push: blocks must be 16-byte aligned multiples of 16 bytes");
  PushArgs a;
  for (int j = 0; j < T2P_MAX_PEERS; ++j) a.base[j] = j < peers->n_peers ? peers->base[j] : nullptr;
  for (int j = 0; j < peers->n_peers; ++j) T2P_REQUIRE(a.base[j] != nullptr, T2P_ERR_INVALID, "peer_push: peer %d has no buffer", j);
  a.src[0] = static_cast<const uint8_t*>(d_src0);
  a.src[1] = static_cast<const uint8_t*>(d_src1 ? d_src1 : d_src0);
  a.src_stride[0] = src_stride0; a.src_stride[1] = src_stride1;
  a.dst_off[0] = dst_off0; a.dst_off[1] = dst_off1;
  a.bytes[0] = bytes0; a.bytes[1] = d_src1 ? bytes1 : 0;
  a.flag_off = flag_off;
  a.epoch = reinterpret_cast<const unsigned long long*>(d_epoch);
  peer_push_kernel<<<peers->n_peers, PUSH_THREADS, 0, as_stream(stream)>>>(a);
  T2P_LAUNCH_CHECK();
  return T2P_OK;
}

int t2p_peer_wait(uint64_t* d_flags, int n_peers, uint64_t* d_epoch, t2p_stream stream) {
  T2P_REQUIRE(d_flags && d_epoch && n_peers >= 1 && n_peers <= T2P_MAX_PEERS, T2P_ERR_INVALID, "peer_wait: bad argument");
  peer_wait_kernel<<<1, 32, 0, as_stream(stream)>>>(reinterpret_cast<volatile unsigned long long*>(d_flags), n_peers,
                                                  reinterpret_cast<unsigned long long*>(d_epoch));
  T2P_LAUNCH_CHECK();
  return T2P_OK;
}

}  // extern "C"
