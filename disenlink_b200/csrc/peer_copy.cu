// peer_copy.cu -- the all-gather of the node-partitioned step as one push kernel over NVLink peer memory.
//
// Every rank owns a contiguous slice of each per-node array (Z, s, H, prob, dH, r) and every rank
// needs the whole array before the next kernel.  NCCL's ring all-gather moves 25.6 GB at C5 in
// 33.4 ms (670 GB/s per rank, whatever the algorithm / protocol / channel settings).  With NVSwitch
// every GPU reaches every peer at full NVLink bandwidth, so the owner simply WRITES its slice into
// the same position of every peer's buffer: one kernel, one 16-byte load from local HBM and one
// 16-byte store per peer, no staging, no ring latency.  Peer buffers are mapped with CUDA IPC by the
// host side (disenlink_b200/partition.py); ordering between ranks is the caller's (a barrier after
// the kernel: stores to peer memory are visible to the peer once the writing kernel has completed).
#include <string.h>

#include "dl_common.cuh"

namespace {

constexpr int DL_MAX_PEERS = 15;

struct PeerDst {
  uint4* p[DL_MAX_PEERS];
  int n;
};

__global__ void __launch_bounds__(256)
k_push_slice(const uint4* __restrict__ src, long long n_vec, PeerDst dst) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n_vec; i += stride) {
    const uint4 v = __ldg(src + i);
#pragma unroll 1
    for (int q = 0; q < dst.n; ++q) dst.p[q][i] = v;
  }
}

__global__ void k_push_tail(const unsigned char* __restrict__ src, long long n_bytes, long long offset,
                            PeerDst dst) {
  const long long i = offset + (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n_bytes)
    for (int q = 0; q < dst.n; ++q) reinterpret_cast<unsigned char*>(dst.p[q])[i] = src[i];
}

}  // namespace

extern "C" int dl_push_slice(const void* src, void* const* peer_dst, int n_peers, int64_t n_bytes,
                             dl_stream_t stream) {
  if (n_peers < 0 || n_peers > DL_MAX_PEERS || n_bytes < 0 || (n_bytes > 0 && !src)) return DL_EINVAL;
  if (n_peers == 0 || n_bytes == 0) return DL_OK;
  if (!peer_dst) return DL_EINVAL;
  if (((uintptr_t)src & 15) != 0) return DL_EINVAL;
  PeerDst d;
  d.n = n_peers;
  for (int q = 0; q < n_peers; ++q) {
    if (!peer_dst[q] || ((uintptr_t)peer_dst[q] & 15) != 0) return DL_EINVAL;
    d.p[q] = (uint4*)peer_dst[q];
  }
  cudaStream_t st = (cudaStream_t)stream;
  const long long n_vec = n_bytes / 16;
  if (n_vec > 0) {
    long long grid = (n_vec + 255) / 256;
    if (grid > 148 * 16) grid = 148 * 16;
    k_push_slice<<<(int)grid, 256, 0, st>>>((const uint4*)src, n_vec, d);
    DL_LAUNCH_CHECK();
  }
  if (n_bytes % 16) {
    k_push_tail<<<1, 16, 0, st>>>((const unsigned char*)src, n_bytes, n_vec * 16, d);
    DL_LAUNCH_CHECK();
  }
  return DL_OK;
}

// enable stores from the current device into `peer_device`'s memory (idempotent)
extern "C" int dl_enable_peer_access(int peer_device) {
  int cur = 0, can = 0;
  DL_CUDA_TRY(cudaGetDevice(&cur));
  if (cur == peer_device) return DL_OK;
  DL_CUDA_TRY(cudaDeviceCanAccessPeer(&can, cur, peer_device));
  if (!can) return DL_EINVAL;
  cudaError_t e = cudaDeviceEnablePeerAccess(peer_device, 0);
  if (e == cudaErrorPeerAccessAlreadyEnabled) { cudaGetLastError(); return DL_OK; }
  return (int)e;
}

// Map a peer process's allocation into THIS device's address space.  handle = the 64 bytes of a
// cudaIpcMemHandle_t exported by the owner (torch: tensor.untyped_storage()._share_cuda_()[1]).
// The handle is opened with the caller's current device, so the mapping is a peer mapping from this
// device (opening it with the owner's device index current, as torch's own IPC path does, gives a
// pointer that kernels of another device cannot dereference).
extern "C" int dl_ipc_open(const void* handle, void** base_out) {
  if (!handle || !base_out) return DL_EINVAL;
  cudaIpcMemHandle_t h;
  memcpy(&h, handle, sizeof(h));
  void* p = nullptr;
  cudaError_t e = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
  if (e != cudaSuccess) {
    cudaGetLastError();                  // not sticky: leave the context clean for the caller's fallback
    return (int)e;
  }
  *base_out = p;
  return DL_OK;
}

extern "C" int dl_ipc_close(void* base) {
  if (!base) return DL_OK;
  DL_CUDA_TRY(cudaIpcCloseMemHandle(base));
  return DL_OK;
}
