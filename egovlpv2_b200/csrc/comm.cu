// NVSwitch peer-to-peer all-gather for the EgoNCE negatives (replaces the NCCL all_gather calls of
// trainer_egoclip.py:25-41 / model.py:385-388 for the [B_local, 4096] embeddings and noun/verb vectors).
//
// Every rank owns a symmetric buffer `slots[W][slot_bytes]` plus `flags[W]` (cudaMalloc + CUDA IPC, mapped
// into every peer).  One kernel launch per rank:
//   1. store this rank's payload into slot `rank` of EVERY peer's buffer (16-byte st.global over NVLink),
//   2. __threadfence_system(), then publish `seq` into flag `rank` of every peer (release),
//   3. spin until all W flags of the local buffer carry `seq` (acquire) -> local buffer holds the gathered rows.
// Payloads here are 32-80 KB per rank, i.e. latency-bound: one launch instead of four NCCL collectives.
#include "common.cuh"
#include <string.h>

#include "host_common.h"

namespace egv {

struct PeerTable {
  uint8_t* slots[16];
  unsigned int* flags[16];
};

// flags buffer layout (32 x uint32, local allocation of every rank):
//   [0..15]  arrival flags, one per source rank (written by that rank, monotonically increasing sequence number)
//   [16..29] local push counters, one per destination peer
//   [30]     local finished-block counter      [31] local sequence number of the last completed gather
// The sequence number lives on the device so that a captured CUDA graph can replay the gather.
__global__ void __launch_bounds__(256) p2p_allgather_kernel(const uint4* __restrict__ src, long long n16, long long slot_bytes,
                                                            long long half_bytes, PeerTable peers, int rank, int world) {
  pdl_enter();
  unsigned int* my_flags = peers.flags[rank];
  const unsigned int seq = *reinterpret_cast<volatile unsigned int*>(my_flags + 31) + 1u;
  const long long parity_off = (long long)(seq & 1u) * half_bytes;   // consecutive gathers alternate slot sets
  // phase 1: blocks p*bpp .. (p+1)*bpp-1 push to peer p
  const int bpp = gridDim.x / world;
  const int peer = blockIdx.x / bpp;
  const int sub = blockIdx.x % bpp;
  uint4* dst = reinterpret_cast<uint4*>(peers.slots[peer] + parity_off + (long long)rank * slot_bytes);
  for (long long i = (long long)sub * blockDim.x + threadIdx.x; i < n16; i += (long long)bpp * blockDim.x) dst[i] = src[i];
  __threadfence_system();
  __syncthreads();
  // phase 2: the last block to finish pushing to `peer` publishes this rank's flag there
  __shared__ bool last;
  if (threadIdx.x == 0) {
    unsigned int* counter = my_flags + 16 + peer;
    const unsigned int done = atomicAdd(counter, 1u) + 1u;
    last = (done == (unsigned int)bpp);
    if (last) *counter = 0u;
  }
  __syncthreads();
  if (last && threadIdx.x == 0) {
    __threadfence_system();
    *reinterpret_cast<volatile unsigned int*>(peers.flags[peer] + rank) = seq;
  }
  // phase 3: block 0 waits for every rank's flag in the local buffer (a peer may already be one gather ahead)
  if (blockIdx.x == 0 && threadIdx.x < world) {
    volatile unsigned int* f = my_flags + threadIdx.x;
    unsigned long long spins = 0;
    while ((int)(*f - seq) < 0) {
      if (++spins > (1ull << 26)) {   // ~10 s: a lost peer must fail fast, not hang the box -- and not kill the context:
        atomicExch(my_flags + 28, 0x100u | (unsigned int)threadIdx.x);   // sticky error word, read by egv_p2p_error()
        break;                        // the gathered rows of that peer are stale; the trainer decides what to do
      }
    }
    __threadfence_system();
  }
  __syncthreads();
  // phase 4: the last block of the grid commits the sequence number (every block has read it by now)
  if (threadIdx.x == 0) {
    const unsigned int fin = atomicAdd(my_flags + 30, 1u) + 1u;
    if (fin == gridDim.x) {
      my_flags[30] = 0u;
      __threadfence();
      *reinterpret_cast<volatile unsigned int*>(my_flags + 31) = seq;
    }
  }
}

// copy the gathered rows of the gather that just completed (parity from the committed sequence number) into `out`
__global__ void __launch_bounds__(256) p2p_collect_kernel(uint4* __restrict__ out, const uint8_t* __restrict__ slots,
                                                          const unsigned int* __restrict__ flags, long long n16,
                                                          long long slot_bytes, long long half_bytes, int world) {
  pdl_enter();
  const unsigned int seq = flags[31];
  const uint8_t* base = slots + (long long)(seq & 1u) * half_bytes;
  const long long total = n16 * world;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / n16, j = i % n16;
    out[i] = reinterpret_cast<const uint4*>(base + r * slot_bytes)[j];
  }
}

}  // namespace egv

using namespace egv;

extern "C" int egv_p2p_alloc(int64_t bytes, void** ptr, void* ipc_handle_64) {
  if (!ptr || !ipc_handle_64 || bytes <= 0) return fail(EGV_ERR_ARG, "p2p_alloc: bad argument");
  cudaError_t e = cudaMalloc(ptr, (size_t)bytes);
  if (e != cudaSuccess) return fail(EGV_ERR_CUDA, "p2p_alloc: %s", cudaGetErrorString(e));
  cudaMemset(*ptr, 0, (size_t)bytes);
  cudaIpcMemHandle_t h;
  e = cudaIpcGetMemHandle(&h, *ptr);
  if (e != cudaSuccess) return fail(EGV_ERR_CUDA, "p2p_alloc ipc handle: %s", cudaGetErrorString(e));
  static_assert(sizeof(h) == 64, "ipc handle size");
  memcpy(ipc_handle_64, &h, 64);
  return EGV_OK;
}
extern "C" int egv_p2p_open(const void* ipc_handle_64, void** ptr) {
  if (!ptr || !ipc_handle_64) return fail(EGV_ERR_ARG, "p2p_open: bad argument");
  cudaIpcMemHandle_t h;
  memcpy(&h, ipc_handle_64, 64);
  cudaError_t e = cudaIpcOpenMemHandle(ptr, h, cudaIpcMemLazyEnablePeerAccess);
  if (e != cudaSuccess) return fail(EGV_ERR_CUDA, "p2p_open: %s", cudaGetErrorString(e));
  return EGV_OK;
}
extern "C" int egv_p2p_close(void* ptr) {
  cudaError_t e = cudaIpcCloseMemHandle(ptr);
  if (e != cudaSuccess) return fail(EGV_ERR_CUDA, "p2p_close: %s", cudaGetErrorString(e));
  return EGV_OK;
}
extern "C" int egv_p2p_free(void* ptr) {
  cudaError_t e = cudaFree(ptr);
  if (e != cudaSuccess) return fail(EGV_ERR_CUDA, "p2p_free: %s", cudaGetErrorString(e));
  return EGV_OK;
}

// slots / flags: host arrays of `world` device pointers (this process's mappings of every rank's buffers; entry `rank`
// is the local allocation).  Every slots buffer holds two sets of world * slot_bytes (consecutive gathers alternate),
// every flags buffer 32 uint32 (see the kernel).  out receives world * bytes, rank-major.
extern "C" int egv_p2p_allgather(const void* src, int64_t bytes, int64_t slot_bytes, void* const* slots, void* const* flags,
                                 int rank, int world, void* out, egv_stream_t stream) {
  if (!src || !slots || !flags || !out) return fail(EGV_ERR_ARG, "p2p_allgather: null pointer");
  if (world < 1 || world > 14 || rank < 0 || rank >= world) return fail(EGV_ERR_ARG, "p2p_allgather: bad rank/world");
  if (bytes % 16 || slot_bytes % 16 || bytes > slot_bytes || (((uintptr_t)src) & 15) || (((uintptr_t)out) & 15))
    return fail(EGV_ERR_ARG, "p2p_allgather: sizes must be multiples of 16 bytes and pointers 16-byte aligned");
  PeerTable t;
  for (int i = 0; i < 16; ++i) {
    t.slots[i] = i < world ? (uint8_t*)slots[i] : nullptr;
    t.flags[i] = i < world ? (unsigned int*)flags[i] : nullptr;
  }
  const long long n16 = bytes / 16;
  const long long half = (long long)world * slot_bytes;
  int bpp = (int)cdiv(n16, 256 * 4);
  if (bpp < 1) bpp = 1;
  if (bpp > 8) bpp = 8;
  cudaStream_t s = (cudaStream_t)stream;
  launch_k(p2p_allgather_kernel, dim3(bpp * world), dim3(256), 0, s, (const uint4*)src, n16, slot_bytes, half, t, rank, world);
  int rc = check_launch("p2p_allgather_kernel");
  if (rc) return rc;
  int cb = (int)cdiv(n16 * world, 256 * 4);
  if (cb < 1) cb = 1;
  if (cb > 64) cb = 64;
  launch_k(p2p_collect_kernel, dim3(cb), dim3(256), 0, s, (uint4*)out, (const uint8_t*)slots[rank], (const unsigned int*)flags[rank], n16, slot_bytes, half, world);
  return check_launch("p2p_collect_kernel");
}

extern "C" int egv_p2p_error(const void* flags_local, int* err) {
  if (!flags_local || !err) return fail(EGV_ERR_ARG, "p2p_error: null pointer");
  unsigned int w = 0;
  cudaError_t e = cudaMemcpy(&w, reinterpret_cast<const unsigned int*>(flags_local) + 28, sizeof(w), cudaMemcpyDeviceToHost);
  if (e != cudaSuccess) return fail(EGV_ERR_CUDA, "p2p_error: %s", cudaGetErrorString(e));
  *err = (int)w;
  return EGV_OK;
}
