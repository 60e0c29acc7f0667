// Host-side helpers shared by the C-ABI translation units: error reporting, launch counting.
#pragma once
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>

#include <atomic>

#include "../../include/egovlp_b200.h"

namespace egv {

extern thread_local char g_err[512];
extern std::atomic<long long> g_launches;

inline int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

inline int check_launch(const char* what) {
  g_launches.fetch_add(1, std::memory_order_relaxed);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail(EGV_ERR_CUDA, "%s: %s", what, cudaGetErrorString(e));
  return EGV_OK;
}

// Every kernel of the library is launched through launch_k: programmatic dependent launch (see pdl_wait in common.cuh) so
// that the launch latency and the prologue of kernel i+1 overlap the tail of kernel i, also inside captured CUDA graphs
// (the capture records a programmatic dependency edge).  EGV_PDL=0 turns the attribute off (plain stream order).
// EGV_PDL is a bit mask over kernel classes (the translation unit defines EGV_PDL_CLASS before including this header):
// 1 = gemm.cu, 2 = xgemm.cu, 4 = attention*.cu, 8 = layernorm.cu, 16 = everything else.
#ifndef EGV_PDL_CLASS
#define EGV_PDL_CLASS 16
#endif
#ifndef EGV_PDL_DEFAULT
#define EGV_PDL_DEFAULT 31
#endif
inline bool pdl_enabled() {
  static int mask = -1;
  if (mask < 0) {
    const char* e = getenv("EGV_PDL");
    mask = e ? atoi(e) : EGV_PDL_DEFAULT;
  }
  return (mask & EGV_PDL_CLASS) != 0;
}

template <typename... KArgs, typename... Args>
inline cudaError_t launch_k(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}

// the same for a kernel of thread-block clusters of `cluster` CTAs
template <typename... KArgs, typename... Args>
inline cudaError_t launch_cluster_k(void (*kern)(KArgs...), int cluster, dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                                    Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = (unsigned)cluster;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 2 : 1;
  return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}

inline int sm_count() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}

inline long long cdiv(long long a, long long b) { return (a + b - 1) / b; }

}  // namespace egv
