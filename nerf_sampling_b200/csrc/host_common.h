// Host-side plumbing shared by the translation units of libb200nerf.so: error string, launch counter, macros.
#pragma once
#include <atomic>
#include <cuda_runtime.h>

int b200_fail(const char* fmt, ...);
extern std::atomic<unsigned long long> g_b200_launches;

#define CUDA_TRY(expr)                                                                                   \
  do {                                                                                                   \
    cudaError_t e__ = (expr);                                                                            \
    if (e__ != cudaSuccess)                                                                              \
      return b200_fail("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e__), __FILE__, __LINE__);     \
  } while (0)
#define LAUNCH_CHECK()            \
  do {                            \
    ++g_b200_launches;            \
    CUDA_TRY(cudaGetLastError()); \
  } while (0)
