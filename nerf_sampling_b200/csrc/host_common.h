// Host-side plumbing shared by the translation units of libb200nerf.so: error string, launch counter, macros.
#pragma once
#include <atomic>
#include <cuda_runtime.h>

int b200_fail(const char* fmt, ...);
extern std::atomic<unsigned long long> g_b200_launches;

// One-time kernel setup (cudaFuncSetAttribute, occupancy-derived grid caps, SM counts) is cached PER DEVICE ORDINAL: function
// attributes and occupancy are per-device state, and the host API accepts tensors on any device of the process.  The caches are
// plain arrays written with idempotent values, so concurrent first calls from two threads are benign.
constexpr int B200_MAX_DEVICES = 64;
static inline int b200_device() {
  int d = 0;
  if (cudaGetDevice(&d) != cudaSuccess || d < 0 || d >= B200_MAX_DEVICES) return -1;
  return d;
}

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
