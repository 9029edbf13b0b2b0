// Host-side plumbing shared by the translation units of libb200nerf.so: error string, launch counter, macros.
#pragma once
#include <atomic>
#include <cstdlib>
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
// B200NERF_SYNC_CHECK=1 (debugging): synchronise the device after every launch of the library so that an asynchronous fault is
// reported at the launch that caused it (file:line in b200nerf_last_error) instead of at some later call.
static inline bool b200_sync_check() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("B200NERF_SYNC_CHECK");
    v = (e && e[0] == '1') ? 1 : 0;
  }
  return v == 1;
}
#define LAUNCH_CHECK()                                        \
  do {                                                        \
    ++g_b200_launches;                                        \
    CUDA_TRY(cudaGetLastError());                             \
    if (b200_sync_check()) CUDA_TRY(cudaDeviceSynchronize()); \
  } while (0)
