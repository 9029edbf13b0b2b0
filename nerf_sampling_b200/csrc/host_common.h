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

// ---- DepthNet's activated cat layers of the training step on the fused split-precision MLP kernel (b200nerf.cu; called by train.cu) ----
// n_layers LeakyReLU layers of 256 x 256 behind cat_layers.0 (+ the depth head for the forward).  The weights change every step, so
// the bf16 hi / lo images the kernel streams (forward: W_j, backward: W_j^T) and the fp32 bias / head block are rebuilt on the device.
constexpr int B200_CATCHAIN_MAX_LAYERS = 11;
size_t b200_catchain_img_bytes(int n_layers);      // one image (forward or transposed)
size_t b200_catchain_aux_floats();
size_t b200_catchain_mask_words(int n_layers, int n_rows);   // 64-bit words
// W[j], b[j]: device pointers of the n_layers weight [256,256] / bias [256] tensors; head_w [256], head_b [1]; in_bias [256] or null:
// the bias of the layer in FRONT of the chain, for b200_catchain_fwd(in_is_preact = true)
int b200_catchain_pack(const float* const* W, const float* const* b, const float* head_w, const float* head_b, const float* in_bias,
                       int n_layers, void* img_fwd, void* img_jac, float* aux, cudaStream_t st);
// in_act [n,256] -> save[j] = activations of layer j [n,256] (j < n_layers), out_z / out_s [n], masks.  in_is_preact: in_act holds the
// pre-activation sums of the layer in front (no bias yet); the kernel's loader adds the bias, applies LeakyReLU and writes the rows back
int b200_catchain_fwd(const void* img_fwd, const float* aux, int n_layers, float* in_act, bool in_is_preact, int n_rows, float near_,
                      float far_, float* const* save, unsigned long long* mask, float* out_z, float* out_s, cudaStream_t st);
// J_last [n,256] = d z / d(pre-activation of the last layer); save[t] = J of layer n_layers - 1 - t's input side, t = 0 .. n_layers - 1
int b200_catchain_jac(const void* img_jac, const float* aux, int n_layers, const float* s, int n_rows, float near_, float far_,
                      float* j_last, float* const* save, const unsigned long long* mask, cudaStream_t st);
