// UMMA self-test kernel: pins the K-major no-swizzle operand layout, the shared-memory descriptors and the TMEM read-back that the
// MLP kernels (mlp_exact.cuh, mlp_fast.cuh) rely on, independently of any network.
#pragma once
#include <cstdint>

#include "ptx.cuh"

namespace b200 {

constexpr int TILE_M = 128;
constexpr int ACT_KC_STRIDE = 2048;   // bytes between 16-byte K chunks of a 128-row operand: 16 row groups x 128 B

// ---------------------------------------------------------------------------------------------
// UMMA self-test: D[128,N] = A[128,K] * B[N,K]^T through exactly the operand layouts, descriptors and
// TMEM read-back used above.  Lets a GPU test pin the layout assumptions independently of the chain.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128, 1)
umma_selftest_kernel(const uint16_t* __restrict__ A, const uint16_t* __restrict__ B, float* __restrict__ D, int K, int N) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t done_bar;
  __shared__ uint32_t tmem_slot;
  const int kc_n = K / 8;
  uint8_t* sa = smem;                                  // [kc][16 row groups][8][16 B]
  uint8_t* sb = smem + kc_n * ACT_KC_STRIDE;           // [kc][N/8 row groups][8][16 B]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < TILE_M * K; i += blockDim.x) {
    const int r = i / K, k = i % K;
    *reinterpret_cast<uint16_t*>(sa + (k >> 3) * ACT_KC_STRIDE + (r >> 3) * 128 + (r & 7) * 16 + (k & 7) * 2) = A[i];
  }
  for (int i = threadIdx.x; i < N * K; i += blockDim.x) {
    const int r = i / K, k = i % K;
    *reinterpret_cast<uint16_t*>(sb + (k >> 3) * (N * 16) + (r >> 3) * 128 + (r & 7) * 16 + (k & 7) * 2) = B[i];
  }
  if (threadIdx.x == 0) {
    mbar_init(&done_bar, 1);
    fence_mbar_init();
  }
  if (warp == 0) {
    tmem_alloc(&tmem_slot, 256);
    tmem_relinquish();
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_slot;
  if (threadIdx.x == 0) {
    const uint32_t idesc = umma_idesc_bf16(TILE_M, N);
    for (int k = 0; k < K / 16; ++k) {
      const uint64_t da = umma_desc_kmajor(smem_u32(sa) + k * 2 * ACT_KC_STRIDE, ACT_KC_STRIDE, 128);
      const uint64_t db = umma_desc_kmajor(smem_u32(sb) + k * 2 * (N * 16), N * 16, 128);
      tc_mma_bf16(tmem_base, da, db, idesc, k > 0 ? 1u : 0u);
    }
    tc_commit(&done_bar);
  }
  mbar_wait(&done_bar, 0);
  tc_fence_after();
  const int row = warp * 32 + lane;
  for (int c = 0; c < N; c += 32) {
    uint32_t v[32];
    tmem_ld_32x32b_x32(tmem_base + (static_cast<uint32_t>(warp * 32) << 16) + c, v);
    tmem_ld_wait();
    for (int j = 0; j < 32; ++j)
      if (c + j < N) D[row * N + c + j] = __uint_as_float(v[j]);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, 256);
}

}  // namespace b200
