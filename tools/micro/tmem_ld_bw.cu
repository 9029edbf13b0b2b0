// Micro-benchmark: TMEM read bandwidth seen by tcgen05.ld.32x32b.x32 with 4 or 8 warps per SM and 1 or 2 loads in flight.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tmem_ld_bw tools/micro/tmem_ld_bw.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../../nerf_sampling_b200/csrc/ptx.cuh"
using namespace b200;

template <int DEPTH>
__global__ void __launch_bounds__(512, 1) k(int iters, int nwarps, long long* out, uint32_t* sink) {
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) { tmem_alloc(&slot, 512); tmem_relinquish(); }
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t base = slot + (static_cast<uint32_t>((warp & 3) * 32) << 16);
  uint32_t acc = 0;
  __syncthreads();
  const long long t0 = clock64();
  if (warp < nwarps) {
    const uint32_t col0 = (warp >> 2) * 128;   // warps 0-3: columns 0..127, 4-7: 128..255, ...
    for (int it = 0; it < iters; ++it) {
      uint32_t va[32], vb[32];
      if (DEPTH == 1) {
        for (int c = 0; c < 4; ++c) {
          tmem_ld_32x32b_x32(base + ((col0 + c * 32) & 511), va);
          tmem_ld_wait();
          acc += va[0] + va[31];
        }
      } else {
        for (int c = 0; c < 4; c += 2) {
          tmem_ld_32x32b_x32(base + ((col0 + c * 32) & 511), va);
          tmem_ld_32x32b_x32(base + ((col0 + c * 32 + 32) & 511), vb);
          tmem_ld_wait();
          acc += va[0] + vb[31];
        }
      }
    }
  }
  const long long t1 = clock64();
  __syncthreads();
  if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
  if (acc == 0x12345678u) sink[threadIdx.x] = acc;
  tc_fence_before(); __syncthreads();
  if (warp == 0) tmem_dealloc(slot, 512);
}

int main() {
  long long* d; uint32_t* s;
  cudaMalloc(&d, 148 * 8); cudaMalloc(&s, 4096);
  const int iters = 2000;
  for (int depth = 1; depth <= 2; ++depth)
    for (int nw : {4, 8, 16}) {
      if (depth == 1) k<1><<<148, 512>>>(iters, nw, d, s); else k<2><<<148, 512>>>(iters, nw, d, s);
      cudaDeviceSynchronize();
      long long h[148];
      cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
      const double bytes = double(iters) * nw * 4 * 4096;   // per SM
      printf("depth %d warps %2d: %lld cycles, %.1f B/cycle/SM, %.1f cycles per 4 KB load per warp  (%s)\n", depth, nw, h[0],
             bytes / h[0], double(h[0]) / (iters * 4), cudaGetErrorString(cudaGetLastError()));
    }
  return 0;
}
