// raw2outputs (trainers/sampling_trainer.py:153-230; raw2alpha nerf_utils.py:27-42) for S = 32 / 64 / 128, staged through
// shared memory by TMA.
//
// composite_kernel (b200nerf.cu) keeps its loads in registers, so the bytes in flight are tied to occupancy: at 64
// registers it holds 80 KB per SM in flight only while a warp is in its load phase, and ncu shows it bound by SM cycles
// (58 % issue, 407k cycles at any clock) rather than by DRAM.  Here a persistent CTA owns a 2-stage ring of 40 KB
// tiles (2048 samples: 32 KB of raw as one 2-D TMA box with the 128-byte swizzle, 8 KB of z as one bulk copy), so
// 160 KB per SM are in flight at all times whatever the warps are doing, and the instruction count per sample drops:
// no global address arithmetic, colour sigmoids as ex2.approx.ftz + rcp.approx.ftz (4 instructions), the four map
// reductions as a value-splitting butterfly (7 shuffles instead of 16).
//
// Thread t of the CTA owns samples 4t..4t+3 of the tile, i.e. 64 contiguous bytes of raw = half a 128-byte row.
// Linear shared memory would make the four float4 reads of a quarter-warp collide 4-way (stride 64 B); the 128-byte
// swizzle (16-byte chunk index ^= row & 7) spreads them over all eight bank groups, and costs one XOR per load.
#pragma once
#include <cstdint>

#include "ptx.cuh"

namespace b200 {
namespace comp {

constexpr int THREADS = 512;
constexpr int TILE_SAMPLES = 4 * THREADS;            // 2048
constexpr int RAW_BYTES = TILE_SAMPLES * 16;         // 32 KB: 256 rows of 128 B
constexpr int Z_BYTES = TILE_SAMPLES * 4;            // 8 KB
constexpr int STAGE_BYTES = RAW_BYTES + Z_BYTES;     // 40 KB (a multiple of 1024: every stage keeps the swizzle alignment)
constexpr int STAGES = 2;
constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /* alignment slack */ + 64 /* barriers */;

struct alignas(64) TMap {
  uint8_t bytes[128];
};

__device__ __forceinline__ float sigmoid_approx(float x) {
  float e, r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(x * -1.4426950408889634f));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(1.0f + e));
  return r;
}

// LPR lanes per ray, four samples per lane: S == 4 * LPR.
template <int LPR>
__global__ void __launch_bounds__(THREADS, 2)
composite_tma_kernel(const __grid_constant__ TMap tm_raw, const float* __restrict__ z, const float* __restrict__ rays_d,
                     int n_rays, int white, float* __restrict__ o_rgb, float* __restrict__ o_disp, float* __restrict__ o_acc,
                     float* __restrict__ o_depth, float* __restrict__ o_w, float* __restrict__ o_alpha, int rgb_stride,
                     int disp_stride) {
  constexpr int S = 4 * LPR;
  static_assert(LPR == 8 || LPR == 16 || LPR == 32, "S must be 32, 64 or 128");
  extern __shared__ uint8_t smem_raw_[];
  const uint32_t smem0 = (smem_u32(smem_raw_) + 1023u) & ~1023u;
  uint8_t* smem = smem_raw_ + (smem0 - smem_u32(smem_raw_));
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES);
  uint64_t* empty = full + STAGES;

  const int tid = threadIdx.x;
  const int lane = tid & 31;
  const int lig = tid & (LPR - 1);
  const long long total = static_cast<long long>(n_rays) * S;
  const int num_tiles = static_cast<int>((total + TILE_SAMPLES - 1) / TILE_SAMPLES);

  auto issue = [&](int tile, int stage) {
    const long long s0 = static_cast<long long>(tile) * TILE_SAMPLES;
    const long long left = (total - s0) * 4;
    const uint32_t zbytes = left < Z_BYTES ? static_cast<uint32_t>(left) : static_cast<uint32_t>(Z_BYTES);
    // rows of the raw box beyond the tensor are zero-filled and still counted in the transaction bytes
    mbar_arrive_expect_tx(&full[stage], RAW_BYTES + zbytes);
    tma_load_2d(smem0 + stage * STAGE_BYTES, &tm_raw, 0, tile * (RAW_BYTES / 128), smem_u32(&full[stage]));
    tma_load_1d(smem + stage * STAGE_BYTES + RAW_BYTES, z + s0, zbytes, &full[stage]);
  };

  if (tid == 0) {
    tma_prefetch_desc(&tm_raw);
#pragma unroll
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], THREADS / 32);
    }
    fence_mbar_init();
  }
  __syncthreads();
  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < STAGES; ++s) {
      const int tile = blockIdx.x + s * gridDim.x;
      if (tile < num_tiles) issue(tile, s);
    }
  }

  // this thread's half row inside a stage: row = tid / 2, chunks (tid & 1) * 4 + i, swizzled by row & 7
  const uint32_t row_off = static_cast<uint32_t>(tid >> 1) * 128u;
  const uint32_t chunk0 = (static_cast<uint32_t>(tid & 1) * 4u) ^ (static_cast<uint32_t>(tid >> 1) & 7u);

  int k = 0;
  for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++k) {
    const int stage = k % STAGES;
    const uint32_t phase = static_cast<uint32_t>(k / STAGES) & 1u;
    const long long g0 = static_cast<long long>(tile) * TILE_SAMPLES + 4 * tid;
    int ray = static_cast<int>(g0 / S);
    const bool live = ray < n_rays;
    if (!live) ray = n_rays - 1;
    const float dx = __ldg(rays_d + ray * 3), dy = __ldg(rays_d + ray * 3 + 1), dz = __ldg(rays_d + ray * 3 + 2);

    mbar_wait(&full[stage], phase);
    const uint8_t* st = smem + stage * STAGE_BYTES;
    float4 rw[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) rw[i] = *reinterpret_cast<const float4*>(st + row_off + ((chunk0 ^ static_cast<uint32_t>(i)) << 4));
    const float4 z4 = *reinterpret_cast<const float4*>(st + RAW_BYTES + tid * 16);
    __syncwarp();
    if (lane == 0) mbar_arrive(&empty[stage]);
    if (tid == 0) {
      const int next = tile + STAGES * gridDim.x;
      if (next < num_tiles) {
        mbar_wait(&empty[stage], phase);   // all 16 warps hold this tile in registers
        issue(next, stage);
      }
    }

    const float nrm = __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz)));
    float zz[5] = {z4.x, z4.y, z4.z, z4.w, 0.f};
    zz[4] = __shfl_down_sync(0xffffffffu, z4.x, 1, LPR);   // the next lane's first depth (unused on the ray's last lane)

    float alpha[4];
    double ex[4];   // exclusive product inside this lane's four samples
    double run = 1.0;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      ex[i] = run;
      const bool last = i == 3 && lig == LPR - 1;
      const float dist = __fmul_rn(last ? 1e10f : __fadd_rn(zz[i + 1], -zz[i]), nrm);
      alpha[i] = __fadd_rn(1.0f, -expf(__fmul_rn(-fmaxf(rw[i].w, 0.f), dist)));
      run *= static_cast<double>(__fadd_rn(__fadd_rn(1.0f, -alpha[i]), 1e-10f));
    }
    // segmented inclusive scan of the lane products over the ray's LPR lanes, in double like the CPU reference's cumprod
    double incl = run;
#pragma unroll
    for (int off = 1; off < LPR; off <<= 1) {
      const double t = __shfl_up_sync(0xffffffffu, incl, off, LPR);
      if (lig >= off) incl *= t;
    }
    double excl = __shfl_up_sync(0xffffffffu, incl, 1, LPR);
    if (lig == 0) excl = 1.0;

    float w[4];
    float s_r = 0.f, s_g = 0.f, s_b = 0.f, s_d = 0.f, s_a = 0.f;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      w[i] = __fmul_rn(alpha[i], static_cast<float>(excl * ex[i]));
      s_r = fmaf(w[i], sigmoid_approx(rw[i].x), s_r);
      s_g = fmaf(w[i], sigmoid_approx(rw[i].y), s_g);
      s_b = fmaf(w[i], sigmoid_approx(rw[i].z), s_b);
      s_d = fmaf(w[i], zz[i], s_d);
      s_a += w[i];
    }
    if (live) {
      if (o_w) *reinterpret_cast<float4*>(o_w + g0) = make_float4(w[0], w[1], w[2], w[3]);
      if (o_alpha) *reinterpret_cast<float4*>(o_alpha + g0) = make_float4(alpha[0], alpha[1], alpha[2], alpha[3]);
    }

    // acc: plain butterfly (every lane needs it).  r, g, b, depth: each step keeps half of the values and ships the other
    // half, so the lanes with (lig & O1, lig & O2) = (0,0) / (1,0) / (0,1) / (1,1) end with the ray's r / g / b / depth.
#pragma unroll
    for (int off = LPR >> 1; off > 0; off >>= 1) s_a += __shfl_xor_sync(0xffffffffu, s_a, off, LPR);
    constexpr int O1 = LPR / 2, O2 = LPR / 4;
    const bool h1 = (lig & O1) != 0, h2 = (lig & O2) != 0;
    float a_keep = h1 ? s_g : s_r, a_send = h1 ? s_r : s_g;
    float b_keep = h1 ? s_d : s_b, b_send = h1 ? s_b : s_d;
    a_keep += __shfl_xor_sync(0xffffffffu, a_send, O1, LPR);
    b_keep += __shfl_xor_sync(0xffffffffu, b_send, O1, LPR);
    float v = h2 ? b_keep : a_keep;
    const float v_send = h2 ? a_keep : b_keep;
    v += __shfl_xor_sync(0xffffffffu, v_send, O2, LPR);
#pragma unroll
    for (int off = LPR >> 3; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off, LPR);

    if (live && (lig & (O2 - 1)) == 0) {
      if (h1 && h2) {
        if (o_depth) o_depth[ray] = v;
        if (o_disp) o_disp[ray * disp_stride] = __fdiv_rn(1.0f, fmaxf(1e-10f, __fdiv_rn(v, __fadd_rn(s_a, 1e-10f))));
      } else {
        const int ch = (h1 ? 1 : 0) + (h2 ? 2 : 0);
        if (o_rgb) o_rgb[ray * rgb_stride + ch] = white ? v + __fadd_rn(1.0f, -s_a) : v;
        if (ch == 0 && o_acc) o_acc[ray] = s_a;
      }
    }
  }
}

}  // namespace comp
}  // namespace b200
