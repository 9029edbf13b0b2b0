// Fused MLP chain on tcgen05 tensor cores (sm_100a).
//
// One persistent CTA per SM walks 128-row tiles.  For each tile the whole layer chain runs without the
// activations ever leaving the SM:
//
//   prologue (8 warps)  : build the network input for 128 rows (positional encoding) straight into the
//                         canonical K-major UMMA operand layout in shared memory, as bf16 hi + bf16 lo planes
//   warp 0 (1 lane)     : streams pre-packed weight slabs (one K16 slab = N x 16 bf16 hi [+ lo]) from L2 with
//                         1-D bulk TMA through an mbarrier ring
//   warp 1 (1 lane)     : issues tcgen05.mma (M=128, N<=256, K=16) into a TMEM accumulator; split precision
//                         = 3 MMAs per K16 block (Ahi*Whi + Ahi*Wlo + Alo*Whi) so the result keeps ~16
//                         mantissa bits with fp32 accumulation
//   epilogue (8 warps)  : tcgen05.ld the accumulator, + bias, activation, re-split to bf16 hi/lo and write the
//                         next layer's A operand in place; heads (N=1 / N=3) are fp32 dot products here
//
// Reference semantics: NeRF.forward (nerf_pytorch/run_nerf_helpers.py:109-134) with the encoding of
// Trainer.run_network (nerf_pytorch/trainers/Trainer.py:789-806) fused in front, and DepthNet.forward
// (depth_nets/depth_net.py:117-169) with its activation-free branches folded into the first layer.
#pragma once
#include "ptx.cuh"

namespace b200 {

constexpr int TILE_M = 128;
constexpr int ACT_KC = 36;                         // 16-byte K chunks (8 bf16) the activation buffer holds: K <= 288
constexpr int ACT_KC_STRIDE = 2048;                // bytes between K chunks: 16 row groups x 128 B
constexpr int ACT_PLANE_BYTES = ACT_KC * ACT_KC_STRIDE;  // 73,728 B per bf16 plane (hi or lo)
constexpr int W_STAGE_BYTES = 16384;               // one K16 slab, N=256: hi 8 KB + lo 8 KB
constexpr int NUM_STAGES = 5;
constexpr int CHAIN_THREADS = 384;                 // warps 0-3: TMA / MMA / TMEM alloc / idle, warps 4-11: epilogue
constexpr int EPI_THREADS = 256;
constexpr int MAX_STEPS = 12;
constexpr int TMEM_COLS = 512;

enum : uint8_t { ACT_NONE = 0, ACT_RELU = 1, ACT_LEAKY = 2 };
enum : uint8_t {
  EPI_NONE = 0,        // no epilogue (accumulator is picked up by a later step)
  EPI_STORE = 1,       // bias + activation -> next layer's operand
  EPI_STORE_ALPHA = 2, // as EPI_STORE, plus the sigma head (N=1) as an fp32 dot product
  EPI_NERF_OUT = 3,    // bias + ReLU, rgb head (N=3), write raw [r,g,b,sigma]
  EPI_DEPTH_OUT = 4,   // bias + LeakyReLU, depth head (N=1), sigmoid, near/far scaling, write z
};
enum : int { IN_NERF = 0, IN_DEPTHNET = 1 };

struct Step {
  uint16_t a_k16_begin;  // first K16 block of the activation buffer this step reads
  uint16_t n_k16;        // K16 blocks in this step
  uint16_t n;            // output features (multiple of 16, <= 256)
  uint16_t acc_col;      // TMEM column of the accumulator
  uint8_t accumulate;    // first MMA adds onto what the accumulator already holds (skip connection)
  uint8_t wait_a;        // operand buffer was rewritten since the previous step: wait for it
  uint8_t epi;           // epilogue kind; != EPI_NONE also means "commit to the epilogue after this step"
  uint8_t act;
  uint32_t bias_off;     // float offset of this step's bias inside `aux`
};

struct ChainParams {
  const uint8_t* wpack;  // weight slabs in streaming order
  const float* aux;      // fp32 biases and head weights
  Step steps[MAX_STEPS];
  int n_steps;
  int n_rows;            // points (NeRF) or rays (DepthNet)
  int S;                 // samples per ray (NeRF)
  const float* rays_o;   // [rays,3]
  const float* rays_d;   // [rays,3]
  const float* viewdirs; // [rays,3]
  const float* z;        // [rays,S]   (NeRF: depth of every sample)  or nullptr
  const float* pts;      // [rows,3]   (NeRF: explicit sample positions) or nullptr
  float* out;            // raw [rows,4] (NeRF) or z [rows] (DepthNet)
  const int* row_index;  // NeRF only, optional: row i of this launch is sample point row_index[i] (input and output)
  const int* n_rows_dev; // optional: the row count lives on the device (capped by n_rows); used with row_index
  uint32_t head_w_off, head_b_off;    // sigma head (NeRF) / depth head (DepthNet)
  uint32_t rgb_w_off, rgb_b_off;      // rgb head [3,128]
  float radius, near, far;
};

struct __align__(16) ChainSmemTail {
  uint64_t full[NUM_STAGES];
  uint64_t empty[NUM_STAGES];
  uint64_t a_ready;
  uint64_t acc_full;
  uint32_t tmem_base;
  uint32_t pad[3];
  float head_part[2][TILE_M];  // per column-half partial of the N=1 head
};

template <bool SPLIT>
constexpr int chain_smem_bytes() {
  return (SPLIT ? 2 : 1) * ACT_PLANE_BYTES + NUM_STAGES * W_STAGE_BYTES + (int)sizeof(ChainSmemTail);
}

// ---------------------------------------------------------------------------------------------
// operand writes
// ---------------------------------------------------------------------------------------------
template <bool SPLIT>
__device__ __forceinline__ void store_chunk8(uint8_t* act, int kc, int row_off, const float (&x)[8]) {
  uint32_t h[4], l[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    h[i] = pack_bf16x2(x[2 * i], x[2 * i + 1]);
    if (SPLIT) {
      const float r0 = x[2 * i] - __uint_as_float(h[i] << 16);
      const float r1 = x[2 * i + 1] - __uint_as_float(h[i] & 0xffff0000u);
      l[i] = pack_bf16x2(r0, r1);
    }
  }
  uint8_t* p = act + kc * ACT_KC_STRIDE + row_off;
  *reinterpret_cast<uint4*>(p) = make_uint4(h[0], h[1], h[2], h[3]);
  if (SPLIT) *reinterpret_cast<uint4*>(p + ACT_PLANE_BYTES) = make_uint4(l[0], l[1], l[2], l[3]);
}

// Columns [C0,C1) of the positional encoding [x, sin(2^0 x), cos(2^0 x), ..., cos(2^(NFREQ-1) x)] of a
// 3-vector (run_nerf_helpers.py:15-63), zero beyond 3+6*NFREQ, written as K chunks kc0.. of the operand.
template <bool SPLIT, int NFREQ, int C0, int C1>
__device__ __forceinline__ void embed_store(const float (&x)[3], uint8_t* act, int kc0, int row_off) {
  constexpr int NCOL = 3 + 6 * NFREQ;
  constexpr int JLO = (C0 <= 3) ? 0 : (C0 - 3) / 6;
  constexpr int JHI_ = (C1 - 1 < 3) ? -1 : (C1 - 1 - 3) / 6;
  constexpr int JHI = JHI_ < NFREQ ? JHI_ : NFREQ - 1;
  constexpr int NJ = (JHI - JLO + 1) > 0 ? (JHI - JLO + 1) : 1;
  float sn[NJ][3], cs[NJ][3];
#pragma unroll
  for (int j = JLO; j <= JHI; ++j) {
    const float f = static_cast<float>(1 << j);
#pragma unroll
    for (int t = 0; t < 3; ++t) sincosf(x[t] * f, &sn[j - JLO][t], &cs[j - JLO][t]);
  }
#pragma unroll
  for (int c = C0; c < C1; c += 8) {
    float v[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int cc = c + i;
      if (cc < 3) {
        v[i] = x[cc];
      } else if (cc < NCOL) {
        const int j = (cc - 3) / 6, t = (cc - 3) % 6;
        v[i] = t < 3 ? sn[j - JLO][t] : cs[j - JLO][t - 3];
      } else {
        v[i] = 0.f;
      }
    }
    store_chunk8<SPLIT>(act, kc0 + (c - C0) / 8, row_off, v);
  }
}

__device__ __forceinline__ float apply_act(float x, uint8_t act) {
  if (act == ACT_RELU) return fmaxf(x, 0.f);
  if (act == ACT_LEAKY) return x >= 0.f ? x : x * 0.01f;
  return x;
}

// ---------------------------------------------------------------------------------------------
// the kernel
// ---------------------------------------------------------------------------------------------
template <bool SPLIT, int INPUT>
__global__ void __launch_bounds__(CHAIN_THREADS, 1) mlp_chain_kernel(const __grid_constant__ ChainParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  constexpr int ACT_BYTES = (SPLIT ? 2 : 1) * ACT_PLANE_BYTES;
  uint8_t* act = smem;
  uint8_t* wst = smem + ACT_BYTES;
  ChainSmemTail* tail = reinterpret_cast<ChainSmemTail*>(smem + ACT_BYTES + NUM_STAGES * W_STAGE_BYTES);

  const int warp = __shfl_sync(0xffffffffu, static_cast<int>(threadIdx.x >> 5), 0);   // warp-uniform for the compiler
  const int lane = threadIdx.x & 31;
  int n_rows = p.n_rows;
  if (p.n_rows_dev != nullptr) n_rows = min(__ldg(p.n_rows_dev), p.n_rows);
  const int num_tiles = (n_rows + TILE_M - 1) / TILE_M;

  if (threadIdx.x == 0) {
    for (int i = 0; i < NUM_STAGES; ++i) {
      mbar_init(&tail->full[i], 1);
      mbar_init(&tail->empty[i], 1);
    }
    mbar_init(&tail->a_ready, EPI_THREADS);
    mbar_init(&tail->acc_full, 1);
    fence_mbar_init();
  }
  if (warp == 2) {
    tmem_alloc(&tail->tmem_base, TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tail->tmem_base;

  const uint32_t full_addr = smem_u32(&tail->full[0]), empty_addr = smem_u32(&tail->empty[0]);
  if (warp == 0) {
    // ===================================================================== weight producer (TMA; whole warp, one lane issues)
    uint32_t stage = 0, phase = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const uint8_t* src = p.wpack;
      for (int s = 0; s < p.n_steps; ++s) {
        const uint32_t bytes = p.steps[s].n * (SPLIT ? 64u : 32u);
        const int nk = p.steps[s].n_k16;
        for (int k = 0; k < nk; ++k) {
          mbar_wait_lean(empty_addr + stage * 8u, phase ^ 1u);
          if (elect_one()) {
            mbar_arrive_expect_tx(&tail->full[stage], bytes);
            tma_load_1d(wst + stage * W_STAGE_BYTES, src, bytes, &tail->full[stage]);
          }
          __syncwarp();
          src += bytes;
          if (++stage == NUM_STAGES) {
            stage = 0;
            phase ^= 1u;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================================================================== MMA issuer (whole warp, one lane issues)
    uint32_t stage = 0, phase = 0, a_cnt = 0;
    const uint32_t act_addr = smem_u32(act);
    const uint32_t wst_addr = smem_u32(wst);
    const uint32_t a_ready_addr = smem_u32(&tail->a_ready);
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      for (int s = 0; s < p.n_steps; ++s) {
        const Step st = p.steps[s];
        if (st.wait_a) {
          mbar_wait_lean(a_ready_addr, a_cnt & 1u);
          ++a_cnt;
        }
        const uint32_t idesc = umma_idesc_bf16(TILE_M, st.n);
        const uint32_t d_tmem = tmem_base + st.acc_col;
        const uint32_t b_lbo = st.n * 16u;  // bytes between the two K core-matrix columns of a slab
        uint32_t a_addr = act_addr + st.a_k16_begin * 2u * ACT_KC_STRIDE;
        for (int k = 0; k < st.n_k16; ++k) {
          mbar_wait_lean(full_addr + stage * 8u, phase);
          tc_fence_after();
          if (elect_one()) {
            const uint32_t b_addr = wst_addr + stage * W_STAGE_BYTES;
            const uint64_t a_hi = umma_desc_kmajor(a_addr, ACT_KC_STRIDE, 128);
            const uint64_t b_hi = umma_desc_kmajor(b_addr, b_lbo, 128);
            tc_mma_bf16(d_tmem, a_hi, b_hi, idesc, (k > 0 || st.accumulate) ? 1u : 0u);
            if (SPLIT) {
              const uint64_t a_lo = umma_desc_kmajor(a_addr + ACT_PLANE_BYTES, ACT_KC_STRIDE, 128);
              const uint64_t b_lo = umma_desc_kmajor(b_addr + st.n * 32u, b_lbo, 128);
              tc_mma_bf16(d_tmem, a_hi, b_lo, idesc, 1u);
              tc_mma_bf16(d_tmem, a_lo, b_hi, idesc, 1u);
            }
            tc_commit(&tail->empty[stage]);  // slab consumed once these MMAs retire
          }
          __syncwarp();
          a_addr += 2u * ACT_KC_STRIDE;
          if (++stage == NUM_STAGES) {
            stage = 0;
            phase ^= 1u;
          }
        }
        if (st.epi != EPI_NONE) {
          if (elect_one()) tc_commit(&tail->acc_full);
          __syncwarp();
        }
      }
    }
  } else if (warp >= 4) {
    // ===================================================================== prologue + epilogue warps
    const int q = warp & 3;            // TMEM lane quarter this warp may access
    const int hsel = (warp - 4) >> 2;  // which half of the columns this thread owns
    const int row = q * 32 + lane;
    const int row_off = (row >> 3) * 128 + (row & 7) * 16;
    const uint32_t t_lane = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
    uint32_t acc_cnt = 0;

    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const bool valid = tile * TILE_M + row < n_rows;
      // sample point this row evaluates (identity unless the launch works through an index list)
      const int grow = (INPUT == IN_NERF && p.row_index != nullptr) ? (valid ? __ldg(p.row_index + tile * TILE_M + row) : 0)
                                                                    : tile * TILE_M + row;

      // ---------------------------------------------------------------- prologue: network input
      if (INPUT == IN_NERF) {
        const int ray = valid ? grow / p.S : 0;
        if (hsel == 0) {
          float x[3] = {0.f, 0.f, 0.f};
          if (valid) {
            if (p.pts != nullptr) {
#pragma unroll
              for (int t = 0; t < 3; ++t) x[t] = __ldg(p.pts + (size_t)grow * 3 + t);
            } else {
              const float zz = __ldg(p.z + grow);
#pragma unroll
              for (int t = 0; t < 3; ++t)  // o + d*z, product and sum rounded separately like torch
                x[t] = __fadd_rn(__ldg(p.rays_o + ray * 3 + t), __fmul_rn(__ldg(p.rays_d + ray * 3 + t), zz));
            }
          }
          embed_store<SPLIT, 10, 0, 32>(x, act, 0, row_off);
        } else {
          float x[3] = {0.f, 0.f, 0.f}, v[3] = {0.f, 0.f, 0.f};
          if (valid) {
            if (p.pts != nullptr) {
#pragma unroll
              for (int t = 0; t < 3; ++t) x[t] = __ldg(p.pts + (size_t)grow * 3 + t);
            } else {
              const float zz = __ldg(p.z + grow);
#pragma unroll
              for (int t = 0; t < 3; ++t)
                x[t] = __fadd_rn(__ldg(p.rays_o + ray * 3 + t), __fmul_rn(__ldg(p.rays_d + ray * 3 + t), zz));
            }
#pragma unroll
            for (int t = 0; t < 3; ++t) v[t] = __ldg(p.viewdirs + ray * 3 + t);
          }
          embed_store<SPLIT, 10, 32, 64>(x, act, 4, row_off);
          embed_store<SPLIT, 4, 0, 32>(v, act, 32, row_off);  // view encoding, K blocks 16..17
        }
      } else {
        // DepthNet: four 64-wide groups  enc(o) | enc(d) | enc(hit_near) | enc(hit_far)
        float o[3] = {0.f, 0.f, 0.f}, d[3] = {0.f, 0.f, 0.f};
        if (valid) {
#pragma unroll
          for (int t = 0; t < 3; ++t) {
            o[t] = __ldg(p.rays_o + (size_t)grow * 3 + t);
            d[t] = __ldg(p.rays_d + (size_t)grow * 3 + t);
          }
        }
        if (hsel == 0) {
          embed_store<SPLIT, 10, 0, 64>(o, act, 0, row_off);
          embed_store<SPLIT, 10, 0, 64>(d, act, 8, row_off);
        } else {
          // ray / sphere(0, radius) intersection, op order of nerf_pytorch/utils.py:159-217
          const float dot_do = __fadd_rn(__fadd_rn(__fmul_rn(d[0], o[0]), __fmul_rn(d[1], o[1])), __fmul_rn(d[2], o[2]));
          const float b = __fmul_rn(2.f, dot_do);
          const float on = __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(o[0], o[0]), __fmul_rn(o[1], o[1])), __fmul_rn(o[2], o[2])));
          const float c = __fadd_rn(__fmul_rn(on, on), -__fmul_rn(p.radius, p.radius));
          const float a = __fadd_rn(__fadd_rn(__fmul_rn(d[0], d[0]), __fmul_rn(d[1], d[1])), __fmul_rn(d[2], d[2]));
          const float delta = __fadd_rn(__fmul_rn(b, b), -__fmul_rn(__fmul_rn(4.f, a), c));
          const float sq = __fsqrt_rn(delta);  // NaN when the ray misses: propagates to the depth like the reference
          const float two_a = __fmul_rn(2.f, a);
          const float t0 = __fdiv_rn(__fadd_rn(-b, -sq), two_a);
          const float t1 = __fdiv_rn(__fadd_rn(-b, sq), two_a);
          float p0[3], p1[3];
#pragma unroll
          for (int t = 0; t < 3; ++t) {
            p0[t] = __fadd_rn(o[t], __fmul_rn(t0, d[t]));
            p1[t] = __fadd_rn(o[t], __fmul_rn(t1, d[t]));
          }
          embed_store<SPLIT, 10, 0, 64>(p0, act, 16, row_off);
          embed_store<SPLIT, 10, 0, 64>(p1, act, 24, row_off);
        }
      }
      fence_proxy_async_smem();
      tc_fence_before();
      mbar_arrive(&tail->a_ready);

      // ---------------------------------------------------------------- epilogues
      for (int s = 0; s < p.n_steps; ++s) {
        const Step st = p.steps[s];
        if (st.epi == EPI_NONE) continue;
        mbar_wait_lean(smem_u32(&tail->acc_full), acc_cnt & 1u);
        ++acc_cnt;
        tc_fence_after();

        const int ncols = st.n >> 1;       // columns owned by this thread
        const int col0 = hsel * ncols;
        const float* bias = p.aux + st.bias_off;
        const bool head1 = st.epi == EPI_STORE_ALPHA || st.epi == EPI_DEPTH_OUT;
        const bool rgbh = st.epi == EPI_NERF_OUT;
        const bool store = st.epi == EPI_STORE || st.epi == EPI_STORE_ALPHA;
        float hsum = 0.f, rs = 0.f, gs = 0.f, bs = 0.f;

        for (int c = 0; c < ncols; c += 32) {
          uint32_t v[32];
          tmem_ld_32x32b_x32(t_lane + st.acc_col + col0 + c, v);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 32; j += 8) {
            const int col = col0 + c + j;
            const float4 b0 = __ldg(reinterpret_cast<const float4*>(bias + col));
            const float4 b1 = __ldg(reinterpret_cast<const float4*>(bias + col + 4));
            float x[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
            for (int i = 0; i < 8; ++i) x[i] = apply_act(__uint_as_float(v[j + i]) + x[i], st.act);
            if (head1) {
              const float4 w0 = __ldg(reinterpret_cast<const float4*>(p.aux + p.head_w_off + col));
              const float4 w1 = __ldg(reinterpret_cast<const float4*>(p.aux + p.head_w_off + col + 4));
              hsum = fmaf(x[0], w0.x, hsum); hsum = fmaf(x[1], w0.y, hsum);
              hsum = fmaf(x[2], w0.z, hsum); hsum = fmaf(x[3], w0.w, hsum);
              hsum = fmaf(x[4], w1.x, hsum); hsum = fmaf(x[5], w1.y, hsum);
              hsum = fmaf(x[6], w1.z, hsum); hsum = fmaf(x[7], w1.w, hsum);
            }
            if (rgbh) {
              const float* wr = p.aux + p.rgb_w_off + col;
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                rs = fmaf(x[i], __ldg(wr + i), rs);
                gs = fmaf(x[i], __ldg(wr + 128 + i), gs);
                bs = fmaf(x[i], __ldg(wr + 256 + i), bs);
              }
            }
            if (store) store_chunk8<SPLIT>(act, col >> 3, row_off, x);
          }
        }

        if (st.epi == EPI_STORE_ALPHA) tail->head_part[hsel][row] = hsum;

        if (st.epi == EPI_NERF_OUT) {
          // operand buffer is dead after the last MMA: borrow it to combine the two column halves
          float* part = reinterpret_cast<float*>(act);
          if (hsel == 1) {
            part[row * 4 + 0] = rs;
            part[row * 4 + 1] = gs;
            part[row * 4 + 2] = bs;
          }
          named_bar_sync(1, EPI_THREADS);
          if (hsel == 0 && valid) {
            const float* rb = p.aux + p.rgb_b_off;
            float4 o4;
            o4.x = rs + part[row * 4 + 0] + __ldg(rb + 0);
            o4.y = gs + part[row * 4 + 1] + __ldg(rb + 1);
            o4.z = bs + part[row * 4 + 2] + __ldg(rb + 2);
            o4.w = tail->head_part[0][row] + tail->head_part[1][row] + __ldg(p.aux + p.head_b_off);
            // fmaxf(NaN, 0) = 0 but torch.relu(NaN) = NaN: a sample whose position or view direction is not finite
            // (a ray that missed the sphere has a NaN depth) yields NaN like the reference
            {
              const int ry = grow / p.S;
              float m = 0.f;
              if (p.pts != nullptr) {
#pragma unroll
                for (int t = 0; t < 3; ++t) m += fabsf(__ldg(p.pts + (size_t)grow * 3 + t));
              } else {
                const float zz = __ldg(p.z + grow);
#pragma unroll
                for (int t = 0; t < 3; ++t)
                  m += fabsf(__fadd_rn(__ldg(p.rays_o + ry * 3 + t), __fmul_rn(__ldg(p.rays_d + ry * 3 + t), zz)));
              }
#pragma unroll
              for (int t = 0; t < 3; ++t) m += fabsf(__ldg(p.viewdirs + ry * 3 + t));
              if (!(m < __int_as_float(0x7f800000))) o4.x = o4.y = o4.z = o4.w = __int_as_float(0x7fc00000);
            }
            reinterpret_cast<float4*>(p.out)[grow] = o4;
          }
          named_bar_sync(1, EPI_THREADS);  // partials consumed before the next prologue rewrites the buffer
        } else if (st.epi == EPI_DEPTH_OUT) {
          float* part = reinterpret_cast<float*>(act);
          if (hsel == 1) part[row] = hsum;
          named_bar_sync(1, EPI_THREADS);
          if (hsel == 0 && valid) {
            const float t = hsum + part[row] + __ldg(p.aux + p.head_b_off);
            const float sg = 1.0f / (1.0f + expf(-t));
            // near*(1-s) + far*s with every product / sum rounded separately (depth_net.py:168)
            p.out[grow] = __fadd_rn(__fmul_rn(p.near, __fadd_rn(1.0f, -sg)), __fmul_rn(p.far, sg));
          }
          named_bar_sync(1, EPI_THREADS);
        }

        if (store) {
          fence_proxy_async_smem();
          tc_fence_before();
          mbar_arrive(&tail->a_ready);
        } else {
          tc_fence_before();
        }
      }
    }
  }

  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_base, TMEM_COLS);
}

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
