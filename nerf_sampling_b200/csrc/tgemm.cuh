// Grouped, K-segmented fp32 GEMM on the tcgen05 tensor cores for the training slice (train.cu): 3xTF32 error-compensated.
//
// One problem:  C[M,N] (row-major, ldc) = (beta ? C : 0) + sum_seg sum_k A_s(m,k) * B_s(k,n) [+ bias[n]] [LeakyReLU] [* dact]
//   A_s(m,k) = A_s[m*sAm + k*sAk], B_s(k,n) = B_s[k*sBk + n*sBn]
// The strides cover X*W^T (Linear forward), dY*W (input gradient) and dY^T*X (weight gradient) without transposes; the K
// SEGMENTS are the pieces of a concatenated input (`Linear(cat([x, e]))`, depth_net.py:139-157, is two segments of one
// product: no cat, no second launch); a GROUP is up to nine independent problems in one launch (the three DepthNet
// branches advance together; a layer's weight gradients and its input gradient share a launch).  The CUDA-core kernel this
// replaces (sgemm_kernel) took 9.5 of the 12.9 ms of a 4096-ray training step, in ~320 launches.
//
// Precision: every fp32 operand x is split in registers into hi = x with the low 13 mantissa bits cleared (exactly a
// TF32 value) and lo = x - hi (exact in fp32); the tile product is accumulated in TMEM as  lo*hi + hi*lo + hi*hi  with
// `tcgen05.mma.kind::tf32`.  Relative error ~2^-20 per product (the dropped lo*lo and the hardware's truncation of lo):
// measured against fp64 it is within 4x of the fp32 FMA kernel's own rounding error, so the gradients stay fp32-grade.
//
// One CTA = one 128 x 64 output tile (optionally one K slice of it: split K for the weight gradients, which reduce over
// thousands of rays into a 256 x ~300 output; partial sums are added atomically into a pre-zeroed C).  K is walked in
// chunks of 32: all 256 threads load the A (128 x 32) and B (64 x 32) pieces with lanes along the operand's contiguous
// dimension -- one chunk AHEAD, into registers -- split them and store 16-byte k-quads into the K-major no-swizzle UMMA
// layout (core matrix = 8 rows x 16 B = 8 x 4 TF32; LBO padded by 16 B so that a warp's stores spread over all bank
// groups) and arrive on the stage's `full` barrier; a ninth warp waits for it, issues 4 K8 steps x 3 MMAs and commits to the
// stage's `mma_done` barrier.  Two smem stages: the loaders never wait for the MMA issue of the chunk they just stored.  Epilogue: tcgen05.ld (8 warps = 4 lane quarters x 2 column halves), + bias,
// LeakyReLU, optional multiplication by the LeakyReLU derivative of a saved activation (fuses the backward's activation
// step into the input-gradient product), store / atomicAdd.  `colsum` (weight-gradient problems): the bias gradient
// db[m] = sum_k A(m,k) falls out of the A loads for free.
#pragma once
#include <cstdint>

#include "ptx.cuh"

namespace b200 {
namespace tg {

constexpr int BM = 128, BN = 64, KC = 32, THREADS = 256;   // THREADS = the loader / epilogue warps
constexpr int CTA_THREADS = THREADS + 32;                   // + one warp that owns TMEM and issues the MMAs
constexpr int SBO = 144;                       // bytes between 8-row groups: 128 + 16, so that stores of rows 4l..4l+3 (the
                                               // 4 x 4 micro-tile mapping) and of consecutive rows both spread over all bank groups
constexpr int LBO_A = (BM / 8) * SBO + 16;     // bytes between consecutive k-quads (4 TF32 = 16 B) of the A plane: 2,320
constexpr int LBO_B = (BN / 8) * SBO + 16;     // 1,168
constexpr int PLANE_A = (KC / 4) * LBO_A;      // 18,560 B
constexpr int PLANE_B = (KC / 4) * LBO_B;      //  9,344 B
constexpr int STAGE_BYTES = 2 * PLANE_A + 2 * PLANE_B;   // hi + lo of both operands: 55,808 B
// fp32 staging ring between global memory and the split: every loader thread owns QA + QB + 1 16-byte slots per stage (its k-quads
// of A and B and its kscale factors), filled by cp.async, NSTG chunks deep
constexpr int NSTG = 3;
constexpr int STG_SLOTS = BM * (KC / 4) / THREADS + BN * (KC / 4) / THREADS + 1;   // 4 + 2 + 1
constexpr int STG_BYTES = STG_SLOTS * THREADS * 16;                                // 28,672 B
constexpr int SMEM_BYTES = 2 * STAGE_BYTES + NSTG * STG_BYTES + 128;
constexpr uint32_t TMEM_COLS = 64;
constexpr int MAX_SEG = 4, MAX_PROB = 16;
constexpr int QA = BM * (KC / 4) / THREADS;    // k-quads per thread and chunk: 4 of A ...
constexpr int QB = BN * (KC / 4) / THREADS;    // ... 2 of B

struct Seg {
  const float* A;
  const float* B;
  long sAm, sAk, sBk, sBn;
  int K;
  int vecA, vecB;   // k-contiguous operand with 16-byte aligned rows: float4 loads
  int pad_;
};
struct Prob {
  Seg seg[MAX_SEG];
  float* C;
  const float* bias;
  float* colsum;        // db[m] += sum_k A(m,k) (A must be k-strided, i.e. sAm == 1); pre-zeroed by the caller
  const float* dact;    // epilogue *= (dact[m*ld_dact + n] > 0 ? 1 : slope)
  const float* kscale;  // A_0(m,k) *= kscale[k] (segment 0 only): the per-ray factor of a weight gradient whose dY is dz[r] * J[r, :]
  int M, N, nseg, ldc, beta, act, ld_dact;
  float slope;
  int tiles_x, tiles_y, splits, k_per, cta_begin;   // splits > 1 only with nseg == 1; k_per a multiple of KC
};
struct Group {
  int nprob;
  long long* dbg;   // diagnostics: CTA 0 writes %globaltimer stamps of its phases here (null in normal use)
  Prob prob[MAX_PROB];
};
static_assert(sizeof(Group) <= 32000, "kernel parameter space (32,764 bytes since CUDA 12.1)");

// kind::tf32 instruction descriptor: D = f32, A = B = TF32 (format 2), both K-major
__host__ __device__ constexpr uint32_t idesc_tf32(uint32_t M, uint32_t N) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((N >> 3) << 17) | ((M >> 4) << 24);
}
__device__ __forceinline__ long long gtimer() {
  long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
__device__ __forceinline__ void mma_tf32(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
      "}\n" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}

__device__ __forceinline__ int row_off(int r) { return (r & 7) * 16 + (r >> 3) * SBO; }

// Per-thread loader of one operand (R rows x KC k per chunk, Q k-quads per thread).  Four mappings, chosen per segment
// (CTA-uniform) from the operand's strides and alignment:
//   KVEC    k contiguous, rows 16-byte aligned: quad (row, kq) = one LDG.128; 8 lanes cover one row's 128 bytes
//   KSCALAR k contiguous, unaligned: the same quads, four LDG.32 each
//   RVEC    rows contiguous (k strided), aligned, full tile (A only): a 4 rows x 4 k micro-tile per thread -- four LDG.128, each
//           a fully coalesced 512-byte warp access -- transposed in registers into the quads of rows 4*lane .. 4*lane+3
//   RSCALAR rows contiguous otherwise: lanes along rows, four LDG.32 per quad
// Interior tiles (all rows valid, whole chunk inside the K range) take a path without any bounds arithmetic: the pointer of
// quad 0 and the constant stride to the thread's next quads are set up once per segment and advanced by one add per chunk.
// (The first version recomputed rows, bounds and 64-bit addresses per element and chunk: ~450 instructions per thread
// and chunk, which made the kernel issue-bound -- ncu: 48 % issue-active with the tensor pipe at 9 %.)
enum : int { KVEC = 0, KSCALAR = 1, RVEC = 2, RSCALAR = 3 };

// cp.async (LDGSTS): global -> shared without a register in between, completion counted per GROUP (wait_group N = all but the N
// youngest groups have landed) -- unlike register loads, whose scoreboards count every outstanding load of the thread, so that
// waiting for the older of two chunks in flight also waits for the younger.  src_bytes < size zero-fills the rest.
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, uint32_t src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async4(uint32_t dst, const void* src, uint32_t src_bytes) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

template <int R, int Q, int LBO>
struct OpLoader {
  const float* p0;   // quad 0 at the current chunk
  const float* src;  // segment base (generic path)
  const float* safe; // a valid, 16-byte aligned address of the operand: source of the zero-filling copies (no byte of it is read)
  long qstride;      // elements between this thread's consecutive quads
  long sk, srow;
  int off0, offstride;   // shared-memory byte offset of quad 0, stride to the next quads
  int mode, row0, row_lim, rows_full;

  __device__ __forceinline__ void setup(const float* base, long srow_, long sk_, int vec, int row0_, int row_lim_, int k0, int tid,
                                        bool allow_rvec) {
    src = base; safe = base; srow = srow_; sk = sk_; row0 = row0_; row_lim = row_lim_;
    rows_full = row0_ + R <= row_lim_;
    const int lane = tid & 31, warp = tid >> 5;
    if (sk_ == 1) {
      mode = vec ? KVEC : KSCALAR;
      const int r = tid >> 3, kq = tid & 7;                   // quad i: row r + 32 i
      p0 = base + (row0_ + r) * srow_ + (k0 + kq * 4);
      qstride = 32 * srow_;
      off0 = row_off(r) + kq * LBO;
      offstride = 4 * SBO;
    } else if (allow_rvec && Q == 4 && R == 128 && srow_ == 1 && (sk_ & 3) == 0 && rows_full &&
               ((reinterpret_cast<uintptr_t>(base + row0_) & 15) == 0)) {
      mode = RVEC;                                             // quad i: row 4 lane + i, kq = warp
      safe = base + (row0_ + 4 * lane);
      p0 = base + (row0_ + 4 * lane) + (k0 + warp * 4) * sk_;
      qstride = 0;
      off0 = row_off(4 * lane) + warp * LBO;
      offstride = 16;
    } else {
      mode = RSCALAR;
      const int r = tid & (R - 1), kq = tid / R;               // quad i: kq + (THREADS / R) i
      p0 = base + (row0_ + r) * srow_ + (k0 + kq * 4) * sk_;
      qstride = (THREADS / R) * 4 * sk_;
      off0 = row_off(r) + kq * LBO;
      offstride = (THREADS / R) * LBO;
    }
  }

  // chunk [k0, k0 + KC) of the segment, valid k < k_end: cp.async of this thread's quads into its staging slots (slot i at
  // stg + i * THREADS * 16; shared-memory byte address).  RVEC stores the micro-tile untransposed: slot kk = rows 4 lane .. + 3 at k = kk.
  __device__ __forceinline__ void issue(uint32_t stg, int k0, int k_end, int tid) {
    constexpr uint32_t SLOT = THREADS * 16;
    if (mode == RVEC) {
      const int kb = k0 + (tid >> 5) * 4;
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) cp_async16(stg + kk * SLOT, kb + kk < k_end ? p0 + kk * sk : safe, kb + kk < k_end ? 16u : 0u);
    } else if (rows_full && k0 + KC <= k_end) {
      if (mode == KVEC) {
#pragma unroll
        for (int i = 0; i < Q; ++i) cp_async16(stg + i * SLOT, p0 + i * qstride, 16u);
      } else if (mode == KSCALAR) {
#pragma unroll
        for (int i = 0; i < Q; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) cp_async4(stg + i * SLOT + j * 4, p0 + i * qstride + j, 4u);
      } else {
#pragma unroll
        for (int i = 0; i < Q; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) cp_async4(stg + i * SLOT + j * 4, p0 + i * qstride + j * sk, 4u);
      }
    } else {
      // edge tile / tail chunk: bounds-checked, same quad mapping
#pragma unroll
      for (int i = 0; i < Q; ++i) {
        const int qi = tid + i * THREADS;
        int r, kq;
        if (sk == 1) { r = qi >> 3; kq = qi & 7; } else { r = qi & (R - 1); kq = qi / R; }
        const int gr = row0 + r, gk = k0 + kq * 4;
        const float* s = src + gr * srow + gk * sk;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const bool ok = gr < row_lim && gk + j < k_end;
          cp_async4(stg + i * SLOT + j * 4, ok ? s + j * sk : src, ok ? 4u : 0u);
        }
      }
    }
    p0 += KC * sk;
  }
  // the chunk's quads from the staging slots (after cp.async.wait_group); `m` = the mapping the chunk was issued with
  __device__ __forceinline__ static void fetch(float (&v)[Q][4], const uint8_t* stg, int m) {
    constexpr int SLOT = THREADS * 16;
    float4 t[Q];
#pragma unroll
    for (int i = 0; i < Q; ++i) t[i] = *reinterpret_cast<const float4*>(stg + i * SLOT);
    if (m == RVEC) {
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) {
        v[0 % Q][kk] = t[kk % Q].x; v[1 % Q][kk] = t[kk % Q].y; v[2 % Q][kk] = t[kk % Q].z; v[3 % Q][kk] = t[kk % Q].w;
      }
    } else {
#pragma unroll
      for (int i = 0; i < Q; ++i) { v[i][0] = t[i].x; v[i][1] = t[i].y; v[i][2] = t[i].z; v[i][3] = t[i].w; }
    }
  }

  // k-quad index (within the chunk) of this thread's quad i, for the mapping `m` the chunk was loaded with
  __device__ __forceinline__ int kquad(int m, int i, int tid) const {
    return m == RVEC ? (tid >> 5) : ((m == KVEC || m == KSCALAR) ? (tid & 7) : (tid / R + (THREADS / R) * i));
  }

  // hi / lo split and 16-byte stores into the UMMA planes
  __device__ __forceinline__ void store(const float (&v)[Q][4], uint8_t* hi_plane, uint8_t* lo_plane) const {
#pragma unroll
    for (int i = 0; i < Q; ++i) {
      const int off = off0 + i * offstride;
      uint32_t h[4], l[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        h[j] = __float_as_uint(v[i][j]) & 0xffffe000u;
        l[j] = __float_as_uint(v[i][j] - __uint_as_float(h[j]));
      }
      *reinterpret_cast<uint4*>(hi_plane + off) = make_uint4(h[0], h[1], h[2], h[3]);
      *reinterpret_cast<uint4*>(lo_plane + off) = make_uint4(l[0], l[1], l[2], l[3]);
    }
  }
};

__global__ void __launch_bounds__(CTA_THREADS, 1) tgemm_kernel(const __grid_constant__ Group g) {
  extern __shared__ __align__(128) uint8_t tg_smem_[];
  __shared__ uint64_t mma_done[2];   // stage s: its MMAs have retired (tcgen05.commit)
  __shared__ uint64_t full[2];       // stage s: all eight loader warps have stored (and proxy-fenced) their quads
  __shared__ uint32_t tmem_slot;
  __shared__ float colsum_red[BM];
  uint8_t* smem = tg_smem_ + ((128u - (smem_u32(tg_smem_) & 127u)) & 127u);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  // Programmatic dependent launch: let the NEXT launch of the stream become resident and run its prologue (barrier init, TMEM
  // allocation, problem decode) under this one's main loop; it blocks at griddepcontrol.wait below until this grid has completed.
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  const bool stamp = g.dbg != nullptr && blockIdx.x == 0 && threadIdx.x == 0;
  if (stamp) g.dbg[0] = gtimer();
  int pi = 0;
  while (pi + 1 < g.nprob && static_cast<int>(blockIdx.x) >= g.prob[pi + 1].cta_begin) ++pi;
  const Prob& P = g.prob[pi];
  const int local = blockIdx.x - P.cta_begin;
  const int tx = local % P.tiles_x, ty = (local / P.tiles_x) % P.tiles_y, sz = local / (P.tiles_x * P.tiles_y);
  const int m0 = ty * BM, n0 = tx * BN;
  const bool split = P.splits > 1;

  // chunk walk: segment `cs`, k range [ck, ce) of it; a split-K CTA owns one slice of segment 0
  int cs = 0;
  int ck = split ? sz * P.k_per : 0;
  int ce = split ? min(P.seg[0].K, ck + P.k_per) : P.seg[0].K;
  int nchunks = 0;
  if (split) {
    nchunks = (ce - ck + KC - 1) / KC;
  } else {
    for (int s = 0; s < P.nseg; ++s) nchunks += (P.seg[s].K + KC - 1) / KC;
  }
  if (nchunks <= 0) return;   // uniform over the CTA (an empty split-K slice adds nothing)

  constexpr int MMA_WARP = THREADS / 32;
  if (tid == 0) {
    mbar_init(&mma_done[0], 1);
    mbar_init(&mma_done[1], 1);
    mbar_init(&full[0], THREADS / 32);
    mbar_init(&full[1], THREADS / 32);
    fence_mbar_init();
  }
  if (warp == MMA_WARP) {
    tmem_alloc(&tmem_slot, TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  // everything above touched only this CTA's own state; the operands (and C / dact) may be outputs of the previous launch
  asm volatile("griddepcontrol.wait;" ::: "memory");
  const uint32_t tmem_base = tmem_slot;
  constexpr uint32_t idesc = idesc_tf32(BM, BN);

  if (warp == MMA_WARP) {
    // ---- MMA issue: decoupled from the loaders, which never wait for the issue of the chunk they just stored
    if (lane == 0) {
      for (int c = 0; c < nchunks; ++c) {
        const int s = c & 1;
        const uint32_t a_hi = smem_u32(smem) + s * STAGE_BYTES, a_lo = a_hi + PLANE_A, b_hi = a_lo + PLANE_A, b_lo = b_hi + PLANE_B;
        mbar_wait(&full[s], static_cast<uint32_t>(c >> 1) & 1u);
        // The generic-proxy -> async-proxy fence of the loaders' st.shared sits HERE, on the consumer side of the release / acquire
        // pair (a proxy fence anywhere along the causality path orders the two proxies).  In the loader threads it compiles to
        // MEMBAR.ALL.CTA + FENCE.VIEW.ASYNC.S, and the MEMBAR drains the thread's in-flight global loads: one full load latency per
        // chunk (1.05 us per chunk; 0.41 us for the one chunk of a CTA that had no load in flight).
        fence_proxy_async_smem();
        tc_fence_after();
#pragma unroll
        for (int j = 0; j < KC / 8; ++j) {
          // one K8 step = two k-quads of each plane
          const uint64_t dah = umma_desc_kmajor(a_hi + j * 2 * LBO_A, LBO_A, SBO);
          const uint64_t dal = umma_desc_kmajor(a_lo + j * 2 * LBO_A, LBO_A, SBO);
          const uint64_t dbh = umma_desc_kmajor(b_hi + j * 2 * LBO_B, LBO_B, SBO);
          const uint64_t dbl = umma_desc_kmajor(b_lo + j * 2 * LBO_B, LBO_B, SBO);
          mma_tf32(tmem_base, dal, dbh, idesc, (c > 0 || j > 0) ? 1u : 0u);
          mma_tf32(tmem_base, dah, dbl, idesc, 1u);
          mma_tf32(tmem_base, dah, dbh, idesc, 1u);
        }
        tc_commit(&mma_done[s]);
      }
    }
    __syncwarp();
    tc_fence_before();
    __syncthreads();
    tmem_dealloc(tmem_base, TMEM_COLS);
    return;
  }

  // ---- loader warps
  if (stamp) g.dbg[1] = gtimer();
  // NSTG - 1 chunks in flight through the staging ring: the copies of chunk c + 2 are issued at the hand-off of chunk c.  Only the
  // mapping of a chunk in flight lives in registers (the segment may change under it).
  struct ChunkMeta {
    int a_mode, a_off0, a_offstride, b_off0, b_offstride, a_k0, a_kend;
  };
  ChunkMeta r0, r1;
  OpLoader<BM, QA, LBO_A> la;
  OpLoader<BN, QB, LBO_B> lb;
  auto open_segment = [&]() {
    const Seg& S = P.seg[cs];
    la.setup(S.A, S.sAm, S.sAk, S.vecA, m0, P.M, ck, tid, true);
    lb.setup(S.B, S.sBn, S.sBk, S.vecB, n0, P.N, ck, tid, false);
  };
  const float* const kscale = P.kscale;
  uint8_t* const stg_base = smem + 2 * STAGE_BYTES + tid * 16;
  const bool ks_vec = kscale != nullptr && (reinterpret_cast<uintptr_t>(kscale) & 15) == 0;
  int issued = 0;
  auto prefetch = [&](ChunkMeta& R) {   // issues the copies of chunk (cs, ck) into staging stage `issued % NSTG` and advances the walk
    const uint32_t stg = smem_u32(stg_base) + static_cast<uint32_t>(issued % NSTG) * STG_BYTES;
    ++issued;
    R.a_mode = la.mode;
    R.a_off0 = la.off0; R.a_offstride = la.offstride; R.b_off0 = lb.off0; R.b_offstride = lb.offstride;
    la.issue(stg, ck, ce, tid);
    lb.issue(stg + QA * THREADS * 16, ck, ce, tid);
    if (kscale) {
      R.a_k0 = ck; R.a_kend = ce;
      if (R.a_mode == RVEC) {   // the four factors of this warp's k-quad ride in slot QA + QB
        const int kb = ck + warp * 4;
        const uint32_t dst = stg + (QA + QB) * THREADS * 16;
        if (ks_vec && kb + 4 <= ce) {
          cp_async16(dst, kscale + kb, 16u);
        } else {
#pragma unroll
          for (int j = 0; j < 4; ++j) cp_async4(dst + j * 4, kb + j < ce ? kscale + kb + j : kscale, kb + j < ce ? 4u : 0u);
        }
      }
    }
    cp_async_commit();
    ck += KC;
    if (ck >= ce && !split && cs + 1 < P.nseg) {
      ++cs;
      ck = 0;
      ce = P.seg[cs].K;
      open_segment();
    }
  };
  if (P.colsum) {
    for (int i = tid; i < BM; i += THREADS) colsum_red[i] = 0.f;
    named_bar_sync(1, THREADS);
  }
  float csum[4] = {0.f, 0.f, 0.f, 0.f};
  int a_mode = KVEC;
  auto hand_off = [&](int c, ChunkMeta& R) {   // chunk c (mapping in R) -> operand stage c & 1; then R describes chunk c + 2
    const int s = c & 1;
    uint8_t* a_hi = smem + s * STAGE_BYTES;
    uint8_t* a_lo = a_hi + PLANE_A;
    uint8_t* b_hi = a_lo + PLANE_A;
    uint8_t* b_lo = b_hi + PLANE_B;
    const uint8_t* stg = stg_base + (c % NSTG) * STG_BYTES;
    if (c + 1 < nchunks) cp_async_wait<1>(); else cp_async_wait<0>();   // this thread's copies of chunk c have landed
    float va[QA][4], vb[QB][4];
    OpLoader<BM, QA, LBO_A>::fetch(va, stg, R.a_mode);
    OpLoader<BN, QB, LBO_B>::fetch(vb, stg + QA * THREADS * 16, KVEC);
    a_mode = R.a_mode;
    if (kscale) {   // before the split and the column sums: dW = sum_r (dz_r J[r, :])^T x[r, :], db = sum_r dz_r J[r, :]
      if (R.a_mode == RVEC) {
        const float4 k4 = *reinterpret_cast<const float4*>(stg + (QA + QB) * THREADS * 16);
        const float ks[4] = {k4.x, k4.y, k4.z, k4.w};
#pragma unroll
        for (int i = 0; i < QA; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) va[i][j] *= ks[j];
      } else {
#pragma unroll
        for (int i = 0; i < QA; ++i) {
          const int kb = R.a_k0 + la.kquad(R.a_mode, i, tid) * 4;
#pragma unroll
          for (int j = 0; j < 4; ++j) va[i][j] *= kb + j < R.a_kend ? __ldg(kscale + kb + j) : 0.f;
        }
      }
    }
    if (c >= 2) mbar_wait(&mma_done[s], static_cast<uint32_t>((c >> 1) - 1) & 1u);   // the MMAs that read this stage have retired
    {
      OpLoader<BM, QA, LBO_A> sa = la;
      sa.off0 = R.a_off0; sa.offstride = R.a_offstride;
      sa.store(va, a_hi, a_lo);
      OpLoader<BN, QB, LBO_B> sb = lb;
      sb.off0 = R.b_off0; sb.offstride = R.b_offstride;
      sb.store(vb, b_hi, b_lo);
    }
    if (P.colsum) {   // k-strided A: RVEC -> quad i is row 4 lane + i; RSCALAR -> every quad is row tid & 127
#pragma unroll
      for (int i = 0; i < QA; ++i) {
        const float t = (va[i][0] + va[i][1]) + (va[i][2] + va[i][3]);
        if (R.a_mode == RVEC) csum[i] += t; else csum[0] += t;
      }
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(&full[s]);   // release: the MMA warp's acquire + proxy fence make the stores visible to the tensor core
    if (stamp && c < 12) g.dbg[4 + c] = gtimer();
    if (c + 2 < nchunks) prefetch(R);   // into the staging stage chunk c - 1 has left
  };
  open_segment();
  prefetch(r0);
  if (nchunks > 1) prefetch(r1);
  for (int c = 0; c < nchunks; c += 2) {
    hand_off(c, r0);
    if (c + 1 < nchunks) hand_off(c + 1, r1);
  }

  if (P.colsum) {
    if (a_mode == RVEC) {
#pragma unroll
      for (int i = 0; i < 4; ++i) atomicAdd(&colsum_red[4 * lane + i], csum[i]);
    } else {
      atomicAdd(&colsum_red[tid & (BM - 1)], csum[0]);
    }
    named_bar_sync(1, THREADS);
    if (tid < BM && tx == 0 && m0 + tid < P.M) atomicAdd(P.colsum + m0 + tid, colsum_red[tid]);
  }

  const int last = nchunks - 1;
  mbar_wait(&mma_done[last & 1], static_cast<uint32_t>(last >> 1) & 1u);
  tc_fence_after();
  if (stamp) g.dbg[2] = gtimer();

  // epilogue: warp w reads TMEM lanes 32*(w&3).. (its sub-partition) and columns 32*(w>>2).. and parks them in shared memory
  // (both operand stages are free: every MMA has retired) as a [128][64 + 4] fp32 tile, so that the global pass below
  // runs along rows: 16 lanes x float4 = one 256-byte row segment.  Thread-per-row stores straight from TMEM cost 6.8 us of a
  // 15 us CTA (4-byte writes to 32 different rows per instruction: 8x the L2 sector operations).
  {
    constexpr int TLD = BN + 4;   // floats per tile row: 272 B keeps the float4 stores of 32 rows conflict-free
    float* tile = reinterpret_cast<float*>(smem);
    {
      const int lq = warp & 3, ch = warp >> 2;
      uint32_t v[32];
      tmem_ld_32x32b_x32(tmem_base + (static_cast<uint32_t>(lq * 32) << 16) + static_cast<uint32_t>(ch * 32), v);
      tmem_ld_wait();
      float* trow = tile + (lq * 32 + lane) * TLD + ch * 32;
#pragma unroll
      for (int j = 0; j < 32; j += 4)
        *reinterpret_cast<float4*>(trow + j) = make_float4(__uint_as_float(v[j]), __uint_as_float(v[j + 1]), __uint_as_float(v[j + 2]), __uint_as_float(v[j + 3]));
    }
    named_bar_sync(1, THREADS);
    const bool vec_c = (P.ldc & 3) == 0 && (reinterpret_cast<uintptr_t>(P.C) & 15) == 0 && !split;
    const bool vec_d = P.dact && (P.ld_dact & 3) == 0 && (reinterpret_cast<uintptr_t>(P.dact) & 15) == 0;
#pragma unroll 2
    for (int it = 0; it < BM * (BN / 4) / THREADS; ++it) {
      const int idx = tid + it * THREADS;
      const int row = idx >> 4, c4 = (idx & 15) * 4;
      const int gm = m0 + row, gn = n0 + c4;
      if (gm >= P.M || gn >= P.N) continue;
      const float4 t4 = *reinterpret_cast<const float4*>(tile + row * TLD + c4);
      float x[4] = {t4.x, t4.y, t4.z, t4.w};
      float* crow = P.C + static_cast<size_t>(gm) * P.ldc + gn;
      const bool full4 = gn + 3 < P.N;
      if (split) {
#pragma unroll
        for (int j = 0; j < 4; ++j)
          if (gn + j < P.N) atomicAdd(crow + j, (P.bias && sz == 0) ? x[j] + __ldg(P.bias + gn + j) : x[j]);
        continue;
      }
      if (P.beta) {
        if (full4 && vec_c) {
          const float4 c4v = *reinterpret_cast<const float4*>(crow);
          x[0] += c4v.x; x[1] += c4v.y; x[2] += c4v.z; x[3] += c4v.w;
        } else {
#pragma unroll
          for (int j = 0; j < 4; ++j)
            if (gn + j < P.N) x[j] += crow[j];
        }
      }
      if (P.bias) {
#pragma unroll
        for (int j = 0; j < 4; ++j)
          if (gn + j < P.N) x[j] += __ldg(P.bias + gn + j);
      }
      if (P.act) {
#pragma unroll
        for (int j = 0; j < 4; ++j) x[j] = x[j] > 0.f ? x[j] : x[j] * P.slope;
      }
      if (P.dact) {
        const float* drow = P.dact + static_cast<size_t>(gm) * P.ld_dact + gn;
        float d[4] = {1.f, 1.f, 1.f, 1.f};
        if (full4 && vec_d) {
          const float4 d4 = __ldg(reinterpret_cast<const float4*>(drow));
          d[0] = d4.x; d[1] = d4.y; d[2] = d4.z; d[3] = d4.w;
        } else {
#pragma unroll
          for (int j = 0; j < 4; ++j)
            if (gn + j < P.N) d[j] = __ldg(drow + j);
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) x[j] = d[j] > 0.f ? x[j] : x[j] * P.slope;
      }
      if (full4 && vec_c) {
        *reinterpret_cast<float4*>(crow) = make_float4(x[0], x[1], x[2], x[3]);
      } else {
#pragma unroll
        for (int j = 0; j < 4; ++j)
          if (gn + j < P.N) crow[j] = x[j];
      }
    }
  }
  if (stamp) g.dbg[3] = gtimer();
  tc_fence_before();
  __syncthreads();   // pairs with the MMA warp's: it frees the accumulator after every epilogue warp has read it
}

}  // namespace tg
}  // namespace b200
