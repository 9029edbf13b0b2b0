// The grouped 3xTF32 GEMM of tgemm.cuh in its THROUGHPUT form: operands through registers (one chunk ahead), two CTAs per SM.
//
// tgemm.cuh's kernel stages its operands through a cp.async ring (three chunks deep, 86 KB), which leaves room for ONE CTA per SM:
// best for the launches of a training step that are a single wave of CTAs (a 4096 x 256 x 256 product = 128 CTAs: 13 -> 11 us).
// A launch of many waves -- the split backward's group of weight gradients, ~2,200 CTAs of 8 chunks -- is bound by chunks per SM
// and microsecond, and there two co-resident CTAs of this form hide each other's load latency better than one staged CTA
// (measured for that group at 4096 rays: 145 us here, 203 us staged).  train.cu picks per launch.  Same problem description
// (tg::Group), same operand layouts, same epilogue; the loaders execute the generic -> async proxy fence themselves.
#pragma once
#include "tgemm.cuh"

namespace b200 {
namespace tgr {
using namespace tg;
constexpr int SMEM_BYTES = 2 * STAGE_BYTES + 128;

template <int R, int Q, int LBO>
struct OpLoader {
  const float* p0;   // quad 0 at the current chunk
  const float* src;  // segment base (generic path)
  long qstride;      // elements between this thread's consecutive quads
  long sk, srow;
  int off0, offstride;   // shared-memory byte offset of quad 0, stride to the next quads
  int mode, row0, row_lim, rows_full;

  __device__ __forceinline__ void setup(const float* base, long srow_, long sk_, int vec, int row0_, int row_lim_, int k0, int tid,
                                        bool allow_rvec) {
    src = base; srow = srow_; sk = sk_; row0 = row0_; row_lim = row_lim_;
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

  // chunk [k0, k0 + KC) of the segment, valid k < k_end
  __device__ __forceinline__ void load(float (&v)[Q][4], int k0, int k_end, int tid) {
    if (mode == RVEC) {
      const int kb = k0 + (tid >> 5) * 4;
      float4 t[4];
#pragma unroll
      for (int kk = 0; kk < 4; ++kk)
        t[kk] = kb + kk < k_end ? __ldg(reinterpret_cast<const float4*>(p0 + kk * sk)) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) {
        v[0 % Q][kk] = t[kk].x; v[1 % Q][kk] = t[kk].y; v[2 % Q][kk] = t[kk].z; v[3 % Q][kk] = t[kk].w;
      }
    } else if (rows_full && k0 + KC <= k_end) {
      if (mode == KVEC) {
#pragma unroll
        for (int i = 0; i < Q; ++i) {
          const float4 q = __ldg(reinterpret_cast<const float4*>(p0 + i * qstride));
          v[i][0] = q.x; v[i][1] = q.y; v[i][2] = q.z; v[i][3] = q.w;
        }
      } else if (mode == KSCALAR) {
#pragma unroll
        for (int i = 0; i < Q; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) v[i][j] = __ldg(p0 + i * qstride + j);
      } else {
#pragma unroll
        for (int i = 0; i < Q; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) v[i][j] = __ldg(p0 + i * qstride + j * sk);
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
        for (int j = 0; j < 4; ++j) v[i][j] = (gr < row_lim && gk + j < k_end) ? __ldg(s + j * sk) : 0.f;
      }
    }
    p0 += KC * sk;
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

__global__ void __launch_bounds__(CTA_THREADS, 2) tgemm_reg_kernel(const __grid_constant__ Group g) {
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
  float va[QA][4], vb[QB][4];
  OpLoader<BM, QA, LBO_A> la;
  OpLoader<BN, QB, LBO_B> lb;
  auto open_segment = [&]() {
    const Seg& S = P.seg[cs];
    la.setup(S.A, S.sAm, S.sAk, S.vecA, m0, P.M, ck, tid, true);
    lb.setup(S.B, S.sBn, S.sBk, S.vecB, n0, P.N, ck, tid, false);
  };
  int a_mode;   // the mapping of the chunk held in va (the segment may change under it)
  int a_off0, a_offstride, b_off0, b_offstride;
  int a_k0 = 0, a_kend = 0;            // k range of the chunk held in va (kscale)
  float ks[4] = {1.f, 1.f, 1.f, 1.f};  // RVEC: the chunk's four factors of this warp's k-quad, fetched with the operands
  const float* const kscale = P.kscale;
  auto prefetch = [&]() {   // loads chunk (cs, ck) into registers and advances the walk
    a_mode = la.mode;
    a_off0 = la.off0; a_offstride = la.offstride; b_off0 = lb.off0; b_offstride = lb.offstride;
    la.load(va, ck, ce, tid);
    lb.load(vb, ck, ce, tid);
    if (kscale) {
      a_k0 = ck; a_kend = ce;
      if (a_mode == RVEC) {
        const int kb = ck + warp * 4;
#pragma unroll
        for (int j = 0; j < 4; ++j) ks[j] = kb + j < ce ? __ldg(kscale + kb + j) : 0.f;
      }
    }
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
  open_segment();
  prefetch();
  float csum[4] = {0.f, 0.f, 0.f, 0.f};

  for (int c = 0; c < nchunks; ++c) {
    const int s = c & 1;
    uint8_t* a_hi = smem + s * STAGE_BYTES;
    uint8_t* a_lo = a_hi + PLANE_A;
    uint8_t* b_hi = a_lo + PLANE_A;
    uint8_t* b_lo = b_hi + PLANE_B;
    if (c >= 2) mbar_wait(&mma_done[s], static_cast<uint32_t>((c >> 1) - 1) & 1u);   // the MMAs that read this stage have retired
    if (kscale) {   // before the split and the column sums: dW = sum_r (dz_r J[r, :])^T x[r, :], db = sum_r dz_r J[r, :]
      if (a_mode == RVEC) {
#pragma unroll
        for (int i = 0; i < QA; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) va[i][j] *= ks[j];
      } else {
#pragma unroll
        for (int i = 0; i < QA; ++i) {
          const int kb = a_k0 + la.kquad(a_mode, i, tid) * 4;
#pragma unroll
          for (int j = 0; j < 4; ++j) va[i][j] *= kb + j < a_kend ? __ldg(kscale + kb + j) : 0.f;
        }
      }
    }
    {
      OpLoader<BM, QA, LBO_A> sa = la;
      sa.off0 = a_off0; sa.offstride = a_offstride;
      sa.store(va, a_hi, a_lo);
      OpLoader<BN, QB, LBO_B> sb = lb;
      sb.off0 = b_off0; sb.offstride = b_offstride;
      sb.store(vb, b_hi, b_lo);
    }
    if (P.colsum) {   // k-strided A: RVEC -> quad i is row 4 lane + i; RSCALAR -> every quad is row tid & 127
#pragma unroll
      for (int i = 0; i < QA; ++i) {
        const float t = (va[i][0] + va[i][1]) + (va[i][2] + va[i][3]);
        if (a_mode == RVEC) csum[i] += t; else csum[0] += t;
      }
    }
    if (c + 1 < nchunks) prefetch();   // in flight while the MMA warp works and the next stage is waited for
    fence_proxy_async_smem();
    __syncwarp();
    if (lane == 0) mbar_arrive(&full[s]);
    if (stamp && c < 12) g.dbg[4 + c] = gtimer();
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

}  // namespace tgr
}  // namespace b200
