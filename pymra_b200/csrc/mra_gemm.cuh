// FP64 tile GEMM core of the MRA kernels (sm_100a).
//
// One primitive: a 64x64 accumulator tile per 128-thread CTA (4 warps, each owning 16 full rows),
//     acc += A(64 x K) * B(64 x K)^T,   both operands K-contiguous ("NT"),
// computed with FP64 tensor-core MMA (mma.sync m8n8k4 -> SASS DMMA.8x8x4, the only FP64 MMA
// shape sm_100a has).  Operands living in global memory are streamed through a 3-stage
// cp.async (LDGSTS) pipeline of 64 x 16 chunks; the chunks are stored with an XOR swizzle of
// their 16-byte columns so the 8x4 fragment loads are bank-conflict free without padding.
// Operands that already live in shared memory (the D / P tiles of the callers) are read in
// place through an element functor.  Because a warp owns complete rows, an accumulator tile can be fed
// straight back as the A operand of the next product (tile_gemm_regA: quad shuffles turn the C fragment
// layout into the A fragment layout), so GEMM -> transform -> GEMM chains never touch shared memory and
// row-wise reductions stay inside a quad.
//
// Row sources are described by a functor rr -> const double* (start of the K range of tile row
// rr, or nullptr for a zero row).  VEC = 2 uses 16-byte copies and needs every row start 16-byte
// aligned (true whenever r is even: all leading dimensions are even by construction); VEC = 1
// uses 8-byte copies and has no alignment requirement (odd r).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace mra {

constexpr int TB = 64;        // tile rows / cols
constexpr int KC = 16;        // k chunk per pipeline stage
constexpr int NT = 128;       // threads per CTA
constexpr int NSTAGE = 3;     // cp.async pipeline depth (4 was measured slower: occupancy)
constexpr int LDB = TB + 4;   // smem row stride of a resident 64x64 block (== 4 mod 16 doubles)

constexpr int MAXSEG = 8;     // K segments per tile_gemm_seg call

// Shared-memory working set of the tile primitive.  NS = number of K segments the row tables hold: kernels that
// only use single-segment products take GemmSmemT<1> (49 KB -> 4 CTAs/SM), the segmented ones GemmSmemT<MAXSEG>.
template <int NS>
struct alignas(16) GemmSmemT {
  static constexpr int NSEG = NS;
  double a[NSTAGE][TB * KC];
  double b[NSTAGE][TB * KC];
  const double* row_a[NS][TB];
  const double* row_b[NS][TB];
  int seg_k[NS];
  int pad_[3];
};
using GemmSmem = GemmSmemT<MAXSEG>;
using GemmSmem1 = GemmSmemT<1>;

__device__ __forceinline__ void dmma884(double (&c)[2], double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(c[0]), "+d"(c[1])
               : "d"(a), "d"(b));
}

// Element (i, j, e) of a thread: tile row 16*warp + 8*i + (lane >> 2), column 8*j + 2*(lane & 3) + e.
// NJ = number of 8-column groups of the tile: 8 for the square 64 x 64 tile; kernels whose output is only r <= 48
// columns wide (r0 = 16 / 32: BASELINE cfg4 / cfg3) use 64 x 16 / 64 x 32 / 64 x 48 tiles (NJ = 2 / 4 / 6) so that no
// DMMA, fragment load or B-operand copy is spent on zero padding.
template <int NJ>
struct AccT {
  static constexpr int kNJ = NJ;
  double v[2][NJ][2];
  __device__ __forceinline__ void zero() {
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
      for (int j = 0; j < NJ; ++j) v[i][j][0] = v[i][j][1] = 0.0;
  }
  __device__ __forceinline__ void negate() {
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
      for (int j = 0; j < NJ; ++j) {
        v[i][j][0] = -v[i][j][0];
        v[i][j][1] = -v[i][j][1];
      }
  }
};
using Acc = AccT<8>;
constexpr int nj_for(int r) { return r <= 16 ? 2 : r <= 32 ? 4 : r <= 48 ? 6 : 8; }

__device__ __forceinline__ void cp_async_16(double* smem, const double* gmem, int src_bytes) {
  unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(s), "l"(gmem), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_8(double* smem, const double* gmem, int src_bytes) {
  unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;\n" ::"r"(s), "l"(gmem), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory");
}

// Swizzled position of element (row, k) of a staged 64 x 16 chunk.
__device__ __forceinline__ int stage_pos(int row, int k) {
  return row * KC + ((((k >> 1) ^ ((row & 3) << 1)) << 1) | (k & 1));
}

// Issue the copies of one chunk (k0 .. k0+KC) of one operand.  `dummy` is any valid global address
// (used with src-size 0, which reads nothing and zero-fills).
template <int VEC, int NROWS = TB>
__device__ __forceinline__ void stage_load(double* st, const double* const* rows, int k0, int K,
                                           const double* dummy) {
  if (VEC == 2) {
    const int kc = (threadIdx.x & 7) * 2, rb = threadIdx.x >> 3;
    const int k = k0 + kc;
    const int nv = min(max(K - k, 0), 2) * 8;
#pragma unroll
    for (int i = 0; i < (NROWS + 15) / 16; ++i) {
      const int row = rb + 16 * i;
      const double* p = rows[row];
      cp_async_16(st + stage_pos(row, kc), p ? p + k : dummy, p ? nv : 0);
    }
  } else {
    const int kc = threadIdx.x & 15, rb = threadIdx.x >> 4;
    const int k = k0 + kc;
    const int nv = (k < K) ? 8 : 0;
#pragma unroll
    for (int i = 0; i < (NROWS + 7) / 8; ++i) {
      const int row = rb + 8 * i;
      const double* p = rows[row];
      cp_async_8(st + stage_pos(row, kc), p ? p + k : dummy, p ? nv : 0);
    }
  }
}

// 64 x 64 x 16 of DMMA on one chunk.  ga(row, kk) / gb(row, kk): operand element at tile row `row`,
// k index kk (0..15) inside the chunk.
// mrows: valid rows of the output tile; warps whose 16 rows lie outside are skipped (their accumulators
// keep their value, zero if the tile was zeroed).  Column-group predication was measured slower (r01g).
// (Skipping the all-zero column groups of a triangular B operand with a per-group predicate was measured SLOWER, twice:
// r01g for ragged column counts, r04e for the Linv / Lp^-1 operands -- predict_fused 43.7 -> 50.3 ms.  The unrolled,
// unpredicated DMMA stream is worth more than the 37 % of the flops such a segment could save.)
// J0: first column group that is computed.  A chunk of a LOWER-TRIANGULAR B operand (B[j][k] = 0 for k > j: Linv,
// Lp^-1) whose k range starts at 16 t has only zeros in the column groups below 2 t; chunk_mma_tri picks the unrolled
// variant for that chunk with one warp-uniform switch, so the DMMA stream itself stays free of predicates.
// KS: k-steps (of 4) that are run; the last chunk of a K segment whose length is not a multiple of 16 holds zero-filled
// columns beyond its end, and chunk_mma_tail picks the variant that stops after the steps that hold data.
template <int NJ, int J0 = 0, int J1 = NJ, int KS = KC / 4, class GA, class GB>
__device__ __forceinline__ void chunk_mma(AccT<NJ>& acc, GA ga, GB gb, int mrows = TB, int ncols = TB, int wg = -1) {
  // wg: the 16-row group this warp owns (default: its index).  Kernels whose warps do unequal work (ragged or
  // triangular tiles) rotate it with the CTA index so that the idle tensor pipe differs between co-resident CTAs.
  const int lane = threadIdx.x & 31, warp = wg >= 0 ? wg : (int)(threadIdx.x >> 5);
  const int wm = warp * 16;
  if (wm >= mrows) return;
  (void)ncols;
  const int g = lane >> 2, q = lane & 3;
#pragma unroll
  for (int ks = 0; ks < 4 * KS; ks += 4) {
    double a[2], b[NJ];
#pragma unroll
    for (int i = 0; i < 2; ++i) a[i] = ga(wm + i * 8 + g, ks + q);
#pragma unroll
    for (int j = J0; j < J1; ++j) b[j] = gb(j * 8 + g, ks + q);
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
      for (int j = J0; j < J1; ++j) dmma884(acc.v[i][j], a[i], b[j]);
  }
}

// kvalid: columns of the chunk that hold data (1 .. 16)
template <int NJ, class GA, class GB>
__device__ __forceinline__ void chunk_mma_tail(AccT<NJ>& acc, GA ga, GB gb, int kvalid, int mrows = TB, int wg = -1) {
  if (kvalid > 12) return chunk_mma<NJ, 0, NJ, 4>(acc, ga, gb, mrows, TB, wg);
  if (kvalid > 8) return chunk_mma<NJ, 0, NJ, 3>(acc, ga, gb, mrows, TB, wg);
  if (kvalid > 4) return chunk_mma<NJ, 0, NJ, 2>(acc, ga, gb, mrows, TB, wg);
  chunk_mma<NJ, 0, NJ, 1>(acc, ga, gb, mrows, TB, wg);
}

// Diagonal tile of a symmetric product (64 x 64 tile, NJ = 8): the warp owning row group wg only needs the column
// groups up to its own rows, 0 .. 2 wg + 1; the caller mirrors the lower part.
template <class GA, class GB>
__device__ __forceinline__ void chunk_mma_lower(AccT<8>& acc, GA ga, GB gb, int mrows, int wg) {
  if (wg == 0) return chunk_mma<8, 0, 2>(acc, ga, gb, mrows, TB, wg);
  if (wg == 1) return chunk_mma<8, 0, 4>(acc, ga, gb, mrows, TB, wg);
  if (wg == 2) return chunk_mma<8, 0, 6>(acc, ga, gb, mrows, TB, wg);
  chunk_mma<8, 0, 8>(acc, ga, gb, mrows, TB, wg);
}

// tl: index of the chunk inside the triangular segment (k range [16 tl, 16 tl + 16)); the segment is at most 8 NJ long.
template <int NJ, class GA, class GB>
__device__ __forceinline__ void chunk_mma_tri(AccT<NJ>& acc, GA ga, GB gb, int tl, int mrows = TB) {
  if constexpr (NJ > 6) {
    if (tl == 3) return chunk_mma<NJ, 6>(acc, ga, gb, mrows);
  }
  if constexpr (NJ > 4) {
    if (tl == 2) return chunk_mma<NJ, 4>(acc, ga, gb, mrows);
  }
  if constexpr (NJ > 2) {
    if (tl == 1) return chunk_mma<NJ, 2>(acc, ga, gb, mrows);
  }
  if (tl == 0) chunk_mma<NJ, 0>(acc, ga, gb, mrows);
}

// acc += A B^T.
//   A_GLOBAL: fa(rr) -> const double* row pointer (nullptr = zero row); else fa(rr, k) -> element
//   (shared-memory resident operand; must return 0 for k >= K).  Same for B.
// Must be called by all 128 threads; safe to call back to back (leading barrier).
template <int VEC, bool A_GLOBAL, bool B_GLOBAL, bool LOWER = false, int NJ, class FA, class FB, class SM>
__device__ __forceinline__ void tile_gemm(AccT<NJ>& acc, int K, FA fa, FB fb, SM& sm, const double* dummy,
                                          int mrows = TB, int ncols = TB, int wg = -1, bool lower = false) {
  constexpr int BR = 8 * NJ;     // rows of the B operand (= columns of the tile) that exist
  __syncthreads();   // previous users of the stages / row tables (and of resident operands) are done
  if (A_GLOBAL) {
    if (threadIdx.x < TB) {
      if constexpr (A_GLOBAL) sm.row_a[0][threadIdx.x] = fa((int)threadIdx.x);
    }
  }
  if (B_GLOBAL) {
    if (threadIdx.x >= NT - TB) {
      if constexpr (B_GLOBAL) {
        const int rr = (int)threadIdx.x - (NT - TB);
        sm.row_b[0][rr] = rr < BR ? fb(rr) : nullptr;
      }
    }
  }
  __syncthreads();
  const int nk = (K + KC - 1) / KC;
#pragma unroll
  for (int s = 0; s < NSTAGE - 1; ++s) {
    if (s < nk) {
      if (A_GLOBAL) stage_load<VEC>(sm.a[s], sm.row_a[0], s * KC, K, dummy);
      if (B_GLOBAL) stage_load<VEC, BR>(sm.b[s], sm.row_b[0], s * KC, K, dummy);
    }
    cp_async_commit();
  }
  int buf = 0;
  for (int kt = 0; kt < nk; ++kt) {
    cp_async_wait<NSTAGE - 2>();
    __syncthreads();
    {
      const int kn = kt + NSTAGE - 1;
      int nb = buf + NSTAGE - 1;
      if (nb >= NSTAGE) nb -= NSTAGE;
      if (kn < nk) {
        if (A_GLOBAL) stage_load<VEC>(sm.a[nb], sm.row_a[0], kn * KC, K, dummy);
        if (B_GLOBAL) stage_load<VEC, BR>(sm.b[nb], sm.row_b[0], kn * KC, K, dummy);
      }
      cp_async_commit();
    }
    const double* sa = sm.a[buf];
    const double* sb = sm.b[buf];
    const int k0 = kt * KC;
    auto ga = [&](int row, int kk) -> double {
      if constexpr (A_GLOBAL) return sa[stage_pos(row, kk)];
      else return fa(row, k0 + kk);
    };
    auto gb = [&](int row, int kk) -> double {
      if constexpr (B_GLOBAL) return sb[stage_pos(row, kk)];
      else return fb(row, k0 + kk);
    };
    if constexpr (LOWER && NJ == 8) {
      if (lower) chunk_mma_lower(acc, ga, gb, mrows, wg);      // diagonal tile of a symmetric product
      else chunk_mma(acc, ga, gb, mrows, ncols, wg);
    } else chunk_mma(acc, ga, gb, mrows, ncols, wg);
    if (++buf == NSTAGE) buf = 0;
  }
  cp_async_wait<0>();
}

// acc += sum_s A_s B_s^T over up to MAXSEG K-segments, all operands in global memory, as ONE pipelined
// stream of chunks (no pipeline drain between segments).  fa(s, rr) / fb(s, rr) -> row pointer of tile row
// rr in segment s (nullptr = zero row), fk(s) -> K of segment s (may be 0).
// GEN: the A operand of the LAST segment is not loaded but computed, fg(row, k) -> element (e.g. a covariance
// tile evaluated on the fly); its chunks are written straight into the pipeline stages.
struct NoGen {
  __device__ double operator()(int, int) const { return 0.0; }
};

// TRI: the chunks [tri_kt0, tri_kt0 + tri_nk) of the stream belong to a segment whose B operand is lower triangular
// (B row j zero beyond column j, j = tile column): their all-zero column groups are skipped (chunk_mma_tri).
// wg >= 0: row group of this warp (see chunk_mma); lower: only the column groups up to the warp's own rows are computed
// (diagonal tile of a symmetric product, NJ = 8).
template <int VEC, bool GEN = false, bool CACHE_PTRS = true, bool TRI = false, int NJ, class FA, class FB, class FK, class SM,
          class FG = NoGen>
__device__ __forceinline__ void tile_gemm_seg(AccT<NJ>& acc, int nseg, FA fa, FB fb, FK fk, SM& sm,
                                              const double* dummy, int mrows = TB, int ncols = TB, FG fg = FG(),
                                              int tri_kt0 = 0, int tri_nk = 0, int wg = -1, bool lower = false) {
  constexpr int BR = 8 * NJ;            // rows of the B operand that exist
  constexpr int BI = (BR + 15) / 16;    // 16-row groups of B a thread copies (16-byte path)
  __syncthreads();
  for (int s = 0; s < nseg; ++s) {
    if (threadIdx.x < TB) {
      if (!(GEN && s == nseg - 1)) sm.row_a[s][threadIdx.x] = fa(s, (int)threadIdx.x);
    } else {
      const int rr = (int)threadIdx.x - TB;
      sm.row_b[s][rr] = rr < BR ? fb(s, rr) : nullptr;
    }
    if (threadIdx.x == 0) sm.seg_k[s] = fk(s);
  }
  __syncthreads();
  int nk = 0;
  for (int s = 0; s < nseg; ++s) nk += (sm.seg_k[s] + KC - 1) / KC;
  // loader cursor; with 16-byte copies the four row pointers per operand of the current segment are kept in
  // registers (re-read from the tables only when the segment changes)
  int lseg = 0, lk0 = 0, cached = -1;
  const double* pa[4];
  const double* pb[4];
  auto load_next = [&](int buf) {
    while (lseg < nseg && lk0 >= sm.seg_k[lseg]) {
      ++lseg;
      lk0 = 0;
    }
    if (lseg < nseg) {
      const int K = sm.seg_k[lseg];
      const bool gen = GEN && lseg == nseg - 1;
      if (VEC == 2 && CACHE_PTRS) {
        const int kc = (threadIdx.x & 7) * 2, rb = threadIdx.x >> 3;
        if (cached != lseg) {
          cached = lseg;
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            if (!gen) pa[i] = sm.row_a[lseg][rb + 16 * i];
            if (i < BI) pb[i] = sm.row_b[lseg][rb + 16 * i];
          }
        }
        const int k = lk0 + kc;
        const int nv = min(max(K - k, 0), 2) * 8;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int row = rb + 16 * i;
          const int pos = stage_pos(row, kc);
          if (gen) {
            double2 v;
            v.x = k < K ? fg(row, k) : 0.0;
            v.y = k + 1 < K ? fg(row, k + 1) : 0.0;
            *reinterpret_cast<double2*>(sm.a[buf] + pos) = v;
          } else {
            cp_async_16(sm.a[buf] + pos, pa[i] ? pa[i] + k : dummy, pa[i] ? nv : 0);
          }
          if (i < BI) cp_async_16(sm.b[buf] + pos, pb[i] ? pb[i] + k : dummy, pb[i] ? nv : 0);
        }
      } else {
        if (gen) {
          const int kc = threadIdx.x & 15, rb = threadIdx.x >> 4;
          const int k = lk0 + kc;
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int row = rb + 8 * i;
            sm.a[buf][stage_pos(row, kc)] = k < K ? fg(row, k) : 0.0;
          }
        } else {
          stage_load<VEC>(sm.a[buf], sm.row_a[lseg], lk0, K, dummy);
        }
        stage_load<VEC, BR>(sm.b[buf], sm.row_b[lseg], lk0, K, dummy);
      }
      lk0 += KC;
    }
  };
#pragma unroll
  for (int s = 0; s < NSTAGE - 1; ++s) {
    if (s < nk) load_next(s);
    cp_async_commit();
  }
  int buf = 0;
  for (int kt = 0; kt < nk; ++kt) {
    cp_async_wait<NSTAGE - 2>();
    __syncthreads();
    {
      int nb = buf + NSTAGE - 1;
      if (nb >= NSTAGE) nb -= NSTAGE;
      if (kt + NSTAGE - 1 < nk) load_next(nb);
      cp_async_commit();
    }
    const double* sa = sm.a[buf];
    const double* sb = sm.b[buf];
    auto ga = [&](int row, int kk) -> double { return sa[stage_pos(row, kk)]; };
    auto gb = [&](int row, int kk) -> double { return sb[stage_pos(row, kk)]; };
    if (TRI && (unsigned)(kt - tri_kt0) < (unsigned)tri_nk) chunk_mma_tri(acc, ga, gb, kt - tri_kt0, mrows);
    else if constexpr (NJ == 8 && !TRI && !GEN) {
      // (stopping the last chunk of a ragged segment after the k-steps that hold data -- chunk_mma_tail, as the fused
      // predict and leaf_ut do -- made assemble_A slower here: 17.8 vs 17.7 ms)
      if (lower) chunk_mma_lower(acc, ga, gb, mrows, wg);
      else chunk_mma(acc, ga, gb, mrows, ncols, wg);
    } else chunk_mma(acc, ga, gb, mrows, ncols, wg);
    if (++buf == NSTAGE) buf = 0;
  }
  cp_async_wait<0>();
}

// Position of element (k-row kr, column n) of a staged K-major 16 x 64 chunk (B given as B[k][n]).
__device__ __forceinline__ int kstage_pos(int kr, int n) { return kr * TB + (n ^ ((kr & 3) << 2)); }

// acc += A * B with A K-contiguous rows in global memory (fa(rr) -> row pointer) and B K-major in global
// memory: B[k][n] = bbase[k * ldb + n], k < K, n < ncols (zero outside).  Used where the contraction index
// is the row index of a stored block (left-multiplication of a block by a small matrix).
// wg: row group of this warp (see chunk_mma).  tri_row0: the A operand is lower triangular, A[tile row i][k] = 0 for
// k > tri_row0 + i; a warp skips the chunks that lie entirely beyond its 16 rows (default: never).
template <int VEC, int NJ, class FA, class SM>
__device__ __forceinline__ void tile_gemm_kmajorB(AccT<NJ>& acc, int K, FA fa, const double* bbase, long long ldb,
                                                  int ncols, SM& sm, const double* dummy, int wg = -1,
                                                  int tri_row0 = 1 << 30) {
  const int klast = tri_row0 + (wg >= 0 ? wg : (int)(threadIdx.x >> 5)) * 16 + 15;     // last k with a nonzero in my rows
  __syncthreads();
  if (threadIdx.x < TB) sm.row_a[0][threadIdx.x] = fa((int)threadIdx.x);
  __syncthreads();
  auto load_b = [&](double* st, int k0) {
    if (VEC == 2) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int cch = threadIdx.x + NT * i;
        const int kr = cch >> 5, n0 = (cch & 31) * 2, k = k0 + kr;
        const int nv = k < K ? min(max(ncols - n0, 0), 2) * 8 : 0;
        cp_async_16(st + kstage_pos(kr, n0), nv ? bbase + (size_t)k * ldb + n0 : dummy, nv);
      }
    } else {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int cch = threadIdx.x + NT * i;
        const int kr = cch >> 6, n = cch & 63, k = k0 + kr;
        const int nv = (k < K && n < ncols) ? 8 : 0;
        cp_async_8(st + kstage_pos(kr, n), nv ? bbase + (size_t)k * ldb + n : dummy, nv);
      }
    }
  };
  const int nk = (K + KC - 1) / KC;
#pragma unroll
  for (int s = 0; s < NSTAGE - 1; ++s) {
    if (s < nk) {
      stage_load<VEC>(sm.a[s], sm.row_a[0], s * KC, K, dummy);
      load_b(sm.b[s], s * KC);
    }
    cp_async_commit();
  }
  int buf = 0;
  for (int kt = 0; kt < nk; ++kt) {
    cp_async_wait<NSTAGE - 2>();
    __syncthreads();
    {
      const int kn = kt + NSTAGE - 1;
      int nb = buf + NSTAGE - 1;
      if (nb >= NSTAGE) nb -= NSTAGE;
      if (kn < nk) {
        stage_load<VEC>(sm.a[nb], sm.row_a[0], kn * KC, K, dummy);
        load_b(sm.b[nb], kn * KC);
      }
      cp_async_commit();
    }
    const double* sa = sm.a[buf];
    const double* sb = sm.b[buf];
    if (kt * KC <= klast)
      chunk_mma(acc, [&](int row, int kk) -> double { return sa[stage_pos(row, kk)]; },
                [&](int col, int kk) -> double { return sb[kstage_pos(kk, col)]; }, TB, TB, wg);
    if (++buf == NSTAGE) buf = 0;
  }
  cp_async_wait<0>();
}

// As tile_gemm_kmajorB, but B's K rows are gathered: rowk[k] -> start of B's row k (column n of the tile is
// rowk[k][col0 + n]), nullptr = zero row.  rowk lives in shared memory and holds at least K entries.
template <int VEC, int NJ, class FA, class SM>
__device__ __forceinline__ void tile_gemm_kmajorB_rows(AccT<NJ>& acc, int K, FA fa, const double* const* rowk, int ncols,
                                                       SM& sm, const double* dummy, int col0 = 0, int wg = -1,
                                                       int tri_row0 = 1 << 30, int mrows = TB) {
  const int klast = tri_row0 + (wg >= 0 ? wg : (int)(threadIdx.x >> 5)) * 16 + 15;
  __syncthreads();
  if (threadIdx.x < TB) sm.row_a[0][threadIdx.x] = fa((int)threadIdx.x);
  __syncthreads();
  auto load_b = [&](double* st, int k0) {
    if (VEC == 2) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int cch = threadIdx.x + NT * i;
        const int kr = cch >> 5, n0 = (cch & 31) * 2, k = k0 + kr;
        const double* p = k < K ? rowk[k] : nullptr;
        const int nv = p ? min(max(ncols - n0, 0), 2) * 8 : 0;
        cp_async_16(st + kstage_pos(kr, n0), nv ? p + col0 + n0 : dummy, nv);
      }
    } else {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int cch = threadIdx.x + NT * i;
        const int kr = cch >> 6, n = cch & 63, k = k0 + kr;
        const double* p = k < K ? rowk[k] : nullptr;
        const int nv = (p && n < ncols) ? 8 : 0;
        cp_async_8(st + kstage_pos(kr, n), nv ? p + col0 + n : dummy, nv);
      }
    }
  };
  const int nk = (K + KC - 1) / KC;
#pragma unroll
  for (int s = 0; s < NSTAGE - 1; ++s) {
    if (s < nk) {
      stage_load<VEC>(sm.a[s], sm.row_a[0], s * KC, K, dummy);
      load_b(sm.b[s], s * KC);
    }
    cp_async_commit();
  }
  int buf = 0;
  for (int kt = 0; kt < nk; ++kt) {
    cp_async_wait<NSTAGE - 2>();
    __syncthreads();
    {
      const int kn = kt + NSTAGE - 1;
      int nb = buf + NSTAGE - 1;
      if (nb >= NSTAGE) nb -= NSTAGE;
      if (kn < nk) {
        stage_load<VEC>(sm.a[nb], sm.row_a[0], kn * KC, K, dummy);
        load_b(sm.b[nb], kn * KC);
      }
      cp_async_commit();
    }
    const double* sa = sm.a[buf];
    const double* sb = sm.b[buf];
    if (kt * KC <= klast) {
      auto ga = [&](int row, int kk) -> double { return sa[stage_pos(row, kk)]; };
      auto gb = [&](int col, int kk) -> double { return sb[kstage_pos(kk, col)]; };
      if (K - kt * KC <= 12) chunk_mma_tail(acc, ga, gb, K - kt * KC, mrows, wg);      // last chunk: K mod 16 columns hold data
      else chunk_mma(acc, ga, gb, mrows, TB, wg);
    }
    if (++buf == NSTAGE) buf = 0;
  }
  cp_async_wait<0>();
}

// out += Areg * B^T where Areg is a 64 x 64 tile held in accumulator layout (columns >= K must be zero
// or K a multiple of 4 covering them) and K <= 64.  B_GLOBAL: fb(rr) -> row pointer, streamed through the
// cp.async stages; else fb(rr, k) -> element of a shared-memory resident operand.
template <int VEC, bool B_GLOBAL, int NJ, int NJA, class FB, class SM>
__device__ __forceinline__ void tile_gemm_regA(AccT<NJ>& out, const AccT<NJA>& A, int K, FB fb, SM& sm,
                                               const double* dummy, int mrows = TB, int ncols = TB) {
  constexpr int BR = 8 * NJ;
  const int lane = threadIdx.x & 31;
  const int g = lane >> 2, q = lane & 3;
  const bool active = (int)(threadIdx.x >> 5) * 16 < mrows;
  (void)ncols;
  __syncthreads();
  if (B_GLOBAL) {
    if (threadIdx.x < TB) {
      if constexpr (B_GLOBAL) sm.row_b[0][threadIdx.x] = (int)threadIdx.x < BR ? fb((int)threadIdx.x) : nullptr;
    }
    __syncthreads();
  }
  const int nk = (K + KC - 1) / KC;      // <= NJA / 2
  if (B_GLOBAL) {
#pragma unroll
    for (int s = 0; s < NSTAGE - 1; ++s) {
      if (s < nk) stage_load<VEC, BR>(sm.b[s], sm.row_b[0], s * KC, K, dummy);
      cp_async_commit();
    }
  }
#pragma unroll
  for (int kt = 0; kt < (8 * NJA + KC - 1) / KC; ++kt) {
    if (kt < nk) {
      const int buf = kt % NSTAGE;
      if (B_GLOBAL) {
        cp_async_wait<NSTAGE - 2>();
        __syncthreads();
        if (kt + NSTAGE - 1 < nk) stage_load<VEC, BR>(sm.b[(kt + NSTAGE - 1) % NSTAGE], sm.row_b[0], (kt + NSTAGE - 1) * KC, K, dummy);
        cp_async_commit();
      }
      const double* sb = sm.b[buf];
      if (active)
#pragma unroll
      for (int ks = 0; ks < KC; ks += 4) {
        const int st = (kt * KC + ks) >> 2;            // k-step index 0..15 -> source tile st>>1, half st&1
        const int src = (lane & ~3) | (((st & 1) << 1) | (q >> 1));
        if ((st >> 1) >= NJA) continue;                // beyond the columns A holds (NJA odd)
        double a[2], b[NJ];
#pragma unroll
        for (int i = 0; i < 2; ++i) {
          const double v0 = __shfl_sync(0xffffffffu, A.v[i][(st >> 1) < NJA ? (st >> 1) : 0][0], src);
          const double v1 = __shfl_sync(0xffffffffu, A.v[i][(st >> 1) < NJA ? (st >> 1) : 0][1], src);
          a[i] = (q & 1) ? v1 : v0;
        }
#pragma unroll
        for (int j = 0; j < NJ; ++j) {
          if constexpr (B_GLOBAL) b[j] = sb[stage_pos(j * 8 + g, ks + q)];
          else b[j] = fb(j * 8 + g, kt * KC + ks + q);
        }
#pragma unroll
        for (int i = 0; i < 2; ++i)
#pragma unroll
          for (int j = 0; j < NJ; ++j) dmma884(out.v[i][j], a[i], b[j]);
      }
    }
  }
  if (B_GLOBAL) cp_async_wait<0>();
}

// acc.v = f(row, col, acc.v) element-wise (transform in registers).
template <int NJ, class F>
__device__ __forceinline__ void tile_transform(AccT<NJ>& acc, F f) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int wm = warp * 16;
  const int g = lane >> 2, q = lane & 3;
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int j = 0; j < NJ; ++j)
#pragma unroll
      for (int e = 0; e < 2; ++e) acc.v[i][j][e] = f(wm + i * 8 + g, j * 8 + q * 2 + e, acc.v[i][j][e]);
}

// f(row, col, value) for every accumulator element owned by this thread.
template <int NJ, class F>
__device__ __forceinline__ void tile_epilogue(const AccT<NJ>& acc, F f, int wg = -1) {
  const int lane = threadIdx.x & 31, warp = wg >= 0 ? wg : (int)(threadIdx.x >> 5);
  const int wm = warp * 16;
  const int g = lane >> 2, q = lane & 3;
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int j = 0; j < NJ; ++j)
#pragma unroll
      for (int e = 0; e < 2; ++e) f(wm + i * 8 + g, j * 8 + q * 2 + e, acc.v[i][j][e]);
}

}  // namespace mra
