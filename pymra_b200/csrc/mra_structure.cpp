// Native host builder of the MRA tree for the large-node 2-D path (every node it meets has > 100
// rows and > 100 knot candidates): bit-exact clone of what pymra_b200/structure.py does with
// NumPy, i.e. of the reference's indexing:
//   * knots   np.random.choice(np.arange(n), size=r, replace=False)        pyMRA/MRANode.py:191-193
//             == legacy RandomState.permutation(n)[:r]: Fisher-Yates from the top with
//             random_interval() = masked rejection on 32-bit MT19937 outputs (NumPy mtrand,
//             not part of /root/reference; validated against NumPy in tests/test_native_builder.py)
//   * splits  four quadrants by <=/> the column means, means = sequential sum / n
//                                                                          pyMRA/MRANode.py:232-239
//   * order   DFS pre-order (the order the reference consumes the global RNG), fork semantics at
//             critDepth                                                    pyMRA/MRANode.py:23-98
// Anything else (1-D, KMeans nodes, empty quadrants) returns MRA_BUILD_UNSUPPORTED and the Python
// builder takes over after restoring the RNG state.
#include "../../include/pymra_b200.h"

#include <algorithm>
#include <atomic>
#include <functional>
#include <chrono>
#include <condition_variable>
#include <cstdio>
#include <memory>
#include <mutex>
#include <thread>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <vector>
#if defined(__x86_64__)
#include <immintrin.h>
#endif

namespace {

// MRA_BUILD_TRACE=1: phase timestamps of the threaded build on stderr (seconds since the build started)
struct Trace {
  bool on = std::getenv("MRA_BUILD_TRACE") != nullptr;
  std::chrono::steady_clock::time_point t0 = std::chrono::steady_clock::now();
  void mark(const char* what, int k = -1) const {
    if (!on) return;
    const double t = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    if (k >= 0) std::fprintf(stderr, "[mra build] %8.4f s  %s %d\n", t, what, k);
    else std::fprintf(stderr, "[mra build] %8.4f s  %s\n", t, what);
  }
};

#if defined(__x86_64__)
// lane permutation that packs the lanes selected by an 8-bit mask to the high end, first selected lane last
struct alignas(32) PackLut {
  uint32_t idx[256][8];
  PackLut() {
    for (int m = 0; m < 256; ++m) {
      int k = 0;
      for (int b = 0; b < 8; ++b) idx[m][b] = 0;
      for (int b = 0; b < 8; ++b)
        if ((m >> b) & 1) idx[m][7 - k++] = (uint32_t)b;
    }
  }
};
#endif

// Legacy NumPy MT19937 (mtrand's rk_state): key[624] + pos.  Outputs are produced 624 at a time,
// already tempered, so the hot loops below read a plain buffer.
struct MT {
  uint32_t key[624];
  int pos;
  uint32_t out[624];
  bool out_valid = false;
  __attribute__((target_clones("avx2", "default"))) void gen() {
    const uint32_t UP = 0x80000000u, LO = 0x7fffffffu, MA = 0x9908b0dfu;
    int kk;
    uint32_t y;
    for (kk = 0; kk < 624 - 397; ++kk) {
      y = (key[kk] & UP) | (key[kk + 1] & LO);
      key[kk] = key[kk + 397] ^ (y >> 1) ^ (-(int32_t)(y & 1) & MA);
    }
    for (; kk < 623; ++kk) {
      y = (key[kk] & UP) | (key[kk + 1] & LO);
      key[kk] = key[kk + (397 - 624)] ^ (y >> 1) ^ (-(int32_t)(y & 1) & MA);
    }
    y = (key[623] & UP) | (key[0] & LO);
    key[623] = key[396] ^ (y >> 1) ^ (-(int32_t)(y & 1) & MA);
    pos = 0;
    out_valid = false;
  }
  __attribute__((target_clones("avx2", "default"))) void temper_all() {
    for (int i = 0; i < 624; ++i) {
      uint32_t y = key[i];
      y ^= (y >> 11);
      y ^= (y << 7) & 0x9d2c5680u;
      y ^= (y << 15) & 0xefc60000u;
      y ^= (y >> 18);
      out[i] = y;
    }
    out_valid = true;
  }
  // makes out[pos .. 624) readable; returns how many outputs are available (>= 1)
  inline int avail() {
    if (pos >= 624) gen();
    if (!out_valid) temper_all();
    return 624 - pos;
  }
  inline uint32_t next32() {
    avail();
    return out[pos++];
  }
  static inline uint32_t smear(uint32_t max) {
    uint32_t mask = max;
    mask |= mask >> 1;
    mask |= mask >> 2;
    mask |= mask >> 4;
    mask |= mask >> 8;
    mask |= mask >> 16;
    return mask;
  }
  inline uint32_t interval(uint32_t max) {   // uniform on [0, max], max <= 0xffffffff (legacy rk_interval)
    if (max == 0) return 0;
    const uint32_t mask = smear(max);
    uint32_t v;
    while ((v = (next32() & mask)) > max) {
    }
    return v;
  }
  // jb[i] = interval(i) for i = hi, hi-1, ..., 1 -- the draw sequence of the legacy Fisher-Yates shuffle.
  // Same stream consumption as calling interval() one by one; the rejection loop is branch-free inside a
  // power-of-two segment of i (the mask is constant there): a rejected value is simply overwritten.
  void shuffle_draws_scalar(int32_t* jb, int64_t hi) {
    int64_t i = hi;
    while (i >= 1) {
      const uint32_t mask = smear((uint32_t)i);
      const int64_t lo = (int64_t)(mask >> 1) + 1;          // smallest i with this mask
      while (i >= lo) {
        int n = avail();
        const uint32_t* o = out + pos;
        int used = 0;
        while (used < n && i >= lo) {
          const uint32_t v = o[used++] & mask;
          jb[i] = (int32_t)v;
          i -= (v <= (uint32_t)i);
        }
        pos += used;
      }
    }
  }
#if defined(__x86_64__)
  // AVX2 version: eight raw outputs per step.  With the running bound i, a masked value v <= i-7 is accepted
  // whatever happened to the lanes before it and v > i is rejected whatever happened; a group with a value in
  // between (probability ~7/mask per lane) is replayed by the scalar loop.  Accepted lanes are packed to the
  // high end in reverse order so that one 32-byte store puts them at jb[i], jb[i-1], ... .
  __attribute__((target("avx2"))) void shuffle_draws_avx2(int32_t* jb, int64_t hi) {
    static const PackLut lut;
    int64_t i = hi;
    while (i >= 1) {
      const uint32_t mask = smear((uint32_t)i);
      const int64_t lo = (int64_t)(mask >> 1) + 1;
      const __m256i vmask = _mm256_set1_epi32((int)mask);
      while (i >= lo) {
        int n = avail();
        const uint32_t* o = out + pos;
        int used = 0;
        while (used + 8 <= n && i >= lo + 8) {
          const __m256i v = _mm256_and_si256(_mm256_loadu_si256(reinterpret_cast<const __m256i*>(o + used)), vmask);
          const __m256i acc = _mm256_cmpgt_epi32(_mm256_set1_epi32((int)(i - 6)), v);      // v <= i - 7
          const __m256i rej = _mm256_cmpgt_epi32(v, _mm256_set1_epi32((int)i));            // v > i
          const int am = _mm256_movemask_ps(_mm256_castsi256_ps(acc));
          const int rm = _mm256_movemask_ps(_mm256_castsi256_ps(rej));
          if ((am | rm) != 0xFF) break;
          const __m256i idx = _mm256_load_si256(reinterpret_cast<const __m256i*>(lut.idx[am]));
          _mm256_storeu_si256(reinterpret_cast<__m256i*>(jb + i - 7), _mm256_permutevar8x32_epi32(v, idx));
          i -= __builtin_popcount((unsigned)am);
          used += 8;
        }
        for (int cnt = 0; cnt < 8 && used < n && i >= lo; ++cnt) {
          const uint32_t v = o[used++] & mask;
          jb[i] = (int32_t)v;
          i -= (v <= (uint32_t)i);
        }
        pos += used;
      }
    }
  }
#endif
  void shuffle_draws(int32_t* jb, int64_t hi) {
#if defined(__x86_64__)
    static const bool avx2 = __builtin_cpu_supports("avx2");
    if (avx2 && hi >= 64) {
      shuffle_draws_avx2(jb, hi);
      return;
    }
#endif
    shuffle_draws_scalar(jb, hi);
  }
};

struct Rec {
  int level, parent, kind;
  int64_t row_start, row_count;
  int first_child, n_child;
  int64_t knot_off;   // into knot_tree_row / kinds_local
};

// An ancestor's (or the node's own) knot lying inside a node: local position and knot slot.
struct KEnt {
  int32_t pos, id;
};

template <class T>
struct RawBuf {   // uninitialised, reused across calls (first-touch page faults are paid once per size)
  T* p = nullptr;
  size_t cap = 0;
  T* get(size_t n) {
    if (n > cap) {
      std::free(p);
      p = static_cast<T*>(std::aligned_alloc(64, (sizeof(T) * n + 63) / 64 * 64));   // cache-line aligned
      cap = p ? n : 0;
    }
    return p;
  }
  ~RawBuf() { std::free(p); }
};

struct Pool {
  RawBuf<int32_t> rows[2], scratch, draws, perm;
  RawBuf<double> xs[2], ys[2];
  RawBuf<uint8_t> code, slot_of;
  RawBuf<uint64_t> bits;
  size_t bits_n = 0;
};

struct Builder {
  int64_t N;
  int r, J, critDepth;
  MT rng;
  int32_t* rows[2];
  double *xs[2], *ys[2];
  int32_t *scratch, *draws, *perm;
  uint8_t *code, *slot_of;
  uint64_t* bits;                        // tracked-position bitmap of the reverse selection (all zero between nodes)
  static constexpr int64_t REVERSE_MAXBUF = 32768;   // size of the small-node draw buffer
  int64_t REVERSE_MIN = 16384;                        // nodes below this run the plain shuffle (MRA_REVERSE_MIN overrides)
  std::vector<Rec> rec;
  std::vector<int32_t> knot_tree_row;    // r per internal node, in reference knot order; resolved at the leaves
  std::vector<int32_t> kinds_local;      // r per internal node
  int status = 0;

#if defined(__x86_64__)
  // Tests eight draws at a time against the tracked-position bitmap (two 4-wide 64-bit gathers); a group with
  // a hit -- rare: about r ln(n/r) of n draws -- is replayed in order by the scalar step, which may change
  // the bitmap.  Returns the first index it did not process.
  template <class F>
  __attribute__((target("avx2"))) int64_t undo_scan_avx2(const int32_t* jb, int64_t from, int64_t n, F step) {
    int64_t i = from;
    const long long* b64 = reinterpret_cast<const long long*>(bits);
    const __m256i one = _mm256_set1_epi64x(1), m63 = _mm256_set1_epi64x(63);
    for (; i + 8 <= n; i += 8) {
      const __m128i j0 = _mm_loadu_si128(reinterpret_cast<const __m128i*>(jb + i));
      const __m128i j1 = _mm_loadu_si128(reinterpret_cast<const __m128i*>(jb + i + 4));
      const __m256i w0 = _mm256_i32gather_epi64(b64, _mm_srli_epi32(j0, 6), 8);
      const __m256i w1 = _mm256_i32gather_epi64(b64, _mm_srli_epi32(j1, 6), 8);
      const __m256i s0 = _mm256_and_si256(_mm256_cvtepu32_epi64(j0), m63);
      const __m256i s1 = _mm256_and_si256(_mm256_cvtepu32_epi64(j1), m63);
      const __m256i h0 = _mm256_and_si256(_mm256_srlv_epi64(w0, s0), one);
      const __m256i h1 = _mm256_and_si256(_mm256_srlv_epi64(w1, s1), one);
      if (!_mm256_testz_si256(_mm256_or_si256(h0, h1), one))
        for (int k = 0; k < 8; ++k) step(i + k);
    }
    return i;
  }
#endif

  // a[0..r) of np.random.permutation(n) on the legacy stream (== np.random.choice(arange(n), r, False))
  void first_r_of_permutation(int64_t n, int32_t* out) {
    if (n < REVERSE_MIN || r > 256) {
      // Small node: the work array is cache resident, run the legacy shuffle as it is.
      int32_t* a = scratch;
      int32_t* jb = draws;
      rng.shuffle_draws(jb, n - 1);
      for (int64_t i = 0; i < n; ++i) a[i] = (int32_t)i;
      for (int64_t i = n - 1; i >= 1; --i) {
        const int32_t j = jb[i];
        const int32_t t = a[i];
        a[i] = a[j];
        a[j] = t;
      }
      for (int p = 0; p < r; ++p) out[p] = a[p];
      return;
    }
    // Large node: only a[0..r) of the shuffled arange is needed.  The draws j_i (i = n-1 .. 1) are
    // generated in the legacy order and buffered; steps i >= r are then undone in reverse time order
    // while tracking where the values that end up in positions < r came from (a bitmap of the r tracked
    // positions stays cache resident, the 4n-byte work array is never touched at random); the last
    // r-1 steps only permute the prefix and are replayed forward.
    int32_t* jb = scratch;                             // jb[i] = j_i
    rng.shuffle_draws(jb, n - 1);
    int32_t key[256];
    for (int p = 0; p < r; ++p) {
      key[p] = p;
      bits[p >> 6] |= 1ull << (p & 63);
      slot_of[p] = (uint8_t)p;
    }
    auto undo_step = [&](int64_t i) {
      const int32_t j = jb[i];
      if ((bits[j >> 6] >> (j & 63)) & 1ull) {
        if (j == i) return;
        const uint8_t sl = slot_of[j];
        bits[j >> 6] &= ~(1ull << (j & 63));
        bits[i >> 6] |= 1ull << (i & 63);
        slot_of[i] = sl;
        key[sl] = (int32_t)i;
      }
    };
    int64_t i0 = r;
#if defined(__x86_64__)
    static const bool avx2 = __builtin_cpu_supports("avx2");
    if (avx2) i0 = undo_scan_avx2(jb, r, n, undo_step);
#endif
    for (int64_t i = i0; i < n; ++i) undo_step(i);
    for (int p = 0; p < r; ++p) {
      out[p] = key[p];                                 // initial array is arange: value == position
      bits[key[p] >> 6] &= ~(1ull << (key[p] & 63));
    }
    for (int i = r - 1; i >= 1; --i) {
      const int32_t j = jb[i];
      const int32_t t = out[i];
      out[i] = out[j];
      out[j] = t;
    }
  }

  // node occupying [s, e) of buffer b; kp = knots of its ancestors lying inside it (sorted by position),
  // (sx, sy) = column sums of its locations accumulated in row order from 0.0 (np.mean's order)
  void visit(int parent, int level, int64_t s, int64_t e, int b, int levels_left, const std::vector<KEnt>& kp,
             double sx, double sy) {
    if (status) return;
    const int me = (int)rec.size();
    rec.push_back(Rec{level, parent, MRA_NODE_LEAF, s, e - s, -1, 0, -1});
    const int64_t n = e - s;
    const int32_t* R = rows[b] + s;
    const double* X = xs[b] + s;
    const double* Y = ys[b] + s;
    const int64_t n_nk = n - (int64_t)kp.size();
    const bool internal = levels_left > 0 && n_nk > std::max(r, J);
    if (!internal) {
      std::memcpy(perm + s, R, sizeof(int32_t) * n);
      for (const KEnt& k : kp) knot_tree_row[k.id] = (int32_t)(s + k.pos);
      return;
    }
    if (n_nk <= 100 || n <= 100) {   // KMeans knot / split paths: not handled natively
      status = 1;
      return;
    }
    // ---- knots: sorted permutation(n_nk)[:r], candidate index -> local position by skipping the
    // ancestors' knots (select on the sorted kp list instead of a scan of the node)
    int32_t pick[256];
    first_r_of_permutation(n_nk, pick);
    std::sort(pick, pick + r);
    rec[me].kind = MRA_NODE_INTERNAL;
    rec[me].knot_off = (int64_t)kinds_local.size();
    std::vector<KEnt> mk;             // kp merged with the new knots, sorted by position
    mk.reserve(kp.size() + r);
    {
      size_t j = 0;
      for (int t = 0; t < r; ++t) {
        int32_t pos = pick[t] + (int32_t)j;
        while (j < kp.size() && kp[j].pos <= pos) {
          mk.push_back(kp[j]);
          ++j;
          ++pos;
        }
        const int32_t id = (int32_t)kinds_local.size();
        kinds_local.push_back(pos);
        knot_tree_row.push_back(-1);
        mk.push_back(KEnt{pos, id});
      }
      for (; j < kp.size(); ++j) mk.push_back(kp[j]);
    }
    // ---- quadrant split by column means (MRANode.py:232-239)
    const double mx = sx / (double)n, my = sy / (double)n;
    uint8_t* C = code + s;
    int64_t nx = 0, ny = 0, nxy = 0;
    for (int64_t i = 0; i < n; ++i) {
      const int gx = X[i] > mx, gy = Y[i] > my;     // !(x <= mx): NaN-free inputs
      C[i] = (uint8_t)(2 * gx + gy);
      nx += gx;
      ny += gy;
      nxy += gx & gy;
    }
    const int64_t cnt[4] = {n - nx - ny + nxy, ny - nxy, nx - nxy, nxy};
    for (int c = 0; c < 4; ++c)
      if (cnt[c] == 0) {
        status = 1;   // the reference would build an empty child here; not handled natively
        return;
      }
    int64_t off[4];
    off[0] = s;
    for (int c = 1; c < 4; ++c) off[c] = off[c - 1] + cnt[c - 1];
    const int nb = b ^ 1;
    std::vector<KEnt> ckp[4];
    double csx[4] = {0.0, 0.0, 0.0, 0.0}, csy[4] = {0.0, 0.0, 0.0, 0.0};
    {
      int64_t w[4] = {off[0], off[1], off[2], off[3]};
      int32_t* R2 = rows[nb];
      double* X2 = xs[nb];
      double* Y2 = ys[nb];
      size_t q = 0;
      int64_t next_k = mk.empty() ? n : mk[0].pos;
      int64_t i = 0;
      while (i < n) {
        const int64_t stop = std::min(next_k, n);
        for (; i < stop; ++i) {                      // plain stretch between two knots
          const int c = C[i];
          const int64_t d = w[c]++;
          R2[d] = R[i];
          const double x = X[i], y = Y[i];
          X2[d] = x;
          Y2[d] = y;
          csx[c] += x;
          csy[c] += y;
        }
        if (i < n) {                                 // i is a knot position: same move, remember where it went
          const int c = C[i];
          const int64_t d = w[c]++;
          R2[d] = R[i];
          const double x = X[i], y = Y[i];
          X2[d] = x;
          Y2[d] = y;
          csx[c] += x;
          csy[c] += y;
          ckp[c].push_back(KEnt{(int32_t)(d - off[c]), mk[q].id});
          ++q;
          next_k = q < mk.size() ? mk[q].pos : n;
          ++i;
        }
      }
    }
    const bool fork = level == critDepth;
    MT saved;
    if (fork) saved = rng;
    rec[me].n_child = 4;
    for (int c = 0; c < 4; ++c) {
      if (fork) rng = saved;
      if (c == 0) rec[me].first_child = (int)rec.size();
      visit(me, level + 1, off[c], off[c] + cnt[c], nb, levels_left - 1, ckp[c], csx[c], csy[c]);
      if (status) return;
    }
    if (fork) rng = saved;
  }
};


// ------------------------------------------------------------------------------------------------
// Two-phase build for regular trees (every node above level M has > 100 rows and four non-empty
// quadrants).  The quadrant partition of the locations does not depend on the RNG, so worker threads
// compute it level by level (BFS) and publish, per level, the quadrant code of every position as two
// bit planes with cumulative block counts.  The calling thread replays the reference's DFS / legacy RNG
// stream on top of that with O(1) rank queries: child ranges, the quadrant of every knot and its position
// inside the child all come from the bit planes, so it never touches the 20 bytes/location row data.
struct LevelBits {
  std::vector<uint64_t> lo, hi;    // code = 2*[x > mean_x] + [y > mean_y]: hi = x bit, lo = y bit
  std::vector<int32_t> cum;        // cum[4*b + c] = #positions in [base, base + 64*b) with code c
  int64_t base = 0;                // first position covered (0 for a whole level, the subtree start otherwise)
  inline int code(int64_t x) const {
    x -= base;
    const int64_t b = x >> 6;
    const int sh = (int)(x & 63);
    return (int)(((hi[b] >> sh) & 1ull) * 2 + ((lo[b] >> sh) & 1ull));
  }
  inline int64_t rank(int c, int64_t x) const {      // #positions in [base, x) with code c
    x -= base;
    const int64_t b = x >> 6;
    const uint64_t wl = (c & 1) ? lo[b] : ~lo[b], wh = (c & 2) ? hi[b] : ~hi[b];
    const uint64_t low = (x & 63) ? ((1ull << (x & 63)) - 1ull) : 0ull;
    return (int64_t)cum[4 * b + c] + __builtin_popcountll(wl & wh & low);
  }
};

template <class F>
void run_parallel(int nthreads, int64_t nitems, F fn) {    // fn(item) for item in [0, nitems)
  if (nitems <= 0) return;
  const int nt = (int)std::min<int64_t>(nthreads, nitems);
  if (nt <= 1) {
    for (int64_t i = 0; i < nitems; ++i) fn(i);
    return;
  }
  std::atomic<int64_t> next{0};
  std::vector<std::thread> th;
  for (int t = 0; t < nt; ++t)
    th.emplace_back([&] {
      for (int64_t i; (i = next.fetch_add(1)) < nitems;) fn(i);
    });
  for (auto& t : th) t.join();
}

// Stable 4-way scatter of rows / coordinates with software write combining: every destination stream collects
// one cache line (8 doubles / 16 row ids) in an L1-resident buffer and writes it with non-temporal stores, so
// the scattered data neither costs a read-for-ownership nor evicts the caches the RNG replay lives in.  The
// destination arrays are 64-byte aligned (RawBuf); partial lines at the two ends of a stream are written with
// plain stores because their other halves belong to the neighbouring chunk.  Measured on the GPU box's host:
// the two cooperative levels finish 1.4 ms earlier (19.2 -> 17.8 ms), the subtree levels and the total do not move.
#if defined(__x86_64__)
struct WcScatter {
  alignas(64) double bx[4][8];
  alignas(64) double by[4][8];
  alignas(64) int32_t br[4][16];
  int64_t w[4], w0[4];
  int32_t* R2;
  double *X2, *Y2;

  void begin(const int64_t start[4], int32_t* r2, double* x2, double* y2) {
    for (int c = 0; c < 4; ++c) w[c] = w0[c] = start[c];
    R2 = r2;
    X2 = x2;
    Y2 = y2;
  }
  __attribute__((target("avx"))) static inline void line_pd(double* dst, const double* src) {
    _mm256_stream_pd(dst, _mm256_load_pd(src));
    _mm256_stream_pd(dst + 4, _mm256_load_pd(src + 4));
  }
  __attribute__((target("avx"))) static inline void line_i32(int32_t* dst, const int32_t* src) {
    _mm256_stream_si256(reinterpret_cast<__m256i*>(dst), _mm256_load_si256(reinterpret_cast<const __m256i*>(src)));
    _mm256_stream_si256(reinterpret_cast<__m256i*>(dst + 8), _mm256_load_si256(reinterpret_cast<const __m256i*>(src + 8)));
  }
  __attribute__((target("avx"))) inline void put(int c, int32_t row, double x, double y) {
    const int64_t d = w[c]++;
    const int s8 = (int)(d & 7), s16 = (int)(d & 15);
    bx[c][s8] = x;
    by[c][s8] = y;
    br[c][s16] = row;
    if (s8 == 7) {
      const int64_t l0 = d - 7;
      if (l0 >= w0[c]) {
        line_pd(X2 + l0, bx[c]);
        line_pd(Y2 + l0, by[c]);
      } else {                                   // first, partial line of this stream
        for (int64_t i = w0[c]; i <= d; ++i) {
          X2[i] = bx[c][i & 7];
          Y2[i] = by[c][i & 7];
        }
      }
      if (s16 == 15) {
        const int64_t r0 = d - 15;
        if (r0 >= w0[c]) line_i32(R2 + r0, br[c]);
        else
          for (int64_t i = w0[c]; i <= d; ++i) R2[i] = br[c][i & 15];
      }
    }
  }
  __attribute__((target("avx"))) void end() {
    for (int c = 0; c < 4; ++c) {
      const int64_t e = w[c];
      for (int64_t i = std::max(w0[c], e & ~(int64_t)7); i < e; ++i) {
        X2[i] = bx[c][i & 7];
        Y2[i] = by[c][i & 7];
      }
      for (int64_t i = std::max(w0[c], e & ~(int64_t)15); i < e; ++i) R2[i] = br[c][i & 15];
    }
    _mm_sfence();
  }
};
#endif

struct Partitioner {
  int64_t N = 0;
  int M = 0, nthreads = 1;
  int32_t* rows[2];
  double *xs[2], *ys[2];
  uint8_t* code = nullptr;
  std::vector<LevelBits> lv;                 // levels < S: whole-level bit planes
  int S = 0;                                 // levels >= S are partitioned subtree by subtree (4^S subtrees)
  std::vector<std::vector<LevelBits>> sub;   // sub[k][L - S]: bit planes of subtree k at level L
  std::unique_ptr<std::atomic<int>[]> sub_ready;   // sub_ready[k] = number of levels (from S) of subtree k published
  std::atomic<int> ready{0};
  std::atomic<int> failed{0};
  std::atomic<int> all_done{0};              // partition finished (and, in stream mode, the output arrays filled)
  struct PNode {
    int64_t s, e;
    double sx, sy;
  };
  std::vector<std::vector<PNode>> level_nodes;   // [L][index inside level]: every node's row range
  std::function<void()> on_done;                 // stream mode: fills the caller's node arrays / perm
  Trace trace;
  bool use_nt = false;                           // write-combined non-temporal scatter (AVX hosts; MRA_BUILD_NT=0 disables)
  static constexpr int64_t BIG = 1 << 18;

  // codes and counts of positions [a, b) of a node with means (mx, my)
  static void pass1(const double* X, const double* Y, uint8_t* C, int64_t a, int64_t b, double mx, double my,
                    int64_t cnt[4]) {
    int64_t nx = 0, ny = 0, nxy = 0;
    for (int64_t i = a; i < b; ++i) {
      const int gx = X[i] > mx, gy = Y[i] > my;
      C[i] = (uint8_t)(2 * gx + gy);
      nx += gx;
      ny += gy;
      nxy += gx & gy;
    }
    cnt[0] = (b - a) - nx - ny + nxy;
    cnt[1] = ny - nxy;
    cnt[2] = nx - nxy;
    cnt[3] = nxy;
  }

  // A node is partitioned in two passes: classify() writes the quadrant codes of its rows and counts them (all
  // the bit planes need, so a level can be published to the RNG replay before any row has moved), scatter() then
  // moves rows / coordinates into the children's ranges and forms the children's column sums.
  struct NodePlan {
    int64_t cnt[4] = {0, 0, 0, 0}, off[4] = {0, 0, 0, 0};
    std::vector<int64_t> cc;      // cooperative path: per-chunk counts
    int nc = 0;
    int64_t step = 0;
  };

  void classify(const PNode& nd, int b, NodePlan& pl) {
    const int64_t n = nd.e - nd.s;
    if (n <= 100) {
      failed.store(1);
      return;
    }
    const double mx = nd.sx / (double)n, my = nd.sy / (double)n;
    const double *X = xs[b], *Y = ys[b];
    if (n < BIG) {
      pass1(X, Y, code, nd.s, nd.e, mx, my, pl.cnt);
    } else {
      // cooperative path: all threads on one node, chunked codes
      pl.nc = nthreads * 4;
      pl.step = (n + pl.nc - 1) / pl.nc;
      pl.cc.assign((size_t)pl.nc * 4, 0);
      run_parallel(nthreads, pl.nc, [&](int64_t k) {
        const int64_t a = nd.s + k * pl.step, bb = std::min(nd.e, a + pl.step);
        if (a < bb) pass1(X, Y, code, a, bb, mx, my, &pl.cc[4 * k]);
      });
      for (int c = 0; c < 4; ++c) {
        pl.cnt[c] = 0;
        for (int k = 0; k < pl.nc; ++k) pl.cnt[c] += pl.cc[4 * k + c];
      }
    }
    pl.off[0] = nd.s;
    for (int c = 1; c < 4; ++c) pl.off[c] = pl.off[c - 1] + pl.cnt[c - 1];
    for (int c = 0; c < 4; ++c)
      if (pl.cnt[c] == 0) failed.store(1);
  }

  void scatter(const PNode& nd, PNode* ch, int b, const NodePlan& pl) {
    const int64_t n = nd.e - nd.s;
    const int32_t* R = rows[b];
    const double *X = xs[b], *Y = ys[b];
    int32_t* R2 = rows[b ^ 1];
    double *X2 = xs[b ^ 1], *Y2 = ys[b ^ 1];
    const int64_t* cnt = pl.cnt;
    const int64_t* off = pl.off;
    if (n < BIG) {
      double csx[4] = {0, 0, 0, 0}, csy[4] = {0, 0, 0, 0};
#if defined(__x86_64__)
      if (use_nt) {
        WcScatter wc;
        wc.begin(off, R2, X2, Y2);
        for (int64_t i = nd.s; i < nd.e; ++i) {
          const int c = code[i];
          const double x = X[i], y = Y[i];
          wc.put(c, R[i], x, y);
          csx[c] += x;
          csy[c] += y;
        }
        wc.end();
      } else
#endif
      {
        int64_t w[4] = {off[0], off[1], off[2], off[3]};
        for (int64_t i = nd.s; i < nd.e; ++i) {
          const int c = code[i];
          const int64_t d = w[c]++;
          R2[d] = R[i];
          const double x = X[i], y = Y[i];
          X2[d] = x;
          Y2[d] = y;
          csx[c] += x;
          csy[c] += y;
        }
      }
      for (int c = 0; c < 4; ++c) ch[c] = PNode{off[c], off[c] + cnt[c], csx[c], csy[c]};
    } else {
      // cooperative path: chunked stable scatter, then the children's sums one thread each, in row order as
      // np.mean needs
      const int nc = pl.nc;
      const int64_t step = pl.step;
      std::vector<int64_t> w0((size_t)nc * 4);
      for (int c = 0; c < 4; ++c) {
        int64_t acc = off[c];
        for (int k = 0; k < nc; ++k) {
          w0[4 * k + c] = acc;
          acc += pl.cc[4 * k + c];
        }
      }
      run_parallel(nthreads, nc, [&](int64_t k) {
        const int64_t a = nd.s + k * step, bb = std::min(nd.e, a + step);
#if defined(__x86_64__)
        if (use_nt) {
          WcScatter wc;
          wc.begin(&w0[4 * k], R2, X2, Y2);
          for (int64_t i = a; i < bb; ++i) wc.put(code[i], R[i], X[i], Y[i]);
          wc.end();
          return;
        }
#endif
        int64_t w[4] = {w0[4 * k], w0[4 * k + 1], w0[4 * k + 2], w0[4 * k + 3]};
        for (int64_t i = a; i < bb; ++i) {
          const int64_t d = w[code[i]]++;
          R2[d] = R[i];
          X2[d] = X[i];
          Y2[d] = Y[i];
        }
      });
      run_parallel(nthreads, 4, [&](int64_t c) {
        double sx = 0.0, sy = 0.0;
        for (int64_t i = off[c]; i < off[c] + cnt[c]; ++i) {
          sx += X2[i];
          sy += Y2[i];
        }
        ch[c] = PNode{off[c], off[c] + cnt[c], sx, sy};
      });
    }
  }

  // bit planes + cumulative block counts of positions [s0, e0) from the code bytes; nt threads
  void build_bits(LevelBits& lb, int64_t s0, int64_t e0, int nt) {
    const int64_t n = e0 - s0;
    const int64_t nblk = (n >> 6) + 2;
    lb.base = s0;
    lb.lo.assign(nblk, 0);
    lb.hi.assign(nblk, 0);
    lb.cum.assign(nblk * 4, 0);
    const int nc = std::max(1, nt * 2);
    const int64_t bstep = (nblk + nc - 1) / nc;
    std::vector<int64_t> cc((size_t)nc * 4, 0);
    const uint8_t* cd0 = code + s0;
    run_parallel(nt, nc, [&](int64_t k) {
      const int64_t b0 = k * bstep, b1 = std::min(nblk, b0 + bstep);
      int64_t c4[4] = {0, 0, 0, 0};
      for (int64_t b = b0; b < b1; ++b) {
        uint64_t wl = 0, wh = 0;
        const int64_t x0 = b << 6, x1 = std::min(n, x0 + 64);
        for (int64_t x = x0; x < x1; ++x) {
          const uint64_t cd = cd0[x];
          wl |= (cd & 1ull) << (x - x0);
          wh |= (cd >> 1) << (x - x0);
        }
        lb.lo[b] = wl;
        lb.hi[b] = wh;
        const int64_t nvalid = std::max<int64_t>(0, x1 - x0);
        const uint64_t valid = nvalid >= 64 ? ~0ull : ((1ull << nvalid) - 1ull);
        lb.cum[4 * b + 0] = (int32_t)__builtin_popcountll(~wl & ~wh & valid);   // per-block counts for now
        lb.cum[4 * b + 1] = (int32_t)__builtin_popcountll(wl & ~wh & valid);
        lb.cum[4 * b + 2] = (int32_t)__builtin_popcountll(~wl & wh & valid);
        lb.cum[4 * b + 3] = (int32_t)__builtin_popcountll(wl & wh & valid);
        for (int c = 0; c < 4; ++c) c4[c] += lb.cum[4 * b + c];
      }
      for (int c = 0; c < 4; ++c) cc[4 * k + c] = c4[c];
    });
    std::vector<int64_t> base((size_t)nc * 4, 0);
    for (int c = 0; c < 4; ++c) {
      int64_t acc = 0;
      for (int k = 0; k < nc; ++k) {
        base[4 * k + c] = acc;
        acc += cc[4 * k + c];
      }
    }
    run_parallel(nt, nc, [&](int64_t k) {
      const int64_t b0 = k * bstep, b1 = std::min(nblk, b0 + bstep);
      int64_t run[4] = {base[4 * k], base[4 * k + 1], base[4 * k + 2], base[4 * k + 3]};
      for (int64_t b = b0; b < b1; ++b)
        for (int c = 0; c < 4; ++c) {
          const int32_t own = lb.cum[4 * b + c];
          lb.cum[4 * b + c] = (int32_t)run[c];
          run[c] += own;
        }
    });
  }

  // levels S .. M-1 of the subtree rooted at `root` (a level-S node), one thread, BFS inside the subtree
  void run_subtree(int k, const PNode& root) {
    std::vector<PNode> cur{root}, next;
    std::vector<NodePlan> plans;
    for (int L = S; L < M; ++L) {
      next.assign(cur.size() * 4, PNode{0, 0, 0.0, 0.0});
      plans.assign(cur.size(), NodePlan());
      for (size_t i = 0; i < cur.size(); ++i) classify(cur[i], L & 1, plans[i]);
      if (failed.load()) return;
      build_bits(sub[k][L - S], root.s, root.e, 1);
      sub_ready[k].store(L - S + 1, std::memory_order_release);     // levels S .. L of this subtree are published
      for (size_t i = 0; i < cur.size(); ++i) scatter(cur[i], &next[4 * i], L & 1, plans[i]);
      cur.swap(next);
      std::copy(cur.begin(), cur.end(), level_nodes[L + 1].begin() + (size_t)k * cur.size());
    }
  }

  void run(double sx0, double sy0) {
    trace.mark("partition: level-0 arrays ready");
    std::vector<PNode> cur{PNode{0, N, sx0, sy0}}, next;
    level_nodes.assign(M + 1, {});
    for (int L = 0; L <= M; ++L) level_nodes[L].resize((size_t)1 << (2 * L));
    for (int L = 0; L < S; ++L) {
      level_nodes[L] = cur;
      next.assign(cur.size() * 4, PNode{0, 0, 0.0, 0.0});
      const int b = L & 1;
      // big nodes one after the other (all threads inside), the rest in parallel; the level's bit planes are
      // published between the two passes, so the RNG replay does not wait for the row moves
      std::vector<int64_t> small;
      std::vector<NodePlan> plans(cur.size());
      for (size_t i = 0; i < cur.size(); ++i) {
        if (cur[i].e - cur[i].s >= BIG) classify(cur[i], b, plans[i]);
        else small.push_back((int64_t)i);
      }
      run_parallel(nthreads, (int64_t)small.size(), [&](int64_t k) {
        const int64_t i = small[k];
        classify(cur[i], b, plans[i]);
      });
      if (failed.load()) break;
      build_bits(lv[L], 0, N, nthreads);
      ready.store(L + 1, std::memory_order_release);
      trace.mark("partition: level published", L);
      for (size_t i = 0; i < cur.size(); ++i)
        if (cur[i].e - cur[i].s >= BIG) scatter(cur[i], &next[4 * i], b, plans[i]);
      run_parallel(nthreads, (int64_t)small.size(), [&](int64_t k) {
        const int64_t i = small[k];
        scatter(cur[i], &next[4 * i], b, plans[i]);
      });
      trace.mark("partition: level done", L);
      cur.swap(next);
    }
    const int nsub = (int)sub.size();
    if (!failed.load()) level_nodes[S] = cur;
    if (!failed.load() && S < M) {
      // the subtrees below level S are independent: workers take them in DFS order, so the RNG replay (which
      // visits them in the same order) rarely has to wait
      std::atomic<int> nextk{0};
      std::vector<std::thread> th;
      for (int t = 0; t < std::min(nthreads, nsub); ++t)
        th.emplace_back([&] {
          for (int k; (k = nextk.fetch_add(1)) < nsub;) {
            if (!failed.load()) run_subtree(k, cur[k]);
            sub_ready[k].store(M + 1, std::memory_order_release);
          }
        });
      for (auto& t : th) t.join();
      trace.mark("partition: subtrees done");
    }
    if (failed.load()) {
      ready.store(M + 1, std::memory_order_release);
      for (int k = 0; k < nsub; ++k) sub_ready[k].store(M + 1, std::memory_order_release);
    } else if (on_done) {
      on_done();
      trace.mark("partition: output arrays filled");
    }
    all_done.store(1, std::memory_order_release);
  }
};

// Stream mode (mra_build_structure_2d_stream): for a regular tree the BFS numbering is known in closed form,
// so node arrays and the permutation are final as soon as the partition is (stage 0), and the knot rows of a
// level-1 subtree are final when the RNG replay leaves it (stage 1) -- the caller can start device work on a
// finished subtree while the replay continues with the next one.
struct StreamOut {
  std::function<void(int)> cb;          // cb(0): stage 0; cb(1 + c): level-1 subtree c finished
  int64_t* knot_rows = nullptr;
  int32_t* kinds_local = nullptr;
  int32_t* dfs_index = nullptr;
  std::vector<int64_t> level_off;      // BFS id of the first node of every level
  bool stage0_fired = false;
  int dfs_counter = 0;
  std::vector<KEnt> root_knots;        // the root's own knots (position at level 0, slot)
};

struct RankBuilder {
  Builder* B;
  Partitioner* P;
  int M;
  StreamOut* so = nullptr;

  const LevelBits& bits_for(int level, int64_t idx) const {
    if (level < P->S) return P->lv[level];
    return P->sub[idx >> (2 * (level - P->S))][level - P->S];
  }

  // tree-order rows of the root's knots by walking the finished bit planes, then the stage-0 callback
  void fire_stage0() {
    so->stage0_fired = true;
    for (const KEnt& k : so->root_knots) {
      int64_t x = k.pos, s0 = 0, e0 = P->N, idx = 0;
      for (int L = 0; L < M; ++L) {
        const LevelBits& lb = bits_for(L, idx);
        const int c = lb.code(x);
        int64_t off = s0;
        for (int cc = 0; cc < c; ++cc) off += lb.rank(cc, e0) - lb.rank(cc, s0);
        const int64_t cnt = lb.rank(c, e0) - lb.rank(c, s0);
        x = off + lb.rank(c, x) - lb.rank(c, s0);
        s0 = off;
        e0 = off + cnt;
        idx = 4 * idx + c;
      }
      so->knot_rows[k.id] = x;
    }
    if (so->cb) so->cb(0);
  }

  // idx: index of the node inside its level (regular 4-ary tree), which also names its level-S subtree
  void visit(int parent, int level, int64_t idx, int64_t s, int64_t e, int levels_left, const std::vector<KEnt>& kp) {
    Builder& b = *B;
    if (b.status) return;
    const int me = (int)b.rec.size();
    b.rec.push_back(Rec{level, parent, MRA_NODE_LEAF, s, e - s, -1, 0, -1});
    if (so) {
      so->dfs_index[so->level_off[level] + idx] = so->dfs_counter++;
      if (!so->stage0_fired && level > 0 && P->all_done.load(std::memory_order_acquire) && !P->failed.load())
        fire_stage0();
    }
    const int64_t n = e - s;
    const int64_t n_nk = n - (int64_t)kp.size();
    const bool internal = levels_left > 0 && n_nk > std::max(b.r, b.J);
    if (!internal) {
      if (levels_left > 0) {       // ragged tree (early leaf): the serial builder handles it
        b.status = 1;
        return;
      }
      if (so) {
        for (const KEnt& k : kp)
          if (k.id >= b.r || !so->stage0_fired) so->knot_rows[k.id] = s + k.pos;   // root slots are final after stage 0
      } else {
        for (const KEnt& k : kp) b.knot_tree_row[k.id] = (int32_t)(s + k.pos);
      }
      return;
    }
    if (n_nk <= 100 || n <= 100) {
      b.status = 1;
      return;
    }
    int32_t pick[256];
    b.first_r_of_permutation(n_nk, pick);
    if (level == 0) P->trace.mark("replay: root draws done");
    std::sort(pick, pick + b.r);
    b.rec[me].kind = MRA_NODE_INTERNAL;
    b.rec[me].knot_off = (int64_t)b.kinds_local.size();
    std::vector<KEnt> mk;
    mk.reserve(kp.size() + b.r);
    {
      size_t j = 0;
      for (int t = 0; t < b.r; ++t) {
        int32_t pos = pick[t] + (int32_t)j;
        while (j < kp.size() && kp[j].pos <= pos) {
          mk.push_back(kp[j]);
          ++j;
          ++pos;
        }
        int32_t id;
        if (so) {            // slot in the caller's arrays: BFS id of this node * r + t
          id = (int32_t)((so->level_off[level] + idx) * b.r + t);
          so->kinds_local[id] = pos;
          if (level == 0) so->root_knots.push_back(KEnt{pos, id});
        } else {
          id = (int32_t)b.kinds_local.size();
          b.kinds_local.push_back(pos);
          b.knot_tree_row.push_back(-1);
        }
        mk.push_back(KEnt{pos, id});
      }
      for (; j < kp.size(); ++j) mk.push_back(kp[j]);
    }
    const LevelBits* lbp;
    if (level < P->S) {
      while (P->ready.load(std::memory_order_acquire) <= level) std::this_thread::yield();
      lbp = &P->lv[level];
    } else {
      const int64_t k = idx >> (2 * (level - P->S));            // ancestor at level S
      while (P->sub_ready[k].load(std::memory_order_acquire) <= level - P->S) std::this_thread::yield();
      lbp = &P->sub[k][level - P->S];
    }
    if (P->failed.load()) {
      b.status = 1;
      return;
    }
    const LevelBits& lb = *lbp;
    int64_t base[4], cnt[4], off[4];
    for (int c = 0; c < 4; ++c) {
      base[c] = lb.rank(c, s);
      cnt[c] = lb.rank(c, e) - base[c];
      if (cnt[c] == 0) {
        b.status = 1;
        return;
      }
    }
    off[0] = s;
    for (int c = 1; c < 4; ++c) off[c] = off[c - 1] + cnt[c - 1];
    std::vector<KEnt> ckp[4];
    for (const KEnt& k : mk) {
      const int64_t x = s + k.pos;
      const int c = lb.code(x);
      ckp[c].push_back(KEnt{(int32_t)(lb.rank(c, x) - base[c]), k.id});
    }
    const bool fork = level == b.critDepth;
    MT saved;
    if (fork) saved = b.rng;
    b.rec[me].n_child = 4;
    for (int c = 0; c < 4; ++c) {
      if (fork) b.rng = saved;
      if (c == 0) b.rec[me].first_child = (int)b.rec.size();
      visit(me, level + 1, 4 * idx + c, off[c], off[c] + cnt[c], levels_left - 1, ckp[c]);
      if (b.status) return;
      if (level == 0) P->trace.mark("replay: level-1 subtree done", c);
      if (so && level == 0) {          // a level-1 subtree is complete: its knot rows are final
        while (!P->all_done.load(std::memory_order_acquire)) std::this_thread::yield();
        if (P->failed.load()) {
          b.status = 1;
          return;
        }
        if (!so->stage0_fired) fire_stage0();
        if (so->cb) so->cb(1 + c);
      }
    }
    if (fork) b.rng = saved;
  }
};

// Scratch buffers shared by all builds of this process (one build at a time).
Pool g_pool;
std::mutex g_pool_mu;

bool prepare_builder(Builder& B, Pool& pool, int64_t N, int r, int J, int critDepth, const uint32_t* mt_key, int mt_pos) {
  B.N = N;
  B.r = r;
  B.J = J;
  B.critDepth = critDepth;
  std::memcpy(B.rng.key, mt_key, sizeof(uint32_t) * 624);
  B.rng.pos = mt_pos;
  for (int b = 0; b < 2; ++b) {
    B.rows[b] = pool.rows[b].get(N);
    B.xs[b] = pool.xs[b].get(N);
    B.ys[b] = pool.ys[b].get(N);
  }
  B.scratch = pool.scratch.get(N);
  if (const char* e = std::getenv("MRA_REVERSE_MIN")) B.REVERSE_MIN = std::max<int64_t>(64, std::min<int64_t>(Builder::REVERSE_MAXBUF, std::atoll(e)));
  B.draws = pool.draws.get(std::min<int64_t>(N, Builder::REVERSE_MAXBUF) + 1);
  B.perm = pool.perm.get(N);
  B.code = pool.code.get(N);
  B.slot_of = pool.slot_of.get(N);
  {
    const size_t nb = (size_t)(N + 63) / 64 + 1;
    const bool fresh = nb > pool.bits.cap;
    B.bits = pool.bits.get(nb);
    if (B.bits && (fresh || pool.bits_n < nb)) std::memset(B.bits, 0, sizeof(uint64_t) * pool.bits.cap);
    pool.bits_n = pool.bits.cap;
  }
  return B.rows[0] && B.rows[1] && B.xs[0] && B.xs[1] && B.ys[0] && B.ys[1] && B.scratch && B.draws && B.perm &&
         B.code && B.slot_of && B.bits;
}

// worker threads of the partition: all cores but one (the RNG replay keeps the calling thread busy), at most 12;
// MRA_HOST_THREADS lowers it (several ranks on one host), MRA_BUILD_THREADS sets it outright
int build_threads() {
  const int hw = (int)std::max(1u, std::thread::hardware_concurrency());
  int nt = std::max(1, std::min(12, hw - 1));
  if (hw <= 8) nt = hw;
  if (const char* e = std::getenv("MRA_HOST_THREADS")) nt = std::max(1, std::min(nt, std::atoi(e)));
  if (const char* e = std::getenv("MRA_BUILD_THREADS")) nt = std::max(1, std::min(64, std::atoi(e)));
  return nt;
}

// level-0 arrays (parallel) and the root's column sums (one thread, sequential: np.mean's order)
void init_level0(Builder& B, const double* locs, int64_t N, int nthreads, double& sx, double& sy) {
  std::thread summer([&] {
    double a = 0.0, b = 0.0;
    for (int64_t i = 0; i < N; ++i) {
      a += locs[2 * i];
      b += locs[2 * i + 1];
    }
    sx = a;
    sy = b;
  });
  const int64_t nchunk = std::max(1, nthreads) * 4, step = (N + nchunk - 1) / nchunk;
  int32_t* R = B.rows[0];
  double *X = B.xs[0], *Y = B.ys[0];
  run_parallel(std::max(1, nthreads - 1), nchunk, [&](int64_t k) {
    const int64_t i0 = k * step, i1 = std::min(N, i0 + step);
    for (int64_t i = i0; i < i1; ++i) {
      R[i] = (int32_t)i;
      X[i] = locs[2 * i];
      Y[i] = locs[2 * i + 1];
    }
  });
  summer.join();
}

void setup_partitioner(Partitioner& P, Builder& B, int64_t N, int M) {
  P.N = N;
  P.M = M;
  P.nthreads = build_threads();
  for (int b = 0; b < 2; ++b) {
    P.rows[b] = B.rows[b];
    P.xs[b] = B.xs[b];
    P.ys[b] = B.ys[b];
  }
  P.code = B.code;
#if defined(__x86_64__)
  P.use_nt = __builtin_cpu_supports("avx");
  if (const char* e = std::getenv("MRA_BUILD_NT")) P.use_nt = P.use_nt && std::atoi(e) != 0;
#endif
  P.S = std::min(2, (int)M);
  if (const char* e = std::getenv("MRA_BUILD_S")) P.S = std::max(1, std::min((int)M, std::atoi(e)));
  P.lv.resize(P.S);
  const int nsub = P.S < M ? 1 << (2 * P.S) : 0;
  P.sub.assign(nsub, std::vector<LevelBits>(M - P.S));
  P.sub_ready.reset(new std::atomic<int>[std::max(1, nsub)]);
  for (int k = 0; k < nsub; ++k) P.sub_ready[k].store(0);
}

}  // namespace

// Streaming build (mra_build_stream_*): the build runs on its own thread and reports progress events.
struct mra_build_job {
  std::thread th;
  std::mutex mu;
  std::condition_variable cv;
  int fired = 0;          // events 0 .. fired-1 have happened (0: partition ready, 1 + c: level-1 subtree c final)
  int status = MRA_OK;    // valid once done
  bool done = false;
  int32_t n_nodes = 0, depth = 0;
  int64_t n_knot_rows = 0;
};

namespace {

struct StreamArgs {
  const double* locs;
  int64_t N;
  int32_t r, M, J, critDepth;
  uint32_t* mt_key;
  int32_t* mt_pos;
  int32_t *node_level, *node_parent, *node_kind;
  int64_t *node_row_start, *node_row_count;
  int32_t *node_child_start, *node_child_count;
  int64_t *node_knot_off, *knot_rows;
  int32_t* kinds_local;
  int64_t* perm;
  int32_t* dfs_index;
};

void stream_body(mra_build_job* job, StreamArgs a) {
  int status = MRA_OK;
  {
    std::lock_guard<std::mutex> pool_lock(g_pool_mu);
    Builder B;
    if (!prepare_builder(B, g_pool, a.N, a.r, a.J, a.critDepth, a.mt_key, *a.mt_pos)) {
      status = MRA_ERR_NOMEM;
    } else {
      const int64_t N = a.N;
      const int M = a.M;
      Partitioner P;
      setup_partitioner(P, B, N, M);
      StreamOut so;
      so.knot_rows = a.knot_rows;
      so.kinds_local = a.kinds_local;
      so.dfs_index = a.dfs_index;
      so.level_off.assign(M + 2, 0);
      for (int L = 0; L <= M; ++L) so.level_off[L + 1] = so.level_off[L] + ((int64_t)1 << (2 * L));
      so.cb = [job](int ev) {
        {
          std::lock_guard<std::mutex> lk(job->mu);
          job->fired = ev + 1;
        }
        job->cv.notify_all();
      };
      // node arrays in closed form (regular 4-ary tree, BFS ids) and the permutation, on the partition thread
      P.on_done = [&] {
        for (int L = 0; L <= M; ++L) {
          const int64_t cnt = (int64_t)1 << (2 * L);
          for (int64_t idx = 0; idx < cnt; ++idx) {
            const int64_t id = so.level_off[L] + idx;
            const Partitioner::PNode& nd = P.level_nodes[L][idx];
            a.node_level[id] = L;
            a.node_parent[id] = L ? (int32_t)(so.level_off[L - 1] + idx / 4) : -1;
            a.node_kind[id] = L < M ? MRA_NODE_INTERNAL : MRA_NODE_LEAF;
            a.node_row_start[id] = nd.s;
            a.node_row_count[id] = nd.e - nd.s;
            a.node_child_start[id] = L < M ? (int32_t)(so.level_off[L + 1] + 4 * idx) : -1;
            a.node_child_count[id] = L < M ? 4 : 0;
            a.node_knot_off[id] = L < M ? id * a.r : -1;
          }
        }
        const int32_t* src = P.rows[M & 1];
        int64_t* dst = a.perm;
        const int64_t nchunk = 64, step = (N + nchunk - 1) / nchunk;
        run_parallel(P.nthreads, nchunk, [&](int64_t k) {
          const int64_t i0 = k * step, i1 = std::min(N, i0 + step);
          for (int64_t i = i0; i < i1; ++i) dst[i] = src[i];
        });
      };
      double sx = 0.0, sy = 0.0;
      std::thread worker([&] {
        init_level0(B, a.locs, N, P.nthreads, sx, sy);
        P.run(sx, sy);
      });
      RankBuilder RB{&B, &P, M, &so};
      RB.visit(-1, 0, 0, 0, N, M, std::vector<KEnt>());
      worker.join();
      if (B.status || P.failed.load()) {
        status = MRA_BUILD_UNSUPPORTED;     // ragged tree: the caller's RNG state has not been touched
      } else {
        std::memcpy(a.mt_key, B.rng.key, sizeof(uint32_t) * 624);
        *a.mt_pos = B.rng.pos;
        job->n_nodes = (int32_t)so.level_off[M + 1];
        job->depth = M;
        job->n_knot_rows = so.level_off[M] * a.r;
      }
    }
  }
  {
    std::lock_guard<std::mutex> lk(job->mu);
    job->status = status;
    job->done = true;
  }
  job->cv.notify_all();
}

}  // namespace

extern "C" {

int mra_build_structure_2d(const double* locs, int64_t n_locs, int32_t r, int32_t M, int32_t J,
                           int32_t critDepth, uint32_t* mt_key, int32_t* mt_pos, int32_t max_nodes,
                           int32_t* n_nodes_out, int32_t* depth_out, int32_t* node_level, int32_t* node_parent,
                           int32_t* node_kind, int64_t* node_row_start, int64_t* node_row_count,
                           int32_t* node_child_start, int32_t* node_child_count, int64_t* node_knot_off,
                           int64_t* knot_rows, int32_t* kinds_local, int64_t* n_knot_rows_out, int64_t* perm,
                           int32_t* dfs_index) {
  if (!locs || n_locs <= 0 || n_locs >= (int64_t(1) << 31) || r < 1 || !mt_key || !mt_pos) return MRA_ERR_ARG;
  if (r > 256) return MRA_BUILD_UNSUPPORTED;
  std::lock_guard<std::mutex> pool_lock(g_pool_mu);
  Pool& pool = g_pool;
  Builder B;
  if (!prepare_builder(B, pool, n_locs, r, J, critDepth, mt_key, *mt_pos)) return MRA_ERR_NOMEM;
  const int64_t N = n_locs;
  double sx = 0.0, sy = 0.0;
  const int nthreads0 = build_threads();
  auto init_level0 = [&] { ::init_level0(B, locs, N, nthreads0, sx, sy); };
  bool done = false;
  if (M >= 1 && N >= (int64_t)1 << 16) {
    // threaded two-phase build; falls back to the serial DFS below when the tree is not regular
    Partitioner P;
    setup_partitioner(P, B, N, M);
    std::thread worker([&] {     // the root's draws need nothing but N: the RNG replay starts right away
      init_level0();
      P.run(sx, sy);
    });
    RankBuilder RB{&B, &P, M};
    RB.visit(-1, 0, 0, 0, N, M, std::vector<KEnt>());
    worker.join();
    if (!B.status && !P.failed.load()) {
      std::memcpy(B.perm, B.rows[M & 1], sizeof(int32_t) * N);
      done = true;
    } else {
      // restart serially from the caller's RNG state and the original row order
      B.status = 0;
      B.rec.clear();
      B.kinds_local.clear();
      B.knot_tree_row.clear();
      std::memcpy(B.rng.key, mt_key, sizeof(uint32_t) * 624);
      B.rng.pos = *mt_pos;
      B.rng.out_valid = false;
      init_level0();
    }
  } else {
    init_level0();
  }
  if (!done) B.visit(-1, 0, 0, N, 0, M, std::vector<KEnt>(), sx, sy);
  if (B.status) return MRA_BUILD_UNSUPPORTED;
  const int nn = (int)B.rec.size();
  if (nn > max_nodes) return MRA_ERR_NOMEM;
  // BFS renumbering: stable by level, DFS pre-order inside a level
  int depth = 0;
  for (auto& rc : B.rec) depth = std::max(depth, rc.level);
  std::vector<int> count(depth + 2, 0), newid(nn);
  for (auto& rc : B.rec) ++count[rc.level + 1];
  for (int l = 0; l <= depth; ++l) count[l + 1] += count[l];
  {
    std::vector<int> next(count.begin(), count.end() - 1);
    for (int i = 0; i < nn; ++i) newid[i] = next[B.rec[i].level]++;
  }
  int64_t koff = 0;
  // internal nodes must receive knot offsets in BFS order (like the Python builder)
  std::vector<int> order(nn);
  for (int i = 0; i < nn; ++i) order[newid[i]] = i;
  for (int id = 0; id < nn; ++id) {
    const Rec& rc = B.rec[order[id]];
    node_level[id] = rc.level;
    node_parent[id] = rc.parent < 0 ? -1 : newid[rc.parent];
    node_kind[id] = rc.kind;
    node_row_start[id] = rc.row_start;
    node_row_count[id] = rc.row_count;
    node_child_start[id] = rc.n_child ? newid[rc.first_child] : -1;
    node_child_count[id] = rc.n_child;
    dfs_index[id] = order[id];
    if (rc.kind == MRA_NODE_INTERNAL) {
      node_knot_off[id] = koff;
      for (int k = 0; k < r; ++k) {
        knot_rows[koff + k] = B.knot_tree_row[rc.knot_off + k];
        kinds_local[koff + k] = B.kinds_local[rc.knot_off + k];
      }
      koff += r;
    } else {
      node_knot_off[id] = -1;
    }
  }
  for (int64_t i = 0; i < N; ++i) perm[i] = B.perm[i];
  *n_nodes_out = nn;
  *depth_out = depth;
  *n_knot_rows_out = koff;
  std::memcpy(mt_key, B.rng.key, sizeof(uint32_t) * 624);
  *mt_pos = B.rng.pos;
  return MRA_OK;
}

int mra_build_stream_start(const double* locs, int64_t n_locs, int32_t r, int32_t M, int32_t J, int32_t critDepth,
                           uint32_t* mt_key, int32_t* mt_pos, int32_t max_nodes, int32_t* node_level,
                           int32_t* node_parent, int32_t* node_kind, int64_t* node_row_start,
                           int64_t* node_row_count, int32_t* node_child_start, int32_t* node_child_count,
                           int64_t* node_knot_off, int64_t* knot_rows, int32_t* kinds_local, int64_t* perm,
                           int32_t* dfs_index, mra_build_job** job_out) {
  if (!job_out) return MRA_ERR_ARG;
  *job_out = nullptr;
  if (!locs || n_locs <= 0 || n_locs >= (int64_t(1) << 31) || r < 1 || !mt_key || !mt_pos || !node_level ||
      !node_parent || !node_kind || !node_row_start || !node_row_count || !node_child_start || !node_child_count ||
      !node_knot_off || !knot_rows || !kinds_local || !perm || !dfs_index)
    return MRA_ERR_ARG;
  if (r > 256 || M < 1 || M > 12 || n_locs < ((int64_t)1 << 16)) return MRA_BUILD_UNSUPPORTED;
  int64_t nn = 0;
  for (int L = 0; L <= M; ++L) nn += (int64_t)1 << (2 * L);
  if (nn > max_nodes) return MRA_ERR_NOMEM;
  mra_build_job* job = new (std::nothrow) mra_build_job();
  if (!job) return MRA_ERR_NOMEM;
  // slots of knots that are not final yet read as row 0 (a valid row id)
  std::memset(knot_rows, 0, sizeof(int64_t) * (size_t)(nn - ((int64_t)1 << (2 * M))) * r);
  StreamArgs a{locs, n_locs, r, M, J, critDepth, mt_key, mt_pos, node_level, node_parent, node_kind,
               node_row_start, node_row_count, node_child_start, node_child_count, node_knot_off, knot_rows,
               kinds_local, perm, dfs_index};
  try {
    job->th = std::thread(stream_body, job, a);
  } catch (...) {
    delete job;
    return MRA_ERR_NOMEM;
  }
  *job_out = job;
  return MRA_OK;
}

int mra_build_stream_wait(mra_build_job* job, int32_t event) {
  if (!job || event < 0) return MRA_ERR_ARG;
  std::unique_lock<std::mutex> lk(job->mu);
  if (event >= 5) {
    job->cv.wait(lk, [&] { return job->done; });
    return job->status;
  }
  job->cv.wait(lk, [&] { return job->fired > event || job->done; });
  if (job->fired > event) return MRA_OK;
  return job->status != MRA_OK ? job->status : MRA_ERR_STATE;
}

int mra_build_stream_finish(mra_build_job* job, int32_t* n_nodes_out, int32_t* depth_out, int64_t* n_knot_rows_out) {
  if (!job) return MRA_ERR_ARG;
  if (job->th.joinable()) job->th.join();
  const int status = job->status;
  if (n_nodes_out) *n_nodes_out = job->n_nodes;
  if (depth_out) *depth_out = job->depth;
  if (n_knot_rows_out) *n_knot_rows_out = job->n_knot_rows;
  delete job;
  return status;
}

}  // extern "C"
