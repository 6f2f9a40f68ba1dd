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
#include <cstdint>
#include <cstring>
#include <vector>

namespace {

struct MT {
  uint32_t key[624];
  int pos;
  void gen() {
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
  }
  inline uint32_t next32() {
    if (pos >= 624) gen();
    uint32_t y = key[pos++];
    y ^= (y >> 11);
    y ^= (y << 7) & 0x9d2c5680u;
    y ^= (y << 15) & 0xefc60000u;
    y ^= (y >> 18);
    return y;
  }
  inline uint32_t interval(uint32_t max) {   // uniform on [0, max], max <= 0xffffffff
    if (max == 0) return 0;
    uint32_t mask = max;
    mask |= mask >> 1;
    mask |= mask >> 2;
    mask |= mask >> 4;
    mask |= mask >> 8;
    mask |= mask >> 16;
    uint32_t v;
    while ((v = (next32() & mask)) > max) {
    }
    return v;
  }
};

struct Rec {
  int level, parent, kind;
  int64_t row_start, row_count;
  int first_child, n_child;
  int64_t knot_off;   // into knots_global
};

struct Builder {
  const double* locs;
  int64_t N;
  int r, J, critDepth;
  MT rng;
  // ping-pong level buffers
  std::vector<int32_t> rows[2];
  std::vector<double> xs[2], ys[2];
  std::vector<uint8_t> nk[2];
  std::vector<int32_t> scratch;       // permutation work array
  std::vector<uint8_t> code;
  std::vector<Rec> rec;
  std::vector<int64_t> knots_global;  // r per internal node, in reference knot order
  std::vector<int32_t> kinds_local;   // r per internal node
  std::vector<int32_t> perm;
  int status = 0;

  // node occupying [s, e) of buffer b
  void visit(int parent, int level, int64_t s, int64_t e, int b, int levels_left) {
    if (status) return;
    const int me = (int)rec.size();
    rec.push_back(Rec{level, parent, MRA_NODE_LEAF, s, e - s, -1, 0, -1});
    const int64_t n = e - s;
    const int32_t* R = rows[b].data() + s;
    const double* X = xs[b].data() + s;
    const double* Y = ys[b].data() + s;
    uint8_t* F = nk[b].data() + s;
    int64_t n_nk = 0;
    for (int64_t i = 0; i < n; ++i) n_nk += F[i];
    const bool internal = levels_left > 0 && n_nk > std::max(r, J);
    if (!internal) {
      std::memcpy(perm.data() + s, R, sizeof(int32_t) * n);
      return;
    }
    if (n_nk <= 100 || n <= 100) {   // KMeans knot / split paths: not handled natively
      status = 1;
      return;
    }
    // ---- knots: permutation(n_nk)[:r] mapped through the not-knot list, sorted
    int32_t* a = scratch.data();
    for (int64_t i = 0; i < n_nk; ++i) a[i] = (int32_t)i;
    {
      // Fisher-Yates from the top; the draws do not depend on the array, so they are generated
      // LA steps ahead (same order, same stream) and their targets prefetched.
      constexpr int LA = 64;
      uint32_t jq[LA];
      int64_t gen_i = n_nk - 1;
      for (int s0 = 0; s0 < LA && gen_i >= 1; ++s0, --gen_i) {
        jq[s0] = rng.interval((uint32_t)gen_i);
        __builtin_prefetch(&a[jq[s0]], 1);
      }
      int slot = 0;
      for (int64_t i = n_nk - 1; i >= 1; --i) {
        const uint32_t j = jq[slot];
        if (gen_i >= 1) {
          jq[slot] = rng.interval((uint32_t)gen_i);
          __builtin_prefetch(&a[jq[slot]], 1);
          --gen_i;
        }
        slot = (slot + 1 == LA) ? 0 : slot + 1;
        const int32_t t = a[i];
        a[i] = a[j];
        a[j] = t;
      }
    }
    std::vector<int32_t> pick(a, a + r);
    std::sort(pick.begin(), pick.end());
    // map candidate index -> local row id
    rec[me].kind = MRA_NODE_INTERNAL;
    rec[me].knot_off = (int64_t)knots_global.size();
    {
      int64_t c = 0;
      int p = 0;
      for (int64_t i = 0; i < n && p < r; ++i) {
        if (F[i]) {
          if (c == pick[p]) {
            kinds_local.push_back((int32_t)i);
            knots_global.push_back(R[i]);
            F[i] = 0;
            ++p;
          }
          ++c;
        }
      }
    }
    // ---- quadrant split by column means (sequential sums, as NumPy reduces axis 0 of an (n,2) array)
    double sx = 0.0, sy = 0.0;
    for (int64_t i = 0; i < n; ++i) {
      sx += X[i];
      sy += Y[i];
    }
    const double mx = sx / (double)n, my = sy / (double)n;
    int64_t cnt[4] = {0, 0, 0, 0};
    uint8_t* C = code.data() + s;
    for (int64_t i = 0; i < n; ++i) {
      uint8_t c = (uint8_t)((X[i] <= mx ? 0 : 2) + (Y[i] <= my ? 0 : 1));
      C[i] = c;
      ++cnt[c];
    }
    for (int c = 0; c < 4; ++c)
      if (cnt[c] == 0) {
        status = 1;   // the reference would build an empty child here; not handled natively
        return;
      }
    int64_t off[4];
    off[0] = s;
    for (int c = 1; c < 4; ++c) off[c] = off[c - 1] + cnt[c - 1];
    const int nb = b ^ 1;
    {
      int64_t w[4] = {off[0], off[1], off[2], off[3]};
      int32_t* R2 = rows[nb].data();
      double* X2 = xs[nb].data();
      double* Y2 = ys[nb].data();
      uint8_t* F2 = nk[nb].data();
      for (int64_t i = 0; i < n; ++i) {
        const int64_t d = w[C[i]]++;
        R2[d] = R[i];
        X2[d] = X[i];
        Y2[d] = Y[i];
        F2[d] = F[i];
      }
    }
    const bool fork = level == critDepth;
    MT saved;
    if (fork) saved = rng;
    rec[me].n_child = 4;
    for (int c = 0; c < 4; ++c) {
      if (fork) rng = saved;
      if (c == 0) rec[me].first_child = (int)rec.size();
      int child = (int)rec.size();
      (void)child;
      visit(me, level + 1, off[c], off[c] + cnt[c], nb, levels_left - 1);
      if (status) return;
    }
    if (fork) rng = saved;
  }
};

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
  Builder B;
  B.locs = locs;
  B.N = n_locs;
  B.r = r;
  B.J = J;
  B.critDepth = critDepth;
  std::memcpy(B.rng.key, mt_key, sizeof(uint32_t) * 624);
  B.rng.pos = *mt_pos;
  const int64_t N = n_locs;
  for (int b = 0; b < 2; ++b) {
    B.rows[b].resize(N);
    B.xs[b].resize(N);
    B.ys[b].resize(N);
    B.nk[b].resize(N);
  }
  B.scratch.resize(N);
  B.code.resize(N);
  B.perm.resize(N);
  for (int64_t i = 0; i < N; ++i) {
    B.rows[0][i] = (int32_t)i;
    B.xs[0][i] = locs[2 * i];
    B.ys[0][i] = locs[2 * i + 1];
    B.nk[0][i] = 1;
  }
  B.visit(-1, 0, 0, N, 0, M);
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
  std::vector<int32_t> inv(N);
  for (int64_t i = 0; i < N; ++i) inv[B.perm[i]] = (int32_t)i;
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
        knot_rows[koff + k] = inv[B.knots_global[rc.knot_off + k]];
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

}  // extern "C"
