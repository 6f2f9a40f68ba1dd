// Host runtime and C ABI of pymra_b200 (see include/pymra_b200.h).
//
// The handle owns no device memory of its own: every device buffer is carved out of the arena
// the caller binds with mra_bind_workspace (a torch tensor in the Python host).
#include "../../include/pymra_b200.h"

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <future>
#include <map>
#include <memory>
#include <mutex>
#include <utility>
#include <string>
#include <thread>
#include <vector>

#include "mra_kernels.cuh"

using namespace mra;

namespace {

// vector storage that is filled right after resize(): skips the zero-fill of the O(N) host tables
template <class T>
struct NoInit : std::allocator<T> {
  template <class U>
  struct rebind { using other = NoInit<U>; };
  template <class U>
  void construct(U* p) noexcept { ::new (static_cast<void*>(p)) U; }
  template <class U, class... Args>
  void construct(U* p, Args&&... args) { ::new (static_cast<void*>(p)) U(std::forward<Args>(args)...); }
};
using IntBuf = std::vector<int, NoInit<int>>;

struct Arena {
  size_t off = 0;
  size_t take(size_t bytes) {
    size_t o = off;
    off += (bytes + 255) & ~size_t(255);
    return o;
  }
};

struct Layout {
  size_t nodes, knot_rows, obs_rows, unobs_rows, perm, xs, ys, yobs, V, S, DI, UT, QT, A, GT, LPINV, VK, LINV, dnode,
      mean, var, vnorm, status, out, stage_locs, stage_obs, out_mean, out_sd, lists, ptiles, pgroups, gather, chunks, ltiles, GTF, UTF, fold, VKL, xidx, LS, UTTN, params;
  size_t total;
};

}  // namespace

struct mra_handle {
  int device = 0;
  std::string err;
  // build_lists runs as a background job from mra_set_structure / mra_set_shard on (it depends on the tree only);
  // mra_plan joins it after its own O(N) passes over the observations
  std::future<void> lists_job;
  bool defer_lists = false;      // mra_expect_shard: mra_set_shard will follow, do not build the unsharded lists first
  bool lists_built = false;
  bool tree_planned = false, tree_bound = false, split_arena = false;      // two-part plan (mra_plan_tree / mra_plan_obs)
  size_t total_A = 0, total_B = 0;
  Layout lay_b{};                                                          // part B, offsets relative to its start
  bool has_structure = false, planned = false, bound = false, uploaded = false, lik_done = false,
       pred_done = false, want_predict = false;
  // structure
  int64_t N = 0;
  int dim = 0, r = 0, depth = 0, n_nodes = 0;
  std::vector<int> level, parent, kind, child_start, child_count, level_off;
  std::vector<int64_t> row_start, row_count, knot_off;
  std::vector<int> knot_rows;
  IntBuf perm;
  // derived lists
  std::vector<std::vector<int>> internal_at;   // node ids per level
  std::vector<int> leaves;                     // node ids of leaves + orphans
  std::vector<std::vector<int4>> ptiles_at;    // prior pass: 64-row tiles per level (+ gathered knot tiles when sharded)
  std::vector<int> gather_rows;                // row ids of the gathered tiles
  std::vector<int2> gather_node;               // (top node, first slot in gather_rows): r slots each
  std::vector<int> n_regular_tiles;            // per level: tiles of ptiles_at that are not gathered tiles
  std::vector<std::vector<int4>> pgroups_at;   // per level: runs of <= PG consecutive regular tiles of a node (k_prior_groups)
  std::vector<std::vector<int>> group_of_tile; // per level: regular tile -> its group
  std::vector<size_t> pgroups_off;             // per level offsets (bytes) inside lay.pgroups
  bool use_groups = true;
  bool chol_mma = true;
  bool keep_t0 = false;
  bool leaf_v2 = true;
  bool leaf_wide = true;
  DevParams hparams{};
  // CUDA graph of a whole evaluation on a frozen tree (mra_run_graph)
  bool capturing = false;
  cudaGraph_t graph = nullptr;
  cudaGraphExec_t gexec = nullptr;
  int graph_predict = -1;
  int64_t graph_launches = 0;
  cudaStream_t gstream = nullptr;
  std::vector<int4> leaf_tiles;                // 64-row tiles of leaves / orphans (fused predict pass)
  std::vector<int4> fold_items;                // (node, ancestor level, row tile, column tile) of k_fold
  std::vector<int2> emit_chunks;               // row ranges whose results this rank emits
  // subtree sharding (mra_set_shard): role per node, 0 = another rank's, 1 = mine, 2 = replicated top,
  // 3 = replicated top whose rows this rank emits
  int shard_level = 0;
  std::vector<int8_t> role;
  std::vector<int> sroots;                     // my subtree roots at the shard level
  // streamed evaluation (mra_stream_*): the subtrees of the root's children ("parts") as ranges of the work lists
  struct Range { int begin, count; };
  int n_parts = 0;
  std::vector<std::vector<Range>> part_nodes, part_tiles;   // [level][part] into internal_at / ptiles_at
  std::vector<std::vector<Range>> part_leaves;              // [part] -> ranges of `leaves`
  std::vector<std::vector<Range>> part_knots;               // [part] -> ranges of knot_rows
  std::vector<Range> part_gtiles;                           // [part] -> gathered tiles of its level-1 node in ptiles_at[0]
  std::vector<int2> part_gather;                            // [part] -> (level-1 top node, first slot in gather_rows) or (-1, 0)
  int stream_parts_done = 0;                                // bit mask of the parts run since mra_stream_begin_async
  int stream_prior_done = 0;                                // parts whose prior levels have run (mra_stream_part_prior_async)
  int my_parts = 0;                                         // bit mask of the parts this rank evaluates (all when unsharded)
  bool stream_open = false, leafq_done = false;
  cudaStream_t copy_stream = nullptr;
  cudaEvent_t copy_event = nullptr;
  int slot_base = 0, n_slots = 0;
  size_t sroots_off = 0;
  std::vector<NodeDev> nodes;
  IntBuf obs_rows, unobs_rows;
  int max_leaf_obs = 0, max_leaf_rows = 0, max_leaf_W = 1, max_leaf_unobs = 0;
  int64_t n_obs_total = 0;
  long long ldv = 0;
  // device
  Layout lay{};
  char* ws = nullptr;
  size_t ws_bytes = 0;
  std::vector<size_t> list_off, ptiles_off;   // per level offsets (bytes) inside lay.lists / lay.ptiles
  size_t leaves_off = 0;
  CovParams cov{0, 1.0, 1.0, 1.0, 1.0, nullptr, 0};
  double R = 1.0;
  bool cov_set = false, R_set = false;
  int64_t launches = 0;
  int warnings = 0;                            // MRA_WARN_* bits seen since the last likelihood pass started
  double flops_lik = 0, flops_pred = 0;
  // per-kernel profiling (CUDA events on the launching stream)
  bool profiling = false;
  struct ProfRec { int kid; cudaEvent_t e0, e1; };
  std::vector<ProfRec> prof_recs;
  std::vector<cudaEvent_t> event_pool;
  std::map<std::string, int> kid_of;
  std::vector<std::string> kname;
  std::vector<double> kflops, kbytes;   // algorithmic work per launch-group, accumulated at plan time
  std::vector<double> kms;
  std::vector<int64_t> klaunch;
};

namespace {

// Splits [0, n) over a few host threads (the host-side O(N) passes of plan / set_structure).
template <class F>
void parallel_for(int64_t n, F fn) {
  unsigned nt = std::min(16u, std::max(1u, std::thread::hardware_concurrency()));
  if (const char* e = std::getenv("MRA_HOST_THREADS")) nt = (unsigned)std::max(1, std::min((int)nt, std::atoi(e)));
  if (n < (int64_t)1 << 18 || nt == 1) {
    fn((int64_t)0, n);
    return;
  }
  std::vector<std::thread> th;
  const int64_t step = (n + nt - 1) / nt;
  for (unsigned t = 0; t < nt; ++t) {
    const int64_t a = (int64_t)t * step, b = std::min(n, a + step);
    if (a < b) th.emplace_back([=] { fn(a, b); });
  }
  for (auto& t : th) t.join();
}

// MRA_HOST_TRACE=1: wall-clock phases of the host-side entry points on stderr (where the pre-GPU time of a construction goes)
struct HostTrace {
  const char* fn;
  bool on;
  std::chrono::steady_clock::time_point t0, last;
  explicit HostTrace(const char* f) : fn(f), on(std::getenv("MRA_HOST_TRACE") != nullptr) {
    t0 = last = std::chrono::steady_clock::now();
  }
  void mark(const char* what) {
    if (!on) return;
    auto now = std::chrono::steady_clock::now();
    fprintf(stderr, "[host] %s: %-22s %7.3f ms\n", fn, what, std::chrono::duration<double, std::milli>(now - last).count());
    last = now;
  }
  ~HostTrace() {
    if (on)
      fprintf(stderr, "[host] %s: total %7.3f ms\n", fn,
              std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count());
  }
};

// The large host-side vectors of a handle (work lists, row lists, permutation, node table: ~50 MB at 4 M locations) are
// recycled across handles: a construction that follows another one in the same process (an MLE loop with the reference's
// semantics builds a tree per evaluation) finds them allocated and page-faulted.  One set is kept.
struct HostBuffers {
  std::vector<std::vector<int4>> ptiles_at, pgroups_at;
  std::vector<std::vector<int>> group_of_tile, internal_at;
  std::vector<int4> leaf_tiles, fold_items;
  std::vector<int> knot_rows, leaves;
  IntBuf perm, obs_rows, unobs_rows;
  std::vector<NodeDev> nodes;
};
std::mutex g_pool_mu;
bool env_on(const char* name) {
  const char* e = std::getenv(name);
  return e && e[0] && e[0] != '0';
}
std::unique_ptr<HostBuffers> g_pool;

template <class H, class B>
void swap_buffers(H* h, B& b) {
  h->ptiles_at.swap(b.ptiles_at);
  h->pgroups_at.swap(b.pgroups_at);
  h->group_of_tile.swap(b.group_of_tile);
  h->internal_at.swap(b.internal_at);
  h->leaf_tiles.swap(b.leaf_tiles);
  h->fold_items.swap(b.fold_items);
  h->knot_rows.swap(b.knot_rows);
  h->leaves.swap(b.leaves);
  h->perm.swap(b.perm);
  h->obs_rows.swap(b.obs_rows);
  h->unobs_rows.swap(b.unobs_rows);
  h->nodes.swap(b.nodes);
}

// outer.resize(n) with every inner vector emptied but its capacity kept
template <class T>
void reset_levels(std::vector<std::vector<T>>& v, size_t n) {
  v.resize(n);
  for (auto& x : v) x.clear();
}

int fail(mra_handle* h, int code, const std::string& msg) {
  if (h) h->err = msg;
  return code;
}

void drop_graph(mra_handle* h) {
  if (h->gexec) cudaGraphExecDestroy(h->gexec);
  if (h->graph) cudaGraphDestroy(h->graph);
  h->gexec = nullptr;
  h->graph = nullptr;
  h->graph_predict = -1;
}

#define CU(call)                                                                                 \
  do {                                                                                           \
    cudaError_t e_ = (call);                                                                     \
    if (e_ != cudaSuccess)                                                                       \
      return fail(h, MRA_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_));          \
  } while (0)

// Every entry point that touches the device selects the handle's device for its own duration and restores the
// caller's current device on the way out (a handle on device k must not change PyTorch's current device).
struct DevGuard {
  int prev = -1;
  bool changed = false;
  cudaError_t err = cudaSuccess;
  explicit DevGuard(int dev) {
    err = cudaGetDevice(&prev);
    if (err == cudaSuccess && prev != dev) {
      err = cudaSetDevice(dev);
      changed = err == cudaSuccess;
    }
  }
  ~DevGuard() {
    if (changed) cudaSetDevice(prev);
  }
};
#define DEVICE_SCOPE(h)          \
  DevGuard dev_guard_((h)->device); \
  CU(dev_guard_.err)

template <class T>
T* at(mra_handle* h, size_t off) {
  return reinterpret_cast<T*>(h->ws + off);
}

DevCtx make_ctx(mra_handle* h) {
  DevCtx c{};
  const Layout& L = h->lay;
  c.nodes = at<NodeDev>(h, L.nodes);
  c.knot_rows = at<int>(h, L.knot_rows);
  c.obs_rows = at<int>(h, L.obs_rows);
  c.unobs_rows = at<int>(h, L.unobs_rows);
  c.fill_qt = (h->want_predict && !h->leaf_v2) ? 1 : 0;
  c.gather_rows = at<int>(h, L.gather);
  // dense covariance matrix: the kernels' "coordinates" are the locations' original row indices
  c.xs = h->cov.family == MRA_COV_DENSE ? at<double>(h, L.xidx) : at<double>(h, L.xs);
  c.ys = h->cov.family == MRA_COV_DENSE ? at<double>(h, L.xidx) : at<double>(h, L.ys);
  c.yobs = at<double>(h, L.yobs);
  c.V = at<double>(h, L.V);
  c.ldv = h->ldv;
  c.r = h->r;
  c.N = (int)h->N;
  c.S = at<double>(h, L.S);
  c.DI = at<double>(h, L.DI);
  c.UT = at<double>(h, L.UT);
  c.QT = at<double>(h, L.QT);
  c.A = at<double>(h, L.A);
  c.GT = at<double>(h, L.GT);
  c.GTF = at<double>(h, L.GTF);
  c.UTF = at<double>(h, L.UTF);
  c.LPINV = at<double>(h, L.LPINV);
  c.VK = at<double>(h, L.VK);
  c.VKL = at<double>(h, L.VKL);
  c.LINV = at<double>(h, L.LINV);
  c.dnode = at<double>(h, L.dnode);
  c.mean = at<double>(h, L.mean);
  c.var = at<double>(h, L.var);
  c.vnorm = at<double>(h, L.vnorm);
  c.status = at<int>(h, L.status);
  c.P = at<DevParams>(h, L.params);
  c.chol_mma = h->chol_mma ? 1 : 0;
  {
    static const int tune = [] { const char* e = std::getenv("MRA_TUNE"); return e ? std::atoi(e) : 0; }();
    c.tune = tune;
  }
  c.keep_t0 = h->keep_t0 ? 1 : 0;
  c.leaf_v2 = h->leaf_v2 ? 1 : 0;
  c.LS = at<double>(h, L.LS);
  c.UTTN = at<double>(h, L.UTTN);
  return c;
}

constexpr size_t GS = sizeof(GemmSmem);     // kernels with segmented products
constexpr size_t GS1 = sizeof(GemmSmem1);   // single-segment kernels
size_t smem_knot(int r) { return GS1 + sizeof(double) * ((size_t)2 * r) + sizeof(int) * r + 16; }   // k_knot_gram
size_t smem_cholinv(int n) {     // the larger of the two chol_inv_block variants
  return sizeof(double) * std::max((size_t)n * (n + 1) + n + NT * 9 + 8, chol_mma_smem_doubles(n));
}
size_t smem_prior(int r) { return sizeof(GemmSmemT<2>) + sizeof(double) * ((size_t)2 * r + 2 * TB) + sizeof(int) * TB; }
// MRA_SMEM_PAD_PRIOR / MRA_SMEM_PAD_PREDICT (bytes): extra dynamic shared memory, an A/B knob that lowers the number of
// co-resident CTAs of the two heaviest kernels without touching their code
size_t env_pad(const char* name) {
  const char* e = std::getenv(name);
  return e ? (size_t)std::max(0, std::atoi(e)) : 0;
}
size_t smem_pgroups(int r) {
  static const size_t pad = env_pad("MRA_SMEM_PAD_PRIOR");
  return sizeof(PriorSmem) + sizeof(double) * ((size_t)2 * r + 2 * PG * TB) + pad;
}
size_t smem_ut2(int max_obs) { return GS1 + sizeof(double*) * (size_t)((max_obs + TB - 1) / TB * TB); }
size_t smem_leafq(int max_obs) {
  return sizeof(GemmSmemT<2>) + sizeof(double) * ((size_t)3 * max_obs + 2 * TB) + sizeof(int) * TB + 16;
}
size_t smem_leafq2(int max_obs) {
  return sizeof(WideSmem) + sizeof(double) * ((size_t)3 * max_obs + 2 * TB + 4 * TB) + sizeof(int) * TB + 16;
}
size_t smem_gram() { return GS1 + sizeof(double) * 4 * TB + sizeof(int) * 2 * TB; }
size_t smem_solve() { return GS1; }
size_t smem_plain() { return sizeof(GemmSmemT<4>); }   // k_assemble_A: up to 4 children per product
size_t smem_predict(int r, int depth) {
  int ldT = ((r + 15) / 16) * 16 + 4;
  static const size_t pad = env_pad("MRA_SMEM_PAD_PREDICT");
  return GS + sizeof(double) * ((size_t)(r > TB ? TB * ldT : 0) + 2 * TB + (size_t)std::max(depth, 1) * r) +
         sizeof(int) * MAX_LEVELS + sizeof(long long) * 2 * MAX_LEVELS + pad;
}

size_t smem_predict2(int r, int depth, int tile_rows) {
  static const size_t pad = env_pad("MRA_SMEM_PAD_PREDICT");
  return pad + sizeof(double) * NSTAGE * (tile_rows + TB) * KC + sizeof(double) * ((size_t)2 * tile_rows + (size_t)std::max(depth, 1) * r) +
         sizeof(long long) * 2 * MAX_LEVELS;
}

// Kernels are instantiated for VEC = 2 (16-byte cp.async, even r) and VEC = 1 (odd r).
#define MRA_FOR_VEC(h, expr)         \
  do {                               \
    if (((h)->r & 1) == 0) {         \
      constexpr int V_ = 2;          \
      expr;                          \
    } else {                         \
      constexpr int V_ = 1;          \
      expr;                          \
    }                                \
  } while (0)

// cudaFuncAttributeMaxDynamicSharedMemorySize is per kernel and device, not per handle: several handles with different r
// or leaf sizes live in one process (sharded emulation, tests, MLE over several trees), so the limit only ever grows.
template <class K>
cudaError_t smem_at_least(K kernel, size_t bytes) {
  static std::mutex mu;
  static std::map<std::pair<int, const void*>, size_t> seen;
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  std::lock_guard<std::mutex> lock(mu);
  size_t& cur = seen[std::make_pair(dev, reinterpret_cast<const void*>(kernel))];
  if (bytes <= cur) return cudaSuccess;
  e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
  if (e == cudaSuccess) cur = bytes;
  return e;
}

// Kernels whose tile columns are the r knots of a node are also instantiated for narrow tiles (NJ 8-column groups,
// mra_gemm.cuh): r0 = 16 / 32 (BASELINE cfg4 / cfg3) run 64 x 16 / 64 x 32 tiles instead of zero-padded 64 x 64 ones.
#define MRA_FOR_VEC_NJ(h, expr)                            \
  do {                                                     \
    if (((h)->r & 1) != 0) {                               \
      constexpr int V_ = 1, J_ = 8;                        \
      expr;                                                \
    } else if ((h)->r <= 16) {                             \
      constexpr int V_ = 2, J_ = 2;                        \
      expr;                                                \
    } else if ((h)->r <= 32) {                             \
      constexpr int V_ = 2, J_ = 4;                        \
      expr;                                                \
    } else if ((h)->r <= 48) {                             \
      constexpr int V_ = 2, J_ = 6;                        \
      expr;                                                \
    } else {                                               \
      constexpr int V_ = 2, J_ = 8;                        \
      expr;                                                \
    }                                                      \
  } while (0)

template <int V_, int J_>
cudaError_t configure_vec_nj(int r, int depth) {
  cudaError_t e;
#define SET_(k, bytes)              \
  e = smem_at_least(k, (bytes));    \
  if (e != cudaSuccess) return e
  SET_((k_knot_gram<V_, J_>), smem_knot(r));
  SET_((k_prior_tiles<V_, J_>), smem_prior(r));
  if (V_ == 2) {
    SET_((k_prior_groups<J_>), smem_pgroups(r));
    SET_((k_prior_groups<J_, true>), smem_pgroups(r));
  }
  SET_((k_node_gt<V_, J_>), GS1);
  SET_((k_predict_fused<V_, J_>), smem_predict(r, depth));
#undef SET_
  return cudaSuccess;
}

template <int V_>
cudaError_t configure_vec(int r, int max_obs) {
  cudaError_t e;
#define SET_(k, bytes)              \
  e = smem_at_least(k, (bytes));    \
  if (e != cudaSuccess) return e
  SET_(k_knot_vkl<V_>, GS1);
  SET_(k_leaf_gram<V_>, smem_gram());
  SET_(k_leaf_upd<V_>, GS1);
  SET_(k_leaf_linv<V_>, GS1);
  SET_(k_leaf_ut2<V_>, smem_ut2(max_obs));
  SET_(k_leaf_q<V_>, smem_leafq(max_obs));
  SET_(k_leaf_trsm<V_>, GS1);
  SET_(k_leaf_solve_ut<V_>, smem_solve());
  SET_(k_leaf_solve_qt<V_>, smem_solve());
  SET_(k_assemble_A<V_>, smem_plain());

  SET_(k_fold<V_>, GS1 + sizeof(long long) * MAX_LEVELS);
#undef SET_
  return cudaSuccess;
}

int configure_kernels(mra_handle* h) {
  MRA_FOR_VEC(h, CU(configure_vec<V_>(h->r, h->max_leaf_obs)));
  MRA_FOR_VEC_NJ(h, CU((configure_vec_nj<V_, J_>(h->r, h->depth))));
  CU(smem_at_least(k_knot_chol, smem_cholinv(h->r)));
  CU(smem_at_least(k_node_chol, smem_cholinv(h->r)));
  CU(smem_at_least(k_leaf_chol, smem_cholinv(TB)));
  CU(smem_at_least(k_leaf_q2, smem_leafq2(h->max_leaf_obs)));
  return MRA_OK;
}

int kid(mra_handle* h, const std::string& name) {
  auto it = h->kid_of.find(name);
  if (it != h->kid_of.end()) return it->second;
  int id = (int)h->kname.size();
  h->kid_of[name] = id;
  h->kname.push_back(name);
  h->kflops.push_back(0.0);
  h->kbytes.push_back(0.0);
  h->kms.push_back(0.0);
  h->klaunch.push_back(0);
  return id;
}

inline void add_w(mra_handle* h, int id, double flops, double bytes) {
  h->kflops[id] += flops;
  h->kbytes[id] += bytes;
}

cudaEvent_t get_event(mra_handle* h) {
  if (!h->event_pool.empty()) {
    cudaEvent_t e = h->event_pool.back();
    h->event_pool.pop_back();
    return e;
  }
  cudaEvent_t e;
  cudaEventCreate(&e);
  return e;
}

struct ProfScope {
  mra_handle* h;
  cudaStream_t st;
  mra_handle::ProfRec rec;
  bool on;
  ProfScope(mra_handle* h_, cudaStream_t st_, const char* name) : h(h_), st(st_), on(h_->profiling) {
    ++h->launches;
    if (!on) return;
    rec.kid = kid(h, name);
    rec.e0 = get_event(h);
    rec.e1 = get_event(h);
    cudaEventRecord(rec.e0, st);
  }
  ~ProfScope() {
    if (!on) return;
    cudaEventRecord(rec.e1, st);
    h->prof_recs.push_back(rec);
  }
};
#define LAUNCH(name, ...)            \
  do {                               \
    ProfScope ps_(h, st, name);      \
    __VA_ARGS__;                     \
  } while (0)

int check_status(mra_handle* h, cudaStream_t st) {
  int flag = 0;
  CU(cudaMemcpyAsync(&flag, at<int>(h, h->lay.status), sizeof(int), cudaMemcpyDeviceToHost, st));
  CU(cudaStreamSynchronize(st));
  h->warnings |= flag & ~1;
  if (flag & 1)
    return fail(h, MRA_ERR_NOT_SPD, "a Cholesky factorisation met a non-positive pivot (covariance not positive definite)");
  return MRA_OK;
}

using Range = mra_handle::Range;

// elimination of the own level of `nn` nodes of level m: Cholesky factor + inverse (latency kernel), then GT (tiles)
void node_factor(mra_handle* h, cudaStream_t st, const DevCtx& c, int m, const int* list, int nn) {
  const int r = h->r;
  const int nct = (r + TB - 1) / TB, nw = (m * r + 1 + TB - 1) / TB;
  LAUNCH("node_chol", k_node_chol<<<nn, NT, smem_cholinv(r), st>>>(c, list));
  MRA_FOR_VEC_NJ(h, LAUNCH("node_gt", k_node_gt<V_, J_><<<(unsigned)nn * nw * nct, NT, GS1, st>>>(c, list, nw * nct, nct)));
}

// assemble_A + node_factor for nodes [rg.begin, rg.begin + rg.count) of level m's list
int upward_level(mra_handle* h, cudaStream_t st, const DevCtx& c, int m, Range rg) {
  const Layout& L = h->lay;
  const int r = h->r;
  const int nn = rg.count;
  if (nn <= 0) return MRA_OK;
  const int* list = reinterpret_cast<const int*>(h->ws + L.lists + h->list_off[m]) + rg.begin;
  const int W = (m + 1) * r + 1, nb = (W - 1 + TB - 1) / TB;
  const int nt = nb * (nb + 1) / 2 + (W + TB - 1) / TB;   // lower tile pairs of the basis block + augmented-row jobs
  MRA_FOR_VEC(h, LAUNCH("assemble_A", k_assemble_A<V_><<<(unsigned)((nn + 15) / 16 * 16) * nt, NT, smem_plain(), st>>>(c, list, nullptr, 0, nn, nt)));
  node_factor(h, st, c, m, list, nn);
  return MRA_OK;
}

int upward_level(mra_handle* h, cudaStream_t st, const DevCtx& c, int m) {
  return upward_level(h, st, c, m, Range{0, (int)h->internal_at[m].size()});
}

// knot_factor + prior_tiles for a range of level m's node list and the matching range of its tile list
int prior_level(mra_handle* h, cudaStream_t st, const DevCtx& c, int m, Range nodes, Range tiles) {
  const Layout& L = h->lay;
  const int r = h->r;
  if (nodes.count <= 0) return MRA_OK;
  const int* list = reinterpret_cast<const int*>(h->ws + L.lists + h->list_off[m]) + nodes.begin;
  {
    const int nt = (r + TB - 1) / TB, npair = nt * (nt + 1) / 2, nkt = (m * r + TB - 1) / TB;
    MRA_FOR_VEC_NJ(h, LAUNCH("knot_gram", k_knot_gram<V_, J_><<<(unsigned)nodes.count * npair, NT, smem_knot(r), st>>>(c, list, npair)));
    LAUNCH("knot_chol", k_knot_chol<<<nodes.count, NT, smem_cholinv(r), st>>>(c, list));
    if (nkt > 0)
      MRA_FOR_VEC(h, LAUNCH("knot_vkl", k_knot_vkl<V_><<<(unsigned)nodes.count * nt * nkt, NT, GS1, st>>>(c, list, nt * nkt, nkt)));
  }
  if (tiles.count <= 0) return MRA_OK;
  // regular tiles run grouped (one chunk stream per <= PG tiles of a node), gathered tiles one CTA each
  const int nreg = h->n_regular_tiles[m];
  int t0 = tiles.begin, t1 = tiles.begin + tiles.count;
  if (h->use_groups && (r & 1) == 0 && t0 < nreg) {
    const int te = std::min(t1, nreg);
    const int g0 = h->group_of_tile[m][t0], g1 = h->group_of_tile[m][te - 1] + 1;
    const int4* gl = reinterpret_cast<const int4*>(h->ws + L.pgroups + h->pgroups_off[m]) + g0;
    // r a multiple of 16 (one column tile): the level's covariance block is evaluated ahead of the product, into the
    // columns of V the product overwrites (k_cov_fill); MRA_TUNE bit 7 keeps the evaluation inside the product (A/B)
    if (r % KC == 0 && r <= TB && !(c.tune & 128)) {
      {
        // the family is a launch-time constant except inside a captured graph (refit may change it between replays)
        const int fam = h->capturing ? -1 : h->cov.family;
        const size_t csm = sizeof(double) * 2 * r;
        ProfScope ps_(h, st, "prior_tiles");
        if (fam == 0) k_cov_fill<0><<<g1 - g0, 256, csm, st>>>(c, gl, m);
        else if (fam == 1) k_cov_fill<1><<<g1 - g0, 256, csm, st>>>(c, gl, m);
        else if (fam == 2) k_cov_fill<2><<<g1 - g0, 256, csm, st>>>(c, gl, m);
        else if (fam == 3) k_cov_fill<3><<<g1 - g0, 256, csm, st>>>(c, gl, m);
        else k_cov_fill<-1><<<g1 - g0, 256, csm, st>>>(c, gl, m);
      }
      MRA_FOR_VEC_NJ(h, LAUNCH("prior_tiles", (k_prior_groups<J_, true><<<g1 - g0, NT, smem_pgroups(r), st>>>(c, gl, m))));
    } else {
      MRA_FOR_VEC_NJ(h, LAUNCH("prior_tiles", k_prior_groups<J_><<<g1 - g0, NT, smem_pgroups(r), st>>>(c, gl, m)));
    }
    t0 = te;
  }
  if (t0 < t1) {
    const int4* tl = reinterpret_cast<const int4*>(h->ws + L.ptiles + h->ptiles_off[m]) + t0;
    MRA_FOR_VEC_NJ(h, LAUNCH("prior_tiles", k_prior_tiles<V_, J_><<<t1 - t0, NT, smem_prior(r), st>>>(c, tl, m)));
  }
  return MRA_OK;
}

// leaf terms of the likelihood (Gram of the observed rows, its Cholesky factor, UT) for a range of `leaves`
int leaf_terms(mra_handle* h, cudaStream_t st, const DevCtx& c, Range rg) {
  const Layout& L = h->lay;
  const int nleaf = rg.count;
  if (nleaf <= 0 || h->max_leaf_obs <= 0) return MRA_OK;
  const int* leaf_list = reinterpret_cast<const int*>(h->ws + L.lists + h->leaves_off) + rg.begin;
  const int nbo = (h->max_leaf_obs + TB - 1) / TB;
  const int nt1 = nbo * (nbo + 1) / 2;
  MRA_FOR_VEC(h, LAUNCH("leaf_gram", k_leaf_gram<V_><<<(unsigned)nleaf * nt1, NT, smem_gram(), st>>>(c, leaf_list, 0, nleaf)));
  for (int p = 0; p < nbo; ++p) {        // left-looking over the 64-wide block columns of S
    if (p > 0) MRA_FOR_VEC(h, LAUNCH("leaf_upd", k_leaf_upd<V_><<<nleaf, NT, GS1, st>>>(c, leaf_list, p)));
    LAUNCH("leaf_chol", k_leaf_chol<<<nleaf, NT, smem_cholinv(TB), st>>>(c, leaf_list, p));
    if (p + 1 < nbo)
      MRA_FOR_VEC(h, LAUNCH("leaf_trsm", k_leaf_trsm<V_><<<(unsigned)nleaf * (nbo - 1 - p), NT, GS1, st>>>(c, leaf_list, p, nbo - 1 - p)));
  }
  const int nt3 = std::max(1, (h->max_leaf_W - 1 + TB - 1) / TB);
  if (h->leaf_v2) {
    MRA_FOR_VEC(h, LAUNCH("leaf_linv", k_leaf_linv<V_><<<nleaf, NT, GS1, st>>>(c, leaf_list)));
    if (h->max_leaf_W > 1)
      MRA_FOR_VEC(h, LAUNCH("leaf_ut", k_leaf_ut2<V_><<<(unsigned)nleaf * nbo, NT, smem_ut2(h->max_leaf_obs), st>>>(
                                           c, leaf_list, nbo, nt3)));
  } else {
    MRA_FOR_VEC(h, LAUNCH("leaf_solve", k_leaf_solve_ut<V_><<<(unsigned)nleaf * nt3, NT, smem_solve(), st>>>(c, leaf_list, nt3)));
  }
  return MRA_OK;
}

// leaf part of the predict pass (QT and the leaf moments) for a range of `leaves`; needs leaf_terms only
int leaf_predict_terms(mra_handle* h, cudaStream_t st, const DevCtx& c, Range rg) {
  const Layout& L = h->lay;
  const int nleaf = rg.count;
  if (nleaf <= 0 || h->max_leaf_obs <= 0) return MRA_OK;
  const int* leaf_list = reinterpret_cast<const int*>(h->ws + L.lists + h->leaves_off) + rg.begin;
  const int nbo = (h->max_leaf_obs + TB - 1) / TB, nbr = (h->max_leaf_rows + TB - 1) / TB;
  const int nbu = (h->max_leaf_unobs + TB - 1) / TB;
  if (h->leaf_v2) {
    if (nbu > 0) {
      if (h->leaf_wide)
      {
        // leaves with at most 128 observations (one super tile): the covariance block is evaluated ahead of the product, into
        // the rows of QT the product overwrites (MRA_TUNE bit 8: always inside the product)
        const int pre = (c.tune & 256) ? 0 : 1;
        if (pre)
          LAUNCH("leaf_q", k_leaf_cov_fill<<<(unsigned)nleaf * nbu, 256, sizeof(double) * 2 * h->max_leaf_obs, st>>>(c, leaf_list, nbu, h->max_leaf_obs));
        LAUNCH("leaf_q", k_leaf_q2<<<(unsigned)nleaf * nbu, NTW, smem_leafq2(h->max_leaf_obs), st>>>(c, leaf_list, nbu, h->max_leaf_obs, pre));
      }
      else
        MRA_FOR_VEC(h, LAUNCH("leaf_q", k_leaf_q<V_><<<(unsigned)nleaf * nbu, NT, smem_leafq(h->max_leaf_obs), st>>>(
                                            c, leaf_list, nbu, h->max_leaf_obs)));
    }
    LAUNCH("leaf_qobs", k_leaf_qobs<<<nleaf, NT, 0, st>>>(c, leaf_list));
    return MRA_OK;
  }
  if (nbu > 0)
    MRA_FOR_VEC(h, LAUNCH("leaf_gram_T", k_leaf_gram<V_><<<(unsigned)nleaf * nbu * nbo, NT, smem_gram(), st>>>(
                                             c, leaf_list, 1, nleaf)));
  MRA_FOR_VEC(h, LAUNCH("leaf_solve_Q", k_leaf_solve_qt<V_><<<(unsigned)nleaf * nbr, NT, smem_solve(), st>>>(
                                            c, leaf_list, nbr)));
  return MRA_OK;
}

// the parameter block of this pass (covariance descriptor, nugget) goes to the device ahead of its kernels
int upload_params(mra_handle* h, cudaStream_t st) {
  h->hparams.cov = h->cov;
  h->hparams.R = h->R;
  CU(cudaMemcpyAsync(h->ws + h->lay.params, &h->hparams, sizeof(DevParams), cudaMemcpyHostToDevice, st));
  return MRA_OK;
}

int reset_pass(mra_handle* h, cudaStream_t st) {
  const Layout& L = h->lay;
  if (!h->capturing) {
    int rc = upload_params(h, st);
    if (rc) return rc;
  }
  h->launches = 0;
  h->warnings = 0;
  h->leafq_done = false;
  CU(cudaMemsetAsync(at<double>(h, L.dnode), 0, sizeof(double) * h->n_nodes, st));
  CU(cudaMemsetAsync(at<int>(h, L.status), 0, sizeof(int), st));
  CU(cudaMemsetAsync(at<double>(h, L.vnorm), 0, sizeof(double) * h->N, st));
  return MRA_OK;
}

// sharded: the summaries (A~_c, d_c) of this rank's subtree roots into its slots of dev_summary
int export_summaries(mra_handle* h, cudaStream_t st, const DevCtx& c, double* dev_summary) {
  if (h->shard_level <= 0 || h->sroots.empty()) return MRA_OK;
  const Layout& L = h->lay;
  const int r = h->r;
  const int* list = reinterpret_cast<const int*>(h->ws + L.lists + h->sroots_off);
  const int W = h->shard_level * r + 1, nb = (W - 1 + TB - 1) / TB;
  const int nt = nb * (nb + 1) / 2 + (W + TB - 1) / TB;
  MRA_FOR_VEC(h, LAUNCH("export_summary", k_assemble_A<V_><<<(unsigned)((h->sroots.size() + 15) / 16 * 16) * nt, NT, smem_plain(), st>>>(
                                              c, list, dev_summary, h->slot_base, (int)h->sroots.size(), nt)));
  return MRA_OK;
}

// Prior pass, leaf terms and the upward pass down to the shard level (level 0 when not sharded).
// Sharded: ends by exporting the summaries of this rank's subtree roots into dev_summary.
int launch_likelihood_local(mra_handle* h, cudaStream_t st, double* dev_summary) {
  DevCtx c = make_ctx(h);
  int rc = reset_pass(h, st);
  if (rc) return rc;
  // ---- prior, top-down
  for (int m = 0; m < (int)h->internal_at.size(); ++m) {
    rc = prior_level(h, st, c, m, Range{0, (int)h->internal_at[m].size()}, Range{0, (int)h->ptiles_at[m].size()});
    if (rc) return rc;
  }
  // ---- leaves
  rc = leaf_terms(h, st, c, Range{0, (int)h->leaves.size()});
  if (rc) return rc;
  // ---- upward, levels >= shard level
  for (int m = (int)h->internal_at.size() - 1; m >= h->shard_level; --m) {
    rc = upward_level(h, st, c, m);
    if (rc) return rc;
  }
  rc = export_summaries(h, st, c, dev_summary);
  if (rc) return rc;
  CU(cudaGetLastError());
  return MRA_OK;
}

// The replicated top of the upward pass (levels < shard level) and the final reduction.
int launch_likelihood_top(mra_handle* h, cudaStream_t st, const double* dev_summary) {
  const Layout& L = h->lay;
  DevCtx c = make_ctx(h);
  const int r = h->r;
  if (h->shard_level > 0) {
    const int m = h->shard_level - 1;
    const int nn = m < (int)h->internal_at.size() ? (int)h->internal_at[m].size() : 0;
    if (nn) {
      const int* list = reinterpret_cast<const int*>(h->ws + L.lists + h->list_off[m]);
      const int W = (m + 1) * r + 1;
      dim3 g(nn, std::min(64, (W * W + 255) / 256));
      LAUNCH("assemble_summary", k_assemble_from_summary<<<g, 256, 0, st>>>(c, list, dev_summary, h->slot_base));
      node_factor(h, st, c, m, list, nn);
    }
    for (int mm = std::min(m - 1, (int)h->internal_at.size() - 1); mm >= 0; --mm) {
      int rc = upward_level(h, st, c, mm);
      if (rc) return rc;
    }
  }
  LAUNCH("finalize", k_finalize<<<1, NT, 0, st>>>(c, at<double>(h, L.out)));
  CU(cudaGetLastError());
  h->lik_done = true;
  h->pred_done = false;
  return MRA_OK;
}

int launch_predict(mra_handle* h, cudaStream_t st, double* dev_mean, double* dev_sd) {
  const Layout& L = h->lay;
  DevCtx c = make_ctx(h);
  const int r = h->r;
  if (!h->pred_done) {
    if (!h->leafq_done) {
      int rc = leaf_predict_terms(h, st, c, Range{0, (int)h->leaves.size()});
      if (rc) return rc;
      h->leafq_done = true;
    }
    if (!h->fold_items.empty())
      MRA_FOR_VEC(h, LAUNCH("fold", k_fold<V_><<<(unsigned)h->fold_items.size(), NT, GS1 + sizeof(long long) * MAX_LEVELS, st>>>(
                                        c, at<int4>(h, L.fold))));
    if (!h->leaf_tiles.empty())
    if (r % KC == 0 && r <= TB && !(c.tune & 512)) {      // table-free loader (MRA_TUNE bit 9: the general kernel)
      const int nt = (int)h->leaf_tiles.size();
      if (!(c.tune & 1024)) {                              // 64-row tiles, 4 CTAs per SM
        const size_t sm2 = smem_predict2(r, h->depth, TB);
        MRA_FOR_VEC_NJ(h, CU(smem_at_least((k_predict_fused2<J_, 64>), sm2)));
        MRA_FOR_VEC_NJ(h, LAUNCH("predict_fused", (k_predict_fused2<J_, 64><<<(unsigned)nt, 128, sm2, st>>>(
                                                   c, at<int4>(h, L.ltiles), nt, h->depth))));
      } else {      // MRA_TUNE bit 10: 128 rows per CTA (pairs of tiles share their B chunks) -- measured slower, 44.8 vs 41.5 ms
        const size_t sm2 = smem_predict2(r, h->depth, 2 * TB);
        MRA_FOR_VEC_NJ(h, CU(smem_at_least((k_predict_fused2<J_, 128>), sm2)));
        MRA_FOR_VEC_NJ(h, LAUNCH("predict_fused", (k_predict_fused2<J_, 128><<<(unsigned)((nt + 1) / 2), 256, sm2, st>>>(
                                                   c, at<int4>(h, L.ltiles), nt, h->depth))));
      }
    } else
      MRA_FOR_VEC_NJ(h, LAUNCH("predict_fused", k_predict_fused<V_, J_><<<(unsigned)h->leaf_tiles.size(), NT, smem_predict(r, h->depth), st>>>(
                                                 c, at<int4>(h, L.ltiles), h->depth)));
    h->pred_done = true;   // V now holds the posterior-updated basis; results stay cached in mean/var
  }
  double* om = dev_mean ? dev_mean : at<double>(h, L.out_mean);
  double* os = dev_sd ? dev_sd : at<double>(h, L.out_sd);
  if (h->shard_level > 0) {   // rows of other ranks stay zero so the caller can sum-reduce the outputs
    CU(cudaMemsetAsync(om, 0, sizeof(double) * h->N, st));
    CU(cudaMemsetAsync(os, 0, sizeof(double) * h->N, st));
  }
  if (!h->emit_chunks.empty())
    LAUNCH("unpermute", k_unpermute<<<(unsigned)h->emit_chunks.size(), 256, 0, st>>>(
                            c.mean, c.var, at<int>(h, L.perm), at<int2>(h, L.chunks), om, os, c.P, c.status));
  CU(cudaGetLastError());
  return MRA_OK;
}

// Work lists of this rank: internal nodes per level, leaves, 64-row tiles for the prior and predict
// passes, gathered knot tiles and emit ranges.  Not sharded: everything.  Sharded at level s: my subtrees
// (role 1) in full, the replicated top (role >= 2) restricted to the rows of my subtrees and of top-level
// leaves, plus the rows of the top nodes' knots that lie in other ranks' subtrees (needed by knot_factor).
void build_lists(mra_handle* h) {
  HostTrace tr("build_lists");
  const int nn = h->n_nodes, s = h->shard_level;
  const size_t nlev = (size_t)h->depth + 1;
  reset_levels(h->internal_at, nlev);
  reset_levels(h->ptiles_at, nlev);
  h->leaves.clear();
  h->gather_rows.clear();
  h->emit_chunks.clear();
  h->sroots.clear();
  {
    std::vector<int64_t> rows_at(nlev, 0);
    for (int n = 0; n < nn; ++n)
      if (h->kind[n] == KIND_INTERNAL) rows_at[h->level[n]] += h->row_count[n] / TB + 1;
    for (size_t m = 0; m < nlev; ++m) h->ptiles_at[m].reserve((size_t)rows_at[m] + 8);
  }
  auto add_tiles = [&](int node, int64_t row0, int64_t cnt) {
    const int lv = h->level[node];
    for (int64_t r0 = 0; r0 < cnt; r0 += TB) {
      int4 t = make_int4(node, (int)(row0 + r0), (int)std::min<int64_t>(TB, cnt - r0), 0);
      h->ptiles_at[lv].push_back(t);
    }
  };
  auto add_emit = [&](int64_t row0, int64_t cnt) {
    for (int64_t r0 = 0; r0 < cnt; r0 += 2048)
      h->emit_chunks.push_back(make_int2((int)(row0 + r0), (int)std::min<int64_t>(2048, cnt - r0)));
  };
  if (s == 0) {
    for (int n = 0; n < nn; ++n) {
      if (h->kind[n] == KIND_INTERNAL) h->internal_at[h->level[n]].push_back(n);
      else h->leaves.push_back(n);
    }
    {
      // the tile lists of the levels are independent of each other: one host thread per level (N / 64 tiles each)
      std::vector<std::thread> th;
      for (size_t m = 0; m < nlev; ++m) {
        if (h->internal_at[m].empty()) continue;
        th.emplace_back([h, m] {
          // sized first and filled through a raw pointer: the vector headers of the levels sit next to each other in memory,
          // and a push_back per tile from seven threads made their cache line bounce (~50 ns per tile)
          std::vector<int4>& tl = h->ptiles_at[m];
          size_t total = 0;
          for (int n : h->internal_at[m]) total += (size_t)((h->row_count[n] + TB - 1) / TB);
          tl.resize(total);
          int4* p = tl.data();
          for (int n : h->internal_at[m]) {
            const int64_t row0 = h->row_start[n], cnt = h->row_count[n];
            for (int64_t r0 = 0; r0 < cnt; r0 += TB) *p++ = make_int4(n, (int)(row0 + r0), (int)std::min<int64_t>(TB, cnt - r0), 0);
          }
        });
      }
      for (auto& t : th) t.join();
    }
    add_emit(0, h->N);
    h->n_regular_tiles.assign(nlev, 0);
    for (size_t m = 0; m < nlev; ++m) h->n_regular_tiles[m] = (int)h->ptiles_at[m].size();
    h->gather_node.clear();
  } else {
    for (int n = 0; n < nn; ++n) {
      const int role = h->role[n], lv = h->level[n];
      if (!role) continue;
      const bool internal = h->kind[n] == KIND_INTERNAL;
      if (internal) h->internal_at[lv].push_back(n);
      else h->leaves.push_back(n);
      if (lv == s) h->sroots.push_back(n);
      const bool piece = (lv == s) || (lv < s && !internal);
      if (piece) {
        for (int a = h->parent[n]; a >= 0; a = h->parent[a]) add_tiles(a, h->row_start[n], h->row_count[n]);
        if (role == 1 || role == 3) add_emit(h->row_start[n], h->row_count[n]);
      }
      if (lv >= s && internal) add_tiles(n, h->row_start[n], h->row_count[n]);
    }
    // The knots of a replicated top node at levels 1 .. s-1 may lie anywhere below it, also in other ranks'
    // subtrees, and knot_factor needs the ancestors' basis at those rows: all r knot rows of every such node are
    // "gathered" into extra tiles of the ancestors' levels.  The lists have a fixed shape (r rows per node), so
    // they can be laid out before the knots are drawn; refresh_gather_rows() fills in the row ids.
    h->n_regular_tiles.assign(nlev, 0);
    for (size_t m = 0; m < nlev; ++m) h->n_regular_tiles[m] = (int)h->ptiles_at[m].size();
    h->gather_node.clear();
    for (int n = 0; n < nn; ++n) {
      const int lv = h->level[n];
      if (h->role[n] < 2 || h->kind[n] != KIND_INTERNAL || lv < 1 || lv >= s) continue;
      const int g0 = (int)h->gather_rows.size();
      h->gather_node.push_back(make_int2(n, g0));
      for (int i = 0; i < h->r; ++i) h->gather_rows.push_back(h->knot_rows[h->knot_off[n] + i]);
      for (int a = h->parent[n]; a >= 0; a = h->parent[a])
        for (int r0 = 0; r0 < h->r; r0 += TB)
          h->ptiles_at[h->level[a]].push_back(make_int4(a, 0, std::min(TB, h->r - r0), g0 + r0 + 1));
    }
  }
  while (!h->internal_at.empty() && h->internal_at.back().empty()) {
    h->internal_at.pop_back();
    h->ptiles_at.pop_back();
  }
  tr.mark("tiles");
  // groups of the regular prior tiles: consecutive full tiles of one node, at most PG per group
  reset_levels(h->pgroups_at, h->ptiles_at.size());
  reset_levels(h->group_of_tile, h->ptiles_at.size());
  std::vector<std::thread> gth;
  for (size_t m = 0; m < h->ptiles_at.size(); ++m) gth.emplace_back([h, m] {
    const std::vector<int4>& tl = h->ptiles_at[m];
    std::vector<int4>& gl = h->pgroups_at[m];
    std::vector<int>& got = h->group_of_tile[m];
    const int nreg = h->n_regular_tiles[m];
    got.resize((size_t)nreg);
    gl.resize((size_t)nreg);            // upper bound; filled through raw pointers (see the tile lists), trimmed below
    int4* gp = gl.data();
    int* gt = got.data();
    int ng = 0;
    for (int i = 0; i < nreg; ++i) {
      const int4 t = tl[i];
      bool extend = false;
      if (ng > 0) {
        const int4& gb = gp[ng - 1];
        extend = gb.x == t.x && gb.y + gb.z == t.y && gb.z % TB == 0 && gb.z < PG * TB;
      }
      if (extend) gp[ng - 1].z += t.z;
      else gp[ng++] = make_int4(t.x, t.y, t.z, 0);
      gt[i] = ng - 1;
    }
    gl.resize((size_t)ng);
  });
  for (auto& t : gth) t.join();
  tr.mark("groups");
  {
    size_t total = 0;
    for (int n : h->leaves) total += (size_t)((h->row_count[n] + TB - 1) / TB);
    h->leaf_tiles.resize(total);
    int4* p = h->leaf_tiles.data();
    for (int n : h->leaves)
      for (int64_t r0 = 0; r0 < h->row_count[n]; r0 += TB)
        *p++ = make_int4(n, (int)(h->row_start[n] + r0), (int)std::min<int64_t>(TB, h->row_count[n] - r0), 0);
  }
  tr.mark("leaf tiles");
  // parts for the streamed evaluation: the subtree of every child of the root is a contiguous range of each list.
  // Unsharded: every part is mine.  Sharded at level 1: the parts are the shards.  Sharded at level 2: a part is
  // mine when one of my subtrees lies in it; its level-1 node is a replicated top node whose gathered knot tiles
  // (level 0) belong to the part as well.
  h->n_parts = 0;
  h->my_parts = 0;
  h->part_nodes.clear();
  h->part_tiles.clear();
  h->part_gtiles.clear();
  h->part_leaves.clear();
  h->part_knots.clear();
  h->part_gather.clear();
  if (s <= 2 && nn > 1 && h->kind[0] == KIND_INTERNAL && !h->internal_at.empty()) {
    const int np = h->child_count[0], c0 = h->child_start[0];
    std::vector<int> part_of((size_t)nn, -1);
    for (int n = 1; n < nn; ++n) part_of[n] = h->parent[n] == 0 ? n - c0 : part_of[h->parent[n]];
    h->n_parts = np;
    const size_t nl = h->internal_at.size();
    h->part_nodes.assign(nl, std::vector<Range>(np, Range{0, 0}));
    h->part_tiles.assign(nl, std::vector<Range>(np, Range{0, 0}));
    h->part_gtiles.assign(np, Range{0, 0});
    h->part_leaves.assign(np, {});
    h->part_knots.assign(np, {});
    h->part_gather.assign(np, make_int2(-1, 0));
    auto extend = [](Range& rg, int i) {
      if (rg.count == 0) rg.begin = i;
      ++rg.count;
    };
    bool ok = np <= 30;
    for (size_t m = 1; m < nl && ok; ++m) {
      int last = -1;
      for (int i = 0; i < (int)h->internal_at[m].size(); ++i) {
        const int n = h->internal_at[m][i], pp = part_of[n];
        if (pp < last) ok = false;
        last = pp;
        extend(h->part_nodes[m][pp], i);
        if (h->role[n] == 1) h->my_parts |= 1 << pp;
        std::vector<Range>& kr = h->part_knots[pp];
        const int ko = (int)h->knot_off[n];
        if (!kr.empty() && kr.back().begin + kr.back().count == ko) kr.back().count += h->r;
        else kr.push_back(Range{ko, h->r});
      }
      last = -1;
      {
        // the tiles of a node are consecutive: walk the list node run by node run (N / 64 tiles per level)
        const std::vector<int4>& tl = h->ptiles_at[m];
        const int nreg = h->n_regular_tiles[m];
        int i = 0;
        while (i < nreg) {
          const int node = tl[i].x;
          // a node's run is at most ceil(rows / 64) tiles long; pieces of a replicated node (sharded) form shorter runs
          int j = i + 1;
          const int64_t cap = i + (h->row_count[node] + TB - 1) / TB;
          if (cap <= nreg && cap > i && tl[cap - 1].x == node && (cap == nreg || tl[cap].x != node)) j = (int)cap;
          else
            while (j < nreg && tl[j].x == node) ++j;
          const int pp = part_of[node];
          if (pp < last) ok = false;
          last = pp;
          Range& rg = h->part_tiles[m][pp];
          if (rg.count == 0) rg.begin = i;
          rg.count += j - i;
          i = j;
        }
      }
      if (h->n_regular_tiles[m] != (int)h->ptiles_at[m].size()) ok = false;   // gathered tiles below level 0: s > 2
    }
    for (int i = 0; i < (int)h->leaves.size() && ok; ++i) {
      const int n = h->leaves[i];
      if (h->role[n] != 1 && s > 0) {       // a replicated leaf above the shard level: not streamed
        ok = false;
        break;
      }
      h->my_parts |= 1 << part_of[n];
      std::vector<Range>& lr = h->part_leaves[part_of[n]];
      if (!lr.empty() && lr.back().begin + lr.back().count == i) ++lr.back().count;
      else lr.push_back(Range{i, 1});
    }
    // gathered tiles of the level-1 top nodes (sharded at level 2) sit behind the regular tiles of level 0
    for (const int2& gn : h->gather_node) {
      if (h->level[gn.x] != 1) {
        ok = false;
        break;
      }
      h->part_gather[part_of[gn.x]] = gn;
    }
    for (int i = h->n_regular_tiles[0]; i < (int)h->ptiles_at[0].size() && ok; ++i) {
      const int slot = h->ptiles_at[0][i].w - 1;
      int pp = -1;
      for (const int2& gn : h->gather_node)
        if (slot >= gn.y && slot < gn.y + h->r) pp = part_of[gn.x];
      if (pp < 0) ok = false;
      else extend(h->part_gtiles[pp], i);
    }
    if (s == 0) h->my_parts = (1 << np) - 1;
    if (!ok) h->n_parts = 0;      // streaming is refused then
  }
}

}  // namespace

// ================================================================================================
extern "C" {

const char* mra_version(void) { return "pymra_b200 0.1 (sm_100a)"; }

int mra_create(mra_handle** out, int device) {
  if (!out) return MRA_ERR_ARG;
  const char* env_groups = std::getenv("MRA_PRIOR_GROUPS");
  *out = nullptr;
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev <= 0 || device < 0 || device >= ndev) return MRA_ERR_CUDA;
  mra_handle* h = new (std::nothrow) mra_handle();
  if (!h) return MRA_ERR_NOMEM;
  h->device = device;
  h->use_groups = !(env_groups && env_groups[0] == '0');      // A/B switch for profiling (default: grouped prior tiles)
  if (const char* e = std::getenv("MRA_CHOL_MMA")) h->chol_mma = e[0] != '0';
  if (const char* e = std::getenv("MRA_LEAF_V2")) h->leaf_v2 = e[0] != '0';
  if (const char* e = std::getenv("MRA_LEAF_WIDE")) h->leaf_wide = e[0] != '0';
  {
    DevGuard g(device);          // validates the ordinal / creates the context; the caller's device is restored
    if (g.err != cudaSuccess) {
      delete h;
      return MRA_ERR_CUDA;
    }
  }
  if (!env_on("MRA_NO_HOST_POOL")) {
    std::lock_guard<std::mutex> lock(g_pool_mu);
    if (g_pool) {
      swap_buffers(h, *g_pool);
      g_pool.reset();
    }
  }
  *out = h;
  return MRA_OK;
}

int mra_destroy(mra_handle* h) {
  if (h && h->lists_job.valid()) h->lists_job.get();
  if (h && !env_on("MRA_NO_HOST_POOL")) {
    std::unique_ptr<HostBuffers> b(new (std::nothrow) HostBuffers());
    if (b) {
      swap_buffers(h, *b);
      std::lock_guard<std::mutex> lock(g_pool_mu);
      g_pool = std::move(b);      // the previous set, if any, is released
    }
  }
  if (h && h->copy_stream) {
    DevGuard g(h->device);
    cudaEventDestroy(h->copy_event);
    cudaStreamDestroy(h->copy_stream);
  }
  if (h) {
    drop_graph(h);
    if (h->gstream) cudaStreamDestroy(h->gstream);
    for (auto& rec : h->prof_recs) {
      cudaEventDestroy(rec.e0);
      cudaEventDestroy(rec.e1);
    }
    for (auto& e : h->event_pool) cudaEventDestroy(e);
  }
  delete h;
  return MRA_OK;
}

const char* mra_last_error(const mra_handle* h) { return h ? h->err.c_str() : "null handle"; }

namespace {
void wait_lists(mra_handle* h) {
  if (h->lists_job.valid()) h->lists_job.get();
}
void start_lists(mra_handle* h) {
  h->lists_built = true;
  if (std::getenv("MRA_SYNC_LISTS")) build_lists(h);
  else h->lists_job = std::async(std::launch::async, [h] { build_lists(h); });
}
}  // namespace

int mra_set_structure(mra_handle* h, const mra_structure* s) {
  if (!h || !s) return MRA_ERR_ARG;
  wait_lists(h);
  if (s->n_locs <= 0 || s->n_locs >= (int64_t(1) << 31)) return fail(h, MRA_ERR_ARG, "n_locs out of range");
  if (s->dim != 1 && s->dim != 2) return fail(h, MRA_ERR_ARG, "dim must be 1 or 2");
  if (s->r < 1 || s->r > 128) return fail(h, MRA_ERR_ARG, "r must be in [1, 128] in this build");
  if (s->n_nodes < 1 || s->depth < 0) return fail(h, MRA_ERR_ARG, "empty tree");
  if (s->depth >= MAX_LEVELS) return fail(h, MRA_ERR_ARG, "tree deeper than MAX_LEVELS in this build");
  HostTrace tr("set_structure");
  h->N = s->n_locs;
  h->dim = s->dim;
  h->r = s->r;
  h->depth = s->depth;
  h->n_nodes = s->n_nodes;
  const int nn = s->n_nodes;
  h->level.assign(s->node_level, s->node_level + nn);
  h->parent.assign(s->node_parent, s->node_parent + nn);
  h->kind.assign(s->node_kind, s->node_kind + nn);
  h->row_start.assign(s->node_row_start, s->node_row_start + nn);
  h->row_count.assign(s->node_row_count, s->node_row_count + nn);
  h->child_start.assign(s->node_child_start, s->node_child_start + nn);
  h->child_count.assign(s->node_child_count, s->node_child_count + nn);
  h->knot_off.assign(s->node_knot_off, s->node_knot_off + nn);
  h->level_off.assign(s->level_off, s->level_off + s->depth + 2);
  h->knot_rows.resize(s->n_knot_rows);
  {
    std::vector<int> bad(1, 0);
    int* badp = bad.data();
    int* dst = h->knot_rows.data();
    const int64_t* src = s->knot_rows;
    const int64_t N = h->N;
    parallel_for(s->n_knot_rows, [=](int64_t a, int64_t b) {
      for (int64_t i = a; i < b; ++i) {
        if (src[i] < 0 || src[i] >= N) *badp = 1;
        dst[i] = (int)src[i];
      }
    });
    if (bad[0]) return fail(h, MRA_ERR_ARG, "knot row out of range");
  }
  tr.mark("node arrays + knots");
  h->perm.resize(h->N);
  {
    std::vector<int> bad(1, 0);
    int* badp = bad.data();
    int* dst = h->perm.data();
    const int64_t* src = s->perm;
    const int64_t N = h->N;
    parallel_for(N, [=](int64_t a, int64_t b) {
      for (int64_t i = a; i < b; ++i) {
        if (src[i] < 0 || src[i] >= N) *badp = 1;
        dst[i] = (int)src[i];
      }
    });
    if (bad[0]) return fail(h, MRA_ERR_ARG, "perm entry out of range");
  }
  for (int n = 0; n < nn; ++n) {
    const int lv = h->level[n];
    if (lv < 0 || lv > s->depth) return fail(h, MRA_ERR_ARG, "node level out of range");
    if (h->row_start[n] < 0 || h->row_start[n] + h->row_count[n] > h->N)
      return fail(h, MRA_ERR_ARG, "node row range out of bounds");
    if (h->parent[n] >= n) return fail(h, MRA_ERR_ARG, "nodes must be numbered level by level (parent before child)");
    if (h->kind[n] == KIND_INTERNAL) {
      if (h->knot_off[n] < 0 || h->knot_off[n] + h->r > s->n_knot_rows)
        return fail(h, MRA_ERR_ARG, "internal node without r knots");
      if (h->child_count[n] <= 0) return fail(h, MRA_ERR_ARG, "internal node without children");
      if (h->child_count[n] > 16) return fail(h, MRA_ERR_ARG, "more than 16 children per node (the reference's IDs allow 9)");
    }
  }
  tr.mark("perm + validation");
  h->shard_level = 0;
  h->role.assign(nn, 1);
  h->ldv = std::max<long long>(2, ((long long)std::max(h->depth, 1) * h->r + 1) / 2 * 2);
  h->has_structure = true;
  h->planned = h->bound = h->uploaded = h->lik_done = h->pred_done = false;
  h->lists_built = false;
  if (!h->defer_lists) start_lists(h);
  h->defer_lists = false;
  tr.mark("build_lists started");
  return MRA_OK;
}

int mra_expect_shard(mra_handle* h) {
  if (!h) return MRA_ERR_ARG;
  h->defer_lists = true;
  return MRA_OK;
}

// The plan has a tree-dependent part (node table of the internal nodes, work lists, the basis slab V and every
// per-internal-node block: "part A" of the arena) and an observation-dependent part (NaN scan, the leaves' observed /
// unobserved row lists, the per-leaf blocks: "part B").  mra_plan lays both out in ONE arena.  mra_plan_tree /
// mra_bind_tree / mra_plan_obs / mra_bind_obs expose the two halves, so that a caller can start the prior pass (which
// needs part A only) while the host is still scanning the observations.
namespace {

struct LeafOffsets {
  long long s_off = 0, di_off = 0, ut_off = 0, qt_off = 0, utt_off = 0;
};

int plan_tree_impl(mra_handle* h, int want_predict) {
  HostTrace tr("plan_tree");
  const int nn = h->n_nodes, r = h->r;
  h->want_predict = want_predict != 0;
  drop_graph(h);
  if (!h->lists_built) start_lists(h);
  wait_lists(h);
  tr.mark("wait for build_lists");
  h->nodes.assign(nn, NodeDev{});
  long long a_off = 0, gt_off = 0, lp_off = 0, vk_off = 0, linv_off = 0;
  const int W_assemble_A = kid(h, "assemble_A");
  const int W_knot_chol = kid(h, "knot_chol");
  const int W_knot_gram = kid(h, "knot_gram");
  const int W_knot_vkl = kid(h, "knot_vkl");
  const int W_node_chol = kid(h, "node_chol");
  const int W_node_gt = kid(h, "node_gt");
  const int W_predict_fused = kid(h, "predict_fused");
  const int W_prior_tiles = kid(h, "prior_tiles");
  const int W_unpermute = kid(h, "unpermute");
  for (auto& f : h->kflops) f = 0.0;
  for (auto& f : h->kbytes) f = 0.0;
  std::vector<double> my_rows_prior((size_t)nn, 0.0), my_rows_pred((size_t)nn, 0.0);
  for (size_t m = 0; m < h->ptiles_at.size(); ++m)
    for (size_t i = 0; i < h->ptiles_at[m].size(); ++i) {
      const int4& t = h->ptiles_at[m][i];
      my_rows_prior[t.x] += t.z;
      if (t.w == 0) my_rows_pred[t.x] += t.z;
    }
  for (int n = 0; n < nn; ++n) {
    NodeDev& d = h->nodes[n];
    d.level = h->level[n];
    d.kind = h->kind[n];
    d.parent = h->parent[n];
    d.child_start = h->child_start[n];
    d.child_count = h->child_count[n];
    d.row_start = (int)h->row_start[n];
    d.row_count = (int)h->row_count[n];
    d.knot_off = (int)h->knot_off[n];
    d.W = d.level * r + 1;
    if (!h->role[n]) continue;       // another rank's subtree
    if (d.kind != KIND_INTERNAL) continue;
    const double Kv = (double)d.level * r;
    const int Wa = (d.level + 1) * r + 1;
    d.lda = (Wa + 3) / 4 * 4;
    d.a_off = a_off;
    a_off += (long long)Wa * d.lda;
    d.gt_off = gt_off;
    gt_off += (long long)d.W * r;
    d.lpinv_off = lp_off;
    lp_off += (long long)r * r;
    d.linv_off = linv_off;
    linv_off += (long long)r * r;
    d.vk_off = vk_off;
    vk_off += (long long)r * d.level * r;
    // rows of this node this rank really works on: all of them unless the node is a replicated top node of a
    // sharded handle, whose tile list holds only this rank's pieces (+ gathered knot rows, prior pass only)
    const double nr = my_rows_prior[n], nrp = my_rows_pred[n], rr = (double)r;
    const double Waf = Kv + rr + 1;
    add_w(h, W_knot_gram, rr * rr * Kv, 8.0 * (2.0 * rr * Kv + rr * rr / 2));
    add_w(h, W_knot_chol, 2.0 * rr * rr * rr / 3.0, 8.0 * rr * rr);
    add_w(h, W_knot_vkl, 2.0 * rr * rr * Kv, 8.0 * (2.0 * rr * Kv + rr * rr));
    add_w(h, W_prior_tiles, 2.0 * nr * rr * Kv + nr * rr * rr, 8.0 * nr * (Kv + rr + 2));
    double fa = 0;     // assemble: symmetric half of W x W, K = n_obs (leaf child) or r (internal child)
    for (int ch = d.child_start; ch < d.child_start + d.child_count; ++ch)
      if (h->kind[ch] == KIND_INTERNAL && !(h->shard_level > 0 && d.level == h->shard_level - 1)) fa += Waf * Waf * rr;
    add_w(h, W_assemble_A, fa, 8.0 * Waf * Waf);
    add_w(h, W_node_chol, 2.0 * rr * rr * rr / 3.0, 8.0 * 2.0 * rr * rr);
    add_w(h, W_node_gt, (Kv + 1) * rr * rr, 8.0 * (rr * rr + 2 * (Kv + 1) * rr));
    add_w(h, W_predict_fused, nrp * rr * rr + 2.0 * nrp * rr * Kv + 4.0 * nrp * rr, 8.0 * nrp * rr);
  }
  add_w(h, W_unpermute, 0.0, 8.0 * 4.0 * (double)h->N);
  tr.mark("node loop (tree)");
  // ---- arena layout, part A
  Arena ar;
  Layout& L = h->lay;
  L = Layout{};
  const size_t N = (size_t)h->N, D = sizeof(double);
  L.nodes = ar.take(sizeof(NodeDev) * nn);
  L.knot_rows = ar.take(sizeof(int) * std::max<size_t>(1, h->knot_rows.size()));
  L.perm = ar.take(sizeof(int) * N);
  L.xs = ar.take(D * N);
  L.ys = ar.take(D * N);
  L.yobs = ar.take(D * N);
  L.V = ar.take(D * N * (size_t)h->ldv);
  L.GTF = ar.take(D * std::max<long long>(1, h->want_predict ? gt_off : 0));
  L.A = ar.take(D * std::max<long long>(1, a_off));
  L.GT = ar.take(D * std::max<long long>(1, gt_off));
  L.LPINV = ar.take(D * std::max<long long>(1, lp_off));
  L.VK = ar.take(D * std::max<long long>(1, vk_off));
  L.VKL = ar.take(D * std::max<long long>(1, vk_off));
  L.LINV = ar.take(D * std::max<long long>(1, linv_off));
  L.dnode = ar.take(D * nn);
  L.mean = ar.take(D * N);
  L.var = ar.take(D * N);
  L.vnorm = ar.take(D * N);
  L.xidx = ar.take(D * N);
  L.status = ar.take(256);
  L.params = ar.take(256);
  L.out = ar.take(256);
  L.stage_locs = ar.take(D * N * h->dim);
  L.stage_obs = ar.take(D * N);
  L.out_mean = ar.take(D * N);
  L.out_sd = ar.take(D * N);
  h->list_off.assign(h->internal_at.size(), 0);
  h->ptiles_off.assign(h->internal_at.size(), 0);
  size_t lo = 0, po = 0;
  for (size_t m = 0; m < h->internal_at.size(); ++m) {
    h->list_off[m] = lo;
    lo += (sizeof(int) * h->internal_at[m].size() + 255) & ~size_t(255);
    h->ptiles_off[m] = po;
    po += (sizeof(int4) * h->ptiles_at[m].size() + 255) & ~size_t(255);
  }
  h->pgroups_off.assign(h->internal_at.size(), 0);
  size_t go = 0;
  for (size_t m = 0; m < h->internal_at.size(); ++m) {
    h->pgroups_off[m] = go;
    go += (sizeof(int4) * h->pgroups_at[m].size() + 255) & ~size_t(255);
  }
  h->leaves_off = lo;
  lo += (sizeof(int) * h->leaves.size() + 255) & ~size_t(255);
  h->sroots_off = lo;
  lo += (sizeof(int) * h->sroots.size() + 255) & ~size_t(255);
  L.lists = ar.take(std::max<size_t>(256, lo));
  L.ptiles = ar.take(std::max<size_t>(256, po));
  L.pgroups = ar.take(std::max<size_t>(256, go));
  L.gather = ar.take(std::max<size_t>(256, sizeof(int) * h->gather_rows.size()));
  L.chunks = ar.take(std::max<size_t>(256, sizeof(int2) * h->emit_chunks.size()));
  L.ltiles = ar.take(std::max<size_t>(256, sizeof(int4) * h->leaf_tiles.size()));
  h->total_A = ar.off;
  h->total_B = 0;
  h->tree_planned = true;
  h->planned = h->bound = h->tree_bound = h->uploaded = h->lik_done = h->pred_done = false;
  return MRA_OK;
}

// Part B.  Offsets are laid out relative to the start of part B; place_obs_part() turns them into offsets from h->ws.
int plan_obs_impl(mra_handle* h, const double* obs) {
  HostTrace tr("plan_obs");
  const int nn = h->n_nodes, r = h->r;
  long long s_off = 0, di_off = 0, ut_off = 0, qt_off = 0, utt_off = 0;
  h->max_leaf_obs = h->max_leaf_rows = h->max_leaf_unobs = 0;
  h->max_leaf_W = 1;
  const int W_assemble_A = kid(h, "assemble_A");
  const int W_fold = kid(h, "fold");
  const int W_leaf_chol = kid(h, "leaf_chol");
  const int W_leaf_gram = kid(h, "leaf_gram");
  const int W_leaf_gram_T = kid(h, "leaf_gram_T");
  const int W_leaf_linv = kid(h, "leaf_linv");
  const int W_leaf_q = kid(h, "leaf_q");
  const int W_leaf_qobs = kid(h, "leaf_qobs");
  const int W_leaf_solve = kid(h, "leaf_solve");
  const int W_leaf_solve_Q = kid(h, "leaf_solve_Q");
  const int W_leaf_trsm = kid(h, "leaf_trsm");
  const int W_leaf_upd = kid(h, "leaf_upd");
  const int W_leaf_ut = kid(h, "leaf_ut");
  const int W_predict_fused = kid(h, "predict_fused");
  std::vector<uint8_t, NoInit<uint8_t>> finite_row((size_t)h->N);     // np.isfinite(obs) in tree order (MRANode.py:415)
  {
    // two passes: the flags in the caller's order (a sequential scan of obs), then a gather of BYTES through the
    // permutation -- the random reads hit a 1-byte-per-location table that stays in cache, not the 8-byte observations
    std::vector<uint8_t, NoInit<uint8_t>> finite_c((size_t)h->N);
    uint8_t* fc = finite_c.data();
    parallel_for(h->N, [=](int64_t a, int64_t b) {
      for (int64_t i = a; i < b; ++i) fc[i] = std::isfinite(obs[i]) ? 1 : 0;
    });
    uint8_t* fr = finite_row.data();
    const int* pm = h->perm.data();
    parallel_for(h->N, [=](int64_t a, int64_t b) {
      for (int64_t i = a; i < b; ++i) fr[i] = fc[pm[i]];
    });
  }
  tr.mark("finite flags");
  // observed / unobserved row lists of every leaf of this rank: count per leaf, prefix, fill -- all in parallel
  std::vector<int> leaf_ids, leaf_obs_off, leaf_unobs_off;
  for (int n = 0; n < nn; ++n)
    if (h->role[n] && h->kind[n] == KIND_LEAF) leaf_ids.push_back(n);
  {
    const int64_t nl = (int64_t)leaf_ids.size();
    std::vector<int> cnt((size_t)nl, 0);
    const uint8_t* fr = finite_row.data();
    const int* lid = leaf_ids.data();
    int* cp = cnt.data();
    const int64_t* rs = h->row_start.data();
    const int64_t* rc = h->row_count.data();
    auto over_leaves = [&](auto fn) {
      // parallel_for splits an index range; give it the leaves
      unsigned nt = std::min(16u, std::max(1u, std::thread::hardware_concurrency()));
      if (const char* e = std::getenv("MRA_HOST_THREADS")) nt = (unsigned)std::max(1, std::min((int)nt, std::atoi(e)));
      if (nl < 1024 || nt == 1) {
        fn((int64_t)0, nl);
        return;
      }
      std::vector<std::thread> th;
      const int64_t step = (nl + nt - 1) / nt;
      for (unsigned t = 0; t < nt; ++t) {
        const int64_t a = (int64_t)t * step, b = std::min(nl, a + step);
        if (a < b) th.emplace_back([=] { fn(a, b); });
      }
      for (auto& t : th) t.join();
    };
    over_leaves([=](int64_t a, int64_t b) {
      for (int64_t k = a; k < b; ++k) {
        int c0 = 0;
        const int64_t s0 = rs[lid[k]], e0 = s0 + rc[lid[k]];
        for (int64_t row = s0; row < e0; ++row) c0 += fr[row];
        cp[k] = c0;
      }
    });
    leaf_obs_off.resize((size_t)nl + 1);
    leaf_unobs_off.resize((size_t)nl + 1);
    int64_t oo = 0, uo = 0;
    for (int64_t k = 0; k < nl; ++k) {
      leaf_obs_off[k] = (int)oo;
      leaf_unobs_off[k] = (int)uo;
      oo += cnt[k];
      if (h->want_predict) uo += (int)rc[lid[k]] - cnt[k];
    }
    leaf_obs_off[nl] = (int)oo;
    leaf_unobs_off[nl] = (int)uo;
    h->obs_rows.resize((size_t)oo);
    h->unobs_rows.resize((size_t)uo);
    int* orow = h->obs_rows.data();
    int* urow = h->unobs_rows.data();
    const int* ooff = leaf_obs_off.data();
    const int* uoff = leaf_unobs_off.data();
    const bool wp = h->want_predict;
    over_leaves([=](int64_t a, int64_t b) {
      for (int64_t k = a; k < b; ++k) {
        int io = ooff[k], iu = uoff[k];
        const int64_t s0 = rs[lid[k]], e0 = s0 + rc[lid[k]];
        for (int64_t row = s0; row < e0; ++row) {
          if (fr[row]) orow[io++] = (int)row;
          else if (wp) urow[iu++] = (int)row;
        }
      }
    });
  }
  tr.mark("leaf row lists");
  int64_t leaf_cursor = 0;
  for (int n = 0; n < nn; ++n) {
    NodeDev& d = h->nodes[n];
    if (!h->role[n] || d.kind == KIND_INTERNAL) continue;
    const double Kv = (double)d.level * r;
    if (d.kind == KIND_LEAF) {
      d.obs_off = leaf_obs_off[leaf_cursor];
      d.unobs_off = leaf_unobs_off[leaf_cursor];
      d.n_obs = leaf_obs_off[leaf_cursor + 1] - d.obs_off;
      d.n_unobs = leaf_unobs_off[leaf_cursor + 1] - d.unobs_off;
      ++leaf_cursor;
    } else {          // orphan rows: no data, no residual term
      d.obs_off = d.unobs_off = 0;
      d.n_obs = d.n_unobs = 0;
    }
    if (d.n_obs > 0) h->max_leaf_unobs = std::max(h->max_leaf_unobs, d.n_unobs);
    d.ldo = std::max(4, (d.n_obs + 3) / 4 * 4);
    const int nb = (d.n_obs + TB - 1) / TB;
    d.s_off = s_off;
    s_off += (long long)d.n_obs * d.ldo;
    d.di_off = di_off;
    di_off += (long long)nb * TB * TB;
    d.ut_off = ut_off;
    ut_off += (long long)d.W * d.ldo;
    d.utt_off = utt_off;
    utt_off += (long long)d.n_obs * std::max(2, (d.W - 1 + 1) / 2 * 2);
    d.qt_off = qt_off;
    if (h->want_predict) qt_off += (long long)d.row_count * d.ldo;
    h->max_leaf_obs = std::max(h->max_leaf_obs, d.n_obs);
    h->max_leaf_rows = std::max(h->max_leaf_rows, d.row_count);
    if (d.n_obs > 0) h->max_leaf_W = std::max(h->max_leaf_W, d.W);
    const double no = d.n_obs, nl = d.row_count, W = Kv + 1;
    add_w(h, W_leaf_gram, no * no * Kv, 8.0 * (no * Kv + no * no / 2));
    for (int pb = 0; pb * TB < d.n_obs; ++pb) {      // blocked factorisation as executed (triangular halves)
      const double nv = std::min(TB, d.n_obs - pb * TB), Kp = (double)pb * TB, below = no - Kp - nv;
      add_w(h, W_leaf_chol, 2.0 * nv * nv * nv / 3.0, 8.0 * 2.0 * nv * nv);
      if (pb > 0) add_w(h, W_leaf_upd, nv * nv * Kp, 8.0 * (nv * Kp + nv * nv));
      if (below > 0) add_w(h, W_leaf_trsm, below * nv * (2.0 * Kp + nv), 8.0 * (below * (Kp + 2 * nv) + nv * Kp + nv * nv));
    }
    if (h->leaf_v2) {
      for (int bj = 0; (bj + 1) * TB < d.n_obs; ++bj)
        for (int bi = bj + 1; bi * TB < d.n_obs; ++bi) {
          const double nvi = std::min(TB, d.n_obs - bi * TB);
          add_w(h, W_leaf_linv, nvi * TB * TB * (bi - bj) + nvi * nvi * TB, 8.0 * (3.0 * nvi * TB + TB * TB * (bi - bj)));
        }
      add_w(h, W_leaf_linv, 0.0, 8.0 * no * no);
      add_w(h, W_leaf_ut, no * no * Kv, 8.0 * (3 * no * Kv + no * no / 2));
    } else {
      add_w(h, W_leaf_solve, no * no * W, 8.0 * (2 * no * W + no * no / 2));
    }
    add_w(h, W_assemble_A, W * W * no, 8.0 * no * W);
    if (d.kind == KIND_LEAF) {
      add_w(h, W_predict_fused, 2.0 * nl * Kv, 8.0 * nl * (Kv + 2));
      if (h->leaf_v2) {
        add_w(h, W_leaf_q, 2.0 * (nl - no) * no * Kv + (nl - no) * no * no, 8.0 * ((nl - no) * (Kv + no) + no * (Kv + no / 2)));
        add_w(h, W_leaf_qobs, 0.0, 8.0 * 2.0 * no * no);
      } else {
        add_w(h, W_leaf_gram_T, 2.0 * (nl - no) * no * Kv, 8.0 * ((nl + no) * Kv + nl * no));
        add_w(h, W_leaf_solve_Q, nl * no * no, 8.0 * (2 * nl * no + no * no / 2));
      }
      add_w(h, W_predict_fused, 2.0 * nl * no * W + 2.0 * nl * no, 8.0 * (nl * no + no * W));
    }
  }
  tr.mark("node loop (leaves)");
  h->n_obs_total = (int64_t)h->obs_rows.size();
  h->fold_items.clear();
  if (h->want_predict) {
    const int nct = (r + TB - 1) / TB;
    h->fold_items.reserve((size_t)nn * (size_t)std::max(1, h->depth) * nct * 2);
    for (int n = 0; n < nn; ++n) {
      const NodeDev& d = h->nodes[n];
      if (!h->role[n] || d.level < 1) continue;
      int nxt;
      if (d.kind == KIND_INTERNAL) nxt = nct;
      else if (d.kind == KIND_LEAF && d.n_obs > 0) nxt = (d.n_obs + TB - 1) / TB;
      else continue;
      for (int ct = 0; ct < nct; ++ct)
        for (int xt = 0; xt < nxt; ++xt) h->fold_items.push_back(make_int4(n, 0, ct, xt));      // the CTA walks the levels j
      const double cols = d.kind == KIND_INTERNAL ? r : d.n_obs;
      add_w(h, W_fold, 2.0 * d.level * (double)r * r * cols, 8.0 * 2.0 * d.level * r * cols);
    }
  }
  h->flops_lik = h->flops_pred = 0.0;
  for (size_t i = 0; i < h->kname.size(); ++i) {
    const std::string& nm = h->kname[i];
    const bool pred = nm == "leaf_gram_T" || nm == "leaf_solve_Q" || nm == "leaf_q" || nm == "leaf_qobs" || nm == "fold" ||
                      nm == "predict_fused" || nm == "unpermute";
    (pred ? h->flops_pred : h->flops_lik) += h->kflops[i];
  }
  tr.mark("fold items");
  // ---- arena layout, part B (offsets relative to its start)
  Arena ar;
  Layout& R = h->lay_b;
  R = Layout{};
  const size_t D = sizeof(double);
  R.obs_rows = ar.take(sizeof(int) * std::max<size_t>(1, h->obs_rows.size()));
  R.unobs_rows = ar.take(sizeof(int) * std::max<size_t>(1, h->unobs_rows.size()));
  R.S = ar.take(D * std::max<long long>(1, s_off));
  R.LS = ar.take(D * std::max<long long>(1, h->leaf_v2 ? s_off : 0));
  R.UTTN = ar.take(D * std::max<long long>(1, h->leaf_v2 ? utt_off : 0));
  R.DI = ar.take(D * std::max<long long>(1, di_off));
  R.UT = ar.take(D * std::max<long long>(1, ut_off));
  R.QT = ar.take(D * std::max<long long>(1, qt_off));
  R.UTF = ar.take(D * std::max<long long>(1, h->want_predict ? ut_off : 0));
  R.fold = ar.take(std::max<size_t>(256, sizeof(int4) * h->fold_items.size()));
  h->total_B = ar.off;
  return MRA_OK;
}

// part B starts `delta` bytes after h->ws (modulo 2^64 when it lives in an allocation of its own)
void place_obs_part(mra_handle* h, size_t delta) {
  Layout& L = h->lay;
  const Layout& R = h->lay_b;
  L.obs_rows = delta + R.obs_rows;
  L.unobs_rows = delta + R.unobs_rows;
  L.S = delta + R.S;
  L.LS = delta + R.LS;
  L.UTTN = delta + R.UTTN;
  L.DI = delta + R.DI;
  L.UT = delta + R.UT;
  L.QT = delta + R.QT;
  L.UTF = delta + R.UTF;
  L.fold = delta + R.fold;
}

}  // namespace

int mra_plan(mra_handle* h, const double* obs, int want_predict, size_t* workspace_bytes) {
  if (!h || !obs || !workspace_bytes) return MRA_ERR_ARG;
  if (!h->has_structure) return fail(h, MRA_ERR_STATE, "mra_set_structure must be called first");
  int rc = plan_tree_impl(h, want_predict);
  if (rc) return rc;
  rc = plan_obs_impl(h, obs);
  if (rc) return rc;
  place_obs_part(h, h->total_A);
  h->lay.total = h->total_A + h->total_B;
  *workspace_bytes = h->lay.total;
  h->planned = true;
  h->split_arena = false;
  h->bound = h->uploaded = h->lik_done = h->pred_done = false;
  return MRA_OK;
}

int mra_plan_tree(mra_handle* h, int want_predict, size_t* tree_bytes) {
  if (!h || !tree_bytes) return MRA_ERR_ARG;
  if (!h->has_structure) return fail(h, MRA_ERR_STATE, "mra_set_structure must be called first");
  if (h->shard_level > 0) return fail(h, MRA_ERR_STATE, "the two-part plan is for unsharded handles");
  int rc = plan_tree_impl(h, want_predict);
  if (rc) return rc;
  *tree_bytes = h->total_A;
  return MRA_OK;
}

int mra_plan_obs(mra_handle* h, const double* obs, size_t* obs_bytes) {
  if (!h || !obs || !obs_bytes) return MRA_ERR_ARG;
  if (!h->tree_planned) return fail(h, MRA_ERR_STATE, "mra_plan_tree must be called first");
  int rc = plan_obs_impl(h, obs);
  if (rc) return rc;
  *obs_bytes = h->total_B;
  h->planned = true;
  return MRA_OK;
}

int mra_bind_workspace(mra_handle* h, void* dev_workspace, size_t bytes) {
  if (!h || !dev_workspace) return MRA_ERR_ARG;
  if (!h->planned) return fail(h, MRA_ERR_STATE, "mra_plan must be called first");
  if (bytes < h->lay.total) return fail(h, MRA_ERR_NOMEM, "workspace smaller than mra_plan reported");
  if (reinterpret_cast<uintptr_t>(dev_workspace) % 256) return fail(h, MRA_ERR_ARG, "workspace must be 256-byte aligned");
  DEVICE_SCOPE(h);
  drop_graph(h);
  h->ws = static_cast<char*>(dev_workspace);
  h->ws_bytes = bytes;
  int rc = configure_kernels(h);
  if (rc) return rc;
  h->bound = true;
  h->uploaded = h->lik_done = h->pred_done = false;
  return MRA_OK;
}

// tree-dependent tables (part A of the arena)
static int upload_tree_tables(mra_handle* h, cudaStream_t st) {
  const Layout& L = h->lay;
  const size_t N = (size_t)h->N;
  if (!h->knot_rows.empty())
    CU(cudaMemcpyAsync(h->ws + L.knot_rows, h->knot_rows.data(), sizeof(int) * h->knot_rows.size(), cudaMemcpyHostToDevice, st));
  CU(cudaMemcpyAsync(h->ws + L.perm, h->perm.data(), sizeof(int) * N, cudaMemcpyHostToDevice, st));
  for (size_t m = 0; m < h->internal_at.size(); ++m) {
    if (!h->internal_at[m].empty())
      CU(cudaMemcpyAsync(h->ws + L.lists + h->list_off[m], h->internal_at[m].data(),
                         sizeof(int) * h->internal_at[m].size(), cudaMemcpyHostToDevice, st));
    {
      // regular tiles run through the group list when it is in use: only the gathered tiles are needed on the device
      const bool grouped = h->use_groups && (h->r & 1) == 0;
      const size_t first = grouped ? (size_t)h->n_regular_tiles[m] : 0, cnt = h->ptiles_at[m].size() - first;
      if (cnt > 0)
        CU(cudaMemcpyAsync(h->ws + L.ptiles + h->ptiles_off[m] + sizeof(int4) * first, h->ptiles_at[m].data() + first,
                           sizeof(int4) * cnt, cudaMemcpyHostToDevice, st));
    }
    if (!h->pgroups_at[m].empty())
      CU(cudaMemcpyAsync(h->ws + L.pgroups + h->pgroups_off[m], h->pgroups_at[m].data(),
                         sizeof(int4) * h->pgroups_at[m].size(), cudaMemcpyHostToDevice, st));
  }
  if (!h->sroots.empty())
    CU(cudaMemcpyAsync(h->ws + L.lists + h->sroots_off, h->sroots.data(), sizeof(int) * h->sroots.size(),
                       cudaMemcpyHostToDevice, st));
  if (!h->gather_rows.empty())
    CU(cudaMemcpyAsync(h->ws + L.gather, h->gather_rows.data(), sizeof(int) * h->gather_rows.size(),
                       cudaMemcpyHostToDevice, st));
  if (!h->leaf_tiles.empty())
    CU(cudaMemcpyAsync(h->ws + L.ltiles, h->leaf_tiles.data(), sizeof(int4) * h->leaf_tiles.size(),
                       cudaMemcpyHostToDevice, st));
  if (!h->emit_chunks.empty())
    CU(cudaMemcpyAsync(h->ws + L.chunks, h->emit_chunks.data(), sizeof(int2) * h->emit_chunks.size(),
                       cudaMemcpyHostToDevice, st));
  if (!h->leaves.empty())
    CU(cudaMemcpyAsync(h->ws + L.lists + h->leaves_off, h->leaves.data(), sizeof(int) * h->leaves.size(),
                       cudaMemcpyHostToDevice, st));
  return MRA_OK;
}

// node table + observation-dependent tables (part B; the node table carries the leaves' offsets into it)
static int upload_obs_tables(mra_handle* h, cudaStream_t st) {
  const Layout& L = h->lay;
  CU(cudaMemcpyAsync(h->ws + L.nodes, h->nodes.data(), sizeof(NodeDev) * h->nodes.size(), cudaMemcpyHostToDevice, st));
  if (!h->obs_rows.empty())
    CU(cudaMemcpyAsync(h->ws + L.obs_rows, h->obs_rows.data(), sizeof(int) * h->obs_rows.size(), cudaMemcpyHostToDevice, st));
  if (!h->unobs_rows.empty())
    CU(cudaMemcpyAsync(h->ws + L.unobs_rows, h->unobs_rows.data(), sizeof(int) * h->unobs_rows.size(),
                       cudaMemcpyHostToDevice, st));
  if (!h->fold_items.empty())
    CU(cudaMemcpyAsync(h->ws + L.fold, h->fold_items.data(), sizeof(int4) * h->fold_items.size(),
                       cudaMemcpyHostToDevice, st));
  return MRA_OK;
}

// inputs into tree order; they come from host buffers (locs, obs) or are already on the device (dev_*)
static int permute_inputs(mra_handle* h, const double* locs, const double* obs, const double* dev_locs,
                          const double* dev_obs, cudaStream_t st) {
  const Layout& L = h->lay;
  const size_t N = (size_t)h->N;
  if (!dev_locs) {
    CU(cudaMemcpyAsync(h->ws + L.stage_locs, locs, sizeof(double) * N * h->dim, cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(h->ws + L.stage_obs, obs, sizeof(double) * N, cudaMemcpyHostToDevice, st));
    dev_locs = at<double>(h, L.stage_locs);
    dev_obs = at<double>(h, L.stage_obs);
  }
  k_permute_inputs<<<(unsigned)((N + 255) / 256), 256, 0, st>>>(
      dev_locs, dev_obs, at<int>(h, L.perm), (int)N, h->dim,
      at<double>(h, L.xs), at<double>(h, L.ys), at<double>(h, L.yobs), at<double>(h, L.xidx));
  CU(cudaGetLastError());
  return MRA_OK;
}

static int upload_impl(mra_handle* h, const double* locs, const double* obs, const double* dev_locs,
                       const double* dev_obs, void* stream) {
  if (!h->bound) return fail(h, MRA_ERR_STATE, "mra_bind_workspace must be called first");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  DEVICE_SCOPE(h);
  int rc = upload_obs_tables(h, st);
  if (rc) return rc;
  rc = upload_tree_tables(h, st);
  if (rc) return rc;
  rc = permute_inputs(h, locs, obs, dev_locs, dev_obs, st);
  if (rc) return rc;
  CU(cudaStreamSynchronize(st));   // host vectors may change after return
  h->uploaded = true;
  h->lik_done = h->pred_done = false;
  return MRA_OK;
}

// Two-part set-up (see mra_plan_tree): part A of the arena + the tree tables + the inputs; after it the prior pass may run
// (mra_set_cov / mra_set_nugget, mra_stream_begin_async, mra_stream_part_prior_async).
int mra_bind_tree(mra_handle* h, void* dev_workspace, size_t bytes, const double* dev_locs, const double* dev_obs,
                  void* stream) {
  if (!h || !dev_workspace || !dev_locs || !dev_obs) return MRA_ERR_ARG;
  if (!h->tree_planned) return fail(h, MRA_ERR_STATE, "mra_plan_tree must be called first");
  if (bytes < h->total_A) return fail(h, MRA_ERR_NOMEM, "workspace smaller than mra_plan_tree reported");
  if (reinterpret_cast<uintptr_t>(dev_workspace) % 256) return fail(h, MRA_ERR_ARG, "workspace must be 256-byte aligned");
  DEVICE_SCOPE(h);
  drop_graph(h);
  h->ws = static_cast<char*>(dev_workspace);
  h->ws_bytes = bytes;
  h->split_arena = true;
  int rc = configure_kernels(h);
  if (rc) return rc;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  // the node table as far as the prior pass reads it (the leaves' fields follow with mra_bind_obs)
  CU(cudaMemcpyAsync(h->ws + h->lay.nodes, h->nodes.data(), sizeof(NodeDev) * h->nodes.size(), cudaMemcpyHostToDevice, st));
  rc = upload_tree_tables(h, st);
  if (rc) return rc;
  rc = permute_inputs(h, nullptr, nullptr, dev_locs, dev_obs, st);
  if (rc) return rc;
  CU(cudaStreamSynchronize(st));
  h->tree_bound = true;
  h->uploaded = true;
  h->bound = false;
  h->lik_done = h->pred_done = false;
  return MRA_OK;
}

// Part B in an allocation of its own; its tables travel on the handle's copy stream (the prior pass may be running on
// `stream`, which only waits for them).  The handle is fully set up afterwards.
int mra_bind_obs(mra_handle* h, void* dev_workspace_obs, size_t bytes, void* stream) {
  if (!h || !dev_workspace_obs) return MRA_ERR_ARG;
  if (!h->tree_bound || !h->planned) return fail(h, MRA_ERR_STATE, "mra_bind_tree and mra_plan_obs must be called first");
  if (bytes < h->total_B) return fail(h, MRA_ERR_NOMEM, "workspace smaller than mra_plan_obs reported");
  if (reinterpret_cast<uintptr_t>(dev_workspace_obs) % 256) return fail(h, MRA_ERR_ARG, "workspace must be 256-byte aligned");
  DEVICE_SCOPE(h);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  place_obs_part(h, (size_t)(reinterpret_cast<uintptr_t>(dev_workspace_obs) - reinterpret_cast<uintptr_t>(h->ws)));
  h->lay.total = h->total_A;
  int rc = configure_kernels(h);      // the leaf kernels' shared-memory sizes depend on the observation counts
  if (rc) return rc;
  if (!h->copy_stream) {
    CU(cudaStreamCreateWithFlags(&h->copy_stream, cudaStreamNonBlocking));
    CU(cudaEventCreateWithFlags(&h->copy_event, cudaEventDisableTiming));
  }
  rc = upload_obs_tables(h, h->copy_stream);
  if (rc) return rc;
  CU(cudaEventRecord(h->copy_event, h->copy_stream));
  CU(cudaStreamWaitEvent(st, h->copy_event, 0));
  CU(cudaStreamSynchronize(h->copy_stream));   // host vectors may change after return
  h->bound = true;
  return MRA_OK;
}

int mra_upload_data(mra_handle* h, const double* locs, const double* obs, void* stream) {
  if (!h || !locs || !obs) return MRA_ERR_ARG;
  return upload_impl(h, locs, obs, nullptr, nullptr, stream);
}

int mra_upload_data_dev(mra_handle* h, const double* dev_locs, const double* dev_obs, void* stream) {
  if (!h || !dev_locs || !dev_obs) return MRA_ERR_ARG;
  return upload_impl(h, nullptr, nullptr, dev_locs, dev_obs, stream);
}

int mra_set_cov(mra_handle* h, int family, double length_scale, double sig) {
  if (!h) return MRA_ERR_ARG;
  if (family < MRA_COV_EXP || family > MRA_COV_GAUSSIAN) return fail(h, MRA_ERR_ARG, "unknown covariance family");
  if (!(length_scale > 0.0) || !(sig > 0.0)) return fail(h, MRA_ERR_ARG, "length scale and sig must be positive");
  if (h->cov.family == MRA_COV_DENSE) drop_graph(h);
  h->cov.dense = nullptr;
  h->cov.n_dense = 0;
  h->cov.family = family;
  h->cov.l = length_scale;
  h->cov.sig = sig;
  h->cov.c0 = h->cov.sig;
  if (family == MRA_COV_GAUSSIAN) h->cov.a = 1.0 / (2.0 * length_scale * length_scale);
  else h->cov.a = (family == MRA_COV_EXP ? 1.0 : family == MRA_COV_MATERN32 ? 1.7320508075688772 : 2.23606797749979) / length_scale;
  h->cov_set = true;
  h->lik_done = h->pred_done = false;
  return MRA_OK;
}

int mra_set_cov_dense(mra_handle* h, const double* dev_cov, int64_t n, double max_diag) {
  if (!h || !dev_cov) return MRA_ERR_ARG;
  if (!h->has_structure) return fail(h, MRA_ERR_STATE, "mra_set_structure must be called first");
  if (n != h->N) return fail(h, MRA_ERR_ARG, "the dense covariance matrix must be N x N");
  if (!(max_diag > 0.0)) return fail(h, MRA_ERR_ARG, "the covariance matrix needs a positive diagonal");
  if (h->cov.family != MRA_COV_DENSE) drop_graph(h);      // the kernels' coordinate arrays change
  h->cov.family = MRA_COV_DENSE;
  h->cov.l = h->cov.sig = h->cov.a = 1.0;
  h->cov.c0 = max_diag;
  h->cov.dense = dev_cov;
  h->cov.n_dense = n;
  h->cov_set = true;
  h->lik_done = h->pred_done = false;
  return MRA_OK;
}

int mra_set_nugget(mra_handle* h, double R) {
  if (!h) return MRA_ERR_ARG;
  if (!(R > 0.0)) return fail(h, MRA_ERR_ARG, "R must be a positive scalar");
  h->R = R;
  h->R_set = true;
  h->lik_done = h->pred_done = false;
  return MRA_OK;
}

static int ready_to_run(mra_handle* h) {
  if (!h->uploaded) return fail(h, MRA_ERR_STATE, "mra_upload_data must be called first");
  if (!h->cov_set || !h->R_set) return fail(h, MRA_ERR_STATE, "mra_set_cov and mra_set_nugget must be called first");
  return MRA_OK;
}

int mra_run_likelihood_async(mra_handle* h, void* stream) {
  if (!h) return MRA_ERR_ARG;
  if (h->shard_level > 0)
    return fail(h, MRA_ERR_STATE, "sharded handle: use mra_run_likelihood_local_async / _top_async");
  DEVICE_SCOPE(h);
  int rc = ready_to_run(h);
  if (rc) return rc;
  rc = launch_likelihood_local(h, static_cast<cudaStream_t>(stream), nullptr);
  if (rc) return rc;
  return launch_likelihood_top(h, static_cast<cudaStream_t>(stream), nullptr);
}

int mra_run_graph(mra_handle* h, void* stream, int with_predict) {
  if (!h) return MRA_ERR_ARG;
  if (h->shard_level > 0) return fail(h, MRA_ERR_STATE, "sharded handle: the exchange between the ranks cannot be captured");
  if (h->profiling) return fail(h, MRA_ERR_STATE, "per-kernel profiling and graph replay exclude each other");
  if (with_predict && !h->want_predict) return fail(h, MRA_ERR_STATE, "mra_plan was called with want_predict = 0");
  DEVICE_SCOPE(h);
  int rc = ready_to_run(h);
  if (rc) return rc;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (st == nullptr || st == cudaStreamLegacy) {
    // the legacy default stream cannot be captured: use an own (blocking) stream, which the legacy stream and every
    // other blocking stream order themselves against implicitly
    if (!h->gstream) CU(cudaStreamCreate(&h->gstream));
    st = h->gstream;
  }
  with_predict = with_predict ? 1 : 0;
  if (!h->gexec || h->graph_predict != with_predict) {
    drop_graph(h);
    h->capturing = true;
    cudaError_t e = cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal);
    if (e == cudaSuccess) {
      rc = launch_likelihood_local(h, st, nullptr);
      if (!rc) rc = launch_likelihood_top(h, st, nullptr);
      if (!rc && with_predict) rc = launch_predict(h, st, nullptr, nullptr);
      cudaGraph_t g = nullptr;
      cudaError_t e2 = cudaStreamEndCapture(st, &g);
      h->graph = g;
      if (e2 != cudaSuccess) e = e2;
    }
    h->capturing = false;
    if (rc) {
      drop_graph(h);
      return rc;
    }
    if (e != cudaSuccess) {
      drop_graph(h);
      return fail(h, MRA_ERR_CUDA, std::string("graph capture: ") + cudaGetErrorString(e));
    }
    h->graph_launches = h->launches;
    e = cudaGraphInstantiate(&h->gexec, h->graph, 0);
    if (e != cudaSuccess) {
      drop_graph(h);
      return fail(h, MRA_ERR_CUDA, std::string("cudaGraphInstantiate: ") + cudaGetErrorString(e));
    }
    h->graph_predict = with_predict;
  }
  rc = upload_params(h, st);
  if (rc) return rc;
  CU(cudaGraphLaunch(h->gexec, st));
  h->launches = h->graph_launches;
  h->warnings = 0;
  h->lik_done = true;
  h->pred_done = with_predict != 0;
  h->leafq_done = with_predict != 0;
  return MRA_OK;
}

int mra_set_shard(mra_handle* h, int32_t shard_level, const int8_t* node_role) {
  if (!h) return MRA_ERR_ARG;
  if (!h->has_structure) return fail(h, MRA_ERR_STATE, "mra_set_structure must be called first");
  wait_lists(h);
  if (shard_level == 0) {
    h->shard_level = 0;
    h->role.assign(h->n_nodes, 1);
  } else {
    if (!node_role) return MRA_ERR_ARG;
    if (shard_level < 1 || shard_level > h->depth) return fail(h, MRA_ERR_ARG, "shard level outside the tree");
    for (int n = 0; n < h->n_nodes; ++n) {
      const int lv = h->level[n], ro = node_role[n];
      const bool ok = lv < shard_level ? (ro == 2 || ro == 3)
                                       : (lv == shard_level ? (ro == 0 || ro == 1) : ro == node_role[h->parent[n]]);
      if (!ok) return fail(h, MRA_ERR_ARG, "inconsistent node roles");
    }
    h->shard_level = shard_level;
    h->role.assign(node_role, node_role + h->n_nodes);
  }
  drop_graph(h);
  h->slot_base = shard_level ? h->level_off[shard_level] : 0;
  h->n_slots = shard_level ? h->level_off[shard_level + 1] - h->level_off[shard_level] : 0;
  start_lists(h);
  h->planned = h->bound = h->uploaded = h->lik_done = h->pred_done = false;
  return MRA_OK;
}

int mra_summary_size(const mra_handle* h, int64_t* n_doubles) {
  if (!h || !n_doubles) return MRA_ERR_ARG;
  const int64_t W = (int64_t)h->shard_level * h->r + 1;
  *n_doubles = h->shard_level ? (int64_t)h->n_slots * (W * W + 1) : 0;
  return MRA_OK;
}

int mra_run_likelihood_local_async(mra_handle* h, void* stream, double* dev_summary) {
  if (!h) return MRA_ERR_ARG;
  if (h->shard_level > 0 && !dev_summary) return fail(h, MRA_ERR_ARG, "sharded handle needs a summary buffer");
  DEVICE_SCOPE(h);
  int rc = ready_to_run(h);
  if (rc) return rc;
  return launch_likelihood_local(h, static_cast<cudaStream_t>(stream), dev_summary);
}

int mra_run_likelihood_top_async(mra_handle* h, void* stream, const double* dev_summary) {
  if (!h) return MRA_ERR_ARG;
  if (h->shard_level > 0 && !dev_summary) return fail(h, MRA_ERR_ARG, "sharded handle needs a summary buffer");
  DEVICE_SCOPE(h);
  int rc = ready_to_run(h);
  if (rc) return rc;
  return launch_likelihood_top(h, static_cast<cudaStream_t>(stream), dev_summary);
}

int mra_stream_parts(const mra_handle* h, int32_t* n_parts) {
  if (!h || !n_parts) return MRA_ERR_ARG;
  *n_parts = h->shard_level <= 2 ? h->n_parts : 0;
  return MRA_OK;
}

int mra_stream_my_parts(const mra_handle* h, int32_t* mask) {
  if (!h || !mask) return MRA_ERR_ARG;
  *mask = h->shard_level <= 2 && h->n_parts > 0 ? h->my_parts : 0;
  return MRA_OK;
}

// uploads ranges of the caller's knot_rows (tree-order row ids) on the copy stream and makes `st` wait for them
static int upload_knot_ranges(mra_handle* h, cudaStream_t st, const int64_t* knot_rows, const std::vector<Range>& ranges,
                              int2 gather = make_int2(-1, 0)) {
  if (!knot_rows || ranges.empty()) return MRA_OK;
  if (!h->copy_stream) {
    CU(cudaStreamCreateWithFlags(&h->copy_stream, cudaStreamNonBlocking));
    CU(cudaEventCreateWithFlags(&h->copy_event, cudaEventDisableTiming));
  }
  for (const Range& rg : ranges) {
    for (int i = rg.begin; i < rg.begin + rg.count; ++i) {
      if (knot_rows[i] < 0 || knot_rows[i] >= h->N) return fail(h, MRA_ERR_ARG, "knot row out of range");
      h->knot_rows[i] = (int)knot_rows[i];
    }
    CU(cudaMemcpyAsync(h->ws + h->lay.knot_rows + sizeof(int) * (size_t)rg.begin, h->knot_rows.data() + rg.begin,
                       sizeof(int) * (size_t)rg.count, cudaMemcpyHostToDevice, h->copy_stream));
  }
  if (gather.x >= 0) {      // the gathered knot rows of a replicated top node follow its knots
    for (int i = 0; i < h->r; ++i) h->gather_rows[gather.y + i] = h->knot_rows[h->knot_off[gather.x] + i];
    CU(cudaMemcpyAsync(h->ws + h->lay.gather + sizeof(int) * (size_t)gather.y, h->gather_rows.data() + gather.y,
                       sizeof(int) * (size_t)h->r, cudaMemcpyHostToDevice, h->copy_stream));
  }
  CU(cudaEventRecord(h->copy_event, h->copy_stream));
  CU(cudaStreamWaitEvent(st, h->copy_event, 0));
  return MRA_OK;
}

int mra_stream_begin_async(mra_handle* h, void* stream, const int64_t* knot_rows) {
  if (!h) return MRA_ERR_ARG;
  if (h->shard_level > 2 || h->n_parts <= 0)
    return fail(h, MRA_ERR_STATE, "streamed evaluation needs an internal root and a handle that is unsharded or sharded at level 1 or 2");
  DEVICE_SCOPE(h);
  int rc = ready_to_run(h);
  if (rc) return rc;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  DevCtx c = make_ctx(h);
  h->lik_done = h->pred_done = false;
  rc = reset_pass(h, st);
  if (rc) return rc;
  rc = upload_knot_ranges(h, st, knot_rows, std::vector<Range>{Range{(int)h->knot_off[0], h->r}});
  if (rc) return rc;
  rc = prior_level(h, st, c, 0, Range{0, (int)h->internal_at[0].size()}, Range{0, h->n_regular_tiles[0]});
  if (rc) return rc;
  CU(cudaGetLastError());
  h->stream_open = true;
  h->stream_parts_done = 0;
  h->stream_prior_done = 0;
  return MRA_OK;
}

// the prior levels >= 1 of one part (needs part A of the arena only)
static int stream_part_prior(mra_handle* h, cudaStream_t st, int32_t part, const int64_t* knot_rows) {
  DevCtx c = make_ctx(h);
  int rc = upload_knot_ranges(h, st, knot_rows, h->part_knots[part], h->part_gather[part]);
  if (rc) return rc;
  const int nl = (int)h->internal_at.size();
  if (h->part_gtiles[part].count > 0) {     // sharded at level 2: the root's basis at the knots of the part's level-1 node
    const int4* tl = reinterpret_cast<const int4*>(h->ws + h->lay.ptiles + h->ptiles_off[0]) + h->part_gtiles[part].begin;
    MRA_FOR_VEC_NJ(h, LAUNCH("prior_tiles", k_prior_tiles<V_, J_><<<h->part_gtiles[part].count, NT, smem_prior(h->r), st>>>(c, tl, 0)));
  }
  for (int m = 1; m < nl; ++m) {
    rc = prior_level(h, st, c, m, h->part_nodes[m][part], h->part_tiles[m][part]);
    if (rc) return rc;
  }
  h->stream_prior_done |= 1 << part;
  return MRA_OK;
}

int mra_stream_part_prior_async(mra_handle* h, void* stream, int32_t part, const int64_t* knot_rows) {
  if (!h) return MRA_ERR_ARG;
  if (!h->stream_open) return fail(h, MRA_ERR_STATE, "mra_stream_begin_async must be called first");
  if (part < 0 || part >= h->n_parts) return fail(h, MRA_ERR_ARG, "part out of range");
  if ((h->stream_parts_done | h->stream_prior_done) & (1 << part)) return fail(h, MRA_ERR_STATE, "part already evaluated");
  if (!(h->my_parts & (1 << part))) return fail(h, MRA_ERR_ARG, "this part belongs to another rank");
  DEVICE_SCOPE(h);
  int rc = stream_part_prior(h, static_cast<cudaStream_t>(stream), part, knot_rows);
  if (rc) return rc;
  CU(cudaGetLastError());
  return MRA_OK;
}

int mra_stream_part_async(mra_handle* h, void* stream, int32_t part, const int64_t* knot_rows) {
  if (!h) return MRA_ERR_ARG;
  if (!h->stream_open) return fail(h, MRA_ERR_STATE, "mra_stream_begin_async must be called first");
  if (!h->bound) return fail(h, MRA_ERR_STATE, "the observation part of the plan is not bound yet (mra_bind_obs)");
  if (part < 0 || part >= h->n_parts) return fail(h, MRA_ERR_ARG, "part out of range");
  if (h->stream_parts_done & (1 << part)) return fail(h, MRA_ERR_STATE, "part already evaluated");
  if (!(h->my_parts & (1 << part))) return fail(h, MRA_ERR_ARG, "this part belongs to another rank");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  DEVICE_SCOPE(h);
  int rc = MRA_OK;
  if (!(h->stream_prior_done & (1 << part))) {
    rc = stream_part_prior(h, st, part, knot_rows);
    if (rc) return rc;
  }
  DevCtx c = make_ctx(h);
  const int nl = (int)h->internal_at.size();
  for (const Range& rg : h->part_leaves[part]) {
    rc = leaf_terms(h, st, c, rg);
    if (rc) return rc;
    if (h->want_predict) {
      rc = leaf_predict_terms(h, st, c, rg);
      if (rc) return rc;
    }
  }
  for (int m = nl - 1; m >= std::max(1, h->shard_level); --m) {
    rc = upward_level(h, st, c, m, h->part_nodes[m][part]);
    if (rc) return rc;
  }
  CU(cudaGetLastError());
  h->stream_parts_done |= 1 << part;
  return MRA_OK;
}

int mra_stream_end_async(mra_handle* h, void* stream) {
  if (!h) return MRA_ERR_ARG;
  if (h->shard_level > 0) return fail(h, MRA_ERR_STATE, "sharded handle: use mra_stream_end_local_async + mra_run_likelihood_top_async");
  if (!h->stream_open) return fail(h, MRA_ERR_STATE, "mra_stream_begin_async must be called first");
  if (h->stream_parts_done != h->my_parts) return fail(h, MRA_ERR_STATE, "not every part has been evaluated");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  DEVICE_SCOPE(h);
  DevCtx c = make_ctx(h);
  int rc = upward_level(h, st, c, 0);
  if (rc) return rc;
  h->stream_open = false;
  h->leafq_done = h->want_predict;
  return launch_likelihood_top(h, st, nullptr);     // shard_level == 0: only the final reduction
}

// After a streamed pass on a sharded handle the device has seen only the knots of the parts this rank ran; later
// plain passes (refit) factor every replicated top node, so the complete knot table is brought over once the
// build has ended (the caller's knot_rows is final by then).
int mra_stream_sync_knots(mra_handle* h, void* stream, const int64_t* knot_rows) {
  if (!h || !knot_rows) return MRA_ERR_ARG;
  if (!h->uploaded) return fail(h, MRA_ERR_STATE, "mra_upload_data must be called first");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  DEVICE_SCOPE(h);
  if (h->knot_rows.empty()) return MRA_OK;
  int rc = upload_knot_ranges(h, st, knot_rows, std::vector<Range>{Range{0, (int)h->knot_rows.size()}});
  if (rc) return rc;
  for (const int2& gn : h->gather_node) {
    rc = upload_knot_ranges(h, st, knot_rows, std::vector<Range>{Range{(int)h->knot_off[gn.x], h->r}}, gn);
    if (rc) return rc;
  }
  return MRA_OK;
}

int mra_stream_end_local_async(mra_handle* h, void* stream, double* dev_summary) {
  if (!h) return MRA_ERR_ARG;
  if (h->shard_level < 1 || h->shard_level > 2)
    return fail(h, MRA_ERR_STATE, "mra_stream_end_local_async is for handles sharded at level 1 or 2");
  if (!dev_summary) return fail(h, MRA_ERR_ARG, "sharded handle needs a summary buffer");
  if (!h->stream_open) return fail(h, MRA_ERR_STATE, "mra_stream_begin_async must be called first");
  if (h->stream_parts_done != h->my_parts) return fail(h, MRA_ERR_STATE, "not every part of this rank has been evaluated");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  DEVICE_SCOPE(h);
  DevCtx c = make_ctx(h);
  int rc = export_summaries(h, st, c, dev_summary);
  if (rc) return rc;
  CU(cudaGetLastError());
  h->stream_open = false;
  h->leafq_done = h->want_predict;
  return MRA_OK;
}

int mra_fetch_likelihood(mra_handle* h, void* stream, double out[2]) {
  if (!h || !out) return MRA_ERR_ARG;
  if (!h->lik_done) return fail(h, MRA_ERR_STATE, "no likelihood pass has been run");
  DEVICE_SCOPE(h);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  CU(cudaMemcpyAsync(out, h->ws + h->lay.out, 2 * sizeof(double), cudaMemcpyDeviceToHost, st));
  return check_status(h, st);
}

int mra_run_likelihood(mra_handle* h, void* stream, double out[2]) {
  int rc = mra_run_likelihood_async(h, stream);
  if (rc) return rc;
  return mra_fetch_likelihood(h, stream, out);
}

int mra_run_predict_dev(mra_handle* h, void* stream, double* dev_mean, double* dev_sd) {
  if (!h) return MRA_ERR_ARG;
  if (!h->lik_done) return fail(h, MRA_ERR_STATE, "mra_run_likelihood must be called first");
  if (!h->want_predict) return fail(h, MRA_ERR_STATE, "mra_plan was called with want_predict = 0");
  DEVICE_SCOPE(h);
  return launch_predict(h, static_cast<cudaStream_t>(stream), dev_mean, dev_sd);
}

// Sharded predict without zero-padded full-length outputs ("gather 16 B / location"): the tree-order row
// ranges this rank emits, its results packed contiguously ([mean of all my rows | var of all my rows]), and the
// un-permutation of complete tree-order arrays on whichever rank has gathered them.
int mra_predict_rows(const mra_handle* h, int64_t* ranges, int32_t max_ranges, int32_t* n_ranges) {
  if (!h || !n_ranges) return MRA_ERR_ARG;
  int n = 0;
  int64_t s0 = -1, e0 = -1;
  auto flush = [&]() {
    if (s0 < 0) return;
    if (ranges && n < max_ranges) {
      ranges[2 * n] = s0;
      ranges[2 * n + 1] = e0 - s0;
    }
    ++n;
  };
  for (const int2& ch : h->emit_chunks) {
    if (s0 >= 0 && ch.x == e0) {
      e0 += ch.y;
      continue;
    }
    flush();
    s0 = ch.x;
    e0 = (int64_t)ch.x + ch.y;
  }
  flush();
  *n_ranges = n;
  return (ranges && n > max_ranges) ? MRA_ERR_NOMEM : MRA_OK;
}

int mra_run_predict_pack_dev(mra_handle* h, void* stream, double* dev_pack, int64_t n_rows) {
  if (!h || !dev_pack) return MRA_ERR_ARG;
  if (!h->lik_done) return fail(h, MRA_ERR_STATE, "mra_run_likelihood must be called first");
  if (!h->want_predict) return fail(h, MRA_ERR_STATE, "mra_plan was called with want_predict = 0");
  DEVICE_SCOPE(h);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  int64_t mine = 0;
  for (const int2& ch : h->emit_chunks) mine += ch.y;
  if (mine != n_rows) return fail(h, MRA_ERR_ARG, "n_rows is not the number of rows this rank emits");
  // the predict kernels proper (no un-permutation: emit list left out)
  {
    std::vector<int2> none;
    none.swap(h->emit_chunks);
    int rc = launch_predict(h, st, at<double>(h, h->lay.out_mean), at<double>(h, h->lay.out_sd));
    none.swap(h->emit_chunks);
    if (rc) return rc;
  }
  int64_t off = 0, s0 = -1, e0 = -1;
  const double* mean = at<double>(h, h->lay.mean);
  const double* var = at<double>(h, h->lay.var);
  auto flush = [&]() -> cudaError_t {
    if (s0 < 0) return cudaSuccess;
    cudaError_t e = cudaMemcpyAsync(dev_pack + off, mean + s0, sizeof(double) * (e0 - s0), cudaMemcpyDeviceToDevice, st);
    if (e == cudaSuccess)
      e = cudaMemcpyAsync(dev_pack + n_rows + off, var + s0, sizeof(double) * (e0 - s0), cudaMemcpyDeviceToDevice, st);
    off += e0 - s0;
    return e;
  };
  for (const int2& ch : h->emit_chunks) {
    if (s0 >= 0 && ch.x == e0) {
      e0 += ch.y;
      continue;
    }
    CU(flush());
    s0 = ch.x;
    e0 = (int64_t)ch.x + ch.y;
  }
  CU(flush());
  return MRA_OK;
}

int mra_unpermute_tree_dev(mra_handle* h, void* stream, const double* dev_mean_tree, const double* dev_var_tree,
                           double* dev_mean, double* dev_sd) {
  if (!h || !dev_mean_tree || !dev_var_tree || !dev_mean || !dev_sd) return MRA_ERR_ARG;
  if (!h->uploaded) return fail(h, MRA_ERR_STATE, "mra_upload_data must be called first");
  DEVICE_SCOPE(h);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  DevCtx c = make_ctx(h);
  k_unpermute_all<<<(unsigned)((h->N + 255) / 256), 256, 0, st>>>(dev_mean_tree, dev_var_tree, at<int>(h, h->lay.perm),
                                                                  (int)h->N, dev_mean, dev_sd, c.P, c.status);
  CU(cudaGetLastError());
  return MRA_OK;
}

int mra_run_predict(mra_handle* h, void* stream, double* mean, double* sd) {
  if (!h || !mean || !sd) return MRA_ERR_ARG;
  DEVICE_SCOPE(h);
  HostTrace tr("run_predict");
  int rc = mra_run_predict_dev(h, stream, nullptr, nullptr);
  if (rc) return rc;
  tr.mark("launches");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (tr.on) {
    cudaStreamSynchronize(st);
    tr.mark("kernels done");
  }
  CU(cudaMemcpyAsync(mean, h->ws + h->lay.out_mean, sizeof(double) * h->N, cudaMemcpyDeviceToHost, st));
  CU(cudaMemcpyAsync(sd, h->ws + h->lay.out_sd, sizeof(double) * h->N, cudaMemcpyDeviceToHost, st));
  tr.mark("copies enqueued");
  rc = check_status(h, st);
  tr.mark("copies done");
  return rc;
}

int mra_set_diagnostics(mra_handle* h, int keep_posterior_basis) {
  if (!h) return MRA_ERR_ARG;
  h->keep_t0 = keep_posterior_basis != 0;
  h->pred_done = false;
  drop_graph(h);
  return MRA_OK;
}

int mra_last_warnings(const mra_handle* h, int32_t* flags) {
  if (!h || !flags) return MRA_ERR_ARG;
  *flags = h->warnings;
  return MRA_OK;
}

int mra_last_launches(const mra_handle* h, int64_t* n) {
  if (!h || !n) return MRA_ERR_ARG;
  *n = h->launches;
  return MRA_OK;
}

int mra_last_flops(const mra_handle* h, double* likelihood_flops, double* predict_flops) {
  if (!h) return MRA_ERR_ARG;
  if (likelihood_flops) *likelihood_flops = h->flops_lik;
  if (predict_flops) *predict_flops = h->flops_pred;
  return MRA_OK;
}

int mra_profile_enable(mra_handle* h, int on) {
  if (!h) return MRA_ERR_ARG;
  h->profiling = on != 0;
  for (auto& r : h->prof_recs) {
    h->event_pool.push_back(r.e0);
    h->event_pool.push_back(r.e1);
  }
  h->prof_recs.clear();
  for (auto& v : h->kms) v = 0.0;
  for (auto& v : h->klaunch) v = 0;
  return MRA_OK;
}

int mra_profile_read(mra_handle* h, char* buf, size_t buflen) {
  if (!h || !buf || buflen == 0) return MRA_ERR_ARG;
  DEVICE_SCOPE(h);
  CU(cudaDeviceSynchronize());
  for (auto& r : h->prof_recs) {
    float ms = 0.f;
    CU(cudaEventElapsedTime(&ms, r.e0, r.e1));
    h->kms[r.kid] += ms;
    h->klaunch[r.kid] += 1;
    h->event_pool.push_back(r.e0);
    h->event_pool.push_back(r.e1);
  }
  h->prof_recs.clear();
  std::string out;
  char line[256];
  for (size_t i = 0; i < h->kname.size(); ++i) {
    snprintf(line, sizeof line, "%s %.6f %lld %.6e %.6e\n", h->kname[i].c_str(), h->kms[i],
             (long long)h->klaunch[i], h->kflops[i], h->kbytes[i]);
    out += line;
  }
  if (out.size() + 1 > buflen) return fail(h, MRA_ERR_NOMEM, "profile buffer too small");
  memcpy(buf, out.c_str(), out.size() + 1);
  return MRA_OK;
}

int64_t mra_debug_fetch(mra_handle* h, const char* what, int node, double* out, int64_t max_doubles) {
  if (!h || !what || !out) return MRA_ERR_ARG;
  if (!h->uploaded) return fail(h, MRA_ERR_STATE, "nothing on the device yet");
  const Layout& L = h->lay;
  const std::string w(what);
  size_t off = 0;
  int64_t cnt = 0;
  const int r = h->r;
  if (w == "V") {
    off = L.V;
    cnt = h->N * h->ldv;
  } else if (w == "dnode") {
    off = L.dnode;
    cnt = h->n_nodes;
  } else if (w == "mean") {
    off = L.mean;
    cnt = h->N;
  } else if (w == "var") {
    off = L.var;
    cnt = h->N;
  } else {
    if (node < 0 || node >= h->n_nodes) return fail(h, MRA_ERR_ARG, "node out of range");
    const NodeDev& d = h->nodes[node];
    if (w == "A") { off = L.A + 8 * d.a_off; cnt = (int64_t)((d.level + 1) * r + 1) * d.lda; }
    else if (w == "GT") { off = L.GT + 8 * d.gt_off; cnt = (int64_t)d.W * r; }
    else if (w == "LPINV") { off = L.LPINV + 8 * d.lpinv_off; cnt = (int64_t)r * r; }
    else if (w == "LINV") { off = L.LINV + 8 * d.linv_off; cnt = (int64_t)r * r; }
    else if (w == "VK") { off = L.VK + 8 * d.vk_off; cnt = (int64_t)r * d.level * r; }
    else if (w == "S") { off = L.S + 8 * d.s_off; cnt = (int64_t)d.n_obs * d.ldo; }
    else if (w == "UT") { off = L.UT + 8 * d.ut_off; cnt = (int64_t)d.W * d.ldo; }
    else if (w == "QT") { off = L.QT + 8 * d.qt_off; cnt = (int64_t)d.row_count * d.ldo; }
    else return fail(h, MRA_ERR_ARG, "unknown buffer name");
    if (d.kind == KIND_INTERNAL && (w == "S" || w == "UT" || w == "QT")) return fail(h, MRA_ERR_ARG, "leaf buffer of an internal node");
    if (d.kind != KIND_INTERNAL && !(w == "S" || w == "UT" || w == "QT")) return fail(h, MRA_ERR_ARG, "internal buffer of a leaf");
  }
  cnt = std::min(cnt, max_doubles);
  DevGuard dev_guard_(h->device);
  if (dev_guard_.err != cudaSuccess) return MRA_ERR_CUDA;
  if (cudaDeviceSynchronize() != cudaSuccess) return fail(h, MRA_ERR_CUDA, "device synchronize failed");
  if (cudaMemcpy(out, h->ws + off, sizeof(double) * cnt, cudaMemcpyDeviceToHost) != cudaSuccess)
    return fail(h, MRA_ERR_CUDA, "debug copy failed");
  return cnt;
}

}  // extern "C"
