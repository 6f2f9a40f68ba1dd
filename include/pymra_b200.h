/*
 * pymra_b200 C ABI -- the drop-in boundary for pyMRA's MRATree hot path on B200 (sm_100a).
 *
 * The reference (marcinjurek/pyMRA) is pure Python and has no FFI of its own; the
 * boundary it exposes is the class MRATree (pyMRA/MRATree.py:20-94).  The host-side
 * mirror pymra_b200.MRATree keeps that class' signature and calls the entry points
 * below through ctypes.  Each entry point names the reference code it replaces.
 *
 * Conventions: every function returns 0 on success and a negative mra_status on
 * failure, never throws across the ABI, and leaves a message retrievable with
 * mra_last_error().  The caller owns every buffer it passes in.  One handle per
 * device; a handle is not thread-safe.  Host pointers unless the name says "dev".
 * There is no CPU fallback: without a CUDA device mra_create fails.
 */
#ifndef PYMRA_B200_H
#define PYMRA_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct mra_handle mra_handle;

enum mra_status {
  MRA_OK = 0,
  MRA_ERR_ARG = -1,        /* bad argument / unsupported configuration */
  MRA_ERR_CUDA = -2,       /* CUDA runtime error */
  MRA_ERR_STATE = -3,      /* call order violated */
  MRA_ERR_NOT_SPD = -4,    /* a Cholesky factorisation met a non-positive pivot */
  MRA_ERR_NOMEM = -5,      /* workspace / output arrays too small */
  MRA_BUILD_UNSUPPORTED = 1 /* mra_build_structure_2d met a node outside its fast path (not an error) */
};

/* Conditions that do not fail a pass but are worth knowing (mra_last_warnings). */
enum mra_warning {
  MRA_WARN_NEGATIVE_VARIANCE = 2 /* a predictive variance fell below -1e-12 C(0) by cancellation and was clamped to 0;
                                    the reference's sum-of-squares form (MRANode.py:504-511) cannot go negative */
};

enum mra_cov_family {      /* pyMRA/MRATools.py */
  MRA_COV_EXP = 0,         /* ExpCovFun  :265-269   exp(-D/l)                          */
  MRA_COV_MATERN32 = 1,    /* Matern32   :289-293   sig*(1+sqrt(3)D/l)*exp(-sqrt(3)D/l) */
  MRA_COV_MATERN52 = 2,    /* Matern52   :281-285   sig*(1+sqrt(5)D/l+5D^2/(3l^2))*exp(-sqrt(5)D/l) */
  MRA_COV_GAUSSIAN = 3,    /* GaussianCovFun :297-301  sig*exp(-D^2/(2 l^2)) */
  MRA_COV_DENSE = 4        /* `cov` given as an N x N np.matrix (pyMRA/MRANode.py:73-75, 381-382); mra_set_cov_dense */
};

enum mra_node_kind { MRA_NODE_INTERNAL = 0, MRA_NODE_LEAF = 1, MRA_NODE_ORPHAN = 2 };

/* Flat tree description in "tree order" (pymra_b200/structure.py).  It carries exactly the
 * information the reference keeps in Node.inds / Node.kInds / Node.children
 * (pyMRA/MRANode.py:32, 38-45, 69, 98). Nodes are numbered level by level, children of a node
 * are consecutive, every node owns rows [row_start, row_start+row_count). */
typedef struct mra_structure {
  int64_t n_locs;                 /* N                                             */
  int32_t dim;                    /* 1 or 2                                        */
  int32_t r;                      /* knots per internal node (MRATree r)           */
  int32_t depth;                  /* deepest level present (root = 0)              */
  int32_t n_nodes;
  const int32_t *node_level;      /* [n_nodes]                                     */
  const int32_t *node_parent;     /* [n_nodes] -1 for the root                     */
  const int32_t *node_kind;       /* [n_nodes] mra_node_kind                       */
  const int64_t *node_row_start;  /* [n_nodes]                                     */
  const int64_t *node_row_count;  /* [n_nodes]                                     */
  const int32_t *node_child_start;/* [n_nodes] -1 if none                          */
  const int32_t *node_child_count;/* [n_nodes]                                     */
  const int64_t *node_knot_off;   /* [n_nodes] offset into knot_rows, -1 for leaves */
  const int64_t *knot_rows;       /* [n_knot_rows] tree-order row ids, r per internal node */
  int64_t n_knot_rows;
  const int32_t *level_off;       /* [depth+2] node id ranges per level            */
  const int64_t *perm;            /* [N] tree position -> caller's row index       */
} mra_structure;

/* Lifetime.  Replaces: construction/destruction of the Python MRATree object. */
int mra_create(mra_handle **out, int device);
int mra_destroy(mra_handle *h);
const char *mra_last_error(const mra_handle *h);
const char *mra_version(void);

/* Native host builder of the tree for 2-D inputs whose internal nodes all have > 100 rows and
 * > 100 knot candidates (random knots + quadrant splits, MRANode.py:36-38, 56-57, 191-193, 232-239).
 * Consumes/advances the legacy NumPy MT19937 state (key[624], pos) exactly like the reference's
 * np.random.choice calls in DFS pre-order, including the fork semantics at critDepth.  All output
 * arrays are caller-allocated (max_nodes entries; knot_rows/kinds_local max_nodes*r; perm n_locs).
 * Returns MRA_BUILD_UNSUPPORTED (RNG state untouched) if a node needs the KMeans/1-D paths. */
int mra_build_structure_2d(const double *locs, int64_t n_locs, int32_t r, int32_t M, int32_t J,
                           int32_t critDepth, uint32_t *mt_key, int32_t *mt_pos, int32_t max_nodes,
                           int32_t *n_nodes_out, int32_t *depth_out, int32_t *node_level,
                           int32_t *node_parent, int32_t *node_kind, int64_t *node_row_start,
                           int64_t *node_row_count, int32_t *node_child_start, int32_t *node_child_count,
                           int64_t *node_knot_off, int64_t *knot_rows, int32_t *kinds_local,
                           int64_t *n_knot_rows_out, int64_t *perm, int32_t *dfs_index);

/* Streaming variant of mra_build_structure_2d for regular trees (every node above level M has > 100 rows,
 * > 100 knot candidates and four non-empty quadrants; M >= 1, n_locs >= 65536): the build runs on its own
 * thread and reports progress, so the caller can start device work while the sequential legacy-RNG replay
 * (MRANode.py:191-193 in DFS pre-order) is still running.  For a regular 4-ary tree the level-by-level node
 * numbering is known in closed form (node k of level L has id (4^L-1)/3 + k, knot_off = id*r), hence
 *   event 0      : node arrays, perm and the ROOT's knot_rows / kinds_local slots are final
 *   event 1 + c  : knot_rows / kinds_local of every node in the subtree of the root's child c are final
 *   event 5      : the build has ended (mt_key / mt_pos advanced exactly like mra_build_structure_2d)
 * Slots that are not final yet read as 0.  mra_build_stream_wait blocks until the event has happened and
 * returns MRA_OK, or the build's failure status -- MRA_BUILD_UNSUPPORTED for a tree outside this path, in
 * which case mt_key / mt_pos are untouched and the output arrays are garbage.  All arrays stay owned by the
 * caller and must outlive mra_build_stream_finish, which joins the thread, frees the job and returns the
 * final status.  mra_build_stream_start itself returns MRA_BUILD_UNSUPPORTED (no job) for small inputs. */
typedef struct mra_build_job mra_build_job;
int mra_build_stream_start(const double *locs, int64_t n_locs, int32_t r, int32_t M, int32_t J,
                           int32_t critDepth, uint32_t *mt_key, int32_t *mt_pos, int32_t max_nodes,
                           int32_t *node_level, int32_t *node_parent, int32_t *node_kind,
                           int64_t *node_row_start, int64_t *node_row_count, int32_t *node_child_start,
                           int32_t *node_child_count, int64_t *node_knot_off, int64_t *knot_rows,
                           int32_t *kinds_local, int64_t *perm, int32_t *dfs_index, mra_build_job **job);
int mra_build_stream_wait(mra_build_job *job, int32_t event);
int mra_build_stream_finish(mra_build_job *job, int32_t *n_nodes_out, int32_t *depth_out,
                            int64_t *n_knot_rows_out);

/* Tree/knot/partition indexing computed on the host (bit-exact to MRANode.py:23-98,
 * 179-242, 289-340); copied into the handle. */
int mra_set_structure(mra_handle *h, const mra_structure *s);
/* Optional hint before mra_set_structure: mra_set_shard will follow, so the (unsharded) work lists need not be built
 * first -- mra_set_structure starts them as a background job otherwise. */
int mra_expect_shard(mra_handle *h);

/* Scans the NaN pattern of obs (MRANode.py:415, np.isfinite) and sizes the device workspace.
 * obs: [N] in the caller's order.  want_predict != 0 reserves the buffers predict() needs. */
int mra_plan(mra_handle *h, const double *obs, int want_predict, size_t *workspace_bytes);

/* Device arena of at least workspace_bytes (256-byte aligned), owned by the caller
 * (a torch uint8 tensor in the Python host). */
int mra_bind_workspace(mra_handle *h, void *dev_workspace, size_t bytes);

/* The same set-up in two parts (unsharded handles), for callers that want the device to start on the prior pass
 * (MRANode.py:73-80, 378-395: knots and locations only) while the host is still scanning the observations
 * (MRANode.py:415): the arena is split into a tree-dependent part (work lists, the basis slab, every per-internal-node
 * block) and an observation-dependent part (the leaves' observed / unobserved row lists and per-leaf blocks), two
 * allocations owned by the caller.
 *   mra_plan_tree  sizes the tree part;
 *   mra_bind_tree  binds it, uploads the tree tables and permutes the inputs (dev_locs / dev_obs as in
 *                  mra_upload_data_dev): mra_set_cov / mra_set_nugget, mra_stream_begin_async and
 *                  mra_stream_part_prior_async may follow;
 *   mra_plan_obs   scans the NaN pattern and sizes the observation part;
 *   mra_bind_obs   binds it and uploads its tables on a side stream (`stream` waits for them): the handle is then in the
 *                  state mra_plan + mra_bind_workspace + mra_upload_data leave it in. */
int mra_plan_tree(mra_handle *h, int want_predict, size_t *tree_bytes);
int mra_bind_tree(mra_handle *h, void *dev_workspace_tree, size_t bytes, const double *dev_locs, const double *dev_obs,
                  void *stream);
int mra_plan_obs(mra_handle *h, const double *obs, size_t *obs_bytes);
int mra_bind_obs(mra_handle *h, void *dev_workspace_obs, size_t bytes, void *stream);

/* locs: [N*dim] row-major (MRATree locs), obs: [N] with NaN = missing (MRATree obs).
 * Host buffers; copied H2D on `stream` and permuted to tree order on the device. */
int mra_upload_data(mra_handle *h, const double *locs, const double *obs, void *stream);
/* Same, with locs / obs already copied to the device by the caller (caller's row order; read on `stream`
 * before the call returns) -- lets the host->device copy start before the tree structure is known. */
int mra_upload_data_dev(mra_handle *h, const double *dev_locs, const double *dev_obs, void *stream);

/* Covariance descriptor introspected from the mt.ExpCovFun / mt.Matern32 closure, and the
 * nugget R (MRATree R / me_scale; must be a scalar, MRANode.py:85-88). */
int mra_set_cov(mra_handle *h, int family, double length_scale, double sig);
int mra_set_nugget(mra_handle *h, double R);
/* `cov` as a dense N x N matrix (row-major, the caller's row order) already on the device and owned by the caller
 * for as long as passes run (the reference slices it: cov[np.ix_(inds, kInds)], MRANode.py:73-75, 381-382).
 * max_diag: its largest diagonal entry.  Meant for N <= ~1e4 (the matrix itself is 8 N^2 bytes). */
int mra_set_cov_dense(mra_handle *h, const double *dev_cov, int64_t n, double max_diag);

/* Prior pass + leaf terms + upward pass (MRANode.py:378-395, 403-480).
 * out[0] = root.d, out[1] = root.u; getLikelihood() = out[0] + out[1] (MRATree.py:82-84). */
int mra_run_likelihood(mra_handle *h, void *stream, double out[2]);

/* Downward pass (MRANode.py:486-520).  mean, sd: [N] host buffers in the caller's order
 * (MRATree.py:90-94: root.mean, sqrt(root.var)).  Requires mra_run_likelihood first. */
int mra_run_predict(mra_handle *h, void *stream, double *mean, double *sd);

/* Same as the two calls above but without the final device->host copies / with results left
 * on the device (used by bench.py's device-resident timing).  dev_mean/dev_sd may be NULL. */
int mra_run_likelihood_async(mra_handle *h, void *stream);
int mra_run_predict_dev(mra_handle *h, void *stream, double *dev_mean, double *dev_sd);
int mra_fetch_likelihood(mra_handle *h, void *stream, double out[2]);

/* Sharded predict() without full-length zero-padded outputs (the outputs are gathered, 16 B per location).
 * mra_predict_rows: the tree-order row ranges [start, count] this rank emits (n_ranges; ranges may be NULL to query).
 * mra_run_predict_pack_dev: runs the downward pass (MRANode.py:486-520) for this rank's subtrees and packs its results
 *   into dev_pack = [mean of my rows | var of my rows] (2 * n_rows doubles, rows in range order); the caller gathers
 *   the packs of all ranks and lays them out in tree order.
 * mra_unpermute_tree_dev: complete tree-order mean / var arrays -> caller's row order, sd = sqrt(var)
 *   (MRANode.py:517-520, MRATree.py:90-94). */
int mra_predict_rows(const mra_handle *h, int64_t *ranges, int32_t max_ranges, int32_t *n_ranges);
int mra_run_predict_pack_dev(mra_handle *h, void *stream, double *dev_pack, int64_t n_rows);
int mra_unpermute_tree_dev(mra_handle *h, void *stream, const double *dev_mean_tree, const double *dev_var_tree,
                           double *dev_mean, double *dev_sd);

/* One whole evaluation on the frozen tree as ONE CUDA graph launch (the Nelder-Mead loop of README.md:96-104
 * re-evaluates the likelihood ~100 times on the same tree): the first call captures the launches of
 * mra_run_likelihood_async (and, with_predict != 0, of mra_run_predict_dev into the handle's own result buffers), later
 * calls replay them.  mra_set_cov / mra_set_nugget between calls take effect: the covariance descriptor and the nugget are
 * read by the kernels from a parameter block in device memory that is refreshed before every launch.  Anything that
 * changes the work lists or the buffers (structure, plan, workspace, sharding, diagnostics, dense <-> closure covariance)
 * drops the graph; the next call captures again.  Same results, bit for bit, as the plain calls.  Unsharded handles only. */
int mra_run_graph(mra_handle *h, void *stream, int with_predict);

/* Multi-GPU subtree sharding (one process / handle per GPU).  Replaces the reference's fork-per-child
 * subtree mode (pyMRA/MRANode.py:90-104, 114-115: mp.Process per child at res == critDepth, the child Node
 * pickled back through a pipe) by one rank per GPU: subtrees rooted at `shard_level` are owned by exactly one
 * rank, the levels above are replicated.  node_role[n_nodes]: 0 = another rank's subtree, 1 = mine,
 * 2 = replicated top, 3 = replicated top whose rows this rank emits in predict.  Call after
 * mra_set_structure and before mra_plan; shard_level == 0 switches sharding off.
 *
 * A sharded likelihood evaluation is three steps: mra_run_likelihood_local_async writes the summaries
 * (A~_c, d_c) of this rank's subtree roots (MRANode.py:474-480, the terms a child hands to its parent) into
 * the caller's zero-initialised device buffer of mra_summary_size doubles; the caller sum-reduces that buffer
 * over the ranks (NCCL all-reduce; slots are disjoint, so the sum is exact); mra_run_likelihood_top_async
 * finishes the replicated levels (MRANode.py:432-468).  mra_run_predict* then produces this rank's rows and
 * leaves the others zero, so outputs can be sum-reduced as well.  With shard_level == 0 the two calls run
 * back to back without a buffer (dev_summary may be NULL). */
int mra_set_shard(mra_handle *h, int32_t shard_level, const int8_t *node_role);
int mra_summary_size(const mra_handle *h, int64_t *n_doubles);
int mra_run_likelihood_local_async(mra_handle *h, void *stream, double *dev_summary);
int mra_run_likelihood_top_async(mra_handle *h, void *stream, const double *dev_summary);

/* Streamed likelihood evaluation on one GPU, for overlapping the device passes with a host build that is still
 * drawing knots (mra_build_stream_*).  The subtrees of the root's children are the "parts" (mra_stream_parts;
 * 0 = not available: a leaf root, or a handle sharded below level 2).  mra_stream_begin_async resets the pass and runs the
 * root's prior level (MRANode.py:378-395 for the root), which needs the root's knots only;
 * mra_stream_part_async(part) runs everything below the root for that subtree -- prior levels >= 1, leaf
 * terms (and, when predictions are planned, the leaf part of the predict pass), upward pass down to level 1
 * (MRANode.py:403-480) -- and mra_stream_end_async finishes the root (MRANode.py:432-468).  The result is
 * bit-identical to mra_run_likelihood_async: the same kernels run on the same nodes, only in more launches.
 * knot_rows: the caller's full knot_rows array (as in mra_structure) whose slots for the root / for the part
 * are final by now; they are copied to the device before the launches (NULL = already uploaded). */
int mra_stream_parts(const mra_handle *h, int32_t *n_parts);
int mra_stream_begin_async(mra_handle *h, void *stream, const int64_t *knot_rows);
int mra_stream_part_async(mra_handle *h, void *stream, int32_t part, const int64_t *knot_rows);
/* Only the prior levels >= 1 of the part (allowed after mra_bind_tree, before the observation part is bound);
 * a later mra_stream_part_async(part) runs the rest. */
int mra_stream_part_prior_async(mra_handle *h, void *stream, int32_t part, const int64_t *knot_rows);
int mra_stream_end_async(mra_handle *h, void *stream);
/* Sharded handles (shard level 1: 2-4 GPUs, the parts ARE the shards; level 2: up to 16 GPUs, a part is shared
 * by the ranks that own subtrees inside it) stream the same way: every rank calls mra_stream_begin_async (root
 * prior level on its own rows), mra_stream_part_async for the parts in its mra_stream_my_parts mask (others
 * are refused) and mra_stream_end_local_async, which exports the summaries like
 * mra_run_likelihood_local_async; the caller all-reduces them and calls mra_run_likelihood_top_async. */
int mra_stream_my_parts(const mra_handle *h, int32_t *mask);
int mra_stream_end_local_async(mra_handle *h, void *stream, double *dev_summary);
/* Brings the complete, final knot table to the device (a sharded streamed pass uploaded only the knots of the
 * parts this rank ran; plain passes such as a later re-fit factor every replicated top node). */
int mra_stream_sync_knots(mra_handle *h, void *stream, const int64_t *knot_rows);

/* Bit mask of mra_warning conditions seen by the passes since the last likelihood pass started; updated when a
 * pass's status is read back (mra_fetch_likelihood, mra_run_likelihood, mra_run_predict). */
int mra_last_warnings(const mra_handle *h, int32_t *flags);

/* Counters for bench.py: kernels launched by the last run_* call, and algorithmic FP64
 * flop of the last likelihood / predict pass as executed. */
int mra_last_launches(const mra_handle *h, int64_t *n);
int mra_last_flops(const mra_handle *h, double *likelihood_flops, double *predict_flops);

/* Per-kernel timing for bench.py's roofline block: when enabled every kernel launch is bracketed
 * by CUDA events on the launching stream.  mra_profile_read synchronises and writes one line per
 * kernel family: "name total_ms launches algorithmic_flops_per_pass algorithmic_bytes_per_pass". */
int mra_profile_enable(mra_handle *h, int on);
int mra_profile_read(mra_handle *h, char *buf, size_t buflen);

/* Opt-in diagnostics (pyMRA/MRATree.py:445-511 getBasisFunctionsMatrix needs per-node state the
 * reference frees, MRANode.py:108-110).  keep_posterior_basis != 0: the next predict pass leaves the folded
 * posterior basis t_j of EVERY level in the basis slab "V" (normally level 0's is consumed on the fly), from which
 * pymra_b200/diagnostics.py rebuilds the reference's BTil blocks with the LINV / LPINV blocks of mra_debug_fetch. */
int mra_set_diagnostics(mra_handle *h, int keep_posterior_basis);

/* Test hook: copies an internal device buffer to the host.
 * what: "V" (N x ldv, node ignored), "A", "GT", "LPINV", "LINV", "VK" (per node), "dnode" (all nodes).
 * Returns the number of doubles written (<= max_doubles) or a negative status. */
int64_t mra_debug_fetch(mra_handle *h, const char *what, int node, double *out, int64_t max_doubles);

#ifdef __cplusplus
}
#endif
#endif /* PYMRA_B200_H */
