/* TEST INFRASTRUCTURE ONLY -- never linked into, loaded by or called from pymra_b200/.
 *
 * Ground truth for the parity tests: the exact Gaussian-process posterior under the covariance the MRA tree
 * implies, evaluated densely in EXTENDED precision (x87 long double, 64-bit mantissa; or __float128 when
 * compiled with -DUSE_QUAD).  SURVEY.md 0.8 / App. A: for any tree the reference's getLikelihood() / predict()
 * (pyMRA/MRATree.py:82-94, computed by the recursion of pyMRA/MRANode.py:378-520) equal
 *
 *     Sigma~ = sum over nodes n of  B_n kInv_n^{-1} B_n^T                       (B_n: MRANode.py:73-80, 378-391)
 *     lik    = logdet(Sigma~_oo + R I) + y_o^T (Sigma~_oo + R I)^{-1} y_o       (d + u, MRANode.py:450-468)
 *     mean   = Sigma~_{.o} (Sigma~_oo + R I)^{-1} y_o                           (MRANode.py:504-520)
 *     var    = diag(Sigma~) - diag(Sigma~_{.o} (Sigma~_oo + R I)^{-1} Sigma~_{o.})
 *
 * in exact arithmetic.  The tree (rows and knots of every node) is an INPUT: it comes from the oracle's node
 * record (oracle/mra_oracle.py, itself pinned bit-exact on the reference's trees), so this file contains no
 * partitioning logic and shares no arithmetic with either the oracle's recursion or the device's dual form:
 * agreement of all three is agreement of three different algorithms.
 *
 * With B_n = V_n L_n^T, L_n L_n^T = kInv_n:  Sigma~(i, j) = sum over the nodes n containing both i and j of
 * V_n(i) . V_n(j), where V_n = (C(X_n, K_n) - sum_{ancestors a} V_a[rows n] V_a[K_n]^T) L_n^{-T}.
 *
 * Covariance families follow pyMRA/MRATools.py:229-245 (Euclidean distance), :265-269 (ExpCovFun),
 * :289-293 (Matern32), :281-285 (Matern52), :297-301 (GaussianCovFun), evaluated in the extended type from the
 * same double-precision inputs; family 4 reads a dense N x N double matrix (`cov` given as np.matrix,
 * pyMRA/MRANode.py:73-75, 381-382).
 *
 * Build: oracle/build_truth.sh (gcc -O2 -fopenmp -shared -fPIC [-DUSE_QUAD]).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#ifdef USE_QUAD
#include <quadmath.h>
typedef __float128 real;
#define R_EXP expq
#define R_SQRT sqrtq
#define R_LOG logq
#define TRUTH_FN mra_dense_truth_q
#else
typedef long double real;
#define R_EXP expl
#define R_SQRT sqrtl
#define R_LOG logl
#define TRUTH_FN mra_dense_truth_l
#endif

typedef struct {
  int n_rows, n_knots, parent;
  const int32_t* rows;    /* global ids */
  const int32_t* knots;   /* global ids, subset of rows */
  int* idx_in_parent;     /* local index of every row inside the parent's row list */
  int* knot_local;        /* local index of every knot */
  real* V;                /* n_rows x n_knots */
  int* obs_local;         /* local rows that are observed */
  int* obs_index;         /* their index in the global observed list */
  int n_obs;
} Node;

static real cov_eval(int family, real l, real sig, const double* a, const double* b, int dim) {
  real d2 = 0;
  for (int k = 0; k < dim; ++k) {
    real t = (real)a[k] - (real)b[k];
    d2 += t * t;
  }
  const real D = R_SQRT(d2);
  switch (family) {
    case 0: return R_EXP(-D / l);
    case 1: { const real t = R_SQRT((real)3) * D / l; return sig * ((1 + t) * R_EXP(-t)); }
    case 2: { const real t = R_SQRT((real)5) * D / l; return sig * ((1 + t + ((real)5 / 3) * (D / l) * (D / l)) * R_EXP(-t)); }
    default: return sig * R_EXP(-d2 / (2 * l * l));
  }
}

/* in-place lower Cholesky of the n x n matrix a (row stride ld); returns 0 or 1 + index of the bad pivot.
 * Blocked right-looking, the trailing update parallel over rows. */
static int cholesky(real* a, int64_t n, int64_t ld) {
  const int64_t NB = 48;
  for (int64_t c0 = 0; c0 < n; c0 += NB) {
    const int64_t c1 = c0 + NB < n ? c0 + NB : n;
    for (int64_t j = c0; j < c1; ++j) {          /* diagonal block, unblocked */
      real d = a[j * ld + j];
      for (int64_t k = c0; k < j; ++k) d -= a[j * ld + k] * a[j * ld + k];
      if (!(d > 0)) return (int)(1 + j);
      d = R_SQRT(d);
      a[j * ld + j] = d;
      for (int64_t i = j + 1; i < c1; ++i) {
        real s = a[i * ld + j];
        for (int64_t k = c0; k < j; ++k) s -= a[i * ld + k] * a[j * ld + k];
        a[i * ld + j] = s / d;
      }
    }
#pragma omp parallel for schedule(static)
    for (int64_t i = c1; i < n; ++i) {            /* panel below the diagonal block */
      for (int64_t j = c0; j < c1; ++j) {
        real s = a[i * ld + j];
        for (int64_t k = c0; k < j; ++k) s -= a[i * ld + k] * a[j * ld + k];
        a[i * ld + j] = s / a[j * ld + j];
      }
    }
#pragma omp parallel for schedule(dynamic, 8)
    for (int64_t i = c1; i < n; ++i) {            /* trailing update, lower triangle */
      const real* li = a + i * ld + c0;
      for (int64_t j = c1; j <= i; ++j) {
        const real* lj = a + j * ld + c0;
        real s = 0;
        for (int64_t k = 0; k < c1 - c0; ++k) s += li[k] * lj[k];
        a[i * ld + j] -= s;
      }
    }
  }
  return 0;
}

/* Returns 0 on success, 1 + node index when that node's kInv is not positive definite, -1 when the dense
 * system is not, -2 on allocation failure. */
int TRUTH_FN(int N, int dim, const double* locs, const double* obs, double R_in, int family, double l_in,
             double sig_in, const double* dense /* family 4: N x N covariance matrix, else NULL */, int n_nodes, const int32_t* node_parent, const int64_t* rows_off, const int32_t* rows,
             const int64_t* knots_off, const int32_t* knots, double* out_lik, double* out_mean, double* out_sd) {
  const real l = (real)l_in, sig = (real)sig_in, R = (real)R_in;
  Node* nd = (Node*)calloc((size_t)n_nodes, sizeof(Node));
  int* deepest = (int*)malloc(sizeof(int) * (size_t)N);
  int* deep_loc = (int*)malloc(sizeof(int) * (size_t)N);
  int* scratch = (int*)malloc(sizeof(int) * (size_t)N);
  int* obs_of = (int*)malloc(sizeof(int) * (size_t)N);     /* global id -> index in the observed list, -1 */
  if (!nd || !deepest || !deep_loc || !scratch || !obs_of) return -2;
  int n_o = 0;
  for (int i = 0; i < N; ++i) obs_of[i] = isfinite(obs[i]) ? n_o++ : -1;
  for (int i = 0; i < N; ++i) deepest[i] = -1;
  int rc = 0;

  /* ---- per-node whitened basis V_n, parents before children */
  for (int n = 0; n < n_nodes && !rc; ++n) {
    Node* q = &nd[n];
    q->parent = node_parent[n];
    q->n_rows = (int)(rows_off[n + 1] - rows_off[n]);
    q->n_knots = (int)(knots_off[n + 1] - knots_off[n]);
    q->rows = rows + rows_off[n];
    q->knots = knots + knots_off[n];
    q->idx_in_parent = (int*)malloc(sizeof(int) * (size_t)(q->n_rows > 0 ? q->n_rows : 1));
    q->knot_local = (int*)malloc(sizeof(int) * (size_t)(q->n_knots > 0 ? q->n_knots : 1));
    q->V = (real*)malloc(sizeof(real) * (size_t)(q->n_rows > 0 ? q->n_rows : 1) * (size_t)(q->n_knots > 0 ? q->n_knots : 1));
    q->obs_local = (int*)malloc(sizeof(int) * (size_t)(q->n_rows > 0 ? q->n_rows : 1));
    q->obs_index = (int*)malloc(sizeof(int) * (size_t)(q->n_rows > 0 ? q->n_rows : 1));
    if (!q->idx_in_parent || !q->knot_local || !q->V || !q->obs_local || !q->obs_index) return -2;
    const int nr = q->n_rows, nk = q->n_knots;
    if (q->parent >= 0) {
      const Node* p = &nd[q->parent];
      for (int i = 0; i < p->n_rows; ++i) scratch[p->rows[i]] = i;
      for (int i = 0; i < nr; ++i) q->idx_in_parent[i] = scratch[q->rows[i]];
    }
    for (int i = 0; i < nr; ++i) {
      scratch[q->rows[i]] = i;
      deepest[q->rows[i]] = n;
      deep_loc[q->rows[i]] = i;
    }
    for (int k = 0; k < nk; ++k) q->knot_local[k] = scratch[q->knots[k]];
    q->n_obs = 0;
    for (int i = 0; i < nr; ++i)
      if (obs_of[q->rows[i]] >= 0) {
        q->obs_local[q->n_obs] = i;
        q->obs_index[q->n_obs] = obs_of[q->rows[i]];
        ++q->n_obs;
      }
    if (nk == 0) continue;
    /* B = C(X_n, K_n) - sum_a V_a[rows] V_a[K]^T */
    real* B = q->V;
#pragma omp parallel for schedule(static)
    for (int i = 0; i < nr; ++i)
      for (int k = 0; k < nk; ++k)
        B[(size_t)i * nk + k] = family == 4 ? (real)dense[(size_t)q->rows[i] * N + q->knots[k]]      /* MRANode.py:73-75, 381-382 */
                                            : cov_eval(family, l, sig, locs + (size_t)q->rows[i] * dim, locs + (size_t)q->knots[k] * dim, dim);
    {
      /* local indices of this node's rows / knots inside every ancestor, by chaining idx_in_parent */
      int* ri = (int*)malloc(sizeof(int) * (size_t)nr);
      int* ki = (int*)malloc(sizeof(int) * (size_t)nk);
      if (!ri || !ki) return -2;
      for (int i = 0; i < nr; ++i) ri[i] = i;
      const Node* cur = q;
      while (cur->parent >= 0) {
        const Node* a = &nd[cur->parent];
        for (int i = 0; i < nr; ++i) ri[i] = cur->idx_in_parent[ri[i]];
        for (int k = 0; k < nk; ++k) ki[k] = ri[q->knot_local[k]];
        const int ra = a->n_knots;
        const real* Va = a->V;
#pragma omp parallel for schedule(static)
        for (int i = 0; i < nr; ++i) {
          const real* vi = Va + (size_t)ri[i] * ra;
          for (int k = 0; k < nk; ++k) {
            const real* vk = Va + (size_t)ki[k] * ra;
            real s = 0;
            for (int t = 0; t < ra; ++t) s += vi[t] * vk[t];
            B[(size_t)i * nk + k] -= s;
          }
        }
        cur = a;
      }
      free(ri);
      free(ki);
    }
    /* kInv = B[K, :], L L^T = kInv, V = B L^{-T} */
    real* Lm = (real*)malloc(sizeof(real) * (size_t)nk * nk);
    if (!Lm) return -2;
    for (int a = 0; a < nk; ++a)
      for (int b = 0; b < nk; ++b) Lm[(size_t)a * nk + b] = B[(size_t)q->knot_local[a] * nk + b];
    if (cholesky(Lm, nk, nk)) {
      rc = 1 + n;
      free(Lm);
      break;
    }
#pragma omp parallel for schedule(static)
    for (int i = 0; i < nr; ++i) {
      real* v = B + (size_t)i * nk;          /* solve v L^T = b: forward substitution over columns */
      for (int k = 0; k < nk; ++k) {
        real s = v[k];
        for (int t = 0; t < k; ++t) s -= v[t] * Lm[(size_t)k * nk + t];
        v[k] = s / Lm[(size_t)k * nk + k];
      }
    }
    free(Lm);
  }

  real* S = NULL;
  real* z = NULL;
  if (!rc) {
    /* ---- S = Sigma~_oo + R I (lower triangle) */
    S = (real*)calloc((size_t)(n_o > 0 ? n_o : 1) * (size_t)(n_o > 0 ? n_o : 1), sizeof(real));
    z = (real*)malloc(sizeof(real) * (size_t)(n_o > 0 ? n_o : 1));
    if (!S || !z) return -2;
    for (int n = 0; n < n_nodes; ++n) {
      const Node* q = &nd[n];
      const int nk = q->n_knots, no = q->n_obs;
      if (!nk) continue;
#pragma omp parallel for schedule(dynamic, 16)
      for (int a = 0; a < no; ++a) {
        const real* va = q->V + (size_t)q->obs_local[a] * nk;
        const int64_t oa = q->obs_index[a];
        for (int b = 0; b <= a; ++b) {       /* obs_index ascends with the local row (rows ascend) or not: use max/min */
          const real* vb = q->V + (size_t)q->obs_local[b] * nk;
          const int64_t ob = q->obs_index[b];
          real s = 0;
          for (int t = 0; t < nk; ++t) s += va[t] * vb[t];
          if (oa >= ob) S[oa * n_o + ob] += s;
          else S[ob * n_o + oa] += s;
        }
      }
    }
    for (int64_t a = 0; a < n_o; ++a) S[a * n_o + a] += R;
    if (n_o > 0 && cholesky(S, n_o, n_o)) rc = -1;
  }
  if (!rc) {
    real logdet = 0, quad = 0;
    {
      int k = 0;
      for (int i = 0; i < N; ++i)
        if (obs_of[i] >= 0) z[k++] = (real)obs[i];
    }
    for (int64_t a = 0; a < n_o; ++a) {
      real s = z[a];
      for (int64_t b = 0; b < a; ++b) s -= S[a * n_o + b] * z[b];
      z[a] = s / S[a * n_o + a];
      quad += z[a] * z[a];
      logdet += R_LOG(S[a * n_o + a]);
    }
    *out_lik = (double)(2 * logdet + quad);

    /* ---- moments, blocks of BS points at a time: W = L^{-1} Sigma~_{o, block} */
    enum { BS = 16 };
    const int nblk = (N + BS - 1) / BS;
    int fail_alloc = 0;
#pragma omp parallel
    {
      real* W = (real*)malloc(sizeof(real) * (size_t)(n_o > 0 ? n_o : 1) * BS);
      if (!W) {
#pragma omp atomic write
        fail_alloc = 1;
      }
#pragma omp barrier
      if (!fail_alloc) {
#pragma omp for schedule(dynamic, 1)
        for (int blk = 0; blk < nblk; ++blk) {
          const int p0 = blk * BS, np = (N - p0 < BS) ? N - p0 : BS;
          memset(W, 0, sizeof(real) * (size_t)n_o * BS);
          real prior[BS];
          for (int p = 0; p < np; ++p) {
            prior[p] = 0;
            int n = deepest[p0 + p], loc = deep_loc[p0 + p];
            while (n >= 0) {
              const Node* q = &nd[n];
              const int nk = q->n_knots;
              const real* vp = q->V + (size_t)loc * nk;
              real s2 = 0;
              for (int t = 0; t < nk; ++t) s2 += vp[t] * vp[t];
              prior[p] += s2;
              for (int a = 0; a < q->n_obs; ++a) {
                const real* va = q->V + (size_t)q->obs_local[a] * nk;
                real s = 0;
                for (int t = 0; t < nk; ++t) s += va[t] * vp[t];
                W[(size_t)q->obs_index[a] * BS + p] += s;
              }
              loc = q->parent >= 0 ? q->idx_in_parent[loc] : 0;
              n = q->parent;
            }
          }
          real mean[BS], red[BS];
          for (int p = 0; p < BS; ++p) mean[p] = red[p] = 0;
          for (int64_t a = 0; a < n_o; ++a) {
            real acc[BS];
            for (int p = 0; p < BS; ++p) acc[p] = W[a * BS + p];
            const real* la = S + a * n_o;
            for (int64_t b = 0; b < a; ++b) {
              const real lab = la[b];
              const real* wb = W + b * BS;
              for (int p = 0; p < BS; ++p) acc[p] -= lab * wb[p];
            }
            const real dinv = 1 / la[a];
            for (int p = 0; p < BS; ++p) {
              const real w = acc[p] * dinv;
              W[a * BS + p] = w;
              mean[p] += w * z[a];
              red[p] += w * w;
            }
          }
          for (int p = 0; p < np; ++p) {
            out_mean[p0 + p] = (double)mean[p];
            const real var = prior[p] - red[p];
            out_sd[p0 + p] = (double)R_SQRT(var > 0 ? var : 0);
          }
        }
      }
      free(W);
    }
    if (fail_alloc) rc = -2;
  }
  for (int n = 0; n < n_nodes; ++n) {
    free(nd[n].idx_in_parent);
    free(nd[n].knot_local);
    free(nd[n].V);
    free(nd[n].obs_local);
    free(nd[n].obs_index);
  }
  free(nd);
  free(deepest);
  free(deep_loc);
  free(scratch);
  free(obs_of);
  free(S);
  free(z);
  return rc;
}
