/*
 * nnk_oracle.c -- CPU restatement of the nnmnkwii MLPG + DTW hot path.
 *
 * TEST INFRASTRUCTURE, NOT PRODUCT.  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load this library.  The shipped path is the CUDA
 * library in nnmnkwii_b200/csrc and fails loudly without it; nothing there links or calls this.
 *
 * Every function cites the reference file:line (relative to r9y9/nnmnkwii v0.1.3) whose
 * arithmetic it restates, in the same operation order, in float64.
 *
 * Parity status
 *   MLPG family   : PINNED.  Checked against the reference itself (oracle/_ref, built by
 *                   oracle/build_ref.sh) and against tests/golden/ (generated from the reference by
 *                   tests/golden/make_golden.py).
 *   DTW / FastDTW : PARITY UNPINNED.  The reference delegates to the third-party package
 *                   `fastdtw` (slaypni/fastdtw, unpinned in setup.py:139; call sites
 *                   preprocessing/alignment.py:50,138) which is not vendored, not installed and
 *                   not installable here.  orc_dtw / orc_fastdtw restate the published algorithm
 *                   (Salvador & Chan 2007, and the package's pure-Python fastdtw.py 0.3.x:
 *                   __fastdtw / __dtw / __reduce_by_half / __expand_window) from its documented
 *                   behaviour; they are cross-checked against a literal pure-Python restatement
 *                   (oracle/fastdtw_py.py) but not against the package itself.
 *
 * Build: gcc -O2 -ffp-contract=off -shared -fPIC nnk_oracle.c -o _build/libnnk_oracle.so -lm
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ORC_F32 0
#define ORC_F64 1

static inline double ld(const void* p, int dtype, size_t i) {
  return dtype == ORC_F32 ? (double)((const float*)p)[i] : ((const double*)p)[i];
}

/* 1 / variance evaluated in the INPUT dtype, then widened (paramgen/_mlpg.py:188, :259). */
static inline double recip_in_dtype(const void* p, int dtype, size_t i) {
  if (dtype == ORC_F32) {
    volatile float r = 1.0f / ((const float*)p)[i];
    return (double)r;
  }
  return 1.0 / ((const double*)p)[i];
}

static inline long lmin(long a, long b) { return a < b ? a : b; }
static inline long lmax(long a, long b) { return a > b ? a : b; }

/* ---------------------------------------------------------------------------------------------
 * Window bookkeeping.  windows = nw triples (l, u, coef[l+u+1]); coef arrays concatenated in
 * `coef`, coef_off[w] the start of window w.   W_w[t, t+k] = coef_w[l_w + k]  (build_win_mats,
 * paramgen/_mlpg.py:42-50 with BandMat.T, _bandmat/core.pyx:69-78).
 * ------------------------------------------------------------------------------------------- */
typedef struct {
  int nw;
  const int* wl;
  const int* wu;
  const double* coef;
  int coef_off[64];
  int sdw;           /* max_w (l_w + u_w)      : build_poe, _mlpg.py:72-73          */
  int max_win_width; /* max_w max(l_w, u_w)    : _mlpg.py:177                       */
} windows_t;

static int windows_init(windows_t* W, int nw, const int* wl, const int* wu, const double* coef) {
  if (nw < 1 || nw > 64) return -1;
  W->nw = nw; W->wl = wl; W->wu = wu; W->coef = coef;
  W->sdw = 0; W->max_win_width = 0;
  int off = 0;
  for (int w = 0; w < nw; ++w) {
    if (wl[w] < 0 || wu[w] < 0) return -1;
    W->coef_off[w] = off;
    off += wl[w] + wu[w] + 1;
    if (wl[w] + wu[w] > W->sdw) W->sdw = wl[w] + wu[w];
    int m = wl[w] > wu[w] ? wl[w] : wu[w];
    if (m > W->max_win_width) W->max_win_width = m;
  }
  return 0;
}

/* b += W_w^T v    : bm.dot_mv_plus_equals(win_mat.T, v, target=b), _bandmat/tensor.pyx:20-64
 * called from build_poe, _mlpg.py:84.  a_bm = win_mat.T has l_a = u_w, u_a = l_w, not transposed,
 * data rows = tiled coefficients, hence a_data[row_a, *] = coef[l_w + o_a].                    */
static void dot_mv_plus_equals_WT(const windows_t* W, int w, const double* v, double* target, long frames) {
  const long l_a = W->wu[w], u_a = W->wl[w];
  const double* c = W->coef + W->coef_off[w];
  for (long o_a = -u_a; o_a <= l_a; ++o_a) {
    const double a = c[u_a + o_a];
    for (long frame = lmax(0, o_a); frame < lmax(0, frames + lmin(0, o_a)); ++frame)
      target[frame] += a * v[frame - o_a];
  }
}

/* P += W_w^T diag(tau) W_w restricted to the (sdw, sdw) band:
 * bm.dot_mm_plus_equals(win_mat.T, win_mat, target_bm=prec, diag=tau), _bandmat/tensor.pyx:82-174
 * called from build_poe, _mlpg.py:85-87 (and mlpg_grad :266, unit_variance_mlpg_matrix :361,:366).
 * c_data is (2*sdw+1, frames), row-major; c_data[sdw + (i-j), j] = P[i][j]
 * (BandMat layout, _bandmat/core.pyx:80-87 + full.pyx band_c).                                  */
static void dot_mm_plus_equals_WTdW(const windows_t* W, int w, const double* tau, double* c_data, long frames) {
  const long l_a = W->wu[w], u_a = W->wl[w]; /* a = win_mat.T, not transposed */
  const long l_b = W->wl[w], u_b = W->wu[w]; /* b = win_mat,   transposed     */
  const long l_c = W->sdw, u_c = W->sdw;
  const double* c = W->coef + W->coef_off[w];
  for (long o_c = -lmin(u_c, u_a + u_b); o_c <= lmin(l_c, l_a + l_b); ++o_c) {
    for (long o_a = -lmin(u_a, l_b - o_c); o_a <= lmin(l_a, u_b + o_c); ++o_a) {
      const long o_b = o_c - o_a;
      const double a = c[u_a + o_a]; /* a_data[row_a = u_a + o_a, .]                */
      const double b = c[l_b - o_b]; /* b_data[row_b = l_b - o_b, .] (transposed_b) */
      const long row_c = u_c + o_c;
      const long d_c = -o_b;
      const long lo = lmax(0, lmax(-o_a, o_b));
      const long hi = lmax(0, frames + lmin(0, lmin(-o_a, o_b)));
      for (long frame = lo; frame < hi; ++frame)
        c_data[row_c * frames + frame + d_c] += a * b * tau[frame];
    }
  }
}

/* Lower banded Cholesky in place on half band `mat` ((depth+1, frames), mat[k, j] = A[j+k][j]).
 * _cholesky_banded(lower=True), _bandmat/linalg.pyx:77-95.  Returns 0, or frame+1 when the
 * leading minor is not positive definite (LinAlgError, linalg.pyx:79-82).                       */
static long cholesky_banded_lower(double* mat, long depth, long frames) {
  double v[64];
  for (long frame = 0; frame < frames; ++frame) {
    const double v0 = mat[frame];
    if (v0 <= 0.0) return frame + 1;
    const double iv0 = 1.0 / v0;
    const double siv0 = sqrt(iv0);
    for (long k = 0; k < depth; ++k) v[k] = mat[(k + 1) * frames + frame];
    mat[frame] = 1.0 / siv0;
    for (long k = 0; k < depth; ++k) mat[(k + 1) * frames + frame] = v[k] * siv0;
    for (long k = 0; k < lmin(depth, frames - frame - 1); ++k)
      for (long l = 0; l < depth - k; ++l)
        mat[l * frames + k + frame + 1] -= v[l + k] * v[k] * iv0;
  }
  return 0;
}

/* L z = b then L^T x = z with the lower factor from above:
 * cho_solve -> 2x _solve_triangular_banded, _bandmat/linalg.pyx:246-262, 145-174
 * (branches lower & !transposed, then lower & transposed).  Returns 0 or -(frame+1) on a zero
 * diagonal (linalg.pyx:170-173).                                                                */
static long cho_solve_lower(const double* chol, long depth, long frames, const double* b, double* x, double* tmp) {
  for (long pos = 0; pos < frames; ++pos) {
    const long frame = pos;
    double diff = b[frame];
    for (long k = 1; k < lmin(depth + 1, pos + 1); ++k) {
      const long framePrev = frame - k;
      diff -= chol[k * frames + framePrev] * tmp[framePrev];
    }
    const double denom = chol[frame];
    if (denom == 0.0) return -(frame + 1);
    tmp[frame] = diff / denom;
  }
  for (long pos = 0; pos < frames; ++pos) {
    const long frame = frames - 1 - pos;
    double diff = tmp[frame];
    for (long k = 1; k < lmin(depth + 1, pos + 1); ++k) {
      const long framePrev = frame + k;
      diff -= chol[k * frames + frame] * x[framePrev];
    }
    const double denom = chol[frame];
    if (denom == 0.0) return -(frame + 1);
    x[frame] = diff / denom;
  }
  return 0;
}

/* Edge rule for dynamic windows (_mlpg.py:190-193, :261-264): Python slices [:m] and [-m:];
 * note [-0:] is the WHOLE column, so m == 0 zeroes every frame of a dynamic window.            */
static void zero_edges(double* col, long T, long m) {
  if (m == 0) { for (long t = 0; t < T; ++t) col[t] = 0.0; return; }
  const long n = lmin(m, T);
  for (long t = 0; t < n; ++t) col[t] = 0.0;
  for (long t = T - n; t < T; ++t) col[t] = 0.0;
}

/* =============================================================================================
 * orc_mlpg: paramgen.mlpg, paramgen/_mlpg.py:166-199 (+ build_poe :53-89, bla.solveh
 * _bandmat/linalg.pyx:290-304).  means/vars row-major (T, D); var_is_1d != 0 => vars is (D,) and
 * is tiled over frames (:169-170).  out (T, D / nw) in the INPUT dtype (:183,:197).
 * Returns 0; >0 = 1-based frame of the non-positive pivot; <0 = argument / singular error.
 * ===========================================================================================*/
long orc_mlpg(const void* means, const void* vars, int dtype, int var_is_1d, long T, long D,
              int nw, const int* wl, const int* wu, const double* coef, void* out) {
  windows_t W;
  if (windows_init(&W, nw, wl, wu, coef)) return -1000000;
  const long sd = D / nw;
  const long sdw = W.sdw;
  double* mu = (double*)malloc(sizeof(double) * (size_t)(T * nw + 1));
  double* prec = (double*)malloc(sizeof(double) * (size_t)(T * nw + 1));
  double* bs = (double*)malloc(sizeof(double) * (size_t)(T + 1));
  double* b = (double*)malloc(sizeof(double) * (size_t)(T + 1));
  double* P = (double*)malloc(sizeof(double) * (size_t)((2 * sdw + 1) * T + 1));
  double* x = (double*)malloc(sizeof(double) * (size_t)(T + 1));
  double* tmp = (double*)malloc(sizeof(double) * (size_t)(T + 1));
  long status = 0;
  for (long d = 0; d < sd && status == 0; ++d) {
    for (int w = 0; w < nw; ++w) { /* column-major workspaces mu[w][t], prec[w][t] */
      const long col = (long)w * sd + d;
      for (long t = 0; t < T; ++t) {
        mu[(long)w * T + t] = ld(means, dtype, (size_t)(t * D + col));
        prec[(long)w * T + t] = recip_in_dtype(vars, dtype, (size_t)(var_is_1d ? col : t * D + col));
      }
      if (w != 0) zero_edges(prec + (long)w * T, T, W.max_win_width);
    }
    memset(b, 0, sizeof(double) * (size_t)T);
    memset(P, 0, sizeof(double) * (size_t)((2 * sdw + 1) * T));
    for (int w = 0; w < nw; ++w) {
      for (long t = 0; t < T; ++t) bs[t] = prec[(long)w * T + t] * mu[(long)w * T + t]; /* :195 */
      dot_mv_plus_equals_WT(&W, w, bs, b, T);
      dot_mm_plus_equals_WTdW(&W, w, prec + (long)w * T, P, T);
    }
    /* cholesky(P, lower=True): half band = rows sdw .. 2*sdw of P.data (linalg.pyx:218-226) */
    double* half = P + sdw * T;
    long r = cholesky_banded_lower(half, sdw, T);
    if (r) { status = r; break; }
    r = cho_solve_lower(half, sdw, T, b, x, tmp);
    if (r) { status = r; break; }
    for (long t = 0; t < T; ++t) {
      if (dtype == ORC_F32) ((float*)out)[t * sd + d] = (float)x[t];
      else ((double*)out)[t * sd + d] = x[t];
    }
  }
  free(mu); free(prec); free(bs); free(b); free(P); free(x); free(tmp);
  return status;
}

/* =============================================================================================
 * orc_mlpg_grad: paramgen.mlpg_grad, paramgen/_mlpg.py:242-281.
 * The reference forms r = W_w^T diag(tau_w) (dense T x T) and solves R G = r with LAPACK
 * (solve_banded, :275), then takes o^T G (:279).  Since R is symmetric this is
 *      grads[:, w*sd+d] = tau_w * (W_w R^{-1} o_d),
 * which is what is evaluated here with the same Cholesky as orc_mlpg (O(T) instead of O(T^2)).
 * It is an algebraic restatement, not an operation-order one; pinned against the reference's
 * output in tests (float32 result, so agreement is to ~1e-7 relative).  out (T, D) float32 always
 * (:248).  grad_output (T, sd) in `go_dtype`.
 * ===========================================================================================*/
long orc_mlpg_grad(const void* vars, int dtype, int var_is_1d, long T, long D, int nw, const int* wl,
                   const int* wu, const double* coef, const void* grad_output, int go_dtype, float* out) {
  windows_t W;
  if (windows_init(&W, nw, wl, wu, coef)) return -1000000;
  const long sd = D / nw;
  const long sdw = W.sdw;
  double* prec = (double*)malloc(sizeof(double) * (size_t)(T * nw + 1));
  double* o = (double*)malloc(sizeof(double) * (size_t)(T + 1));
  double* P = (double*)malloc(sizeof(double) * (size_t)((2 * sdw + 1) * T + 1));
  double* x = (double*)malloc(sizeof(double) * (size_t)(T + 1));
  double* tmp = (double*)malloc(sizeof(double) * (size_t)(T + 1));
  long status = 0;
  memset(out, 0, sizeof(float) * (size_t)(T * D));
  for (long d = 0; d < sd && status == 0; ++d) {
    memset(P, 0, sizeof(double) * (size_t)((2 * sdw + 1) * T));
    for (int w = 0; w < nw; ++w) {
      const long col = (long)w * sd + d;
      for (long t = 0; t < T; ++t)
        prec[(long)w * T + t] = recip_in_dtype(vars, dtype, (size_t)(var_is_1d ? col : t * D + col));
      if (w != 0) zero_edges(prec + (long)w * T, T, W.max_win_width);
      dot_mm_plus_equals_WTdW(&W, w, prec + (long)w * T, P, T);
    }
    for (long t = 0; t < T; ++t) o[t] = ld(grad_output, go_dtype, (size_t)(t * sd + d));
    double* half = P + sdw * T;
    long r = cholesky_banded_lower(half, sdw, T);
    if (r) { status = r; break; }
    r = cho_solve_lower(half, sdw, T, o, x, tmp);
    if (r) { status = r; break; }
    for (int w = 0; w < nw; ++w) {
      const double* c = W.coef + W.coef_off[w];
      const long l = wl[w], u = wu[w];
      for (long t = 0; t < T; ++t) {
        double s = 0.0; /* (W_w x)[t] = sum_k coef[l+k] x[t+k] */
        for (long k = -l; k <= u; ++k)
          if (t + k >= 0 && t + k < T) s += c[l + k] * x[t + k];
        out[t * D + (long)w * sd + d] = (float)(prec[(long)w * T + t] * s);
      }
    }
  }
  free(prec); free(o); free(P); free(x); free(tmp);
  return status;
}

/* cholesky_inv_banded(L_full, w): util/_linalg.pyx:45-71.  R is the dense lower Cholesky factor
 * (T x T row-major); Pout receives the dense inverse of R R^T.  Statement-for-statement.        */
static void cholesky_inv_banded_dense(const double* R, long T, long w, double* Pout) {
  double* g = (double*)calloc((size_t)(T * T), sizeof(double));
  double* hold = (double*)calloc((size_t)T, sizeof(double));
  memset(Pout, 0, sizeof(double) * (size_t)(T * T));
  g[0] = 1.0 / R[0];
  for (long t = 1; t < T; ++t) {
    for (long i = 0; i < T; ++i) hold[i] *= 0.0;
    for (long j = 1; j < w; ++j)
      if (t - j >= 0 && R[t * T + (t - j)] != 0.0)
        for (long i = 0; i <= t; ++i) hold[i] += R[t * T + (t - j)] * g[(t - j) * T + i];
    hold[t] -= 1.0;
    for (long i = 0; i <= t; ++i) g[t * T + i] = -hold[i] / R[t * T + t];
  }
  for (long i = 0; i < T; ++i) Pout[(T - 1) * T + i] = g[(T - 1) * T + i] / R[(T - 1) * T + (T - 1)];
  /* R = R.T : below R^T[t, t+j] = R[t+j, t] */
  for (long t = T - 2; t >= 0; --t) {
    for (long i = 0; i < T; ++i) hold[i] *= 0.0;
    for (long j = 1; j < w; ++j)
      if (t + j < T && R[(t + j) * T + t] != 0.0)
        for (long i = 0; i < T; ++i) hold[i] += R[(t + j) * T + t] * Pout[(t + j) * T + i];
    for (long i = 0; i < T; ++i) Pout[t * T + i] = (g[t * T + i] - hold[i]) / R[t * T + t];
  }
  free(g); free(hold);
}

/* =============================================================================================
 * orc_unit_variance_mlpg_matrix: paramgen.unit_variance_mlpg_matrix, paramgen/_mlpg.py:346-373.
 * Rout (T, nw*T) float32.   P = sum_w Wtilde_w^T W_w with Wtilde_w = diag(edge mask) W_w for
 * w != 0 (:350-367); chol (:369); dense inverse via cholesky_inv_banded (:370);
 * R = Pinv . full_window_mat(mod_win_mats)^T (:372-373; full_window_mat = mlpg_helper.pyx:10-32).
 * The final product is a BLAS dgemm in the reference (order of summation unspecified); here each
 * entry is summed over the <= l+u+1 non-zeros of the window row in increasing column order.
 * ===========================================================================================*/
long orc_unit_variance_mlpg_matrix(long T, int nw, const int* wl, const int* wu, const double* coef, float* Rout) {
  windows_t W;
  if (windows_init(&W, nw, wl, wu, coef)) return -1000000;
  const long sdw = W.sdw, m = W.max_win_width;
  double* mask = (double*)calloc((size_t)T + 1, sizeof(double));
  double* ones = (double*)malloc(sizeof(double) * ((size_t)T + 1));
  /* precisions.data[:, m:-m] += 1.0 (:354): m == 0 gives the empty slice [0:0] -> all zeros */
  if (m > 0) for (long t = m; t < T - m; ++t) mask[t] = 1.0;
  for (long t = 0; t < T; ++t) ones[t] = 1.0;
  double* P = (double*)calloc((size_t)((2 * sdw + 1) * T) + 1, sizeof(double));
  for (int w = 0; w < nw; ++w) dot_mm_plus_equals_WTdW(&W, w, w != 0 ? mask : ones, P, T);
  double* half = P + sdw * T;
  long r = cholesky_banded_lower(half, sdw, T);
  if (r) { free(mask); free(ones); free(P); return r; }
  double* Lf = (double*)calloc((size_t)(T * T) + 1, sizeof(double));
  for (long k = 0; k <= sdw; ++k)
    for (long j = 0; j + k < T; ++j) Lf[(j + k) * T + j] = half[k * T + j];
  double* Pinv = (double*)malloc(sizeof(double) * ((size_t)(T * T) + 1));
  cholesky_inv_banded_dense(Lf, T, sdw + 1, Pinv);
  for (long i = 0; i < T; ++i)
    for (int w = 0; w < nw; ++w) {
      const double* c = W.coef + W.coef_off[w];
      const long l = wl[w], u = wu[w];
      for (long rr = 0; rr < T; ++rr) {
        const double mk = (w != 0) ? mask[rr] : 1.0;
        double s = 0.0; /* sum_j Pinv[i][j] * Wtilde_w[rr][j] */
        for (long k = -l; k <= u; ++k)
          if (rr + k >= 0 && rr + k < T) s += Pinv[i * T + rr + k] * (mk * c[l + k] * 1.0);
        Rout[i * ((long)nw * T) + (long)w * T + rr] = (float)s;
      }
    }
  free(mask); free(ones); free(P); free(Lf); free(Pinv);
  return 0;
}

/* =============================================================================================
 * DTW local cost.  As DTW `dist` the reference calls a Python callable per cell on two float64
 * rows (fastdtw coerces inputs with np.asanyarray(..., dtype='float')):
 *   kind 1: metrics.melcd(x, y)  = _logdb_const * float(sqrt(((x-y)*(x-y)).sum(-1)))
 *           (metrics/__init__.py:5, :52-57), the sum being numpy's pairwise add-reduce
 *           (numpy/core/src/umath/loops_utils.h.src DOUBLE_pairwise_sum: n<8 serial; n<=128 eight
 *           strided accumulators then ((r0+r1)+(r2+r3))+((r4+r5)+(r6+r7)) then the tail serially;
 *           n>128 split at n/2 rounded down to a multiple of 8).  Restated exactly below.
 *   kind 0: default lambda x, y: norm(x - y) (alignment.py:35) -> sqrt(dot(z, z)) via BLAS ddot,
 *           whose summation order is implementation-defined; restated with the SAME pairwise
 *           order as kind 1 (documented deviation: costs may differ from a given BLAS by <= 1ulp
 *           scale, paths only on exact near-ties).
 * ===========================================================================================*/
static double pairwise_sum(const double* a, long n) {
  if (n < 8) {
    double res = -0.0;
    for (long i = 0; i < n; ++i) res += a[i];
    return res;
  } else if (n <= 128) {
    double r[8];
    long i;
    for (int j = 0; j < 8; ++j) r[j] = a[j];
    for (i = 8; i < n - (n % 8); i += 8)
      for (int j = 0; j < 8; ++j) r[j] += a[i + j];
    double res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]));
    for (; i < n; ++i) res += a[i];
    return res;
  } else {
    long n2 = n / 2;
    n2 -= n2 % 8;
    return pairwise_sum(a, n2) + pairwise_sum(a + n2, n - n2);
  }
}

#define ORC_LOGDB_CONST_EXPR (10.0 / log(10.0) * sqrt(2.0))

double orc_logdb_const(void) { return ORC_LOGDB_CONST_EXPR; }

double orc_cost(const double* x, const double* y, long D, int kind) {
  double zz[4096];
  double* buf = D <= 4096 ? zz : (double*)malloc(sizeof(double) * (size_t)D);
  for (long k = 0; k < D; ++k) { const double z = x[k] - y[k]; buf[k] = z * z; }
  const double r = sqrt(pairwise_sum(buf, D));
  if (buf != zz) free(buf);
  return kind == 1 ? ORC_LOGDB_CONST_EXPR * r : r;
}

/* Windowed DTW.  fastdtw.py __dtw: for (i, j) in window (row-major, 1-based after the +1 shift):
 *     dt = dist(x[i-1], y[j-1])
 *     D[i, j] = min((D[i-1, j] + dt, i-1, j), (D[i, j-1] + dt, i, j-1), (D[i-1, j-1] + dt, i-1, j-1),
 *                   key=lambda a: a[0])            # first minimum wins: up, left, diagonal
 * with D[0, 0] = 0 and +inf outside the window; backtrack from (len_x, len_y) to (0, 0).
 * The window is given per row as [lo[i], hi[i]) (0-based columns); lo == NULL => full rectangle.
 * path_i / path_j receive the 0-based path from (0, 0) to (Tx-1, Ty-1); returns its length, or <0.
 * Cells: *cells (if non-NULL) += number of DP cells evaluated.                                   */
long orc_dtw_window(const double* x, const double* y, long Tx, long Ty, long D, int kind, const int32_t* lo,
                    const int32_t* hi, int32_t* path_i, int32_t* path_j, double* dist_out, int64_t* cells) {
  if (Tx <= 0 || Ty <= 0) return -1;
  int64_t* off = (int64_t*)malloc(sizeof(int64_t) * (size_t)(Tx + 1));
  off[0] = 0;
  for (long i = 0; i < Tx; ++i) {
    const long a = lo ? lo[i] : 0, b = lo ? hi[i] : Ty;
    off[i + 1] = off[i] + (b > a ? b - a : 0);
  }
  const int64_t ncell = off[Tx];
  double* Dm = (double*)malloc(sizeof(double) * (size_t)(ncell + 1));
  uint8_t* bp = (uint8_t*)malloc((size_t)(ncell + 1));
  const double INF = INFINITY;
#define DGET(i, j) (((i) < 0 || (j) < 0) ? (((i) == -1 && (j) == -1) ? 0.0 : INF) \
                    : (((j) < (lo ? lo[i] : 0) || (j) >= (lo ? hi[i] : Ty)) ? INF : Dm[off[i] + (j) - (lo ? lo[i] : 0)]))
  for (long i = 0; i < Tx; ++i) {
    const long a = lo ? lo[i] : 0, b = lo ? hi[i] : Ty;
    for (long j = a; j < b; ++j) {
      const double dt = orc_cost(x + i * D, y + j * D, D, kind);
      const double up = DGET(i - 1, j) + dt;       /* (i-1, j)   */
      const double left = DGET(i, j - 1) + dt;     /* (i, j-1)   */
      const double diag = DGET(i - 1, j - 1) + dt; /* (i-1, j-1) */
      double best = up; uint8_t dir = 0;
      if (left < best) { best = left; dir = 1; }
      if (diag < best) { best = diag; dir = 2; }
      Dm[off[i] + j - a] = best;
      bp[off[i] + j - a] = dir;
    }
  }
  if (cells) *cells += ncell;
  long n = 0, i = Tx - 1, j = Ty - 1;
  long status = 0;
  const long cap = Tx + Ty;
  while (!(i == -1 && j == -1)) {
    if (i < 0 || j < 0 || n >= cap || j < (lo ? lo[i] : 0) || j >= (lo ? hi[i] : Ty)) { status = -2; break; }
    path_i[n] = (int32_t)i; path_j[n] = (int32_t)j; ++n;
    const uint8_t dir = bp[off[i] + j - (lo ? lo[i] : 0)];
    if (dir == 0) i -= 1; else if (dir == 1) j -= 1; else { i -= 1; j -= 1; }
  }
#undef DGET
  if (status == 0) {
    for (long a = 0, b = n - 1; a < b; ++a, --b) {
      int32_t t = path_i[a]; path_i[a] = path_i[b]; path_i[b] = t;
      t = path_j[a]; path_j[a] = path_j[b]; path_j[b] = t;
    }
    if (dist_out) *dist_out = Dm[off[Tx - 1] + (Ty - 1) - (lo ? lo[Tx - 1] : 0)];
  }
  free(off); free(Dm); free(bp);
  return status ? status : n;
}

/* fastdtw.py __expand_window(path, len_x, len_y, radius) in closed form.  The literal version
 * dilates the coarse path by `radius` in both axes, doubles every cell into a 2x2 block and, per
 * fine row, keeps the first contiguous run of columns.  Because the coarse path is monotone and
 * connected, the dilated cells of a coarse row a are one interval
 *   [min j of path cells with |i-a|<=r] - r  ..  [max j of those] + r,
 * so the fine row i (coarse a = i/2; rows a with no path cell within r do not occur for r >= 1)
 * gets columns [2*(jmin-r), 2*(jmax+r)+2) clipped to [0, len_y).  oracle/fastdtw_py.py holds the
 * literal set-based version and the test-suite checks they agree.                               */
static void expand_window(const int32_t* pi, const int32_t* pj, long n, long len_x, long len_y, long radius,
                          int32_t* lo, int32_t* hi) {
  const long cx = pi[n - 1] + 1; /* coarse length in x */
  long* jmin = (long*)malloc(sizeof(long) * (size_t)cx);
  long* jmax = (long*)malloc(sizeof(long) * (size_t)cx);
  for (long a = 0; a < cx; ++a) { jmin[a] = (1L << 40); jmax[a] = -1; }
  for (long k = 0; k < n; ++k) {
    if (pj[k] < jmin[pi[k]]) jmin[pi[k]] = pj[k];
    if (pj[k] > jmax[pi[k]]) jmax[pi[k]] = pj[k];
  }
  for (long i = 0; i < len_x; ++i) {
    const long a = i / 2;
    long mn = (1L << 40), mx = -1;
    for (long aa = lmax(0, a - radius); aa <= lmin(cx - 1, a + radius); ++aa) {
      if (jmin[aa] < mn) mn = jmin[aa];
      if (jmax[aa] > mx) mx = jmax[aa];
    }
    long l = 2 * (mn - radius), h = 2 * (mx + radius) + 2;
    if (mx < 0) { l = 0; h = 0; }
    lo[i] = (int32_t)lmax(0, l);
    hi[i] = (int32_t)lmin(len_y, h);
  }
  free(jmin); free(jmax);
}

static long fastdtw_rec(const double* x, const double* y, long Tx, long Ty, long D, int kind, long radius,
                        int32_t* path_i, int32_t* path_j, double* dist_out, int64_t* cells) {
  const long min_time_size = radius + 2; /* fastdtw.py __fastdtw */
  if (Tx < min_time_size || Ty < min_time_size)
    return orc_dtw_window(x, y, Tx, Ty, D, kind, NULL, NULL, path_i, path_j, dist_out, cells);
  /* __reduce_by_half: [(x[i] + x[i+1]) / 2 for i in range(0, len(x) - len(x) % 2, 2)] */
  const long hx = Tx / 2, hy = Ty / 2;
  double* xs = (double*)malloc(sizeof(double) * (size_t)(hx * D + 1));
  double* ys = (double*)malloc(sizeof(double) * (size_t)(hy * D + 1));
  for (long i = 0; i < hx; ++i)
    for (long k = 0; k < D; ++k) xs[i * D + k] = (x[(2 * i) * D + k] + x[(2 * i + 1) * D + k]) / 2;
  for (long i = 0; i < hy; ++i)
    for (long k = 0; k < D; ++k) ys[i * D + k] = (y[(2 * i) * D + k] + y[(2 * i + 1) * D + k]) / 2;
  int32_t* ci = (int32_t*)malloc(sizeof(int32_t) * (size_t)(hx + hy + 2));
  int32_t* cj = (int32_t*)malloc(sizeof(int32_t) * (size_t)(hx + hy + 2));
  double dcoarse;
  long n = fastdtw_rec(xs, ys, hx, hy, D, kind, radius, ci, cj, &dcoarse, cells);
  free(xs); free(ys);
  if (n < 0) { free(ci); free(cj); return n; }
  int32_t* lo = (int32_t*)malloc(sizeof(int32_t) * (size_t)Tx);
  int32_t* hi = (int32_t*)malloc(sizeof(int32_t) * (size_t)Tx);
  expand_window(ci, cj, n, Tx, Ty, radius, lo, hi);
  free(ci); free(cj);
  n = orc_dtw_window(x, y, Tx, Ty, D, kind, lo, hi, path_i, path_j, dist_out, cells);
  free(lo); free(hi);
  return n;
}

/* fastdtw(x, y, radius, dist): x (Tx, D), y (Ty, D) float64 row-major.  radius < 0 selects the
 * exact DTW (`dtw(x, y, dist)` of the same package: full window).  Returns path length or <0.   */
long orc_fastdtw(const double* x, const double* y, long Tx, long Ty, long D, int kind, long radius,
                 int32_t* path_i, int32_t* path_j, double* dist_out, int64_t* cells) {
  if (cells) *cells = 0;
  if (radius < 0) return orc_dtw_window(x, y, Tx, Ty, D, kind, NULL, NULL, path_i, path_j, dist_out, cells);
  return fastdtw_rec(x, y, Tx, Ty, D, kind, radius, path_i, path_j, dist_out, cells);
}

/* expose the window expansion for tests (coarse path -> per-row fine windows) */
void orc_expand_window(const int32_t* pi, const int32_t* pj, long n, long len_x, long len_y, long radius,
                       int32_t* lo, int32_t* hi) {
  expand_window(pi, pj, n, len_x, len_y, radius, lo, hi);
}

/* Direct entry points to the banded kernels for the known-answer tests the reference holds
 * (tests/bandmat/test_linalg.py:77-114: 4x4 SPD tridiagonal, lower band storage).               */
long orc_cholesky_banded_lower(double* half_band, long depth, long frames) {
  return cholesky_banded_lower(half_band, depth, frames);
}
long orc_cho_solve_lower(const double* chol, long depth, long frames, const double* b, double* x) {
  double* tmp = (double*)malloc(sizeof(double) * (size_t)(frames + 1));
  long r = cho_solve_lower(chol, depth, frames, b, x, tmp);
  free(tmp);
  return r;
}
void orc_cholesky_inv_banded(const double* L_full, long T, long width, double* Pinv) {
  cholesky_inv_banded_dense(L_full, T, width, Pinv);
}
