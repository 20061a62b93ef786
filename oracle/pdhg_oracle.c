/*
 * oracle/pdhg_oracle.c -- CPU fp64 restatement of the primal-dual LP iteration.
 *
 * THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may load it.  The product path
 * (mllp_b200/) never links or calls anything in oracle/.
 *
 * PARITY UNPINNED: the reference (HAHHHD/mllp) contains no primal-dual iteration at all
 * (linear_program_methods.py holds GNN basis predictors and max-covering
 * routines only; its only LP solves are third-party GLOP/Gurobi calls in dead code,
 * linear_program_methods.py:477-539, :542-610).  So this file follows the *frozen spec*
 * of SURVEY.md section 8(c), not a reference file.  What the reference does pin, and what this
 * oracle is checked against in tests/test_oracle.py:
 *   - the input contract of linear_program_data.py:58-80 (CSR float64 data / int32
 *     indices+indptr from "<name>_constrs.npz", c from "_coefs.npy", b from "_rhs.npy");
 *   - independent optimal objectives of the LPs those arrays define (HiGHS, BASELINE.md
 *     section 4; fixtures in tests/golden/highs_objectives.json);
 *   - an independent numpy/scipy restatement of the same spec (oracle/pdhg_numpy.py).
 *
 * LP form (dataset "_norm" variant is l = 0, u = +inf, all rows equalities):
 *     min c'x   s.t.  A x - b in K_row,   l <= x <= u
 * with the dual variable of row i boxed in [ylo_i, yhi_i]
 *     equality row : (-inf, +inf)     ">=" row : [0, +inf)     "<=" row : (-inf, 0]
 *
 * Parity mode (fixed step, no data-dependent branches), per iteration:
 *     g    = c - A' y
 *     x+   = clip(x - tau g, l, u)
 *     xbar = 2 x+ - x
 *     y+   = clip(y + sigma (b - A xbar), ylo, yhi)
 *
 * Solve mode: reflected, restarted Halpern PDHG with a fixed step eta = 0.99/||A||_2 and
 * primal weight w (tau = eta / w, sigma = eta * w); see oracle_pdhg_solve below.
 *
 * Plain C99 + optional OpenMP (row-parallel SpMV; every row is summed sequentially in
 * index order, so results do not depend on the thread count).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef struct {
    int m, n;
    const int32_t *indptr, *indices; /* CSR of A, caller owned */
    const double *values;
    int32_t *tptr, *tidx;            /* CSR of A' (built here) */
    double *tval;
} csr_pair;

static int build_transpose(csr_pair *P)
{
    const int m = P->m, n = P->n;
    const int64_t nnz = P->indptr[m];
    P->tptr = (int32_t *)calloc((size_t)n + 1, sizeof(int32_t));
    P->tidx = (int32_t *)malloc((size_t)(nnz > 0 ? nnz : 1) * sizeof(int32_t));
    P->tval = (double *)malloc((size_t)(nnz > 0 ? nnz : 1) * sizeof(double));
    if (!P->tptr || !P->tidx || !P->tval) return -1;
    for (int64_t k = 0; k < nnz; ++k) P->tptr[P->indices[k] + 1]++;
    for (int j = 0; j < n; ++j) P->tptr[j + 1] += P->tptr[j];
    int32_t *fill = (int32_t *)malloc((size_t)(n > 0 ? n : 1) * sizeof(int32_t));
    if (!fill) return -1;
    memcpy(fill, P->tptr, (size_t)n * sizeof(int32_t));
    for (int i = 0; i < m; ++i)
        for (int32_t k = P->indptr[i]; k < P->indptr[i + 1]; ++k) {
            int32_t j = P->indices[k], q = fill[j]++;
            P->tidx[q] = i;
            P->tval[q] = P->values[k];
        }
    free(fill);
    return 0;
}

static void free_transpose(csr_pair *P)
{
    free(P->tptr); free(P->tidx); free(P->tval);
    P->tptr = P->tidx = NULL; P->tval = NULL;
}

static inline double row_dot(const int32_t *ptr, const int32_t *idx, const double *val,
                             int r, const double *v)
{
    double s = 0.0;
    for (int32_t k = ptr[r]; k < ptr[r + 1]; ++k) s += val[k] * v[idx[k]];
    return s;
}

static inline double clip(double v, double lo, double hi)
{
    return v < lo ? lo : (v > hi ? hi : v);
}

int oracle_num_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

static void set_threads(int nthreads)
{
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#else
    (void)nthreads;
#endif
}

/* out = A v (trans = 0, v has n entries) or A' v (trans = 1, v has m entries). */
int oracle_spmv(int m, int n, const int32_t *indptr, const int32_t *indices,
                const double *values, int trans, const double *v, double *out, int nthreads)
{
    set_threads(nthreads);
    if (!trans) {
#pragma omp parallel for schedule(static)
        for (int i = 0; i < m; ++i) out[i] = row_dot(indptr, indices, values, i, v);
        return 0;
    }
    csr_pair P = { m, n, indptr, indices, values, NULL, NULL, NULL };
    if (build_transpose(&P)) return -1;
#pragma omp parallel for schedule(static)
    for (int j = 0; j < n; ++j) out[j] = row_dot(P.tptr, P.tidx, P.tval, j, v);
    free_transpose(&P);
    return 0;
}

/* Parity mode: K fixed-step PDHG iterations in place on x (n) and y (m).
 * lb/ub/ylo/yhi may be NULL (=> l = 0, u = +inf; rows are equalities). */
int oracle_pdhg_run(int m, int n, const int32_t *indptr, const int32_t *indices,
                    const double *values, const double *b, const double *c,
                    const double *lb, const double *ub, const double *ylo, const double *yhi,
                    double *x, double *y, double tau, double sigma, int num_iters, int nthreads)
{
    set_threads(nthreads);
    csr_pair P = { m, n, indptr, indices, values, NULL, NULL, NULL };
    if (build_transpose(&P)) return -1;
    double *xbar = (double *)malloc((size_t)(n > 0 ? n : 1) * sizeof(double));
    if (!xbar) { free_transpose(&P); return -1; }
    for (int it = 0; it < num_iters; ++it) {
#pragma omp parallel for schedule(static)
        for (int j = 0; j < n; ++j) {
            double g = c[j] - row_dot(P.tptr, P.tidx, P.tval, j, y);
            double lo = lb ? lb[j] : 0.0, hi = ub ? ub[j] : INFINITY;
            double xn = clip(x[j] - tau * g, lo, hi);
            xbar[j] = 2.0 * xn - x[j];
            x[j] = xn;
        }
#pragma omp parallel for schedule(static)
        for (int i = 0; i < m; ++i) {
            double yn = y[i] + sigma * (b[i] - row_dot(indptr, indices, values, i, xbar));
            if (ylo) yn = clip(yn, ylo[i], yhi[i]);
            y[i] = yn;
        }
    }
    free(xbar);
    free_transpose(&P);
    return 0;
}

/*
 * KKT scalars at (x, y).  out[0..9]:
 *  0 pobj = c'x
 *  1 dobj = b'y + sum_j (l_j r_j^+ + u_j r_j^-)  over finite bounds, r = c - A'y
 *  2 ||primal residual||_2 : row i contributes (Ax-b)_i unless its sign is allowed
 *       (equality: always; ">=" row [ylo=0]: only if (Ax-b)_i < 0; "<=" row: only if > 0),
 *       and column j contributes x_j - clip(x_j, l_j, u_j): the distance of x from its box (zero at
 *       every projected point; the Halpern combinations of solve mode can leave the box slightly)
 *  3 ||dual residual||_2   : r_j^- if u_j = +inf (must be >= 0 there) and r_j^+ if l_j = -inf,
 *       and row i contributes y_i - clip(y_i, ylo_i, yhi_i): the distance of y from its cone
 *  4 ||b||_2   5 ||c||_2   6 ||x||_2   7 ||y||_2
 *  8 relative KKT error = max(out2/(1+out4), out3/(1+out5), |pobj-dobj|/(1+|pobj|+|dobj|))
 *  9 |pobj - dobj|
 */
static void kkt_eval(const csr_pair *P, const double *b, const double *c,
                     const double *lb, const double *ub, const double *ylo, const double *yhi,
                     const double *x, const double *y, double *out)
{
    const int m = P->m, n = P->n;
    double pobj = 0, dobj = 0, pr2 = 0, dr2 = 0, nb2 = 0, nc2 = 0, nx2 = 0, ny2 = 0, pr2x = 0, dr2y = 0;
#pragma omp parallel for schedule(static) reduction(+ : pobj, dobj, dr2, nc2, nx2, pr2x)
    for (int j = 0; j < n; ++j) {
        double r = c[j] - row_dot(P->tptr, P->tidx, P->tval, j, y);
        double lo = lb ? lb[j] : 0.0, hi = ub ? ub[j] : INFINITY;
        double rp = r > 0 ? r : 0.0, rn = r < 0 ? r : 0.0, viol = 0.0;
        if (isinf(hi)) viol += rn * rn; else dobj += hi * rn;
        if (isinf(lo)) viol += rp * rp; else dobj += lo * rp;
        dr2 += viol;
        double xv = x[j] - clip(x[j], lo, hi);
        pr2x += xv * xv;
        pobj += c[j] * x[j];
        nc2 += c[j] * c[j];
        nx2 += x[j] * x[j];
    }
#pragma omp parallel for schedule(static) reduction(+ : dobj, pr2, nb2, ny2, dr2y)
    for (int i = 0; i < m; ++i) {
        double res = row_dot(P->indptr, P->indices, P->values, i, x) - b[i];
        if (ylo) {
            if (res > 0 && isinf(yhi[i]) && ylo[i] == 0.0) res = 0.0; /* ">=" row satisfied */
            if (res < 0 && isinf(ylo[i]) && yhi[i] == 0.0) res = 0.0; /* "<=" row satisfied */
            double yv = y[i] - clip(y[i], ylo[i], yhi[i]);
            dr2y += yv * yv;
        }
        pr2 += res * res;
        dobj += b[i] * y[i];
        nb2 += b[i] * b[i];
        ny2 += y[i] * y[i];
    }
    pr2 += pr2x; dr2 += dr2y;
    out[0] = pobj; out[1] = dobj; out[2] = sqrt(pr2); out[3] = sqrt(dr2);
    out[4] = sqrt(nb2); out[5] = sqrt(nc2); out[6] = sqrt(nx2); out[7] = sqrt(ny2);
    double gap = fabs(pobj - dobj);
    double e = out[2] / (1.0 + out[4]);
    double e2 = out[3] / (1.0 + out[5]);
    double e3 = gap / (1.0 + fabs(pobj) + fabs(dobj));
    if (e2 > e) e = e2;
    if (e3 > e) e = e3;
    out[8] = e; out[9] = gap;
}

int oracle_kkt(int m, int n, const int32_t *indptr, const int32_t *indices,
               const double *values, const double *b, const double *c,
               const double *lb, const double *ub, const double *ylo, const double *yhi,
               const double *x, const double *y, double *out, int nthreads)
{
    set_threads(nthreads);
    csr_pair P = { m, n, indptr, indices, values, NULL, NULL, NULL };
    if (build_transpose(&P)) return -1;
    kkt_eval(&P, b, c, lb, ub, ylo, yhi, x, y, out);
    free_transpose(&P);
    return 0;
}

/* sigma_max(A) estimate: `iters` steps of power iteration on A'A from v = 1/sqrt(n).
 * Returns sqrt(||A'A v|| / ||v||) of the last step, i.e. sqrt of the Rayleigh-type ratio. */
int oracle_power_iteration(int m, int n, const int32_t *indptr, const int32_t *indices,
                           const double *values, int iters, double *sigma_max, int nthreads)
{
    set_threads(nthreads);
    csr_pair P = { m, n, indptr, indices, values, NULL, NULL, NULL };
    if (build_transpose(&P)) return -1;
    double *v = (double *)malloc((size_t)(n > 0 ? n : 1) * sizeof(double));
    double *w = (double *)malloc((size_t)(m > 0 ? m : 1) * sizeof(double));
    double *z = (double *)malloc((size_t)(n > 0 ? n : 1) * sizeof(double));
    if (!v || !w || !z) return -1;
    for (int j = 0; j < n; ++j) v[j] = 1.0 / sqrt((double)n);
    double lam = 0.0;
    for (int it = 0; it < iters; ++it) {
#pragma omp parallel for schedule(static)
        for (int i = 0; i < m; ++i) w[i] = row_dot(indptr, indices, values, i, v);
        double nz2 = 0.0;
#pragma omp parallel for schedule(static) reduction(+ : nz2)
        for (int j = 0; j < n; ++j) {
            z[j] = row_dot(P.tptr, P.tidx, P.tval, j, w);
            nz2 += z[j] * z[j];
        }
        double nz = sqrt(nz2);
        lam = nz; /* ||A'A v|| with ||v|| = 1 */
        if (nz == 0.0) break;
        for (int j = 0; j < n; ++j) v[j] = z[j] / nz;
    }
    *sigma_max = sqrt(lam);
    free(v); free(w); free(z);
    free_transpose(&P);
    return 0;
}

/*
 * Solve mode: reflected restarted Halpern PDHG, fixed step.
 *
 *   z = (x, y), anchor z0, inner counter k (since last restart), weight w:
 *     tau = eta / w, sigma = eta * w
 *     x'  = clip(x - tau (c - A'y), l, u)          xbar = 2x' - x
 *     y'  = clip(y + sigma (b - A xbar), ylo, yhi)
 *     lam = (k+1)/(k+2)
 *     x  <- lam * xbar        + (1-lam) * x0        (xbar = 2x'-x is the reflection)
 *     y  <- lam * (2y' - y)   + (1-lam) * y0
 *   fixed-point error of the step (before the Halpern combination):
 *     fpe = sqrt( w ||x'-x||^2 + ||y'-y||^2 / w )
 *   Every `check_every` iterations (k counts iterations since the last restart):
 *     - KKT scalars at the current z; stop if rel KKT error <= tol.
 *     - restart if  fpe <= 0.2 fpe_at_restart                       (sufficient)
 *                or fpe <= 0.8 fpe_at_restart and fpe > fpe_prev    (necessary + no progress)
 *                or k >= 0.36 * total_iterations                    (artificial)
 *       on restart: w <- exp(0.5 log(dy/dx) + 0.5 log w) with dx = ||x - x0||, dy = ||y - y0||
 *                   (only if both > 1e-10), z0 <- z, k <- 0, and fpe_at_restart is the fpe
 *                   of the first step after the restart.
 *   Initial primal weight: w0 > 0 as given; w0 <= 0 selects the PDLP default ||c||_2 / ||b||_2 (1 if either is zero).
 *   info[0] iterations done, info[1] restarts, info[2] converged flag, info[3] final w.
 */
int oracle_pdhg_solve(int m, int n, const int32_t *indptr, const int32_t *indices,
                      const double *values, const double *b, const double *c,
                      const double *lb, const double *ub, const double *ylo, const double *yhi,
                      double *x, double *y, double eta, double w0, int max_iters,
                      int check_every, double tol, double *kkt_out, double *info, int nthreads)
{
    set_threads(nthreads);
    csr_pair P = { m, n, indptr, indices, values, NULL, NULL, NULL };
    if (build_transpose(&P)) return -1;
    const size_t nn = (size_t)(n > 0 ? n : 1), mm = (size_t)(m > 0 ? m : 1);
    double *xbar = (double *)malloc(nn * sizeof(double));
    double *x0 = (double *)malloc(nn * sizeof(double));
    double *y0 = (double *)malloc(mm * sizeof(double));
    if (!xbar || !x0 || !y0) return -1;
    memcpy(x0, x, (size_t)n * sizeof(double));
    memcpy(y0, y, (size_t)m * sizeof(double));
    if (w0 <= 0.0) {
        double nb2 = 0.0, nc2 = 0.0;
        for (int i = 0; i < m; ++i) nb2 += b[i] * b[i];
        for (int j = 0; j < n; ++j) nc2 += c[j] * c[j];
        w0 = (nb2 > 0.0 && nc2 > 0.0) ? sqrt(nc2 / nb2) : 1.0;
    }
    double w = w0, fpe_restart = -1.0, fpe_prev = INFINITY;
    int k = 0, it = 0, restarts = 0, converged = 0;
    double kk[10];
    kkt_eval(&P, b, c, lb, ub, ylo, yhi, x, y, kk);
    for (it = 0; it < max_iters && !converged;) {
        const double tau = eta / w, sigma = eta * w;
        const double lam = (double)(k + 1) / (double)(k + 2);
        double dx2 = 0.0, dy2 = 0.0;
#pragma omp parallel for schedule(static) reduction(+ : dx2)
        for (int j = 0; j < n; ++j) {
            double g = c[j] - row_dot(P.tptr, P.tidx, P.tval, j, y);
            double lo = lb ? lb[j] : 0.0, hi = ub ? ub[j] : INFINITY;
            double xn = clip(x[j] - tau * g, lo, hi);
            double d = xn - x[j];
            dx2 += d * d;
            xbar[j] = 2.0 * xn - x[j];
            x[j] = lam * xbar[j] + (1.0 - lam) * x0[j];
        }
#pragma omp parallel for schedule(static) reduction(+ : dy2)
        for (int i = 0; i < m; ++i) {
            double yn = y[i] + sigma * (b[i] - row_dot(indptr, indices, values, i, xbar));
            if (ylo) yn = clip(yn, ylo[i], yhi[i]);
            double d = yn - y[i];
            dy2 += d * d;
            y[i] = lam * (2.0 * yn - y[i]) + (1.0 - lam) * y0[i];
        }
        ++it; ++k;
        double fpe = sqrt(w * dx2 + dy2 / w);
        if (fpe_restart < 0.0) fpe_restart = fpe;
        if (it % check_every == 0 || it == max_iters) {
            kkt_eval(&P, b, c, lb, ub, ylo, yhi, x, y, kk);
            if (kk[8] <= tol) { converged = 1; break; }
            int do_restart = (fpe <= 0.2 * fpe_restart) ||
                             (fpe <= 0.8 * fpe_restart && fpe > fpe_prev) ||
                             ((double)k >= 0.36 * (double)it);
            fpe_prev = fpe;
            if (do_restart) {
                double ddx = 0.0, ddy = 0.0;
                for (int j = 0; j < n; ++j) { double d = x[j] - x0[j]; ddx += d * d; }
                for (int i = 0; i < m; ++i) { double d = y[i] - y0[i]; ddy += d * d; }
                ddx = sqrt(ddx); ddy = sqrt(ddy);
                if (ddx > 1e-10 && ddy > 1e-10) w = exp(0.5 * log(ddy / ddx) + 0.5 * log(w));
                memcpy(x0, x, (size_t)n * sizeof(double));
                memcpy(y0, y, (size_t)m * sizeof(double));
                k = 0; fpe_restart = -1.0; fpe_prev = INFINITY;
                ++restarts;
            }
        }
    }
    memcpy(kkt_out, kk, sizeof(kk));
    info[0] = it; info[1] = restarts; info[2] = converged; info[3] = w;
    free(xbar); free(x0); free(y0);
    free_transpose(&P);
    return 0;
}
