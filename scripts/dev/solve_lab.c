/* dev lab (not product, not oracle): variants of the restart / primal-weight rules of solve mode, run on the CPU to
 * choose what goes into the kernels.  Includes the oracle's helpers. */
#include "../../oracle/pdhg_oracle.c"

int lab_solve(int m, int n, const int32_t *indptr, const int32_t *indices,
              const double *values, const double *b, const double *c,
              const double *lb, const double *ub, const double *ylo, const double *yhi,
              double *x, double *y, double eta, double w0, int max_iters,
              int check_every, double tol, double *kkt_out, double *info, int nthreads,
              double kp, double ki, double kd, int cross, double bsuf, double bnec, double bart)
{
    set_threads(nthreads);
    csr_pair P = { m, n, indptr, indices, values, NULL, NULL, NULL };
    if (build_transpose(&P)) return -1;
    const size_t nn = (size_t)(n > 0 ? n : 1), mm = (size_t)(m > 0 ? m : 1);
    double *xbar = (double *)malloc(nn * sizeof(double));
    double *x0 = (double *)malloc(nn * sizeof(double));
    double *y0 = (double *)malloc(mm * sizeof(double));
    double *dxv = (double *)malloc(nn * sizeof(double));
    memcpy(x0, x, (size_t)n * sizeof(double));
    memcpy(y0, y, (size_t)m * sizeof(double));
    if (w0 <= 0.0) {
        double nb2 = 0.0, nc2 = 0.0;
        for (int i = 0; i < m; ++i) nb2 += b[i] * b[i];
        for (int j = 0; j < n; ++j) nc2 += c[j] * c[j];
        w0 = (nb2 > 0.0 && nc2 > 0.0) ? sqrt(nc2 / nb2) : 1.0;
    }
    double w = w0, fpe_restart = -1.0, fpe_prev = INFINITY, integ = 0.0, eprev = 0.0;
    int k = 0, it = 0, restarts = 0, converged = 0;
    double kk[10];
    kkt_eval(&P, b, c, lb, ub, ylo, yhi, x, y, kk);
    for (it = 0; it < max_iters && !converged;) {
        const double tau = eta / w, sigma = eta * w;
        const double lam = (double)(k + 1) / (double)(k + 2);
        double dx2 = 0.0, dy2 = 0.0, cr = 0.0;
        for (int j = 0; j < n; ++j) {
            double g = c[j] - row_dot(P.tptr, P.tidx, P.tval, j, y);
            double lo = lb ? lb[j] : 0.0, hi = ub ? ub[j] : INFINITY;
            double xn = clip(x[j] - tau * g, lo, hi);
            double d = xn - x[j];
            dx2 += d * d;
            dxv[j] = d;
            xbar[j] = 2.0 * xn - x[j];
            x[j] = lam * xbar[j] + (1.0 - lam) * x0[j];
        }
        for (int i = 0; i < m; ++i) {
            double yn = y[i] + sigma * (b[i] - row_dot(indptr, indices, values, i, xbar));
            if (ylo) yn = clip(yn, ylo[i], yhi[i]);
            double d = yn - y[i];
            dy2 += d * d;
            if (cross) cr += d * row_dot(indptr, indices, values, i, dxv);
            y[i] = lam * (2.0 * yn - y[i]) + (1.0 - lam) * y0[i];
        }
        ++it; ++k;
        double f2 = w * dx2 + dy2 / w;
        if (cross) f2 -= 2.0 * eta * cr;
        double fpe = sqrt(f2 > 0 ? f2 : 0);
        if (fpe_restart < 0.0) fpe_restart = fpe;
        if (it % check_every == 0 || it == max_iters) {
            kkt_eval(&P, b, c, lb, ub, ylo, yhi, x, y, kk);
            if (kk[8] <= tol) { converged = 1; break; }
            int do_restart = (fpe <= bsuf * fpe_restart) ||
                             (fpe <= bnec * fpe_restart && fpe > fpe_prev) ||
                             ((double)k >= bart * (double)it);
            fpe_prev = fpe;
            if (do_restart) {
                double ddx = 0.0, ddy = 0.0;
                for (int j = 0; j < n; ++j) { double d = x[j] - x0[j]; ddx += d * d; }
                for (int i = 0; i < m; ++i) { double d = y[i] - y0[i]; ddy += d * d; }
                ddx = sqrt(ddx); ddy = sqrt(ddy);
                if (ddx > 1e-10 && ddy > 1e-10) {
                    double e = log(w * ddx / ddy);
                    integ += e;
                    w = exp(log(w) - (kp * e + ki * integ + kd * (e - eprev)));
                    eprev = e;
                }
                memcpy(x0, x, (size_t)n * sizeof(double));
                memcpy(y0, y, (size_t)m * sizeof(double));
                k = 0; fpe_restart = -1.0; fpe_prev = INFINITY;
                ++restarts;
            }
        }
    }
    memcpy(kkt_out, kk, sizeof(kk));
    info[0] = it; info[1] = restarts; info[2] = converged; info[3] = w;
    free(xbar); free(x0); free(y0); free(dxv);
    free_transpose(&P);
    return 0;
}
