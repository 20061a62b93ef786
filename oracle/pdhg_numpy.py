"""oracle/pdhg_numpy.py -- independent numpy/scipy restatement of the frozen PDHG spec.

TEST INFRASTRUCTURE ONLY (see oracle/pdhg_oracle.c header): imported by tests/ to
cross-check the C oracle; never imported by mllp_b200/.

PARITY UNPINNED by the reference: HAHHHD/mllp has no primal-dual iteration
(SURVEY.md section 0).  The spec restated here is SURVEY.md section 8(c); the inputs follow
linear_program_data.py:58-80 (scipy CSR from "<name>_constrs.npz", "_coefs.npy", "_rhs.npy").
"""
import numpy as np
import scipy.sparse as sp


def as_csr(m, n, indptr, indices, values):
    return sp.csr_matrix((np.asarray(values, dtype=np.float64),
                          np.asarray(indices, dtype=np.int32),
                          np.asarray(indptr, dtype=np.int32)), shape=(m, n))


def pdhg_run(A, b, c, x, y, tau, sigma, num_iters, lb=None, ub=None, ylo=None, yhi=None):
    """Parity mode: g=c-A'y; x+=clip(x-tau g,l,u); xbar=2x+-x; y+=clip(y+sigma(b-A xbar))."""
    A = sp.csr_matrix(A)
    At = A.T.tocsr()
    x = np.array(x, dtype=np.float64)
    y = np.array(y, dtype=np.float64)
    lo = np.zeros_like(x) if lb is None else lb
    hi = np.full_like(x, np.inf) if ub is None else ub
    for _ in range(num_iters):
        g = c - At @ y
        xn = np.clip(x - tau * g, lo, hi)
        xbar = 2.0 * xn - x
        x = xn
        y = y + sigma * (b - A @ xbar)
        if ylo is not None:
            y = np.clip(y, ylo, yhi)
    return x, y


def kkt(A, b, c, x, y, lb=None, ub=None, ylo=None, yhi=None):
    """Same ten scalars as oracle_kkt (oracle/pdhg_oracle.c)."""
    A = sp.csr_matrix(A)
    lo = np.zeros_like(x) if lb is None else lb
    hi = np.full_like(x, np.inf) if ub is None else ub
    r = c - A.T @ y
    rp, rn = np.maximum(r, 0.0), np.minimum(r, 0.0)
    fin_hi, fin_lo = np.isfinite(hi), np.isfinite(lo)
    dobj = b @ y + np.sum(np.where(fin_hi, hi, 0.0) * np.where(fin_hi, rn, 0.0)) \
        + np.sum(np.where(fin_lo, lo, 0.0) * np.where(fin_lo, rp, 0.0))
    dres2 = np.sum(np.where(fin_hi, 0.0, rn) ** 2) + np.sum(np.where(fin_lo, 0.0, rp) ** 2)
    res = A @ x - b
    if ylo is not None:
        ge = np.isinf(yhi) & (ylo == 0.0)
        le = np.isinf(ylo) & (yhi == 0.0)
        res = np.where(ge & (res > 0), 0.0, res)
        res = np.where(le & (res < 0), 0.0, res)
        dres2 += np.sum((y - np.clip(y, ylo, yhi)) ** 2)      # distance of y from its cone
    pres2 = np.sum(res ** 2) + np.sum((x - np.clip(x, lo, hi)) ** 2)   # + distance of x from its box
    pobj = c @ x
    out = np.zeros(10)
    out[0], out[1] = pobj, dobj
    out[2], out[3] = np.sqrt(pres2), np.sqrt(dres2)
    out[4], out[5] = np.linalg.norm(b), np.linalg.norm(c)
    out[6], out[7] = np.linalg.norm(x), np.linalg.norm(y)
    gap = abs(pobj - dobj)
    out[8] = max(out[2] / (1 + out[4]), out[3] / (1 + out[5]), gap / (1 + abs(pobj) + abs(dobj)))
    out[9] = gap
    return out


def power_iteration(A, iters=50):
    A = sp.csr_matrix(A)
    n = A.shape[1]
    v = np.full(n, 1.0 / np.sqrt(n))
    lam = 0.0
    for _ in range(iters):
        z = A.T @ (A @ v)
        nz = np.linalg.norm(z)
        lam = nz
        if nz == 0.0:
            break
        v = z / nz
    return np.sqrt(lam)


def pdhg_solve(A, b, c, x, y, eta, w0=1.0, max_iters=100000, check_every=64, tol=1e-6,
               lb=None, ub=None, ylo=None, yhi=None):
    """Solve mode: reflected restarted Halpern PDHG, spec in oracle_pdhg_solve's comment."""
    A = sp.csr_matrix(A)
    At = A.T.tocsr()
    x = np.array(x, dtype=np.float64)
    y = np.array(y, dtype=np.float64)
    lo = np.zeros_like(x) if lb is None else lb
    hi = np.full_like(x, np.inf) if ub is None else ub
    x0, y0 = x.copy(), y.copy()
    if w0 <= 0:   # the PDLP default
        nb2, nc2 = float(np.dot(b, b)), float(np.dot(c, c))
        w0 = np.sqrt(nc2 / nb2) if nb2 > 0 and nc2 > 0 else 1.0
    w, fpe_restart, fpe_prev = w0, -1.0, np.inf
    k = it = restarts = 0
    converged = False
    kk = kkt(A, b, c, x, y, lb, ub, ylo, yhi)
    while it < max_iters:
        tau, sigma = eta / w, eta * w
        lam = (k + 1) / (k + 2)
        xn = np.clip(x - tau * (c - At @ y), lo, hi)
        dx2 = np.sum((xn - x) ** 2)
        xbar = 2.0 * xn - x
        x = lam * xbar + (1 - lam) * x0
        yn = y + sigma * (b - A @ xbar)
        if ylo is not None:
            yn = np.clip(yn, ylo, yhi)
        dy2 = np.sum((yn - y) ** 2)
        y = lam * (2.0 * yn - y) + (1 - lam) * y0
        it += 1
        k += 1
        fpe = np.sqrt(w * dx2 + dy2 / w)
        if fpe_restart < 0:
            fpe_restart = fpe
        if it % check_every == 0 or it == max_iters:
            kk = kkt(A, b, c, x, y, lb, ub, ylo, yhi)
            if kk[8] <= tol:
                converged = True
                break
            do_restart = (fpe <= 0.2 * fpe_restart) or \
                (fpe <= 0.8 * fpe_restart and fpe > fpe_prev) or (k >= 0.36 * it)
            fpe_prev = fpe
            if do_restart:
                ddx, ddy = np.linalg.norm(x - x0), np.linalg.norm(y - y0)
                if ddx > 1e-10 and ddy > 1e-10:
                    w = np.exp(0.5 * np.log(ddy / ddx) + 0.5 * np.log(w))
                x0, y0 = x.copy(), y.copy()
                k, fpe_restart, fpe_prev = 0, -1.0, np.inf
                restarts += 1
    return x, y, kk, dict(iters=it, restarts=restarts, converged=converged, w=w)
