"""Diagonal preconditioning for solve mode on raw MPS inputs.

Host-side, once per instance (not on the per-iteration path): Ruiz equilibration followed by
one Pock-Chambolle (alpha = 1) pass, the PDLP recipe.  The scaled LP

    A~ = Dr A Dc,  b~ = Dr b,  c~ = Dc c,  l~ = l / Dc,  u~ = u / Dc      (x = Dc x~,  y = Dr y~)

has the same optimal objective; row senses (dual boxes) are unchanged because Dr > 0.
The reference's own `_norm` arrays are a different, one-sided row scaling.
"""
import numpy as np
import scipy.sparse as sp


def ruiz_pock_chambolle(A, ruiz_iters=10):
    A = sp.csr_matrix(A, dtype=np.float64)
    m, n = A.shape
    dr, dc = np.ones(m), np.ones(n)
    absA = abs(A)
    for _ in range(ruiz_iters):
        B = sp.diags(dr) @ absA @ sp.diags(dc)
        rn = np.sqrt(np.asarray(B.max(axis=1).todense()).ravel())
        cn = np.sqrt(np.asarray(B.max(axis=0).todense()).ravel())
        rn[rn == 0] = 1.0
        cn[cn == 0] = 1.0
        dr /= rn
        dc /= cn
    B = sp.diags(dr) @ absA @ sp.diags(dc)
    rn = np.sqrt(np.asarray(B.sum(axis=1)).ravel())
    cn = np.sqrt(np.asarray(B.sum(axis=0)).ravel())
    rn[rn == 0] = 1.0
    cn[cn == 0] = 1.0
    return dr / rn, dc / cn


def scale_lp(A, b, c, lb=None, ub=None, ruiz_iters=10):
    dr, dc = ruiz_pock_chambolle(A, ruiz_iters)
    As = (sp.diags(dr) @ sp.csr_matrix(A) @ sp.diags(dc)).tocsr()
    As.sort_indices()
    out = {"A": As, "b": dr * np.asarray(b, dtype=np.float64), "c": dc * np.asarray(c, dtype=np.float64), "dr": dr, "dc": dc,
           "lb": None if lb is None else np.asarray(lb, dtype=np.float64) / dc,
           "ub": None if ub is None else np.asarray(ub, dtype=np.float64) / dc}
    return out


def solve_scaled(A, b, c, *, lb=None, ub=None, ylo=None, yhi=None, tol=1e-6, max_iters=400000, check_every=64, device=0,
                 scale=True):
    """Precondition (Ruiz + Pock-Chambolle), solve on the GPU in solve mode and return (objective, x, y, info) in the
    ORIGINAL variables; info['rel_kkt_original'] is the KKT error re-evaluated on the unscaled LP on the device."""
    from .linear_program_methods import DeviceLP, pdhg_linear_program, solve_linear_program
    A = sp.csr_matrix(A)
    m, n = A.shape
    if scale:
        s = scale_lp(A, b, c, lb, ub)
    else:
        s = {"A": A, "b": b, "c": c, "lb": lb, "ub": ub, "dr": np.ones(m), "dc": np.ones(n)}
    h = DeviceLP(s["A"], s["A"].data, m, n, lb=s["lb"], ub=s["ub"], ylo=ylo, yhi=yhi, device=device)
    obj, xs, ys, info = solve_linear_program(s["A"], s["A"].data, s["b"], s["c"], tol=tol, max_iters=max_iters,
                                             check_every=check_every, handle=h)
    x, y = s["dc"] * xs, s["dr"] * ys
    h.close()
    h0 = DeviceLP(A, A.data, m, n, lb=lb, ub=ub, ylo=ylo, yhi=yhi, device=device)
    _, _, _, i0 = pdhg_linear_program(A, A.data, b, c, num_iters=0, tau=1.0, sigma=1.0, x0=x, y0=y, handle=h0)
    h0.close()
    info = dict(info)
    info.pop("handle", None)
    info["rel_kkt_original"] = i0["rel_kkt"]
    return i0["pobj"], x, y, info


def solve_mps(path, *, tol=1e-6, max_iters=400000, check_every=64, device=0, scale=True):
    """Read an MPS file (with row senses, bounds, ranges), precondition, solve on the GPU in
    solve mode and return (objective incl. offset, x, y, info) in the ORIGINAL variables.
    info['rel_kkt_original'] is the KKT error re-evaluated on the unscaled LP on the device."""
    from .mps import read_mps
    lp = read_mps(path)
    pobj, x, y, info = solve_scaled(lp["A"], lp["b"], lp["c"], lb=lp["lb"], ub=lp["ub"], ylo=lp["ylo"], yhi=lp["yhi"], tol=tol,
                                    max_iters=max_iters, check_every=check_every, device=device, scale=scale)
    info["offset"] = lp["offset"]
    objective = pobj + lp["offset"]
    if lp["maximize"]:
        objective = -objective
    return objective, x, y, info
