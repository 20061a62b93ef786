"""Dev tool: the Netlib files solve mode does not bring to 1e-6 (original-LP KKT) within 2e6 iterations, under a few
variants of the solve-mode knobs the C ABI exposes (initial primal weight, check interval, iteration cap)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import mllp_b200 as M
from mllp_b200.mps import read_mps

names = [a for a in sys.argv[1:] if not a.startswith("-")] or ["bnl1", "pilot4", "perold", "pilot.we", "pilot.ja", "greenbea", "pilot"]
for nm in names:
    lp = read_mps(os.path.join("data", "netlib_mps_gz", nm + ".mps.gz"))
    A = lp["A"].tocsr(); m, n = A.shape
    h = M.DeviceLP(A, A.data, m, n, lb=lp["lb"], ub=lp["ub"], ylo=lp["ylo"], yhi=lp["yhi"], precondition=True, flags=M._cabi.F_NO_TUNE)
    dr, dc = h.scaling()
    nb, nc = np.linalg.norm(dr * lp["b"]), np.linalg.norm(dc * lp["c"])
    w_pdlp = nc / nb if nb > 0 and nc > 0 else 1.0
    for tag, kw in (("w0=1 ce=64", dict(primal_weight=1.0, check_every=64)),
                    ("w0=|c|/|b| ce=64", dict(primal_weight=w_pdlp, check_every=64)),
                    ("w0=|c|/|b| ce=256", dict(primal_weight=w_pdlp, check_every=256)),
                    ("w0=1 ce=512", dict(primal_weight=1.0, check_every=512))):
        t = time.time()
        obj, x, y, info = M.solve_linear_program(A, A.data, lp["b"], lp["c"], tol=1e-6, max_iters=4000000, handle=h, **kw)
        print("%-9s %-18s w0 %.3g: conv %s iters %d restarts %d kkt %.2e obj %.9g  (%.1f s)" % (
            nm, tag, kw["primal_weight"], info["converged"], info["iters"], info["restarts"], info["rel_kkt"], obj + lp["offset"], time.time() - t), flush=True)
    h.close()
