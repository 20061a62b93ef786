"""Dev tool: us / iteration and oracle parity of the persistent kernel per launch geometry (MLLP_GEOM):
cooperative grid (0), one cluster of 16 / 8 / 4 CTAs, one CTA (1), and what mllp_lp_create picks on its own (auto)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import mllp_b200 as M
from oracle import pdhg_oracle as O


def timed(lp, A, bt, ct, eta, K):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    M.pdhg_linear_program(A, A.data, bt, ct, num_iters=K, tau=eta, sigma=eta, handle=lp)
    torch.cuda.synchronize()
    e0.record(); M.pdhg_linear_program(A, A.data, bt, ct, num_iters=K, tau=eta, sigma=eta, handle=lp); e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / K


def main(names, geoms):
    for name in names:
        A, b, c = M.load_csr(name); m, n = A.shape
        eta = 0.9 / O.power_iteration(A, 50)
        K = 300
        xo, yo = O.pdhg_run(A, b, c, np.zeros(n), np.zeros(m), eta, eta, K)
        bt, ct = torch.tensor(b, device="cuda"), torch.tensor(c, device="cuda")
        line = "%-8s nnz %7d:" % (name, A.nnz)
        for g in geoms:
            if g == "auto":
                os.environ.pop("MLLP_GEOM", None)
            else:
                os.environ["MLLP_GEOM"] = g
            try:
                lp = M.DeviceLP(A, A.data, m, n)
            except RuntimeError as e:
                line += "  %s: %s" % (g, str(e)[:60]); continue
            obj, x, y, info = M.pdhg_linear_program(A, A.data, b, c, num_iters=K, tau=eta, sigma=eta, handle=lp)
            ex = np.linalg.norm(x - xo) / max(np.linalg.norm(xo), 1e-300)
            ey = np.linalg.norm(y - yo) / max(np.linalg.norm(yo), 1e-300)
            us = timed(lp, A, bt, ct, eta, 2000)
            geo = lp.geometry()
            line += "  %s[%s%d] %.2f us (err %.0e)" % (g, geo["mode"][:2], geo["ctas"], us, max(ex, ey))
            if g == "auto":
                line += " " + str({k: round(v) for k, v in geo["ns_per_iter"].items()}) + " tune " + str(lp.tune_info())
            # solve mode on the same handle
            if name in ("afiro", "sc50a", "25fv47") and g in ("auto", "0"):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(); obj, x, y, si = M.solve_linear_program(A, A.data, b, c, handle=lp, tol=1e-6); e1.record()
                torch.cuda.synchronize()
                line += " solve: obj %.8g iters %d conv %s %.1f ms" % (obj, si["iters"], si["converged"], e0.elapsed_time(e1))
            lp.close()
        print(line, flush=True)


if __name__ == "__main__":
    names = sys.argv[1].split(",") if len(sys.argv) > 1 else ["afiro", "sc105", "25fv47", "d2q06c", "dfl001", "pilot87", "pds-20"]
    geoms = sys.argv[2].split(",") if len(sys.argv) > 2 else ["0", "16", "8", "4", "1", "auto"]
    main(names, geoms)
