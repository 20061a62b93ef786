"""Dev tool: the block kernel (blocks.cu, MLLP_BLOCKS=1) against the grid kernel (MLLP_BLOCKS=0) and what mllp_lp_create
picks on its own: us / iteration and oracle parity."""
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


def main(names):
    for name in names:
        A, b, c = M.load_csr(name); m, n = A.shape
        eta = 0.9 / O.power_iteration(A, 50)
        K = 300
        xo, yo = O.pdhg_run(A, b, c, np.zeros(n), np.zeros(m), eta, eta, K)
        bt, ct = torch.tensor(b, device="cuda"), torch.tensor(c, device="cuda")
        line = "%-8s nnz %7d:" % (name, A.nnz)
        for mode in ("0", "1", "auto"):
            if mode == "auto":
                os.environ.pop("MLLP_BLOCKS", None)
            else:
                os.environ["MLLP_BLOCKS"] = mode
            lp = M.DeviceLP(A, A.data, m, n)
            obj, x, y, info = M.pdhg_linear_program(A, A.data, b, c, num_iters=K, tau=eta, sigma=eta, handle=lp)
            ex = np.linalg.norm(x - xo) / max(np.linalg.norm(xo), 1e-300)
            ey = np.linalg.norm(y - yo) / max(np.linalg.norm(yo), 1e-300)
            us = timed(lp, A, bt, ct, eta, 2000)
            bi = lp.blocks_info()
            line += "  blocks=%s[%d] %.2f us (err %.0e)" % (mode, bi["used"], us, max(ex, ey))
            if mode == "auto":
                line += " " + str(bi)
            lp.close()
        print(line, flush=True)


if __name__ == "__main__":
    main(sys.argv[1:] or ["ken-18", "pds-20", "osa-60", "dfl001"])
