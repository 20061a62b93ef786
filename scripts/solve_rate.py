"""Dev tool: us / iteration of solve mode (reflected restarted Halpern PDHG, KKT check every `check_every` iterations)
next to parity mode on the same handle."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import mllp_b200 as M


def main(names):
    for name in names:
        A, b, c = M.load_csr(name); m, n = A.shape
        lp = M.DeviceLP(A, A.data, m, n)
        eta = 0.9 / lp.sigma_max()
        bt, ct = torch.tensor(b, device="cuda"), torch.tensor(c, device="cuda")
        K = 2000

        def timed(fn):
            fn(); torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); fn(); e1.record(); torch.cuda.synchronize()
            return e0.elapsed_time(e1) * 1e3 / K
        par = timed(lambda: M.pdhg_linear_program(A, A.data, bt, ct, num_iters=K, tau=eta, sigma=eta, handle=lp))
        line = "%-8s parity %.2f us/it" % (name, par)
        for ce in (64, 256, 1000000):
            us = timed(lambda: M.solve_linear_program(A, A.data, bt, ct, tol=1e-30, max_iters=K, check_every=ce, handle=lp))
            line += " | solve check_every=%d: %.2f" % (ce, us)
        print(line, flush=True)
        lp.close()


if __name__ == "__main__":
    main(sys.argv[1:] or ["afiro", "25fv47", "pilot87", "pds-20", "ken-18", "osa-60"])
