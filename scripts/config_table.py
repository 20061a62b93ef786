"""Parity-mode time per iteration of every BASELINE.json config instance on one GPU (the geometry mllp_lp_create picks),
algorithmic bytes per iteration (24 nnz + 36 m + 44 n + 8) and the fraction of the measured HBM copy peak; the CPU oracle
(OpenMP, all host threads) beside it.  Writes a markdown table."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import mllp_b200 as M
from oracle import pdhg_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
NAMES = ["afiro", "sc50a", "sc105", "adlittle", "blend", "share2b", "kb2", "25fv47", "pilot87", "d2q06c", "dfl001", "pds-20", "ken-18", "osa-60"]


def main(out_path):
    try:
        hbm = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:
        hbm = 6531.6
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    rows = ["| instance | m x n | nnz | geometry | us / iteration | iterations/s | bytes / iteration | algorithmic GB/s | of measured HBM %.1f GB/s | solve mode us / iteration (check every 64) | create s | CPU oracle it/s (%d thr) | x rel. err vs oracle (K=200) |" % (hbm, os.cpu_count()),
            "|---|---|---:|---|---:|---:|---:|---:|---:|---:|---:|---:|---:|"]
    for name in NAMES:
        A, b, c = M.load_csr(name); m, n = A.shape
        lp = M.DeviceLP(A, A.data, m, n)
        eta = 0.9 / lp.sigma_max()
        bt, ct = torch.tensor(b, device="cuda"), torch.tensor(c, device="cuda")
        K = 2000
        M.pdhg_linear_program(A, A.data, bt, ct, num_iters=K, tau=eta, sigma=eta, handle=lp)
        ts = []
        for _ in range(5):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); M.pdhg_linear_program(A, A.data, bt, ct, num_iters=K, tau=eta, sigma=eta, handle=lp); e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        us = float(np.median(ts)) * 1e3 / K
        _, x, y, _ = M.pdhg_linear_program(A, A.data, b, c, num_iters=200, tau=eta, sigma=eta, handle=lp)
        t0 = time.perf_counter()
        kc = max(50, min(2000, int(3e8 / max(A.nnz, 1))))
        xo, yo = O.pdhg_run(A, b, c, np.zeros(n), np.zeros(m), eta, eta, kc)
        cpu = kc / (time.perf_counter() - t0)
        xo2, _ = O.pdhg_run(A, b, c, np.zeros(n), np.zeros(m), eta, eta, 200)
        err = np.linalg.norm(x - xo2) / max(np.linalg.norm(xo2), 1e-300)
        g = lp.geometry()
        bpi = lp.info()["bytes_per_iter"]
        # solve mode (k_solve_persistent / k_solve_cluster: Halpern combination, fixed-point error, KKT check + restart test
        # every 64 iterations, all on the device) at a tolerance that is never met: time per iteration of the solve loop
        ss = []
        for _ in range(3):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); M.solve_linear_program(A, A.data, bt, ct, tol=0.0, max_iters=K, check_every=64, eta=eta, handle=lp); e1.record()
            torch.cuda.synchronize()
            ss.append(e0.elapsed_time(e1))
        us_solve = float(np.median(ss)) * 1e3 / K
        rows.append("| %s | %dx%d | %d | %s x%d | %.2f | %.3g | %d | %.0f | %.3f | %.2f | %.2f | %.3g | %.1e |" % (
            name, m, n, A.nnz, "blocks" if lp.blocks_info()["used"] else g["mode"], g["ctas"], us, 1e6 / us, bpi, bpi / us / 1e3, bpi / us / 1e3 / hbm, us_solve, lp.create_s, cpu, err))
        print(rows[-1], flush=True)
        lp.close()
    open(out_path, "w").write("\n".join(rows) + "\n")


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/config_table.md")
