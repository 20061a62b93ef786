"""Solve mode on the BASELINE.json config instances (the reference's `_norm` dataset arrays: min c'x, Ax = b, x >= 0)
against the HiGHS optimum of the same arrays (tests/golden/highs_objectives.json).  Writes a markdown table.
ken-18 is unbounded in this form and is skipped."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import mllp_b200 as M
from mllp_b200.scaling import solve_scaled

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HIGHS = json.load(open(os.path.join(ROOT, "tests", "golden", "highs_objectives.json")))
NAMES = ["afiro", "sc50a", "sc105", "adlittle", "blend", "share2b", "25fv47", "pilot87", "d2q06c", "dfl001", "osa-60", "pds-20"]


def main(out_path, max_iters):
    rows = ["| instance | m x n | nnz | objective (this build) | HiGHS on the same arrays | rel. error | iterations | converged (rel KKT <= 1e-6) | rel KKT (unscaled LP) | s |",
            "|---|---|---:|---:|---:|---:|---:|---|---:|---:|"]
    for name in NAMES:
        A, b, c = M.load_csr(name)
        t0 = time.perf_counter()
        obj, x, y, info = solve_scaled(A, b, c, tol=1e-6, max_iters=max_iters)
        dt = time.perf_counter() - t0
        ref = HIGHS[name]
        rows.append("| %s | %dx%d | %d | %.9g | %.9g | %.1e | %d | %s | %.1e | %.2f |" % (
            name, A.shape[0], A.shape[1], A.nnz, obj, ref, abs(obj - ref) / (1 + abs(ref)), info["iters"],
            "yes" if info["converged"] else "no", info["rel_kkt_original"], dt))
        print(rows[-1], flush=True)
    open(out_path, "w").write("\n".join(rows) + "\n")


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/solve_configs.md", int(sys.argv[2]) if len(sys.argv) > 2 else 2000000)
