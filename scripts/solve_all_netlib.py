"""Solve mode on every Netlib MPS instance the repo carries (data/netlib_mps_gz/, the 97 files of the
reference's netlib_mps/) through solve_mps(): objective vs the HiGHS optimum of the same file
(tests/golden/mps_models_all.json).  Writes a markdown table + JSON.

    python scripts/solve_all_netlib.py [--names a,b,c] [--max-iters N] [--tol T] [--out PREFIX] [--budget SECONDS]
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--names", default="")
    ap.add_argument("--max-iters", type=int, default=2000000)
    ap.add_argument("--tol", type=float, default=1e-6)
    ap.add_argument("--check-every", type=int, default=64)
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "netlib_all"))
    ap.add_argument("--budget", type=float, default=1e9, help="stop starting new instances after this many seconds")
    ap.add_argument("--portfolio", action="store_true", help="try the settings of mllp_b200.scaling.PORTFOLIO in turn")
    a = ap.parse_args()
    from mllp_b200.scaling import solve_mps
    gold = json.load(open(os.path.join(ROOT, "tests", "golden", "mps_models_all.json")))
    d = os.path.join(ROOT, "data", "netlib_mps_gz")
    names = a.names.split(",") if a.names else sorted(gold, key=lambda k: gold[k]["nnz"])
    rows, t0 = [], time.time()
    for nm in names:
        if time.time() - t0 > a.budget:
            print("budget reached before", nm, flush=True)
            break
        g = gold[nm]
        t = time.time()
        try:
            obj, x, y, info = solve_mps(os.path.join(d, nm + ".mps.gz"), tol=a.tol, max_iters=a.max_iters,
                                        check_every=a.check_every, portfolio=a.portfolio)
        except Exception as e:  # keep going: the table must show failures too
            print(nm, "FAILED", repr(e), flush=True)
            rows.append({"name": nm, "error": repr(e)})
            continue
        dt = time.time() - t
        ref = g["objective"]
        r = {"name": nm, "m": g["m"], "n": g["n"], "nnz": g["nnz"], "objective": obj, "highs": ref,
             "rel_err": abs(obj - ref) / (1 + abs(ref)), "iters": int(info["iters"]), "restarts": int(info["restarts"]),
             "converged": bool(info["converged"]), "rel_kkt": float(info["rel_kkt"]),
             "rel_kkt_original": float(info["rel_kkt_original"]), "seconds": dt, "attempt": int(info.get("attempt", 0)),
             "iters_all_attempts": int(info.get("iters_all_attempts", info["iters"]))}
        rows.append(r)
        print("%-10s obj %.9g ref %.9g relerr %.2e iters %d conv %s kkt %.1e/%.1e %.2fs" % (
            nm, obj, ref, r["rel_err"], r["iters"], r["converged"], r["rel_kkt"], r["rel_kkt_original"], dt), flush=True)
    os.makedirs(os.path.dirname(a.out), exist_ok=True)
    json.dump({"tol": a.tol, "max_iters": a.max_iters, "rows": rows}, open(a.out + ".json", "w"), indent=1)
    ok = [r for r in rows if "error" not in r]
    with open(a.out + ".md", "w") as f:
        f.write("| instance | m x n | nnz | objective (this build) | HiGHS on the MPS file | rel. error | iterations | "
                "converged (rel KKT <= %g) | rel KKT (original LP) | s |\n|---|---|---:|---:|---:|---:|---:|---|---:|---:|\n" % a.tol)
        for r in ok:
            f.write("| %s | %dx%d | %d | %.9g | %.9g | %.1e | %d | %s | %.1e | %.2f |\n" % (
                r["name"], r["m"], r["n"], r["nnz"], r["objective"], r["highs"], r["rel_err"], r["iters"],
                "yes" if r["converged"] else "no", r["rel_kkt_original"], r["seconds"]))
        n4 = sum(1 for r in ok if r["rel_err"] <= 1e-4)
        n6 = sum(1 for r in ok if r["rel_err"] <= 1e-6)
        f.write("\n%d instances attempted, %d converged, %d within 1e-4 and %d within 1e-6 (relative, 1+|ref|) of the HiGHS "
                "objective; total %.1f s.\n" % (len(rows), sum(1 for r in ok if r["converged"]), n4, n6, time.time() - t0))
    print(open(a.out + ".md").read().splitlines()[-1])


if __name__ == "__main__":
    main()
