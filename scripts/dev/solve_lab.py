"""dev lab: restart / primal-weight rule variants of solve mode on the hard Netlib files, on the CPU (scripts/dev/solve_lab.c).
usage: python scripts/dev/solve_lab.py [max_iters]"""
import ctypes, os, subprocess, sys, tempfile, time
import numpy as np, scipy.sparse as sp
from multiprocessing import Pool
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from mllp_b200.mps import read_mps
from oracle.scaling_numpy import ruiz_pock_chambolle
from oracle import pdhg_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
_SO = os.path.join(tempfile.gettempdir(), "mllp_solve_lab.so")
if not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(os.path.join(ROOT, "scripts", "dev", "solve_lab.c")):
    subprocess.check_call(["gcc", "-O3", "-march=native", "-fopenmp", "-shared", "-fPIC", "-o", _SO,
                           os.path.join(ROOT, "scripts", "dev", "solve_lab.c"), "-lm"])
LAB = ctypes.CDLL(_SO)
dp = ctypes.POINTER(ctypes.c_double); ip = ctypes.POINTER(ctypes.c_int32)

def prep(name):
    lp = read_mps(os.path.join(ROOT, "data", "netlib_mps_gz", name + ".mps.gz"))
    A = sp.csr_matrix(lp["A"]); A.sort_indices()
    dr, dc = ruiz_pock_chambolle(A)
    As = sp.csr_matrix(sp.diags(dr) @ A @ sp.diags(dc)); As.sort_indices()
    f = lambda v, s: None if v is None else np.ascontiguousarray(np.asarray(v, dtype=np.float64) / s)
    return dict(A=A, As=As, b=lp["b"], c=lp["c"], bs=dr * lp["b"], cs=dc * lp["c"], lb=f(lp["lb"], dc), ub=f(lp["ub"], dc),
                ylo=f(lp["ylo"], dr), yhi=f(lp["yhi"], dr), dr=dr, dc=dc, lp=lp)

def run(job):
    name, var, max_iters = job
    P = prep(name)
    As = O.CSR(P["As"])
    sig = O.power_iteration(As, iters=400, nthreads=1) * 1.02
    eta = 0.99 / sig
    x = np.zeros(As.n); y = np.zeros(As.m); kk = np.zeros(10); info = np.zeros(4)
    d = lambda a: None if a is None else a.ctypes.data_as(dp)
    t = time.time()
    LAB.lab_solve(*As.args(), d(np.ascontiguousarray(P["bs"])), d(np.ascontiguousarray(P["cs"])), d(P["lb"]), d(P["ub"]), d(P["ylo"]), d(P["yhi"]),
                  d(x), d(y), ctypes.c_double(eta), ctypes.c_double(var.get("w0", 1.0)), ctypes.c_int(max_iters),
                  ctypes.c_int(var.get("ce", 64)), ctypes.c_double(1e-6), d(kk), d(info), ctypes.c_int(1),
                  ctypes.c_double(var.get("kp", 0.5)), ctypes.c_double(var.get("ki", 0.0)), ctypes.c_double(var.get("kd", 0.0)),
                  ctypes.c_int(var.get("cross", 0)), ctypes.c_double(var.get("bsuf", 0.2)), ctypes.c_double(var.get("bnec", 0.8)),
                  ctypes.c_double(var.get("bart", 0.36)))
    xo, yo = x * P["dc"], y * P["dr"]
    lp = P["lp"]
    ko = O.kkt(P["A"], lp["b"], lp["c"], xo, yo, lb=lp["lb"], ub=lp["ub"], ylo=lp["ylo"], yhi=lp["yhi"], nthreads=1)
    return name, var, int(info[0]), int(info[1]), bool(info[2]), kk[8], ko[8], time.time() - t

if __name__ == "__main__":
    max_iters = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
    names = sys.argv[2].split(",") if len(sys.argv) > 2 else ["bnl1", "pilot4", "perold", "pilot.we", "pilot.ja", "greenbea", "pilot"]
    variants = [dict(), dict(kp=0.99, ki=0.96), dict(kp=0.99, ki=0.0), dict(cross=1), dict(kp=0.99, ki=0.96, cross=1)]
    if len(sys.argv) > 3:
        variants = eval(sys.argv[3])
    jobs = [(n, v, max_iters) for n in names for v in variants]
    with Pool(8) as pool:
        for name, var, it, rs, conv, ks, ko, dt in pool.imap_unordered(run, jobs):
            print("%-9s %-40s iters %8d restarts %4d conv %d kkt_scaled %.2e kkt_orig %.2e (%.0f s)" % (name, var, it, rs, conv, ks, ko, dt), flush=True)
