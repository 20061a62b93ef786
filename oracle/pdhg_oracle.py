"""ctypes front-end of oracle/pdhg_oracle.c (TEST INFRASTRUCTURE ONLY).

Inputs follow linear_program_data.py:58-80 of the reference: CSR (float64 data, int32
indices / indptr), rhs b (m,), coefs c (n,).  PARITY UNPINNED -- see the C header.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libpdhg_oracle.so")
_lib = None

_dp = ctypes.POINTER(ctypes.c_double)
_ip = ctypes.POINTER(ctypes.c_int32)


def build(force=False):
    src = os.path.join(_HERE, "pdhg_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-B", "-C", _HERE, "libpdhg_oracle.so"])
    return _SO


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(_SO)
        _lib.oracle_num_threads.restype = ctypes.c_int
    return _lib


def num_threads():
    return int(lib().oracle_num_threads())


def _d(a):
    return None if a is None else a.ctypes.data_as(_dp)


def _i(a):
    return a.ctypes.data_as(_ip)


def _f64(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.float64)


class CSR:
    """Holds contiguous CSR arrays of A in the reference's on-disk dtypes."""

    def __init__(self, A):
        import scipy.sparse as sp
        A = sp.csr_matrix(A)
        A.sort_indices()
        self.m, self.n = A.shape
        self.indptr = np.ascontiguousarray(A.indptr, dtype=np.int32)
        self.indices = np.ascontiguousarray(A.indices, dtype=np.int32)
        self.values = np.ascontiguousarray(A.data, dtype=np.float64)

    def args(self):
        return (ctypes.c_int(self.m), ctypes.c_int(self.n), _i(self.indptr), _i(self.indices),
                _d(self.values))


def spmv(A, v, trans=False, nthreads=0):
    A = A if isinstance(A, CSR) else CSR(A)
    v = _f64(v)
    out = np.empty(A.n if trans else A.m)
    rc = lib().oracle_spmv(*A.args(), ctypes.c_int(int(trans)), _d(v), _d(out), ctypes.c_int(nthreads))
    assert rc == 0
    return out


def pdhg_run(A, b, c, x, y, tau, sigma, num_iters, lb=None, ub=None, ylo=None, yhi=None,
             nthreads=0):
    A = A if isinstance(A, CSR) else CSR(A)
    b, c, lb, ub, ylo, yhi = map(_f64, (b, c, lb, ub, ylo, yhi))
    x = np.array(x, dtype=np.float64)
    y = np.array(y, dtype=np.float64)
    rc = lib().oracle_pdhg_run(*A.args(), _d(b), _d(c), _d(lb), _d(ub), _d(ylo), _d(yhi),
                               _d(x), _d(y), ctypes.c_double(tau), ctypes.c_double(sigma),
                               ctypes.c_int(num_iters), ctypes.c_int(nthreads))
    assert rc == 0
    return x, y


def kkt(A, b, c, x, y, lb=None, ub=None, ylo=None, yhi=None, nthreads=0):
    A = A if isinstance(A, CSR) else CSR(A)
    b, c, lb, ub, ylo, yhi, x, y = map(_f64, (b, c, lb, ub, ylo, yhi, x, y))
    out = np.zeros(10)
    rc = lib().oracle_kkt(*A.args(), _d(b), _d(c), _d(lb), _d(ub), _d(ylo), _d(yhi),
                          _d(x), _d(y), _d(out), ctypes.c_int(nthreads))
    assert rc == 0
    return out


def power_iteration(A, iters=50, nthreads=0):
    A = A if isinstance(A, CSR) else CSR(A)
    s = ctypes.c_double(0.0)
    rc = lib().oracle_power_iteration(*A.args(), ctypes.c_int(iters), ctypes.byref(s),
                                      ctypes.c_int(nthreads))
    assert rc == 0
    return s.value


def pdhg_solve(A, b, c, x, y, eta, w0=1.0, max_iters=100000, check_every=64, tol=1e-6,
               lb=None, ub=None, ylo=None, yhi=None, nthreads=0):
    A = A if isinstance(A, CSR) else CSR(A)
    b, c, lb, ub, ylo, yhi = map(_f64, (b, c, lb, ub, ylo, yhi))
    x = np.array(x, dtype=np.float64)
    y = np.array(y, dtype=np.float64)
    kk = np.zeros(10)
    info = np.zeros(4)
    rc = lib().oracle_pdhg_solve(*A.args(), _d(b), _d(c), _d(lb), _d(ub), _d(ylo), _d(yhi),
                                 _d(x), _d(y), ctypes.c_double(eta), ctypes.c_double(w0),
                                 ctypes.c_int(max_iters), ctypes.c_int(check_every),
                                 ctypes.c_double(tol), _d(kk), _d(info), ctypes.c_int(nthreads))
    assert rc == 0
    return x, y, kk, dict(iters=int(info[0]), restarts=int(info[1]), converged=bool(info[2]),
                          w=float(info[3]))
