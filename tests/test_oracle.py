"""CPU tests of the oracle (test infrastructure) against its independent pins."""
import json
import os

import numpy as np
import pytest

import mllp_b200.linear_program_data as D
from oracle import pdhg_numpy as P
from oracle import pdhg_oracle as O

GOLD = os.path.join(os.path.dirname(__file__), "golden")
HIGHS = json.load(open(os.path.join(GOLD, "highs_objectives.json")))
KAT = json.load(open(os.path.join(GOLD, "basis_kat.json")))


def rel(a, b):
    return np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300)


@pytest.mark.parametrize("name", ["afiro", "sc50a", "blend", "25fv47"])
def test_c_oracle_matches_numpy_restatement(name):
    A, b, c = D.load_csr(name)
    m, n = A.shape
    assert abs(O.power_iteration(A, 50) - P.power_iteration(A, 50)) < 1e-12
    eta = 0.9 / O.power_iteration(A, 50)
    rng = np.random.default_rng(3)
    x0, y0 = np.abs(rng.standard_normal(n)), rng.standard_normal(m)
    x1, y1 = O.pdhg_run(A, b, c, x0, y0, eta, 2 * eta, 300)
    x2, y2 = P.pdhg_run(A, b, c, x0, y0, eta, 2 * eta, 300)
    assert rel(x1, x2) < 1e-12 and rel(y1, y2) < 1e-12
    assert np.allclose(O.kkt(A, b, c, x1, y1), P.kkt(A, b, c, x1, y1), rtol=1e-10, atol=1e-10)
    v = rng.standard_normal(n)
    w = rng.standard_normal(m)
    assert rel(O.spmv(A, v), A @ v) < 1e-14 and rel(O.spmv(A, w, trans=True), A.T @ w) < 1e-14


def test_oracle_thread_count_does_not_change_results():
    A, b, c = D.load_csr("25fv47")
    m, n = A.shape
    eta = 0.9 / O.power_iteration(A, 50)
    x1, y1 = O.pdhg_run(A, b, c, np.zeros(n), np.zeros(m), eta, eta, 100, nthreads=1)
    x2, y2 = O.pdhg_run(A, b, c, np.zeros(n), np.zeros(m), eta, eta, 100, nthreads=4)
    assert np.array_equal(x1, x2) and np.array_equal(y1, y2)


def test_oracle_general_form_matches_numpy():
    """bounds l<=x<=u and row senses (dual boxes), incl. infinite sides."""
    A, b, c = D.load_csr("sc105")
    m, n = A.shape
    rng = np.random.default_rng(5)
    lb = np.where(rng.random(n) < 0.3, -np.inf, rng.uniform(-1, 0, n))
    ub = np.where(rng.random(n) < 0.3, np.inf, rng.uniform(0.5, 2, n))
    kind = rng.integers(0, 3, m)
    ylo = np.where(kind == 1, 0.0, -np.inf)
    yhi = np.where(kind == 2, 0.0, np.inf)
    eta = 0.5
    x1, y1 = O.pdhg_run(A, b, c, np.zeros(n), np.zeros(m), eta, eta, 200, lb, ub, ylo, yhi)
    x2, y2 = P.pdhg_run(A, b, c, np.zeros(n), np.zeros(m), eta, eta, 200, lb, ub, ylo, yhi)
    assert rel(x1, x2) < 1e-12 and rel(y1, y2) < 1e-12
    assert np.allclose(O.kkt(A, b, c, x1, y1, lb, ub, ylo, yhi), P.kkt(A, b, c, x1, y1, lb, ub, ylo, yhi),
                       rtol=1e-10, atol=1e-10)


def test_golden_afiro_iterates():
    g = np.load(os.path.join(GOLD, "afiro_parity_K200.npz"))
    A, b, c = D.load_csr("afiro")
    eta = float(g["eta"])
    assert abs(eta - 0.9 / O.power_iteration(A, 50)) < 1e-15
    x, y = O.pdhg_run(A, b, c, np.zeros(A.shape[1]), np.zeros(A.shape[0]), eta, eta, 200)
    assert rel(x, g["x"]) < 1e-13 and rel(y, g["y"]) < 1e-13
    assert np.allclose(O.kkt(A, b, c, x, y), g["kkt"], rtol=1e-11, atol=1e-12)


@pytest.mark.parametrize("name", ["afiro", "sc50a", "sc105", "blend"])
def test_solve_mode_reaches_highs_objective(name):
    """The frozen solve-mode spec converges to the independent HiGHS optimum (1e-4 rel at
    KKT tol 1e-6), on the LP the reference's arrays define."""
    A, b, c = D.load_csr(name)
    m, n = A.shape
    eta = 0.99 / O.power_iteration(A, 50)
    x, y, kk, info = O.pdhg_solve(A, b, c, np.zeros(n), np.zeros(m), eta, max_iters=100000, tol=1e-6)
    assert info["converged"]
    assert abs(kk[0] - HIGHS[name]) <= 1e-4 * (1 + abs(HIGHS[name]))
    x2, y2, kk2, info2 = P.pdhg_solve(A, b, c, np.zeros(n), np.zeros(m), eta, max_iters=100000, tol=1e-6)
    assert info2["iters"] == info["iters"] and info2["restarts"] == info["restarts"]
    assert abs(kk2[0] - kk[0]) < 1e-9


def test_basis_labels_reproduce_highs_objective():
    """The reference's optimal-basis labels (linear_program_data.py:72) and HiGHS agree."""
    n_checked = 0
    for name, val in KAT.items():
        if HIGHS.get(name) is not None:
            assert abs(val - HIGHS[name]) <= 1e-6 * (1 + abs(HIGHS[name])), name
            n_checked += 1
    assert n_checked >= 6


def test_oracle_edge_cases():
    import scipy.sparse as sp
    # empty rows and columns, zero iterations
    A = sp.csr_matrix(np.array([[1.0, 0, 2.0, 0], [0, 0, 0, 0], [0, 0, 3.0, 0]]))
    b = np.array([1.0, 0.0, 2.0])
    c = np.array([1.0, 1.0, 1.0, 1.0])
    x, y = O.pdhg_run(A, b, c, np.ones(4), np.ones(3), 0.1, 0.1, 0)
    assert np.array_equal(x, np.ones(4)) and np.array_equal(y, np.ones(3))
    x, y = O.pdhg_run(A, b, c, np.zeros(4), np.zeros(3), 0.1, 0.1, 50)
    x2, y2 = P.pdhg_run(A, b, c, np.zeros(4), np.zeros(3), 0.1, 0.1, 50)
    assert rel(x, x2) < 1e-13 and rel(y, y2) < 1e-13
    assert y[1] == 0.0  # empty row with b = 0 never moves


def test_config5_perturbation_keeps_the_lp_bounded_where_the_survey_recipe_does_not():
    """BASELINE.json configs[4]: the generator of the 4096-instance batch (bench.config5_batches).  SURVEY 8d's cost
    perturbation c (1 + 0.1 U(-1, 1)) makes 25fv47 (`_norm` form: no upper bounds) unbounded; the variant the bench solves
    moves every cost away from the zero-cost recession directions and keeps the LP feasible and bounded (HiGHS)."""
    import os
    import sys
    import scipy.optimize
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import bench
    import mllp_b200.linear_program_data as D
    A, b, c = D.load_csr("25fv47")
    bs, cs = bench.config5_batches(A, b, c, 0, 1, variant="survey")
    bb, cb = bench.config5_batches(A, b, c, 0, 1, variant="bounded")
    assert np.array_equal(bs, bb)                                     # same right-hand sides, same random stream
    st_survey = [scipy.optimize.linprog(cs[i], A_eq=A, b_eq=bs[i], bounds=(0, None), method="highs").status for i in range(1)]
    res = [scipy.optimize.linprog(cb[i], A_eq=A, b_eq=bb[i], bounds=(0, None), method="highs") for i in range(1)]
    assert st_survey == [3]                                           # unbounded
    assert [r.status for r in res] == [0] and all(30.0 < r.fun < 45.0 for r in res)     # optimum of 25fv47: 35.204
    sr = bench.slack_rows(A, c)
    assert sr.size == 305 and np.all(bb[:, sr] >= b[sr] - 1e-15 * np.abs(b[sr])) or np.all(np.sign(bb[:, sr]) == np.sign(b[sr]))
