"""MLLP_F_PRECONDITION: the diagonal preconditioner computed on the device inside mllp_lp_create / mllp_batch_create
(SURVEY.md section 8f rank 4).  The caller keeps speaking the ORIGINAL LP."""
import numpy as np
import pytest
import scipy.sparse as sp

import mllp_b200 as M
import mllp_b200.linear_program_data as D
from oracle import pdhg_oracle as O
from oracle.scaling_numpy import ruiz_pock_chambolle

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", ["afiro", "25fv47", "pilot87", "pds-20", "osa-60"])
def test_device_scaling_vectors_equal_the_numpy_restatement(name):
    A, b, c = D.load_csr(name)
    m, n = A.shape
    lp = M.DeviceLP(A, A.data, m, n, precondition=True, flags=M._cabi.F_NO_TUNE)
    dr, dc = lp.scaling()
    rr, rc = ruiz_pock_chambolle(A)
    assert np.max(np.abs(dr - rr) / rr) < 1e-12 and np.max(np.abs(dc - rc) / rc) < 1e-12
    # mllp_spmv stays a product with the ORIGINAL matrix
    import torch
    v = torch.tensor(np.random.default_rng(0).standard_normal(n), device="cuda")
    w = torch.tensor(np.random.default_rng(1).standard_normal(m), device="cuda")
    assert np.linalg.norm(lp.spmv(v).cpu().numpy() - A @ v.cpu().numpy()) <= 1e-12 * np.linalg.norm(A @ v.cpu().numpy())
    assert np.linalg.norm(lp.spmv(w, trans=True).cpu().numpy() - A.T @ w.cpu().numpy()) <= 1e-12 * np.linalg.norm(A.T @ w.cpu().numpy())
    lp.close()


@pytest.mark.parametrize("name", ["afiro", "sc105", "25fv47"])
def test_parity_run_on_a_preconditioned_handle_is_pdhg_on_the_scaled_lp(name):
    """mllp_pdhg_run on a preconditioned handle = the frozen iteration on Dr A Dc with b~, c~, mapped back: checked against
    the oracle run on the explicitly scaled LP; the KKT scalars are those of the ORIGINAL LP at the returned point."""
    A, b, c = D.load_csr(name)
    m, n = A.shape
    lp = M.DeviceLP(A, A.data, m, n, precondition=True)
    dr, dc = lp.scaling()
    As = (sp.diags(dr) @ A @ sp.diags(dc)).tocsr()
    eta = 0.9 / O.power_iteration(As, 50)
    obj, x, y, info = M.pdhg_linear_program(A, A.data, b, c, num_iters=300, tau=eta, sigma=eta, handle=lp)
    xs, ys = O.pdhg_run(As, dr * b, dc * c, np.zeros(n), np.zeros(m), eta, eta, 300)
    rel = lambda a, r: np.linalg.norm(a - r) / max(np.linalg.norm(r), 1e-300)
    assert rel(x, dc * xs) < 1e-9 and rel(y, dr * ys) < 1e-9
    kk = O.kkt(A, b, c, x, y)
    for k in range(10):
        assert abs(info[M.SCALAR_NAMES[k]] - kk[k]) <= 1e-6 * (1 + abs(kk[k])), (k, info[M.SCALAR_NAMES[k]], kk[k])
    lp.close()


@pytest.mark.parametrize("name,target", [("afiro", -46.2784021), ("sc105", -52.20206121), ("25fv47", 35.20396755), ("d2q06c", 237.9009627)])
def test_solve_terminates_on_the_original_lp(name, target):
    A, b, c = D.load_csr(name)
    obj, x, y, info = M.solve_linear_program(A, A.data, b, c, tol=1e-6, max_iters=1000000, precondition=True)
    assert info["converged"] and info["rel_kkt"] <= 1e-6
    kk = O.kkt(A, b, c, x, y)                       # the checker's KKT error of the returned point, original LP
    assert kk[8] <= 1.001e-6 and abs(kk[8] - info["rel_kkt"]) <= 1e-9
    assert abs(obj - target) <= 1e-5 * (1 + abs(target))
    # distance from the box is part of the residual: what is left of it is below the tolerance
    assert np.linalg.norm(np.minimum(x, 0.0)) <= 1e-6 * (1 + np.linalg.norm(b))


def test_handle_cache_does_not_mix_boxes():
    """two calls with the same weights array and different bounds must not share a handle"""
    A, b, c = D.load_csr("afiro")
    m, n = A.shape
    eta = 0.5
    lo, hi1, hi2 = np.zeros(n), np.full(n, 10.0), np.full(n, 1.0)
    _, x1, _, _ = M.pdhg_linear_program(A, A.data, b, c, num_iters=400, tau=eta, sigma=eta, lb=lo, ub=hi1)
    _, x2, _, _ = M.pdhg_linear_program(A, A.data, b, c, num_iters=400, tau=eta, sigma=eta, lb=lo, ub=hi2)
    xo1, _ = O.pdhg_run(A, b, c, np.zeros(n), np.zeros(m), eta, eta, 400, lo, hi1)
    xo2, _ = O.pdhg_run(A, b, c, np.zeros(n), np.zeros(m), eta, eta, 400, lo, hi2)
    assert np.linalg.norm(x1 - xo1) <= 1e-9 * np.linalg.norm(xo1) and np.linalg.norm(x2 - xo2) <= 1e-9 * np.linalg.norm(xo2)
    assert x2.max() <= 1.0 and x1.max() > 1.0


def test_pdlp_initial_primal_weight_option_follows_the_oracle():
    """primal_weight=None (w0 = 0 in the C ABI): ||c|| / ||b|| computed on the device; same rule in the oracle"""
    A, b, c = D.load_csr("sc50a")
    m, n = A.shape
    obj, x, y, info = M.solve_linear_program(A, A.data, b, c, tol=1e-30, max_iters=256, check_every=32, primal_weight=None)
    xs, ys, ks, si = O.pdhg_solve(A, b, c, np.zeros(n), np.zeros(m), info["eta"], w0=0.0, max_iters=256, tol=1e-30, check_every=32)
    rel = lambda a, r: np.linalg.norm(a - r) / max(np.linalg.norm(r), 1e-300)
    assert rel(x, xs) < 1e-7 and rel(y, ys) < 1e-7
    assert abs(info["primal_weight"] - si["w"]) <= 1e-9 * si["w"]
    res = M.solve_linear_program_batch([(A, A.data, b, c)] * 3, tol=1e-30, max_iters=256, check_every=32, primal_weight=None,
                                       eta=info["eta"])
    for r in res:
        assert rel(r[1], xs) < 1e-7 and rel(r[2], ys) < 1e-7


def test_portfolio_walks_the_settings_until_one_converges(monkeypatch):
    import os
    import mllp_b200.scaling as S
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "data", "netlib_mps_gz", "afiro.mps.gz")
    monkeypatch.setattr(S, "PORTFOLIO", ((64, 1.0, 128), (512, 1.0, 100000), (64, None, 100000)))   # the first cap is too small
    obj, x, y, info = S.solve_mps(path, portfolio=True)
    assert info["converged"] and info["attempt"] == 1 and info["setting"]["check_every"] == 512
    assert info["iters_all_attempts"] == 128 + info["iters"]
    assert abs(obj - (-464.7531429)) <= 1e-5 * 465
