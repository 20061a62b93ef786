"""GPU parity tests (run with -m gpu on a B200): the CUDA path, called through the C ABI,
against the CPU oracle on the same seeded inputs.

Tolerances (north_star): iterates within 1e-9 relative L2 after a fixed iteration count;
objective / KKT scalars within 1e-6 relative.  fp64 throughout.
"""
import json
import os

import numpy as np
import pytest
import scipy.sparse as sp

import mllp_b200 as M
import mllp_b200.linear_program_data as D
from mllp_b200 import _cabi
from mllp_b200.linear_program_methods import SCALAR_NAMES
from oracle import pdhg_oracle as O

pytestmark = pytest.mark.gpu

GOLD = os.path.join(os.path.dirname(__file__), "golden")
HIGHS = json.load(open(os.path.join(GOLD, "highs_objectives.json")))
ITER_TOL = 1e-9
SCAL_TOL = 1e-6

# (instance, iterations): every BASELINE.json config instance
CASES = [("afiro", 5000), ("sc50a", 1000), ("sc105", 1000), ("adlittle", 1000), ("blend", 1000), ("share2b", 1000),
         ("kb2", 1000), ("25fv47", 1000), ("pilot87", 1000), ("d2q06c", 1000), ("dfl001", 1000), ("ken-18", 1000),
         ("osa-60", 1000), ("pds-20", 1000)]


def rel(a, b):
    return np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300)


def scal_close(info, kk):
    for k in range(10):
        ref = kk[k]
        assert abs(info[SCALAR_NAMES[k]] - ref) <= SCAL_TOL * (1 + abs(ref)), (SCALAR_NAMES[k], info[SCALAR_NAMES[k]], ref)


@pytest.fixture(scope="module")
def torch_cuda():
    import torch
    assert torch.cuda.is_available(), "these tests need a GPU"
    return torch


@pytest.mark.parametrize("name", ["afiro", "25fv47", "pilot87", "ken-18", "osa-60", "pds-20"])
def test_spmv_parity_and_adjointness(name, torch_cuda):
    torch = torch_cuda
    A, _, _ = D.load_csr(name)
    m, n = A.shape
    lp = M.DeviceLP(A, A.data, m, n)
    rng = np.random.default_rng(1)
    v, w = rng.standard_normal(n), rng.standard_normal(m)
    Av = lp.spmv(torch.tensor(v, device="cuda")).cpu().numpy()
    Atw = lp.spmv(torch.tensor(w, device="cuda"), trans=True).cpu().numpy()
    assert rel(Av, O.spmv(A, v)) < 1e-13 and rel(Atw, O.spmv(A, w, trans=True)) < 1e-13
    # size-independent properties: adjointness and linearity
    assert abs(w @ Av - v @ Atw) <= 1e-11 * (np.linalg.norm(w) * np.linalg.norm(Av) + 1)
    v2 = rng.standard_normal(n)
    lin = lp.spmv(torch.tensor(2.0 * v - 3.0 * v2, device="cuda")).cpu().numpy()
    Av2 = lp.spmv(torch.tensor(v2, device="cuda")).cpu().numpy()
    assert rel(lin, 2.0 * Av - 3.0 * Av2) < 1e-12
    lp.close()


@pytest.mark.parametrize("name,K", CASES)
def test_parity_mode_iterates(name, K):
    A, b, c = D.load_csr(name)
    m, n = A.shape
    constrs = np.split(A.indices, A.indptr)[1:-1]
    lp = M.DeviceLP(constrs, A.data, m, n)
    sig = O.power_iteration(A, 50)
    assert abs(lp.sigma_max() - sig) <= 1e-10 * sig
    eta = 0.9 / sig
    obj, x, y, info = M.pdhg_linear_program(constrs, A.data, b, c, num_iters=K, tau=eta, sigma=eta, handle=lp)
    xo, yo = O.pdhg_run(A, b, c, np.zeros(n), np.zeros(m), eta, eta, K)
    assert rel(x, xo) < ITER_TOL and rel(y, yo) < ITER_TOL
    scal_close(info, O.kkt(A, b, c, xo, yo))
    assert info["iters"] == K and abs(obj - c @ xo) <= SCAL_TOL * (1 + abs(c @ xo))
    lp.close()


def test_golden_fixture_afiro():
    g = np.load(os.path.join(GOLD, "afiro_parity_K200.npz"))
    A, b, c = D.load_csr("afiro")
    eta = float(g["eta"])
    obj, x, y, info = M.pdhg_linear_program(A, A.data, b, c, num_iters=200, tau=eta, sigma=eta)
    assert rel(x, g["x"]) < ITER_TOL and rel(y, g["y"]) < ITER_TOL
    scal_close(info, g["kkt"])


@pytest.mark.parametrize("name", ["sc105", "25fv47", "pilot87"])
def test_general_form_bounds_and_row_senses(name):
    A, b, c = D.load_csr(name)
    m, n = A.shape
    rng = np.random.default_rng(7)
    lb = np.where(rng.random(n) < 0.3, -np.inf, rng.uniform(-1, 0, n))
    ub = np.where(rng.random(n) < 0.3, np.inf, rng.uniform(0.5, 2, n))
    kind = rng.integers(0, 3, m)
    ylo = np.where(kind == 1, 0.0, -np.inf)
    yhi = np.where(kind == 2, 0.0, np.inf)
    x0, y0 = rng.uniform(0, 0.4, n), np.clip(rng.standard_normal(m), ylo, yhi)
    eta = 0.9 / O.power_iteration(A, 50)
    obj, x, y, info = M.pdhg_linear_program(A, A.data, b, c, num_iters=400, tau=eta, sigma=0.5 * eta, lb=lb, ub=ub,
                                            ylo=ylo, yhi=yhi, x0=x0, y0=y0, handle=None)
    xo, yo = O.pdhg_run(A, b, c, x0, y0, eta, 0.5 * eta, 400, lb, ub, ylo, yhi)
    assert rel(x, xo) < ITER_TOL and rel(y, yo) < ITER_TOL
    scal_close(info, O.kkt(A, b, c, xo, yo, lb, ub, ylo, yhi))
    assert np.all(x >= lb) and np.all(x <= ub) and np.all(y >= ylo) and np.all(y <= yhi)


def test_zero_iterations_and_warm_start(torch_cuda):
    A, b, c = D.load_csr("blend")
    m, n = A.shape
    rng = np.random.default_rng(2)
    x0, y0 = np.abs(rng.standard_normal(n)), rng.standard_normal(m)
    obj, x, y, info = M.pdhg_linear_program(A, A.data, b, c, num_iters=0, tau=0.1, sigma=0.1, x0=x0, y0=y0)
    assert np.array_equal(x, x0) and np.array_equal(y, y0)
    scal_close(info, O.kkt(A, b, c, x0, y0))
    # K iterations == K/2 + K/2 with a warm start (bitwise: same kernels, same order)
    _, xa, ya, _ = M.pdhg_linear_program(A, A.data, b, c, num_iters=100, tau=0.1, sigma=0.1, x0=x0, y0=y0)
    _, xh, yh, _ = M.pdhg_linear_program(A, A.data, b, c, num_iters=50, tau=0.1, sigma=0.1, x0=x0, y0=y0)
    _, xb, yb, _ = M.pdhg_linear_program(A, A.data, b, c, num_iters=50, tau=0.1, sigma=0.1, x0=xh, y0=yh)
    assert np.array_equal(xa, xb) and np.array_equal(ya, yb)


@pytest.mark.parametrize("name", ["afiro", "pilot87", "osa-60"])
def test_graph_mode_is_bitwise_identical_to_persistent(name):
    A, b, c = D.load_csr(name)
    m, n = A.shape
    eta = 0.9 / O.power_iteration(A, 50)
    K = 70  # not a multiple of the graph unroll
    _, x1, y1, i1 = M.pdhg_linear_program(A, A.data, b, c, num_iters=K, tau=eta, sigma=eta)
    lpg = M.DeviceLP(A, A.data, m, n, flags=_cabi.F_GRAPH_MODE)
    _, x2, y2, i2 = M.pdhg_linear_program(A, A.data, b, c, num_iters=K, tau=eta, sigma=eta, handle=lpg)
    # iterates are bitwise equal (the summation order inside a row never depends on the kernel or on the
    # dealing of the tiles to CTAs); the KKT scalars are sums of per-CTA partials, so they can differ in the
    # last bits between two dealings (the persistent handle is tuned, the graph-mode one is not)
    assert np.array_equal(x1, x2) and np.array_equal(y1, y2)
    assert abs(i1["pobj"] - i2["pobj"]) <= 1e-13 * (1 + abs(i2["pobj"]))


def test_torch_tensor_interface_no_host_sync(torch_cuda):
    torch = torch_cuda
    A, b, c = D.load_csr("25fv47")
    m, n = A.shape
    eta = 0.9 / O.power_iteration(A, 50)
    bt, ct = torch.tensor(b, device="cuda"), torch.tensor(c, device="cuda")
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        obj, x, y, info = M.pdhg_linear_program(A, A.data, bt, ct, num_iters=300, tau=eta, sigma=eta)
    s.synchronize()
    assert x.is_cuda and y.is_cuda and obj.is_cuda
    xo, yo = O.pdhg_run(A, b, c, np.zeros(n), np.zeros(m), eta, eta, 300)
    assert rel(x.cpu().numpy(), xo) < ITER_TOL and rel(y.cpu().numpy(), yo) < ITER_TOL


@pytest.mark.parametrize("name", ["afiro", "sc50a", "sc105", "adlittle", "blend", "share2b"])
def test_solve_mode_objective_matches_highs_and_oracle(name):
    A, b, c = D.load_csr(name)
    m, n = A.shape
    obj, x, y, info = M.solve_linear_program(A, A.data, b, c, tol=1e-6, max_iters=400000)
    assert info["converged"] and info["rel_kkt"] <= 1e-6
    assert abs(obj - HIGHS[name]) <= 1e-4 * (1 + abs(HIGHS[name]))
    # returned scalars describe the returned point
    kk = O.kkt(A, b, c, x, y)
    assert abs(kk[0] - obj) <= SCAL_TOL * (1 + abs(obj)) and abs(kk[8] - info["rel_kkt"]) <= 1e-9
    if name in ("afiro", "sc50a", "sc105"):
        xs, ys, ks, si = O.pdhg_solve(A, b, c, np.zeros(n), np.zeros(m), info["eta"], max_iters=400000, tol=1e-6)
        assert abs(ks[0] - obj) <= SCAL_TOL * (1 + abs(obj))
        assert abs(ks[8] - info["rel_kkt"]) <= SCAL_TOL


def test_solve_mode_max_iters_reports_nonconvergence():
    A, b, c = D.load_csr("share2b")
    obj, x, y, info = M.solve_linear_program(A, A.data, b, c, tol=1e-12, max_iters=200)
    assert not info["converged"] and info["iters"] == 200


def test_error_paths():
    A, b, c = D.load_csr("afiro")
    with pytest.raises(ValueError):
        M.pdhg_linear_program(A, A.data, b[:-1], c, num_iters=1, tau=0.1, sigma=0.1)
    bad = A.copy()
    with pytest.raises(ValueError):
        M.DeviceLP(bad, bad.data, 27, 50)  # wrong column count
    lpg = M.DeviceLP(A, A.data, 27, 51, flags=_cabi.F_GRAPH_MODE)
    with pytest.raises(RuntimeError, match="persistent"):
        M.solve_linear_program(A, A.data, b, c, handle=lpg)


def test_random_skewed_matrix_with_empty_rows():
    """ragged input: empty rows/columns, a few very long rows (split path), random values."""
    rng = np.random.default_rng(11)
    rows = [sp.random(1, 5000, density=d, random_state=int(10000 * d) + 3, format="csr")
            for d in (0.0, 0.0004, 0.002, 0.01, 0.05, 0.2, 0.8)]
    A = sp.vstack(rows * 9).tocsr()
    A.data[:] = rng.standard_normal(A.nnz)
    A.sort_indices()
    m, n = A.shape
    b, c = rng.standard_normal(m), rng.standard_normal(n)
    eta = 0.9 / O.power_iteration(A, 50)
    obj, x, y, info = M.pdhg_linear_program(A, A.data, b, c, num_iters=200, tau=eta, sigma=eta)
    xo, yo = O.pdhg_run(A, b, c, np.zeros(n), np.zeros(m), eta, eta, 200)
    assert rel(x, xo) < ITER_TOL and rel(y, yo) < ITER_TOL


def test_device_graph_builder_matches_reference_loop(torch_cuda):
    """same tensors as the reference's build_graph_from_weights_sets
    (linear_program_methods.py:89-103), restated here as the reference's Python loop."""
    torch = torch_cuda
    from mllp_b200.graph import build_graph_from_weights_sets
    for name in ("afiro", "25fv47"):
        A, b, c = D.load_csr(name)
        constrs = np.split(A.indices, A.indptr)[1:-1]
        g = build_graph_from_weights_sets(constrs, A.data, b, c, device=0)
        index_1, index_2 = [], []
        for constr_idx, vars_ in enumerate(constrs):       # the reference's loop
            for var_index in vars_:
                index_1.append(var_index)
                index_2.append(constr_idx)
        ref_edge = torch.tensor([index_1, index_2])
        # the reference's BipartiteData attributes (:60-66)
        assert torch.equal(g.edge_index.cpu(), ref_edge)
        assert torch.equal(g.edge_attr.cpu(), torch.tensor(A.data, dtype=torch.float).unsqueeze(-1))
        assert torch.equal(g.x1.cpu(), torch.tensor(c, dtype=torch.float).unsqueeze(-1))
        assert torch.equal(g.x2.cpu(), torch.tensor(b, dtype=torch.float).unsqueeze(-1))
        assert g["x_src"] is g.x1 and g["x_tgt"] is g.x2
        assert torch.equal(g.__inc__("edge_index", None), torch.tensor([[c.shape[0]], [b.shape[0]]]))
        # and the message-passing forward takes it as the reference's model takes its graph (:238-239)
        import mllp_b200.gnn as GN
        from oracle import gnn_numpy as GO
        st = GO.init_state(3)
        logits = GN.GNNModel(st, device=0)(g).cpu().numpy()
        ref = GO.gnn_forward(st, A, b, c)
        assert np.max(np.abs(logits - ref)) <= 2e-4 * max(1.0, np.max(np.abs(ref)))


# Launch geometries of the persistent kernels (mllp_lp_geometry): cooperative grid, one cluster (hardware cluster
# barrier), one CTA, and the broadcast cluster (vector copies in distributed shared memory).  mllp_lp_create picks
# one by measurement; here each is forced and held to the same parity bar.
GEOMS = [("0", "grid"), ("16", "cluster"), ("4", "cluster"), ("1", "cta"), ("116", "bcast"), ("108", "bcast"), ("101", "bcast")]


@pytest.mark.parametrize("geom,mode", GEOMS)
def test_launch_geometries_parity_and_solve(geom, mode, monkeypatch):
    monkeypatch.setenv("MLLP_GEOM", geom)
    # parity mode, standard form
    A, b, c = D.load_csr("25fv47")
    m, n = A.shape
    lp = M.DeviceLP(A, A.data, m, n)
    assert lp.geometry()["mode"] == mode
    eta = 0.9 / O.power_iteration(A, 50)
    obj, x, y, info = M.pdhg_linear_program(A, A.data, b, c, num_iters=500, tau=eta, sigma=eta, handle=lp)
    xo, yo = O.pdhg_run(A, b, c, np.zeros(n), np.zeros(m), eta, eta, 500)
    assert rel(x, xo) < ITER_TOL and rel(y, yo) < ITER_TOL
    scal_close(info, O.kkt(A, b, c, xo, yo))
    # warm start on the same handle continues bitwise
    _, xa, ya, _ = M.pdhg_linear_program(A, A.data, b, c, num_iters=100, tau=eta, sigma=eta, x0=x, y0=y, handle=lp)
    _, xh, yh, _ = M.pdhg_linear_program(A, A.data, b, c, num_iters=50, tau=eta, sigma=eta, x0=x, y0=y, handle=lp)
    _, xb, yb, _ = M.pdhg_linear_program(A, A.data, b, c, num_iters=50, tau=eta, sigma=eta, x0=xh, y0=yh, handle=lp)
    assert np.array_equal(xa, xb) and np.array_equal(ya, yb)
    lp.close()
    # general form (bounds and row senses)
    A, b, c = D.load_csr("sc105")
    m, n = A.shape
    rng = np.random.default_rng(7)
    lb = np.where(rng.random(n) < 0.3, -np.inf, rng.uniform(-1, 0, n))
    ub = np.where(rng.random(n) < 0.3, np.inf, rng.uniform(0.5, 2, n))
    kind = rng.integers(0, 3, m)
    ylo, yhi = np.where(kind == 1, 0.0, -np.inf), np.where(kind == 2, 0.0, np.inf)
    x0, y0 = rng.uniform(0, 0.4, n), np.clip(rng.standard_normal(m), ylo, yhi)
    eta = 0.9 / O.power_iteration(A, 50)
    obj, x, y, info = M.pdhg_linear_program(A, A.data, b, c, num_iters=400, tau=eta, sigma=0.5 * eta, lb=lb, ub=ub,
                                            ylo=ylo, yhi=yhi, x0=x0, y0=y0)
    xo, yo = O.pdhg_run(A, b, c, x0, y0, eta, 0.5 * eta, 400, lb, ub, ylo, yhi)
    assert rel(x, xo) < ITER_TOL and rel(y, yo) < ITER_TOL
    # solve mode: same answer as the oracle's solve loop and HiGHS
    A, b, c = D.load_csr("afiro")
    m, n = A.shape
    obj, x, y, info = M.solve_linear_program(A, A.data, b, c, tol=1e-6, max_iters=400000)
    assert info["converged"] and abs(obj - HIGHS["afiro"]) <= 1e-4 * (1 + abs(HIGHS["afiro"]))
    xs, ys, ks, si = O.pdhg_solve(A, b, c, np.zeros(n), np.zeros(m), info["eta"], max_iters=400000, tol=1e-6)
    assert abs(ks[0] - obj) <= SCAL_TOL * (1 + abs(obj))
    kk = O.kkt(A, b, c, x, y)
    assert abs(kk[8] - info["rel_kkt"]) <= 1e-9


@pytest.mark.parametrize("geom", ["8", "108"])
def test_launch_geometries_split_rows(geom, monkeypatch):
    """rows cut into chunks and rows spanning several CTAs of the cluster (polled join) in the cluster geometries"""
    monkeypatch.setenv("MLLP_GEOM", geom)
    rng = np.random.default_rng(11)
    rows = [sp.random(1, 5000, density=d, random_state=int(10000 * d) + 3, format="csr")
            for d in (0.0, 0.0004, 0.002, 0.01, 0.05, 0.2, 0.8)]
    A = sp.vstack(rows * 9).tocsr()
    A.data[:] = rng.standard_normal(A.nnz)
    A.sort_indices()
    m, n = A.shape
    b, c = rng.standard_normal(m), rng.standard_normal(n)
    eta = 0.9 / O.power_iteration(A, 50)
    obj, x, y, info = M.pdhg_linear_program(A, A.data, b, c, num_iters=200, tau=eta, sigma=eta)
    xo, yo = O.pdhg_run(A, b, c, np.zeros(n), np.zeros(m), eta, eta, 200)
    assert rel(x, xo) < ITER_TOL and rel(y, yo) < ITER_TOL
    obj, x, y, info = M.solve_linear_program(A, A.data, b, c, tol=1e-30, max_iters=300, check_every=32)
    xs, ys, ks, si = O.pdhg_solve(A, b, c, np.zeros(n), np.zeros(m), info["eta"], max_iters=300, tol=1e-30, check_every=32)
    assert rel(x, xs) < 1e-7 and rel(y, ys) < 1e-7


def test_geometry_is_chosen_by_measurement():
    A, b, c = D.load_csr("afiro")
    lp = M.DeviceLP(A, A.data, *A.shape)
    g = lp.geometry()
    assert g["mode"] != "grid" and "grid" in g["ns_per_iter"] and min(g["ns_per_iter"].values()) < g["ns_per_iter"]["grid"]
    lp.close()
    A, b, c = D.load_csr("ken-18")       # too large for one cluster: not even tried
    lp = M.DeviceLP(A, A.data, *A.shape)
    assert lp.geometry()["mode"] == "grid" and lp.geometry()["ctas"] >= 148
    lp.close()


@pytest.mark.parametrize("name", ["dfl001", "pds-20"])
def test_preconditioned_solve_reaches_highs_optimum_on_config_instances(name):
    """mid-size / large config instances: Ruiz + Pock-Chambolle scaling on the device
    (MLLP_F_PRECONDITION), solve mode terminating on the KKT error of the ORIGINAL LP"""
    from mllp_b200.scaling import solve_scaled
    A, b, c = D.load_csr(name)
    obj, x, y, info = solve_scaled(A, b, c, tol=1e-6, max_iters=400000)
    assert info["converged"]
    assert abs(obj - HIGHS[name]) <= 1e-5 * (1 + abs(HIGHS[name]))
    assert info["rel_kkt"] <= 1e-6 and info["rel_kkt_original"] <= 1e-6
    # the returned point is a Halpern combination of reflected iterates: bounds hold to the tolerance, not exactly
    assert np.linalg.norm(np.minimum(x, 0.0)) <= 1e-6 * (1 + np.linalg.norm(x))
    kk = O.kkt(A, b, c, x, y)
    assert abs(kk[0] - obj) <= SCAL_TOL * (1 + abs(obj))
    assert kk[8] <= 1.001e-6 and abs(kk[8] - info["rel_kkt"]) <= 1e-9     # the oracle's KKT error of the returned point


def _block_angular(nblocks=400, seed=5):
    """random block-angular LP: nblocks independent blocks (5 x 8, one of them with an empty row and an unused column)
    coupled by 6 dense linking rows; a few columns occur in linking rows only"""
    rng = np.random.default_rng(seed)
    blocks = []
    for k in range(nblocks):
        Bk = sp.random(5, 8, density=0.45, random_state=seed + k, format="csr")
        Bk.data[:] = rng.standard_normal(Bk.nnz)
        blocks.append(Bk)
    D0 = sp.block_diag(blocks, format="csr")
    extra = sp.csr_matrix((D0.shape[0], 12))                       # columns seen by the linking rows only
    top = sp.hstack([D0, extra]).tocsr()
    link = sp.random(6, top.shape[1], density=0.3, random_state=seed + 9999, format="csr")
    link.data[:] = rng.standard_normal(link.nnz)
    A = sp.vstack([top[:700], link[:3], top[700:], link[3:]]).tocsr()   # linking rows in the middle and at the end
    A.sort_indices()
    return A, rng.standard_normal(A.shape[0]), rng.standard_normal(A.shape[1])


@pytest.mark.parametrize("name,K", [("ken-18", 600), ("synthetic", 300)])
def test_block_angular_kernel_matches_oracle_and_grid_kernel(name, K, monkeypatch):
    """blocks.cu: components dealt to CTAs, iterates in shared memory, linking rows through tagged words -- against the
    oracle and against the grid kernel of the same handle type; warm start continues bitwise"""
    A, b, c = _block_angular() if name == "synthetic" else D.load_csr(name)
    m, n = A.shape
    eta = 0.9 / O.power_iteration(A, 50)
    res = {}
    for mode in ("1", "0"):
        monkeypatch.setenv("MLLP_BLOCKS", mode)
        monkeypatch.setenv("MLLP_GEOM", "0")
        lp = M.DeviceLP(A, A.data, m, n)
        bi = lp.blocks_info()
        assert bi["used"] == (mode == "1") and (mode == "0" or (bi["found"] and bi["blocks"] >= 296 and bi["linking_rows"] >= 6))
        _, x, y, info = M.pdhg_linear_program(A, A.data, b, c, num_iters=K, tau=eta, sigma=eta, handle=lp)
        _, xh, yh, _ = M.pdhg_linear_program(A, A.data, b, c, num_iters=K // 2, tau=eta, sigma=eta, handle=lp)
        _, x2, y2, _ = M.pdhg_linear_program(A, A.data, b, c, num_iters=K - K // 2, tau=eta, sigma=eta, x0=xh, y0=yh, handle=lp)
        assert np.array_equal(x, x2) and np.array_equal(y, y2)
        res[mode] = (x, y, info)
        lp.close()
    xo, yo = O.pdhg_run(A, b, c, np.zeros(n), np.zeros(m), eta, eta, K)
    x, y, info = res["1"]
    assert rel(x, xo) < ITER_TOL and rel(y, yo) < ITER_TOL
    scal_close(info, O.kkt(A, b, c, xo, yo))
    assert rel(x, res["0"][0]) < 1e-12 and rel(y, res["0"][1]) < 1e-12


def test_block_angular_kernel_is_chosen_by_measurement_and_only_where_it_applies(monkeypatch):
    monkeypatch.delenv("MLLP_BLOCKS", raising=False)
    A, b, c = D.load_csr("ken-18")                  # 475 blocks + 151 linking rows: built and timed against the grid kernel
    lp = M.DeviceLP(A, A.data, *A.shape)
    bi = lp.blocks_info()
    assert bi["found"] and bi["blocks"] == 475 and bi["linking_rows"] == 151 and bi["ns_per_iter_grid"] > 0
    assert bi["used"] == (bi["ns_per_iter_blocks"] < 0.97 * bi["ns_per_iter_grid"])
    lp.close()
    A, b, c = D.load_csr("pds-20")                  # one giant component: no block structure
    lp = M.DeviceLP(A, A.data, *A.shape)
    assert not lp.blocks_info()["found"] and not lp.blocks_info()["used"]
    lp.close()
    A, b, c = D.load_csr("sc105")                   # general form keeps the grid / cluster kernels
    lp = M.DeviceLP(A, A.data, *A.shape, lb=np.zeros(A.shape[1]), ub=np.ones(A.shape[1]))
    assert not lp.blocks_info()["used"]
    lp.close()


def test_block_angular_kernel_general_form(monkeypatch):
    """boxes on x and sign cones on y (linking rows included) on the block kernel, warm start, against the oracle"""
    monkeypatch.setenv("MLLP_BLOCKS", "1")
    monkeypatch.setenv("MLLP_GEOM", "0")
    A, b, c = _block_angular(seed=8)
    m, n = A.shape
    rng = np.random.default_rng(7)
    lb = np.where(rng.random(n) < 0.3, -np.inf, rng.uniform(-1, 0, n))
    ub = np.where(rng.random(n) < 0.3, np.inf, rng.uniform(0.5, 2, n))
    kind = rng.integers(0, 3, m)
    ylo, yhi = np.where(kind == 1, 0.0, -np.inf), np.where(kind == 2, 0.0, np.inf)
    x0, y0 = rng.uniform(0, 0.4, n), np.clip(rng.standard_normal(m), ylo, yhi)
    eta = 0.9 / O.power_iteration(A, 50)
    lp = M.DeviceLP(A, A.data, m, n, lb=lb, ub=ub, ylo=ylo, yhi=yhi)
    assert lp.blocks_info()["used"]
    obj, x, y, info = M.pdhg_linear_program(A, A.data, b, c, num_iters=300, tau=eta, sigma=0.5 * eta, lb=lb, ub=ub,
                                            ylo=ylo, yhi=yhi, x0=x0, y0=y0, handle=lp)
    xo, yo = O.pdhg_run(A, b, c, x0, y0, eta, 0.5 * eta, 300, lb, ub, ylo, yhi)
    assert rel(x, xo) < ITER_TOL and rel(y, yo) < ITER_TOL
    scal_close(info, O.kkt(A, b, c, xo, yo, lb, ub, ylo, yhi))
    assert np.all(x >= lb) and np.all(x <= ub) and np.all(y >= ylo) and np.all(y <= yhi)
    lp.close()
