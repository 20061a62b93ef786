"""GPU parity tests of the batched multi-instance mode (BASELINE.json configs[1] and [4])."""
import json
import os

import numpy as np
import pytest

import mllp_b200 as M
import mllp_b200.linear_program_data as D
from oracle import pdhg_oracle as O

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")
HIGHS = json.load(open(os.path.join(GOLD, "highs_objectives.json")))
SMALL = ["sc50a", "sc105", "adlittle", "blend", "share2b", "kb2"]


def rel(a, b):
    return np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300)


def load_batch(names):
    insts, mats = [], []
    for nm in names:
        A, b, c = D.load_csr(nm)
        constrs = np.split(A.indices, A.indptr)[1:-1]   # loader representation
        insts.append((constrs, A.data, b, c))
        mats.append((A, b, c))
    return insts, mats


def test_small_netlib_batch_parity_one_launch():
    """configs[1]: sc50a, sc105, adlittle, blend, share2b, kb2 packed into one launch, fp64."""
    insts, mats = load_batch(SMALL)
    bt = M.BatchLP(insts)
    info = bt.info()
    assert info["count"] == 6 and info["sum_m"] == 424 and info["sum_n"] == 723 and info["sum_nnz"] == 2536
    sig = bt.sigma_max().cpu().numpy()
    K = 1000
    res = M.pdhg_linear_program_batch(insts, num_iters=K, handle=bt)
    for (A, b, c), (obj, x, y, inf), s in zip(mats, res, sig):
        assert abs(s - O.power_iteration(A, 50)) <= 1e-10 * s
        eta = 0.9 / s
        xo, yo = O.pdhg_run(A, b, c, np.zeros(A.shape[1]), np.zeros(A.shape[0]), eta, eta, K)
        assert rel(x, xo) < 1e-9 and rel(y, yo) < 1e-9
        kk = O.kkt(A, b, c, xo, yo)
        assert abs(obj - kk[0]) <= 1e-6 * (1 + abs(kk[0])) and abs(inf["rel_kkt"] - kk[8]) <= 1e-6 * (1 + kk[8])


def test_batch_warm_start_and_per_instance_steps():
    insts, mats = load_batch(["afiro", "blend", "sc50a"])
    rng = np.random.default_rng(4)
    x0 = [np.abs(rng.standard_normal(A.shape[1])) for A, _, _ in mats]
    y0 = [rng.standard_normal(A.shape[0]) for A, _, _ in mats]
    tau, sigma = np.array([0.05, 0.1, 0.2]), np.array([0.3, 0.2, 0.1])
    res = M.pdhg_linear_program_batch(insts, num_iters=150, x0=x0, y0=y0, tau=tau, sigma=sigma)
    for k, ((A, b, c), (obj, x, y, inf)) in enumerate(zip(mats, res)):
        xo, yo = O.pdhg_run(A, b, c, x0[k], y0[k], tau[k], sigma[k], 150)
        assert rel(x, xo) < 1e-9 and rel(y, yo) < 1e-9
    res0 = M.pdhg_linear_program_batch(insts, num_iters=0, x0=x0, y0=y0, tau=0.1, sigma=0.1)
    for k in range(3):
        assert np.array_equal(res0[k][1], x0[k]) and np.array_equal(res0[k][2], y0[k])


def test_batch_solve_mode_matches_highs_and_single_instance_path():
    names = ["sc50a", "sc105", "adlittle", "blend", "share2b"]   # kb2 is unbounded without its MPS bounds
    insts, mats = load_batch(names)
    res = M.solve_linear_program_batch(insts, tol=1e-6, max_iters=400000)
    for nm, (A, b, c), (obj, x, y, inf) in zip(names, mats, res):
        assert inf["converged"] and inf["rel_kkt"] <= 1e-6
        assert abs(obj - HIGHS[nm]) <= 1e-4 * (1 + abs(HIGHS[nm]))
        kk = O.kkt(A, b, c, x, y)
        assert abs(kk[0] - obj) <= 1e-6 * (1 + abs(obj))
    # same algorithm as the grid-wide solver: identical iteration counts on a well-conditioned case
    A, b, c = mats[0]
    _, _, _, single = M.solve_linear_program(A, A.data, b, c, tol=1e-6)
    assert single["iters"] == res[0][3]["iters"] and single["restarts"] == res[0][3]["restarts"]


def test_shared_matrix_batch_of_perturbed_instances():
    """configs[4] (scaled down): random b/c perturbations of a fixed Netlib A (25fv47)."""
    A, b, c = D.load_csr("25fv47")
    m, n = A.shape
    B = 300   # more instances than CTAs: exercises the grid-stride loop
    cb = np.empty((B, n))
    bb = np.empty((B, m))
    for i in range(B):
        g = np.random.default_rng(1234 + i)
        cb[i] = c * (1 + 0.1 * g.uniform(-1, 1, n))
        bb[i] = b * (1 + 0.1 * g.uniform(0, 1, m))
    bt = M.BatchLP([(A, A.data, b, c)], shared=True, count=B)
    assert bt.info()["sum_nnz"] == A.nnz and bt.info()["count"] == B
    res = M.pdhg_linear_program_batch([(A, A.data, b, c)], num_iters=200, handle=bt, shared=True, rhs_batch=bb,
                                      coefs_batch=cb)
    eta = 0.9 / O.power_iteration(A, 50)
    for k in (0, 1, 147, 148, 149, B - 1):
        xo, yo = O.pdhg_run(A, bb[k], cb[k], np.zeros(n), np.zeros(m), eta, eta, 200)
        assert rel(res[k][1], xo) < 1e-9 and rel(res[k][2], yo) < 1e-9
    # linearity in (b, c) of the unprojected part is not available (projection), but instances are
    # independent: permuting the batch permutes the results
    perm = np.random.default_rng(0).permutation(B)
    res2 = M.pdhg_linear_program_batch([(A, A.data, b, c)], num_iters=200, handle=bt, shared=True,
                                       rhs_batch=bb[perm], coefs_batch=cb[perm])
    for k in (0, 5, B - 1):
        assert np.array_equal(res2[k][1], res[perm[k]][1]) and np.array_equal(res2[k][2], res[perm[k]][2])


@pytest.mark.parametrize("name,B,RS", [("sc105", 37, (1, 2, 3, 4)), ("25fv47", 13, (1, 2, 3))])   # 4 x 25fv47 exceeds smem
def test_shared_matrix_multi_rhs_is_bitwise_independent_of_group_size(name, B, RS, monkeypatch):
    """R instances per CTA share every matrix step (multi-RHS); per instance the summation order is the one of
    the one-instance walker, so iterates are bitwise equal for R = 1, 2, 3, 4 (B is not a multiple of R: tail group)."""
    A, b, c = D.load_csr(name)
    m, n = A.shape
    rng = np.random.default_rng(7)
    cb = c * (1 + 0.1 * rng.uniform(-1, 1, (B, n)))
    bb = b * (1 + 0.1 * rng.uniform(0, 1, (B, m)))
    out = {}
    for R in RS:
        monkeypatch.setenv("MLLP_BATCH_R", str(R))
        bt = M.BatchLP([(A, A.data, b, c)], shared=True, count=B)
        assert bt.info()["instances_per_cta"] == R
        out[R] = M.pdhg_linear_program_batch([(A, A.data, b, c)], num_iters=120, handle=bt, shared=True, rhs_batch=bb,
                                             coefs_batch=cb)
        bt.close()
    eta = 0.9 / O.power_iteration(A, 50)
    for k in (0, B // 2, B - 1):
        xo, yo = O.pdhg_run(A, bb[k], cb[k], np.zeros(n), np.zeros(m), eta, eta, 120)
        assert rel(out[1][k][1], xo) < 1e-9 and rel(out[1][k][2], yo) < 1e-9
        kk = O.kkt(A, bb[k], cb[k], xo, yo)
        for R in RS:
            inf = out[R][k][3]
            assert abs(out[R][k][0] - kk[0]) <= 1e-6 * (1 + abs(kk[0])) and abs(inf["rel_kkt"] - kk[8]) <= 1e-6 * (1 + kk[8])
    for R in RS[1:]:
        for k in range(B):
            assert np.array_equal(out[R][k][1], out[1][k][1]) and np.array_equal(out[R][k][2], out[1][k][2])


def test_shared_matrix_solve_mode_multi_rhs(monkeypatch):
    """solve mode, lockstep groups of R instances with per-instance restarts / termination vs the one-instance solver"""
    A, b, c = D.load_csr("sc105")
    m, n = A.shape
    B = 9
    rng = np.random.default_rng(11)
    cb = c * (1 + 0.05 * rng.uniform(-1, 1, (B, n)))
    bb = np.tile(b, (B, 1))
    res = {}
    for R in (1, 2, 4):
        monkeypatch.setenv("MLLP_BATCH_R", str(R))
        bt = M.BatchLP([(A, A.data, b, c)], shared=True, count=B)
        assert bt.info()["instances_per_cta_solve"] == R
        res[R] = M.solve_linear_program_batch([(A, A.data, b, c)], tol=1e-6, max_iters=200000, handle=bt, shared=True,
                                              rhs_batch=bb, coefs_batch=cb)
        bt.close()
    for k in range(B):
        for R in (1, 2, 4):
            obj, x, y, inf = res[R][k]
            assert inf["converged"] and inf["rel_kkt"] <= 1e-6
            kk = O.kkt(A, bb[k], cb[k], x, y)
            assert abs(kk[0] - obj) <= 1e-6 * (1 + abs(obj)) and kk[8] <= 1.01e-6
            assert abs(obj - res[1][k][0]) <= 1e-5 * (1 + abs(res[1][k][0]))


def test_batch_errors():
    A, b, c = D.load_csr("afiro")
    with pytest.raises(ValueError):
        M.BatchLP([])
    big_n = 40000   # 8(4n+3m) exceeds shared memory
    import scipy.sparse as sp
    Abig = sp.random(10, big_n, density=1e-3, format="csr", random_state=1)
    with pytest.raises(RuntimeError, match="shared memory"):
        M.BatchLP([(Abig, Abig.data, np.zeros(10), np.zeros(big_n))])


def test_preconditioned_batch_solve_matches_highs_with_fewer_iterations():
    """scale=True: Ruiz + Pock-Chambolle per distinct matrix on the device (MLLP_F_PRECONDITION), one solve launch; results,
    KKT error and termination refer to the ORIGINAL LPs"""
    names = ["sc50a", "sc105", "adlittle", "blend", "share2b"]
    insts, mats = load_batch(names)
    plain = M.solve_linear_program_batch(insts, tol=1e-6, max_iters=400000)
    res = M.solve_linear_program_batch(insts, tol=1e-6, max_iters=400000, scale=True)
    for nm, (A, b, c), (obj, x, y, inf) in zip(names, mats, res):
        assert inf["converged"] and inf["rel_kkt"] <= 1e-6 and inf["rel_kkt_original"] <= 1e-6
        assert abs(obj - HIGHS[nm]) <= 1e-5 * (1 + abs(HIGHS[nm]))
        kk = O.kkt(A, b, c, x, y)
        assert abs(kk[0] - obj) <= 1e-6 * (1 + abs(obj)) and abs(kk[8] - inf["rel_kkt_original"]) <= 1e-9
    assert sum(r[3]["iters"] for r in res) < sum(r[3]["iters"] for r in plain)
    # shared matrix: perturbed costs of one LP
    A, b, c = mats[1]
    rng = np.random.default_rng(3)
    cb = c * (1 + 0.05 * rng.uniform(-1, 1, (8, c.shape[0])))
    bb = np.tile(b, (8, 1))
    res = M.solve_linear_program_batch([insts[1]], tol=1e-6, max_iters=400000, shared=True, rhs_batch=bb, coefs_batch=cb, scale=True)
    ref = M.solve_linear_program_batch([insts[1]], tol=1e-6, max_iters=400000, shared=True, rhs_batch=bb, coefs_batch=cb)
    for (obj, x, y, inf), (obj0, _, _, inf0) in zip(res, ref):
        assert inf["converged"] and abs(obj - obj0) <= 1e-4 * (1 + abs(obj0)) and np.all(x >= -1e-12)


def test_config5_shared_matrix_data_parallel_path_runs_the_real_kernels():
    """BASELINE.json configs[4] through the public data-parallel call (one process = the whole batch is this rank's shard):
    perturbed instances of one Netlib matrix, solve mode with the library preconditioner, HOST batches in, per-instance
    results out; every instance is checked by the oracle's KKT evaluation on its ORIGINAL LP and against HiGHS-level
    objectives of the unperturbed LP's neighbourhood (perturbation 10 %)."""
    import sys
    sys.path.insert(0, __import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.abspath(__file__))))
    import bench
    from mllp_b200.distributed import solve_batch_data_parallel
    A, b, c = D.load_csr("sc105")
    bb, cb = bench.config5_batches(A, b, c, 0, 24)
    assert bench.slack_rows(A, c).size > 0 and np.array_equal(bb[:, np.setdiff1d(np.arange(A.shape[0]), bench.slack_rows(A, c))][0],
                                                               b[np.setdiff1d(np.arange(A.shape[0]), bench.slack_rows(A, c))])
    res = solve_batch_data_parallel([(A, A.data, b, c)], mode="solve", device=0, shared=True, rhs_batch=bb, coefs_batch=cb, count=24,
                                    tol=1e-6, max_iters=400000, scale=True, single_process=True)
    assert len(res) == 24
    for i, (obj, x, y, info) in enumerate(res):
        assert info["converged"] and info["rel_kkt"] <= 1e-6
        kk = O.kkt(A, bb[i], cb[i], x, y)
        assert kk[8] <= 1.001e-6 and abs(kk[8] - info["rel_kkt"]) <= 1e-9 and abs(kk[0] - obj) <= 1e-9 * (1 + abs(obj))
    # the same instances one by one through the single-LP solve: same optimum
    o1, _, _, i1 = M.solve_linear_program(A, A.data, bb[3], cb[3], tol=1e-6, max_iters=400000, precondition=True)
    assert i1["converged"] and abs(o1 - res[3][0]) <= 2e-5 * (1 + abs(o1))


def test_warp_per_instance_kernels_match_the_cta_kernels_and_the_oracle(monkeypatch):
    """Large batches of small LPs run with ONE WARP per instance (k_batch_run_warp / k_batch_solve_warp): forced here on a small
    heterogeneous batch (37 instances: not a multiple of the 4 warps of a CTA).  Per row the summation order is the CTA
    kernels', so parity-mode iterates are bitwise equal; solve mode reaches the same optimum (its sums over rows are added in
    another order, so iteration counts may differ)."""
    names = ["sc50a", "sc105", "blend", "afiro", "adlittle"]
    base, mats0 = load_batch(names)
    rng = np.random.default_rng(11)
    insts, mats = [], []
    for k in range(37):
        constrs, w, b, c = base[k % len(base)]
        A = mats0[k % len(base)][0]
        ck = c * (1 + 0.05 * rng.uniform(-1, 1, c.shape[0]))
        insts.append((constrs, w, b, ck))
        mats.append((A, b, ck))
    out = {}
    for warp in ("0", "1"):
        monkeypatch.setenv("MLLP_BATCH_WARP", warp)
        bt = M.BatchLP(insts)
        assert (bt.info()["threads"] == 128) == (warp == "1") and (bt.info()["instances_per_cta_solve"] == 4) == (warp == "1")
        out[warp] = (M.pdhg_linear_program_batch(insts, num_iters=300, handle=bt),
                     M.solve_linear_program_batch(insts, tol=1e-6, max_iters=400000, handle=bt))
        sig = bt.sigma_max().cpu().numpy()
        bt.close()
    for k, (A, b, c) in enumerate(mats):
        r0, r1 = out["0"][0][k], out["1"][0][k]
        assert np.array_equal(r0[1], r1[1]) and np.array_equal(r0[2], r1[2])
        if k < 10:
            eta = 0.9 / sig[k]
            xo, yo = O.pdhg_run(A, b, c, np.zeros(A.shape[1]), np.zeros(A.shape[0]), eta, eta, 300)
            assert rel(r1[1], xo) < 1e-9 and rel(r1[2], yo) < 1e-9
            kk = O.kkt(A, b, c, xo, yo)
            assert abs(r1[3]["rel_kkt"] - kk[8]) <= 1e-6 * (1 + kk[8])
        s0, s1 = out["0"][1][k], out["1"][1][k]
        assert s1[3]["converged"] and s1[3]["rel_kkt"] <= 1e-6
        assert abs(s1[0] - s0[0]) <= 1e-5 * (1 + abs(s0[0]))
        kk = O.kkt(A, b, c, s1[1], s1[2])
        assert abs(kk[0] - s1[0]) <= 1e-6 * (1 + abs(s1[0])) and abs(kk[8] - s1[3]["rel_kkt"]) <= 1e-9


def test_warp_kernels_are_the_default_for_large_batches_of_small_lps_only(monkeypatch):
    monkeypatch.delenv("MLLP_BATCH_WARP", raising=False)
    A, b, c = D.load_csr("sc105")
    m, n = A.shape
    big = M.BatchLP([(A, A.data, b, c)], shared=True, count=2048)
    small = M.BatchLP([(A, A.data, b, c)], shared=True, count=64)
    # solve mode on warps (4 instances per CTA, one per warp); parity mode stays on the CTA kernels (measured faster there)
    assert big.info()["instances_per_cta_solve"] == 4 and big.info()["threads"] != 128
    assert small.info()["instances_per_cta_solve"] != 4
    # shared matrix, default kernels: instances 0, 1, last against the oracle (parity) and against HiGHS-checked objectives (solve)
    B = 2048
    rng = np.random.default_rng(3)
    cb = c * (1 + 0.05 * rng.uniform(-1, 1, (B, n)))
    bb = np.tile(b, (B, 1))
    res = M.pdhg_linear_program_batch([(A, A.data, b, c)], num_iters=100, handle=big, shared=True, rhs_batch=bb, coefs_batch=cb)
    eta = 0.9 / O.power_iteration(A, 50)
    for k in (0, 1, 1023, B - 1):
        xo, yo = O.pdhg_run(A, bb[k], cb[k], np.zeros(n), np.zeros(m), eta, eta, 100)
        assert rel(res[k][1], xo) < 1e-9 and rel(res[k][2], yo) < 1e-9
    sol = M.solve_linear_program_batch([(A, A.data, b, c)], tol=1e-6, max_iters=200000, handle=big, shared=True, rhs_batch=bb, coefs_batch=cb)
    for k in (0, 1, 1023, B - 1):
        assert sol[k][3]["converged"]
        kk = O.kkt(A, bb[k], cb[k], sol[k][1], sol[k][2])
        assert kk[8] <= 1.0001e-6 and abs(kk[0] - sol[k][0]) <= 1e-6 * (1 + abs(sol[k][0]))
    big.close(); small.close()
    # 25fv47's vectors (80 KB per instance) do not fit 16 times into an SM: CTA kernels
    A2, b2, c2 = D.load_csr("25fv47")
    mid = M.BatchLP([(A2, A2.data, b2, c2)], shared=True, count=2048)
    assert mid.info()["instances_per_cta_solve"] != 4
    mid.close()
