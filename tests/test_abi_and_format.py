"""CPU tests: the C-ABI library loads and exports what the header declares; the host-side
format builder is correct (replayed on the CPU by mllp_format_selfcheck); host logic."""
import ctypes
import os
import re

import numpy as np
import pytest
import scipy.sparse as sp

import mllp_b200 as M
import mllp_b200.linear_program_data as D
from mllp_b200 import _cabi
from mllp_b200.linear_program_methods import csr_from_constrs

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "mllp_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(mllp_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    L = ctypes.CDLL(_cabi.SO_PATH)
    syms = declared_symbols()
    assert len(syms) >= 15
    for s in syms:
        assert hasattr(L, s), s
    assert sorted(_cabi.SIGNATURES) == syms  # the binding mirrors the header one to one
    assert _cabi.lib().mllp_version() >= 100


def test_no_gpu_means_loud_failure():
    """There is no CPU fallback: creating a handle without a CUDA device raises."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    A, b, c = D.load_csr("afiro")
    with pytest.raises(RuntimeError, match="mllp_lp_create failed"):
        M.DeviceLP(A, A.data, *A.shape)
    with pytest.raises(RuntimeError):
        M.pdhg_linear_program(A, A.data, b, c, num_iters=10, tau=0.1, sigma=0.1)


def selfcheck(A, G=296, pref=4, mx=4):
    A = sp.csr_matrix(A)
    A.sort_indices()
    m, n = A.shape
    ip = np.ascontiguousarray(A.indptr, dtype=np.int32)
    ii = np.ascontiguousarray(A.indices, dtype=np.int32)
    vv = np.ascontiguousarray(A.data, dtype=np.float64)
    out = np.zeros(8)
    rc = _cabi.lib().mllp_format_selfcheck(m, n, A.nnz, ip.ctypes.data, ii.ctypes.data, vv.ctypes.data, G, pref, mx,
                                           out.ctypes.data)
    return rc, out


@pytest.mark.parametrize("name", ["afiro", "kb2", "25fv47", "pilot87", "dfl001", "ken-18", "osa-60", "pds-20"])
def test_format_builder_on_netlib(name):
    A, _, _ = D.load_csr(name)
    rc, out = selfcheck(A)
    assert rc == 0
    assert out[0] < 1e-12            # every row dot product reproduced
    assert out[3] < 9 and out[4] < 9  # padding bounded (tiny instances pad to whole warp-steps)
    if name == "osa-60":
        assert out[5] == 82 and out[3] < 1.1 and out[4] < 1.05


@pytest.mark.parametrize("G,pref,mx", [(1, 1, 1), (7, 2, 3), (148, 4, 4), (296, 4, 8), (592, 8, 8), (296, 3, 16)])
def test_format_builder_parameters(G, pref, mx):
    A, _, _ = D.load_csr("pilot87")
    rc, out = selfcheck(A, G, pref, mx)
    assert rc == 0 and out[0] < 1e-12


def test_format_builder_edge_shapes():
    rng = np.random.default_rng(0)
    # empty rows/cols, one dense very long row (split), duplicates of every length class
    rows = [sp.random(1, 3000, density=d, random_state=int(1000 * d) + 1, format="csr")
            for d in (0.0, 0.0003, 0.001, 0.003, 0.01, 0.03, 0.1, 0.3, 0.9, 1.0)]
    A = sp.vstack(rows * 7).tocsr()
    A.data[:] = rng.standard_normal(A.nnz)
    for G in (3, 148):
        rc, out = selfcheck(A, G)
        assert rc == 0 and out[0] < 1e-12
    rc, out = selfcheck(sp.csr_matrix((5, 9)))   # all-zero matrix
    assert rc == 0
    rc, out = selfcheck(sp.csr_matrix((0, 4)))   # no rows
    assert rc == 0
    rc, out = selfcheck(sp.csr_matrix(np.ones((1, 1))))
    assert rc == 0 and out[0] == 0.0


def test_loader_contract_matches_reference():
    """Same tuple as linear_program_data.py:78 of the reference."""
    dataset, train_dict = D.get_netlib_dataset(True, names=["afiro", "sc50a"])
    assert len(dataset) == 2 and train_dict["obj"] == [] and "afiro.mps" in train_dict
    file, constrs, constrs_weights, coefs, rhs, basis_opt = dataset[0]
    assert file == "afiro.mps" and len(constrs) == 27 and coefs.shape == (51,) and rhs.shape == (27,)
    assert constrs_weights.dtype == np.float64 and constrs[0].dtype == np.int32
    assert sum(len(r) for r in constrs) == constrs_weights.shape[0] == 102
    indptr, indices, values = csr_from_constrs(constrs, constrs_weights, 51)
    A, _, _ = D.load_csr("afiro")
    assert np.array_equal(indptr, A.indptr) and np.array_equal(indices, A.indices) and np.array_equal(values, A.data)
    assert "osa-60.mps" in D.list_instances()  # MPS-less instances are enumerated too


def test_host_argument_errors():
    A, b, c = D.load_csr("afiro")
    constrs = np.split(A.indices, A.indptr)[1:-1]
    with pytest.raises(ValueError):
        csr_from_constrs(constrs, A.data[:-1], 51)
    with pytest.raises(ValueError):
        csr_from_constrs(constrs, A.data, 10)  # column index out of range
    with pytest.raises(ValueError):
        M.pdhg_linear_program(constrs, A.data, b, c, num_iters=-1)
    with pytest.raises(ValueError):
        M.DeviceLP(constrs, A.data, 27, 51, device="cpu")


def _blocks_selfcheck(A, G):
    import ctypes
    A = A.tocsr()
    A.sort_indices()
    ip = np.ascontiguousarray(A.indptr, dtype=np.int32)
    ii = np.ascontiguousarray(A.indices, dtype=np.int32)
    vv = np.ascontiguousarray(A.data, dtype=np.float64)
    out = (ctypes.c_double * 8)()
    rc = _cabi.lib().mllp_blocks_selfcheck(A.shape[0], A.shape[1], int(A.nnz), ip.ctypes.data, ii.ctypes.data, vv.ctypes.data, G, out)
    assert rc == 0, _cabi.last_error()
    return list(out)


def test_block_angular_structure_detection_and_images_on_the_host():
    """blocks.cu, host side (no GPU): linking rows, connected components, groups; the groups' lists reproduce A xbar and A'y"""
    import scipy.sparse as sp
    import mllp_b200.linear_program_data as D
    A, b, c = D.load_csr("ken-18")                      # 475 blocks of <= 801 nodes + 151 linking rows
    found, ncomp, nlink, link_nnz, ea, et, smem, max_n = _blocks_selfcheck(A, 148)
    assert found == 1 and ncomp == 475 and nlink == 151 and link_nnz == 49075
    assert ea < 1e-13 and et < 1e-13 and smem < 227 * 1024 and 154699 / 148 <= max_n < 1600
    for name in ("pds-20", "osa-60", "25fv47"):         # one giant component / too few pieces: keep the grid kernel
        A, b, c = D.load_csr(name)
        assert _blocks_selfcheck(A, 148)[0] == 0
    # synthetic: 120 blocks, 5 dense linking rows, columns seen by linking rows only, an all-zero row, 16 groups
    rng = np.random.default_rng(3)
    blocks = []
    for k in range(120):
        Bk = sp.random(4, 6, density=0.5, random_state=k, format="csr")
        Bk.data[:] = rng.standard_normal(Bk.nnz)
        blocks.append(Bk)
    top = sp.hstack([sp.block_diag(blocks, format="csr"), sp.csr_matrix((480, 9))]).tocsr()
    link = sp.random(5, top.shape[1], density=0.15, random_state=77, format="csr")   # (linking rows may hold at most half of the nonzeros)
    link.data[:] = rng.standard_normal(link.nnz)
    A = sp.vstack([top[:100], link[:2], sp.csr_matrix((1, top.shape[1])), top[100:], link[2:]]).tocsr()
    found, ncomp, nlink, link_nnz, ea, et, smem, max_n = _blocks_selfcheck(A, 16)
    assert found == 1 and nlink == 5 and link_nnz == link.nnz and ncomp >= 120 and ea < 1e-13 and et < 1e-13
    assert _blocks_selfcheck(A, 148)[0] == 0            # fewer than 2 x 148 pieces


@pytest.mark.parametrize("name", ["afiro", "sc50a", "sc105", "blend", "adlittle", "share2b", "kb2", "25fv47"])
def test_row_per_lane_images_replay_on_the_cpu(name):
    """the row-per-lane images (HostEll) the warp-per-instance solve kernel walks: the lane walk replayed on the host against
    the plain CSR products, for A and A' in the batch builder's internal orders"""
    import ctypes
    import mllp_b200.linear_program_data as D
    from mllp_b200 import _cabi
    A, _, _ = D.load_csr(name)
    A = A.tocsr()
    A.sort_indices()
    ip = np.ascontiguousarray(A.indptr, dtype=np.int32)
    ii = np.ascontiguousarray(A.indices, dtype=np.int32)
    vv = np.ascontiguousarray(A.data, dtype=np.float64)
    out = (ctypes.c_double * 4)()
    rc = _cabi.lib().mllp_ell_selfcheck(A.shape[0], A.shape[1], A.nnz, ip.ctypes.data, ii.ctypes.data, vv.ctypes.data, out)
    assert rc == 0 and out[0] < 1e-14
    assert out[3] < 3.0          # padding factor: the internal order sorts the rows by length


def test_row_per_lane_images_edge_cases():
    import ctypes
    import scipy.sparse as sp
    from mllp_b200 import _cabi
    rng = np.random.default_rng(5)
    for m, n, dens in ((1, 1, 1.0), (33, 65, 0.1), (64, 32, 0.5), (40, 7, 0.0)):
        Ad = (rng.random((m, n)) < dens) * rng.standard_normal((m, n))
        if m > 2:
            Ad[1] = 0.0                      # an empty row
        A = sp.csr_matrix(Ad)
        A.sort_indices()
        ip = np.ascontiguousarray(A.indptr, dtype=np.int32)
        ii = np.ascontiguousarray(A.indices if A.nnz else np.zeros(1), dtype=np.int32)
        vv = np.ascontiguousarray(A.data if A.nnz else np.zeros(1), dtype=np.float64)
        out = (ctypes.c_double * 4)()
        rc = _cabi.lib().mllp_ell_selfcheck(m, n, A.nnz, ip.ctypes.data, ii.ctypes.data, vv.ctypes.data, out)
        assert rc == 0 and out[0] < 1e-14, (m, n, dens, rc)
