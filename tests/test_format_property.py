"""Property test (CPU): for random sparse matrices of random shape/skew the tiled format reproduces
every row dot product of A and A' exactly once, for random grid sizes and step parameters, with
both tile dealings, and the row partition covers every row exactly once for random rank counts."""
import numpy as np
import scipy.sparse as sp
from hypothesis import given, settings, strategies as st

from mllp_b200 import _cabi


def _arrays(A):
    A = sp.csr_matrix(A)
    A.sort_indices()
    return (A, np.ascontiguousarray(A.indptr, dtype=np.int32), np.ascontiguousarray(A.indices, dtype=np.int32),
            np.ascontiguousarray(A.data, dtype=np.float64))


@st.composite
def sparse_matrices(draw):
    m = draw(st.integers(1, 120))
    n = draw(st.integers(1, 400))
    seed = draw(st.integers(0, 2 ** 31 - 1))
    rng = np.random.default_rng(seed)
    # skewed row lengths: mostly short rows, a few long ones, some empty
    lens = np.minimum(n, (rng.pareto(1.2, m) * draw(st.integers(1, 6))).astype(np.int64))
    lens[rng.random(m) < 0.15] = 0
    if draw(st.booleans()):
        lens[rng.integers(0, m)] = n          # one dense row -> split path when n is large
    rows, cols = [], []
    for i, L in enumerate(lens):
        if L:
            cols.append(rng.choice(n, size=int(L), replace=False))
            rows.append(np.full(int(L), i))
    if rows:
        r, c = np.concatenate(rows), np.concatenate(cols)
        A = sp.csr_matrix((rng.standard_normal(r.shape[0]), (r, c)), shape=(m, n))
    else:
        A = sp.csr_matrix((m, n))
    return A


@settings(max_examples=60, deadline=None)
@given(A=sparse_matrices(), G=st.integers(1, 40), pref=st.integers(1, 6), mx=st.integers(1, 9))
def test_format_selfcheck_random(A, G, pref, mx):
    A, ip, ii, vv = _arrays(A)
    out = np.zeros(8)
    rc = _cabi.lib().mllp_format_selfcheck(A.shape[0], A.shape[1], A.nnz, ip.ctypes.data, ii.ctypes.data, vv.ctypes.data, G,
                                           pref, mx, out.ctypes.data)
    assert rc == 0 and out[0] < 1e-12


@settings(max_examples=30, deadline=None)
@given(A=sparse_matrices(), G=st.integers(1, 12), R=st.integers(1, 8))
def test_rowpart_selfcheck_random(A, G, R):
    A, ip, ii, vv = _arrays(A)
    out = np.zeros(4)
    rc = _cabi.lib().mllp_rowpart_selfcheck(A.shape[0], A.shape[1], A.nnz, ip.ctypes.data, ii.ctypes.data, vv.ctypes.data, G,
                                            R, out.ctypes.data)
    assert rc == 0 and out[0] < 1e-12


@st.composite
def block_angular_matrices(draw):
    """independent blocks of random size (some with empty rows / unused columns), a few dense linking rows, columns that only
    the linking rows touch; rows shuffled"""
    seed = draw(st.integers(0, 2 ** 31 - 1))
    rng = np.random.default_rng(seed)
    nblocks = draw(st.integers(20, 70))
    nlink = draw(st.integers(0, 4))
    blocks = []
    for k in range(nblocks):
        mb, nb = int(rng.integers(1, 6)), int(rng.integers(1, 9))
        B = sp.random(mb, nb, density=float(rng.uniform(0.2, 0.9)), random_state=int(rng.integers(1 << 30)), format="csr")
        B.data[:] = rng.standard_normal(B.nnz)
        blocks.append(B)
    top = sp.hstack([sp.block_diag(blocks, format="csr"), sp.csr_matrix((sum(b.shape[0] for b in blocks), int(rng.integers(0, 5))))]).tocsr()
    parts = [top]
    if nlink:
        link = sp.random(nlink, top.shape[1], density=float(rng.uniform(0.5, 0.9)), random_state=int(rng.integers(1 << 30)), format="csr")
        link.data[:] = rng.standard_normal(link.nnz)
        parts.append(link)
    A = sp.vstack(parts).tocsr()
    A = A[rng.permutation(A.shape[0])].tocsr()
    G = draw(st.integers(2, 10))
    return A, G, nlink


@settings(max_examples=40, deadline=None)
@given(block_angular_matrices())
def test_block_images_reproduce_both_products(case):
    """blocks.cu host side: whatever structure is found, every row / column is placed once and the groups' lists
    reproduce A xbar and A'y (mllp_blocks_selfcheck)"""
    import ctypes
    A, G, nlink = case
    A, ip, ii, vv = _arrays(A)
    out = (ctypes.c_double * 8)()
    rc = _cabi.lib().mllp_blocks_selfcheck(A.shape[0], A.shape[1], int(A.nnz), ip.ctypes.data, ii.ctypes.data, vv.ctypes.data, G, out)
    assert rc == 0, _cabi.last_error()
    if out[0] == 1:
        assert out[4] < 1e-12 and out[5] < 1e-12 and out[1] >= 2 * G
        lens = np.diff(ip)
        thr = max(32.0, 8.0 * A.nnz / A.shape[0])
        assert out[2] == int((lens > int(thr)).sum()) and out[3] == int(lens[lens > int(thr)].sum())
