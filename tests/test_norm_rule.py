"""Pins of the loader-side data contract against what the reference itself holds:

* the MPS reader reproduces the reference's raw arrays dataset/netlib_mps/* bit for bit (range_form='dataset'),
* the restated `_norm` rule (oracle/norm_rule.py) reproduces dataset/netlib_mps_norm/* from the MPS text,
* the device kernel (mllp_norm_scale) equals the restated rule.

The reference's arrays exist in the build container only (/root/reference); their digests / checksums are committed as
tests/golden/ref_arrays.json (made by tests/golden/make_ref_pins.py), so the first two also run where the reference is
absent; the element-wise comparison with the arrays themselves runs wherever they are present."""
import hashlib
import json
import os

import numpy as np
import pytest
import scipy.sparse as sp

from mllp_b200.mps import read_mps
from oracle.norm_rule import norm_rule

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GZ = os.path.join(ROOT, "data", "netlib_mps_gz")
NAMES = sorted(f[:-7] for f in os.listdir(GZ) if f.endswith(".mps.gz"))
PINS = json.load(open(os.path.join(ROOT, "tests", "golden", "ref_arrays.json")))
REF = "/root/reference/dataset"


def digest(*arrays):
    h = hashlib.sha256()
    for a in arrays:
        a = np.ascontiguousarray(a)
        h.update(str(a.dtype).encode() + str(a.shape).encode())
        h.update(a.tobytes())
    return h.hexdigest()


def raw_of(name):
    lp = read_mps(os.path.join(GZ, name + ".mps.gz"), range_form="dataset")
    A = lp["A"].tocsr()
    A.sort_indices()
    return lp, A


def test_all_97_present():
    assert len(NAMES) == 97 and set(NAMES) == set(PINS["raw"])


@pytest.mark.parametrize("name", NAMES)
def test_mps_reader_equals_reference_raw_arrays_bitwise(name):
    lp, A = raw_of(name)
    pin = PINS["raw"][name]
    assert list(A.shape) == pin["shape"] and A.nnz == pin["nnz"]
    assert digest(A.indptr.astype(np.int32), A.indices.astype(np.int32), A.data) == pin["constrs"]
    assert digest(lp["b"]) == pin["rhs"] and digest(lp["c"]) == pin["coefs"]
    if os.path.isdir(REF):   # the arrays themselves, element by element
        R = sp.load_npz("%s/netlib_mps/%s.mps_constrs.npz" % (REF, name)).tocsr()
        R.sort_indices()
        assert np.array_equal(R.indptr, A.indptr) and np.array_equal(R.indices, A.indices) and np.array_equal(R.data, A.data)
        assert np.array_equal(np.load("%s/netlib_mps/%s.mps_rhs.npy" % (REF, name)), lp["b"])
        assert np.array_equal(np.load("%s/netlib_mps/%s.mps_coefs.npy" % (REF, name)), lp["c"])


@pytest.mark.parametrize("name", NAMES)
def test_norm_rule_reproduces_reference_norm_arrays(name):
    lp, A = raw_of(name)
    An, rhs, coefs, info = norm_rule(A, lp["b"], lp["c"], lp["row_sense"])
    pin = PINS["norm"][name]
    assert list(An.shape) == pin["shape"] and An.nnz == pin["nnz"]
    assert digest(An.indptr.astype(np.int32), An.indices.astype(np.int32)) == pin["structure"]      # exact sparsity
    rows = np.repeat(np.arange(An.shape[0]), np.diff(An.indptr))
    div = info["divided"][rows]
    assert digest(An.data[div]) == pin["divided_rows_data"]                 # rows scaled by 1/r: bit-identical
    for got, key in ((An.data, "sum_abs_data"), (rhs, "sum_abs_rhs"), (coefs, "sum_abs_coefs")):
        assert abs(np.abs(got).sum() - pin[key]) <= 1e-13 * max(1.0, pin[key])
    assert abs(info["c_norm2"] - pin["c_norm2"]) <= 1e-15 * pin["c_norm2"]
    if os.path.isdir(REF):
        R = sp.load_npz("%s/netlib_mps_norm/%s.mps_constrs.npz" % (REF, name)).tocsr()
        R.sort_indices()
        assert np.array_equal(R.indptr, An.indptr) and np.array_equal(R.indices, An.indices)
        assert np.array_equal(R.data[div], An.data[div])
        rb, rc = np.load("%s/netlib_mps_norm/%s.mps_rhs.npy" % (REF, name)), np.load("%s/netlib_mps_norm/%s.mps_coefs.npy" % (REF, name))
        tol = 2e-15   # rows scaled by 5/b, right-hand sides, c/||c||: a few units in the last place (oracle header)
        assert np.max(np.abs(R.data - An.data) / np.abs(R.data)) <= tol
        assert np.max(np.abs(rb - rhs) / np.maximum(np.abs(rb), 1e-300)) <= tol
        assert np.max(np.abs(rc - coefs) / np.maximum(np.abs(rc), 1e-300)) <= tol


def test_objective_in_netlib_units():
    """objective of the `_norm` LP x ||c_raw|| (+ offset) is the Netlib optimum."""
    from mllp_b200.scaling import netlib_objective
    lp, A = raw_of("afiro")
    _, _, _, info = norm_rule(A, lp["b"], lp["c"], lp["row_sense"])
    assert abs(netlib_objective(-46.2784021, {"c_norm2": info["c_norm2"], "offset": lp["offset"]}) - (-464.7531429)) < 1e-5
    lp, A = raw_of("e226")
    _, _, _, info = norm_rule(A, lp["b"], lp["c"], lp["row_sense"])
    assert abs(lp["offset"] - 7.113) < 1e-12
    h = json.load(open(os.path.join(ROOT, "tests", "golden", "mps_models_all.json")))["e226"]["objective"]
    assert abs(h - (-11.63892907)) < 1e-6


@pytest.mark.gpu
@pytest.mark.parametrize("name", NAMES)
def test_device_norm_kernel_equals_the_rule(name):
    from mllp_b200.scaling import netlib_norm
    lp, A = raw_of(name)
    An, rhs, coefs, info = norm_rule(A, lp["b"], lp["c"], lp["row_sense"])
    Gn, grhs, gcoefs, ginfo = netlib_norm(A, lp["b"], lp["c"], lp["row_sense"], device=0)
    assert np.array_equal(Gn.indptr, An.indptr) and np.array_equal(Gn.indices, An.indices)
    assert np.array_equal(Gn.data, An.data) and np.array_equal(grhs, rhs)           # bit for bit (same operation order)
    assert np.max(np.abs(gcoefs - coefs)) <= 1e-15 * max(1e-300, np.max(np.abs(coefs)))   # ||c||: two-stage device sum vs BLAS
    assert abs(ginfo["c_norm2"] - info["c_norm2"]) <= 1e-15 * info["c_norm2"]
    pin = PINS["norm"][name]
    assert digest(Gn.indptr.astype(np.int32), Gn.indices.astype(np.int32)) == pin["structure"]


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["afiro", "sc50a", "sc105", "adlittle", "blend", "share2b", "kb2", "25fv47", "pilot87", "d2q06c", "dfl001"])
def test_device_norm_kernel_vs_carried_norm_arrays(name):
    """The `_norm` arrays carried under data/netlib_mps_norm (copies of the reference's) against the device kernel run on
    the MPS text of the same instance."""
    import mllp_b200 as M
    from mllp_b200.scaling import netlib_norm_from_mps
    A, b, c = M.load_csr(name)
    (fname, constrs, weights, coefs, rhs, _), info = netlib_norm_from_mps(os.path.join(GZ, name + ".mps.gz"), device=0)
    assert fname == name + ".mps"
    assert np.array_equal(np.concatenate(constrs) if len(constrs) else np.zeros(0), A.indices)
    assert np.max(np.abs(weights - A.data) / np.abs(A.data)) <= 2e-15
    assert np.max(np.abs(rhs - b) / np.maximum(np.abs(b), 1e-300)) <= 2e-15
    assert np.max(np.abs(coefs - c) / np.maximum(np.abs(c), 1e-300)) <= 2e-15
