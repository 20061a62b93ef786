"""Generates tests/golden/ref_arrays.json and tests/golden/ref_loader.json from the REFERENCE ITSELF (run in the build
container, where /root/reference exists; the GPU box only reads the committed outputs).

ref_arrays.json: per instance with an MPS file (97), digests / checksums of the reference-held arrays
  raw  : dataset/netlib_mps/<name>.mps_{constrs.npz,rhs.npy,coefs.npy}      (what the MPS reader must reproduce bit for bit)
  norm : dataset/netlib_mps_norm/<name>.mps_{constrs.npz,rhs.npy,coefs.npy} (what the `_norm` rule must reproduce)
ref_loader.json: digests of the 6-tuples returned by the reference's own get_netlib_dataset(normalize=True)
  (/root/reference/linear_program_data.py:58-80), imported from where it lies and run from a scratch directory that
  links netlib_mps/ and dataset/ as the function expects them (relative to the working directory).
"""
import hashlib
import importlib.util
import json
import os
import sys
import tempfile

import numpy as np
import scipy.sparse as sp

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
REFROOT = "/root/reference"


def digest(*arrays):
    h = hashlib.sha256()
    for a in arrays:
        a = np.ascontiguousarray(a)
        h.update(str(a.dtype).encode() + str(a.shape).encode())
        h.update(a.tobytes())
    return h.hexdigest()


def reference_dataset():
    """The reference loader's own return value (all 97 listed instances)."""
    import scipy.sparse  # noqa: F401  (the reference does `import scipy` and uses scipy.sparse)
    spec = importlib.util.spec_from_file_location("_mllp_reference_data", os.path.join(REFROOT, "linear_program_data.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    cwd = os.getcwd()
    with tempfile.TemporaryDirectory() as tmp:
        os.symlink(os.path.join(REFROOT, "netlib_mps"), os.path.join(tmp, "netlib_mps"))
        os.symlink(os.path.join(REFROOT, "dataset"), os.path.join(tmp, "dataset"))
        os.chdir(tmp)
        try:
            return mod.get_netlib_dataset(normalize=True)
        finally:
            os.chdir(cwd)


def tuple_digest(inst):
    file, constrs, weights, coefs, rhs, basis = inst
    lens = np.array([len(r) for r in constrs], dtype=np.int64)
    flat = np.concatenate([np.asarray(r) for r in constrs]) if len(constrs) else np.zeros(0, dtype=np.int32)
    return {"rows": len(constrs), "nnz": int(lens.sum()), "constrs": digest(lens, flat.astype(np.int32)), "weights": digest(weights),
            "coefs": digest(coefs), "rhs": digest(rhs), "basis": digest(basis)}


def main():
    from oracle.norm_rule import norm_rule
    from mllp_b200.mps import read_mps
    names = sorted(f[:-4] for f in os.listdir(os.path.join(REFROOT, "netlib_mps")) if f.endswith(".mps"))
    out = {"raw": {}, "norm": {}}
    for nm in names:
        R = sp.load_npz("%s/dataset/netlib_mps/%s.mps_constrs.npz" % (REFROOT, nm)).tocsr()
        R.sort_indices()
        rb, rc = np.load("%s/dataset/netlib_mps/%s.mps_rhs.npy" % (REFROOT, nm)), np.load("%s/dataset/netlib_mps/%s.mps_coefs.npy" % (REFROOT, nm))
        out["raw"][nm] = {"shape": list(R.shape), "nnz": int(R.nnz), "constrs": digest(R.indptr.astype(np.int32), R.indices.astype(np.int32), R.data),
                          "rhs": digest(rb), "coefs": digest(rc)}
        N = sp.load_npz("%s/dataset/netlib_mps_norm/%s.mps_constrs.npz" % (REFROOT, nm)).tocsr()
        N.sort_indices()
        nb, nc = np.load("%s/dataset/netlib_mps_norm/%s.mps_rhs.npy" % (REFROOT, nm)), np.load("%s/dataset/netlib_mps_norm/%s.mps_coefs.npy" % (REFROOT, nm))
        # which rows were divided by their norm is a property of the raw data (|b| / r <= 5): take it from the rule
        lp = read_mps(os.path.join(ROOT, "data", "netlib_mps_gz", nm + ".mps.gz"), range_form="dataset")
        _, _, _, info = norm_rule(lp["A"], lp["b"], lp["c"], lp["row_sense"])
        rows = np.repeat(np.arange(N.shape[0]), np.diff(N.indptr))
        out["norm"][nm] = {"shape": list(N.shape), "nnz": int(N.nnz), "structure": digest(N.indptr.astype(np.int32), N.indices.astype(np.int32)),
                           "divided_rows_data": digest(N.data[info["divided"][rows]]),
                           "sum_abs_data": float(np.abs(N.data).sum()), "sum_abs_rhs": float(np.abs(nb).sum()),
                           "sum_abs_coefs": float(np.abs(nc).sum()), "c_norm2": float(np.linalg.norm(rc))}
    json.dump(out, open(os.path.join(HERE, "ref_arrays.json"), "w"), indent=0, sort_keys=True)
    dataset, train_dict = reference_dataset()
    ld = {"order_independent": True, "keys_of_train_dict": sorted(train_dict.keys()), "instances": {inst[0]: tuple_digest(inst) for inst in dataset}}
    json.dump(ld, open(os.path.join(HERE, "ref_loader.json"), "w"), indent=0, sort_keys=True)
    print("wrote ref_arrays.json (%d instances) and ref_loader.json (%d tuples)" % (len(names), len(dataset)))


if __name__ == "__main__":
    main()
