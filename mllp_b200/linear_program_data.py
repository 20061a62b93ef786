"""Netlib LP loader with the reference's data contract, plus device-resident formats.

Mirrors ``get_netlib_dataset`` of the reference (linear_program_data.py:58-80): for each
instance it loads ``<root>/netlib_mps_norm/<name>.mps_{constrs.npz,coefs.npy,rhs.npy}`` (and
``_basis.npy`` when present) and returns the same 6-tuple
``(file, constrs, constrs_weights, coefs, rhs, basis_opt)`` and ``train_dict``.

Differences, all additive:
* instances are enumerated from the dataset directory (or ``names``), not from
  ``os.listdir("netlib_mps")``, so the 12 MPS-less instances (ken-18, osa-60, pds-20, ...) load;
* ``device=`` builds the device-resident tiled formats of A and A' once per instance
  (mllp_b200.linear_program_methods.DeviceLP) and registers them so that
  ``pdhg_linear_program(constrs, constrs_weights, rhs, coefs, ...)`` finds them.
"""
import os

import numpy as np
import scipy.sparse

_HERE = os.path.dirname(os.path.abspath(__file__))


def data_root(root=None):
    """Directory that holds ``netlib_mps_norm/`` (reference layout: ``dataset/``)."""
    cands = [root, os.environ.get("MLLP_DATA_ROOT"), "dataset", os.path.join(os.path.dirname(_HERE), "data"),
             "/root/reference/dataset"]
    for c in cands:
        if c and os.path.isdir(os.path.join(c, "netlib_mps_norm")):
            return c
    raise FileNotFoundError("no dataset root with netlib_mps_norm/ found (tried %r)" % (cands,))


def list_instances(normalize=True, root=None):
    sub = "netlib_mps_norm" if normalize else "netlib_mps"
    d = os.path.join(data_root(root), sub)
    suffix = "_constrs.npz"
    return sorted(f[:-len(suffix)] for f in os.listdir(d) if f.endswith(suffix))


def load_instance(file, normalize=True, root=None):
    """One instance in the reference's representation.  ``file`` is e.g. 'afiro.mps' (the
    reference's key) or 'afiro'."""
    if not file.endswith(".mps"):
        file = file + ".mps"
    file_path = os.path.join(data_root(root), "netlib_mps_norm" if normalize else "netlib_mps") + os.sep
    coefs = np.load(file_path + file + "_coefs.npy")
    rhs = np.load(file_path + file + "_rhs.npy")
    constrs_sp_matrix = scipy.sparse.load_npz(file_path + file + "_constrs.npz").tocsr()
    constrs = np.split(constrs_sp_matrix.indices, constrs_sp_matrix.indptr)[1:-1]
    constrs_weights = constrs_sp_matrix.data
    basis_file = file_path + file + "_basis.npy"
    basis_opt = np.load(basis_file) if os.path.exists(basis_file) else None
    return (file, constrs, constrs_weights, coefs, rhs, basis_opt)


def get_netlib_dataset(normalize=True, names=None, root=None, device=None):
    """Reference signature ``get_netlib_dataset(normalize=True)`` plus optional ``names``
    (instance subset), ``root`` and ``device`` (build device formats in the loader)."""
    files = list_instances(normalize, root) if names is None else \
        [n if n.endswith(".mps") else n + ".mps" for n in names]
    dataset = []
    train_dict = {}
    train_dict["obj"] = []
    for file in files:
        inst = load_instance(file, normalize, root)
        dataset.append(inst)
        train_dict[file] = []
        if device is not None:
            from .linear_program_methods import device_lp
            _, constrs, constrs_weights, coefs, rhs, _ = inst
            device_lp(constrs, constrs_weights, rhs, coefs, device=device)
    return dataset, train_dict


def load_csr(file, normalize=True, root=None):
    """(scipy CSR A, b, c) -- convenience for benchmarks and tests."""
    if not file.endswith(".mps"):
        file = file + ".mps"
    file_path = os.path.join(data_root(root), "netlib_mps_norm" if normalize else "netlib_mps") + os.sep
    A = scipy.sparse.load_npz(file_path + file + "_constrs.npz").tocsr()
    A.sort_indices()
    return A, np.load(file_path + file + "_rhs.npy"), np.load(file_path + file + "_coefs.npy")
