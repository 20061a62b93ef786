"""Scaling of LPs for the B200 path, all computed ON THE DEVICE by the library:

* ``netlib_norm`` / ``netlib_norm_from_mps``: the reference's own `_norm` rule (one-sided row scaling + c / ||c||, SURVEY
  App. A.3; ``mllp_norm_scale``) and objectives in the MPS file's units (``netlib_objective``);
* ``solve_scaled`` / ``solve_mps``: solve mode on a handle created with ``MLLP_F_PRECONDITION`` -- Ruiz equilibration + one
  Pock-Chambolle (alpha = 1) pass, the PDLP recipe, by device kernels inside ``mllp_lp_create``:

      A~ = Dr A Dc,  b~ = Dr b,  c~ = Dc c,  l~ = l / Dc,  u~ = u / Dc      (x = Dc x~,  y = Dr y~)

  The caller only ever sees the ORIGINAL LP: vectors are scaled at the library boundary and the KKT error that ends the
  solve is evaluated on the original LP inside the kernel.
"""
import ctypes

import numpy as np
import scipy.sparse as sp

from . import _cabi

SENSE_CODE = {"E": 0, "L": 1, "G": -1}


def netlib_norm(A, b, c, sense, device=0, return_device=False):
    """The reference's `_norm` form of a raw LP, computed on the device (``mllp_norm_scale``): standard form with one
    slack column per inequality row (``sense`` = 'E' / 'L' / 'G' per row, RANGES rows already equalities), rows scaled
    by 1 / ||row||_2 or 5 / b_i, c by 1 / ||c||_2 -- the arrays ``dataset/netlib_mps_norm/<name>_{constrs,rhs,coefs}``
    hold for the same file (reference linear_program_data.py:66-77; rule: oracle/norm_rule.py).

    Returns ``(A_norm, rhs_norm, coefs_norm, info)`` with ``A_norm`` a scipy CSR matrix; ``info`` carries
    ``row_scale`` (y_raw = row_scale * y_norm), ``c_norm2`` (objective in the file's units = objective of the `_norm`
    LP * c_norm2 + offset) and ``num_slack``.  ``return_device=True`` adds the device tensors under ``info['device']``."""
    import torch
    from .linear_program_methods import _device_index, _torch_stream
    A = sp.csr_matrix(A, dtype=np.float64)
    A.sort_indices()
    m, n = A.shape
    codes = np.asarray([SENSE_CODE[s] for s in sense], dtype=np.int8) if len(sense) and isinstance(sense[0], str) \
        else np.ascontiguousarray(sense, dtype=np.int8)
    if codes.shape[0] != m or np.shape(b)[0] != m or np.shape(c)[0] != n:
        raise ValueError("sense / rhs must have one entry per row and coefs one per column")
    nslack = int(np.count_nonzero(codes))
    dev = torch.device("cuda", _device_index(device))
    t = lambda a, dt: torch.as_tensor(np.ascontiguousarray(a, dtype=dt), device=dev)
    d_ip, d_ii, d_vv = t(A.indptr, np.int32), t(A.indices, np.int32), t(A.data, np.float64)
    d_se, d_b, d_c = t(codes, np.int8), t(b, np.float64), t(c, np.float64)
    o_ip = torch.empty(m + 1, dtype=torch.int32, device=dev)
    o_ii = torch.empty(A.nnz + nslack, dtype=torch.int32, device=dev)
    o_vv = torch.empty(A.nnz + nslack, dtype=torch.float64, device=dev)
    o_b = torch.empty(m, dtype=torch.float64, device=dev)
    o_c = torch.empty(n + nslack, dtype=torch.float64, device=dev)
    o_d = torch.empty(m, dtype=torch.float64, device=dev)
    o_cn = torch.zeros(1, dtype=torch.float64, device=dev)
    L = _cabi.lib()
    work = torch.empty(int(L.mllp_norm_scale_work_bytes(m)) // 8 + 1, dtype=torch.float64, device=dev)
    p = lambda x: ctypes.c_void_p(x.data_ptr())
    _cabi.check(L.mllp_norm_scale(m, n, A.nnz, nslack, p(d_ip), p(d_ii), p(d_vv), p(d_se), p(d_b), p(d_c), p(o_ip), p(o_ii),
                                  p(o_vv), p(o_b), p(o_c), p(o_d), p(o_cn), p(work), _torch_stream(dev)), "mllp_norm_scale")
    An = sp.csr_matrix((o_vv.cpu().numpy(), o_ii.cpu().numpy(), o_ip.cpu().numpy()), shape=(m, n + nslack))
    info = {"row_scale": o_d.cpu().numpy(), "c_norm2": float(o_cn.cpu()[0]), "num_slack": nslack}
    if return_device:
        info["device"] = {"indptr": o_ip, "indices": o_ii, "values": o_vv, "rhs": o_b, "coefs": o_c, "row_scale": o_d}
    return An, o_b.cpu().numpy(), o_c.cpu().numpy(), info


def netlib_norm_from_mps(path, device=0):
    """MPS text -> the reference's loader tuple ``(file, constrs, constrs_weights, coefs, rhs, None)`` in `_norm` form
    (what ``get_netlib_dataset(normalize=True)`` hands out for the same file, without the pre-converted arrays) plus
    ``info`` (``c_norm2``, ``offset``, ``maximize``, ``row_scale``)."""
    import os
    from .mps import read_mps
    lp = read_mps(path, range_form="dataset")
    An, rhs, coefs, info = netlib_norm(lp["A"], lp["b"], lp["c"], lp["row_sense"], device=device)
    info.update(offset=lp["offset"], maximize=lp["maximize"])
    name = os.path.basename(str(path))
    name = name[:-3] if name.endswith(".gz") else name
    constrs = np.split(An.indices, An.indptr)[1:-1]
    return (name, constrs, An.data, coefs, rhs, None), info


def netlib_objective(objective_norm, info):
    """Objective of the `_norm` LP in the MPS file's units: x ||c_raw||_2 + offset, sign restored."""
    v = objective_norm * info["c_norm2"] + info.get("offset", 0.0)
    return -v if info.get("maximize") else v


# Solve-mode settings tried in turn by ``portfolio=True`` (each from the zero start, on the same preconditioned handle):
# (check_every, initial primal weight, iteration cap).  Measured on the 97 Netlib MPS files (profiles/r02_netlib_all.md,
# r02_hard_instances.md): the first setting brings 90 to 1e-6 within 2e6 iterations; the restart test every 512
# iterations adds bnl1, pilot4, pilot (93); PDLP's initial weight ||c|| / ||b|| adds greenbea and pilot.we (95).
# perold and pilot.ja reach 7e-5 / 2e-4 at best.
PORTFOLIO = ((64, 1.0, 2000000), (512, 1.0, 4000000), (64, None, 4000000))


def solve_scaled(A, b, c, *, lb=None, ub=None, ylo=None, yhi=None, tol=1e-6, max_iters=400000, check_every=64, device=0,
                 scale=True, primal_weight=1.0, portfolio=False):
    """Solve mode on a handle preconditioned by the library (``MLLP_F_PRECONDITION``: Ruiz + Pock-Chambolle computed on
    the device at create time).  Returns (objective, x, y, info) of the ORIGINAL LP; the in-kernel termination test is
    the KKT error of the original LP, ``info['rel_kkt_original']`` repeats it (kept for round-1 callers).
    ``portfolio=True``: the settings of ``PORTFOLIO`` are tried in turn until one converges (``info['attempt']``,
    ``info['iters_all_attempts']``); ``max_iters`` / ``check_every`` / ``primal_weight`` are then ignored."""
    from .linear_program_methods import DeviceLP, solve_linear_program
    A = sp.csr_matrix(A)
    m, n = A.shape
    h = DeviceLP(A, A.data, m, n, lb=lb, ub=ub, ylo=ylo, yhi=yhi, device=device, precondition=bool(scale))
    settings = PORTFOLIO if portfolio else ((check_every, primal_weight, max_iters),)
    total = 0
    try:
        for k, (ce, w0, cap) in enumerate(settings):
            obj, x, y, info = solve_linear_program(A, A.data, b, c, tol=tol, max_iters=int(cap), check_every=int(ce),
                                                   primal_weight=w0, handle=h)
            total += info["iters"]
            if info["converged"]:
                break
    finally:
        h.close()
    info = dict(info)
    info.pop("handle", None)
    info["rel_kkt_original"] = info["rel_kkt"]
    info["attempt"] = k
    info["iters_all_attempts"] = total
    info["setting"] = {"check_every": int(ce), "primal_weight": w0, "max_iters": int(cap)}
    return obj, x, y, info


def solve_mps(path, *, tol=1e-6, max_iters=400000, check_every=64, device=0, scale=True, primal_weight=1.0, portfolio=False):
    """Read an MPS file (with row senses, bounds, ranges), precondition, solve on the GPU in
    solve mode and return (objective incl. offset, x, y, info) in the ORIGINAL variables.
    info['rel_kkt_original'] is the KKT error re-evaluated on the unscaled LP on the device."""
    from .mps import read_mps
    lp = read_mps(path)
    pobj, x, y, info = solve_scaled(lp["A"], lp["b"], lp["c"], lb=lp["lb"], ub=lp["ub"], ylo=lp["ylo"], yhi=lp["yhi"], tol=tol,
                                    max_iters=max_iters, check_every=check_every, device=device, scale=scale,
                                    primal_weight=primal_weight, portfolio=portfolio)
    info["offset"] = lp["offset"]
    objective = pobj + lp["offset"]
    if lp["maximize"]:
        objective = -objective
    return objective, x, y, info
