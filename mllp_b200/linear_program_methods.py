"""Host side of the B200 primal-dual LP path, in the calling convention of the reference's
``linear_program_methods.py``.

The reference has no LP iteration of its own; these functions are the new
entry points the north star asks for, shaped after what the reference does have:

* argument order ``(constrs, constr_weights, rhs, coefs, ...)`` as in
  ``build_graph_from_weights_sets(constrs, constr_weights, rhs, coefs, device)``
  (reference linear_program_methods.py:89), fed by the loader tuple
  ``(file, constrs, constrs_weights, coefs, rhs, basis_opt)`` (linear_program_data.py:78);
* return convention ``(objective, solution, ...)`` and "tensors in => tensors out on the same
  device" as in ``gurobi_max_covering`` (linear_program_methods.py:542-553, :603-607);
* non-convergence is reported, not raised (linear_program_methods.py:537-539); bad arguments
  raise ``ValueError`` (linear_program_experiment.py:39).

All arithmetic happens in hand-written sm_100a kernels behind the C ABI
(include/mllp_b200.h); there is no CPU or PyTorch fallback -- a missing library or GPU raises.
"""
import ctypes
import weakref

import numpy as np

from . import _cabi

__all__ = [
    "DeviceLP", "device_lp", "pdhg_linear_program", "solve_linear_program", "estimate_step_size",
    "csr_from_constrs", "SCALAR_NAMES", "BatchLP", "pdhg_linear_program_batch", "solve_linear_program_batch",
]

SCALAR_NAMES = ("pobj", "dobj", "primal_res", "dual_res", "norm_b", "norm_c", "norm_x", "norm_y",
                "rel_kkt", "gap", "iters", "restarts", "converged", "primal_weight", "fixed_point_err",
                "reserved")


def _ptr(a):
    return None if a is None else ctypes.c_void_p(a.ctypes.data)


def _np_f64(a, size, name):
    a = np.ascontiguousarray(a, dtype=np.float64).reshape(-1)
    if a.shape[0] != size:
        raise ValueError("%s has %d entries, expected %d" % (name, a.shape[0], size))
    return a


def csr_from_constrs(constrs, constr_weights, num_cols):
    """CSR arrays from the loader's representation (linear_program_data.py:75-77):
    ``constrs`` = per-row column-index arrays (np.split of scipy's indices by indptr) or a scipy
    sparse matrix, ``constr_weights`` = the flat CSR data."""
    if hasattr(constrs, "tocsr"):
        A = constrs.tocsr()
        A.sort_indices()
        if A.shape[1] != num_cols:
            raise ValueError("constraint matrix has %d columns, coefs has %d" % (A.shape[1], num_cols))
        return (np.ascontiguousarray(A.indptr, dtype=np.int32), np.ascontiguousarray(A.indices, dtype=np.int32),
                np.ascontiguousarray(A.data, dtype=np.float64))
    m = len(constrs)
    lens = np.fromiter((len(r) for r in constrs), dtype=np.int64, count=m)
    indptr = np.zeros(m + 1, dtype=np.int64)
    np.cumsum(lens, out=indptr[1:])
    if indptr[-1] >= 2 ** 31:
        raise ValueError("more than 2^31-1 nonzeros")
    indices = (np.concatenate([np.asarray(r, dtype=np.int32) for r in constrs])
               if m and indptr[-1] else np.zeros(0, dtype=np.int32))
    values = np.ascontiguousarray(constr_weights, dtype=np.float64).reshape(-1)
    if values.shape[0] != indptr[-1]:
        raise ValueError("constr_weights has %d entries, constrs lists %d" % (values.shape[0], indptr[-1]))
    if indices.size and (indices.min() < 0 or indices.max() >= num_cols):
        raise ValueError("column index out of range")
    return indptr.astype(np.int32), np.ascontiguousarray(indices, dtype=np.int32), values


def _device_index(device):
    if device is None:
        return 0
    if isinstance(device, int):
        return device
    s = str(device)
    if s == "cuda":
        return 0
    if s.startswith("cuda:"):
        return int(s.split(":")[1])
    raise ValueError("mllp_b200 runs on CUDA devices only (got device=%r); there is no CPU path" % (device,))


class DeviceLP:
    """Device-resident tiled formats of A and A' for one LP instance (built once, in the
    loader).  Wraps an ``mllp_lp_t`` handle."""

    def __init__(self, constrs, constr_weights, num_rows, num_cols, lb=None, ub=None, ylo=None, yhi=None,
                 device=0, flags=_cabi.F_DEFAULT, precondition=False):
        """``precondition=True`` (flag MLLP_F_PRECONDITION): Ruiz + Pock-Chambolle scaling computed on the device at create
        time; the handle iterates on the scaled LP, every vector the caller passes or receives and the KKT scalars
        (incl. the termination test of solve mode) stay those of the ORIGINAL LP."""
        import time
        t_create = time.perf_counter()
        if precondition:
            flags = int(flags) | _cabi.F_PRECONDITION
        L = _cabi.lib()
        self.m, self.n = int(num_rows), int(num_cols)
        indptr, indices, values = csr_from_constrs(constrs, constr_weights, self.n)
        if indptr.shape[0] != self.m + 1:
            raise ValueError("constrs has %d rows, rhs has %d" % (indptr.shape[0] - 1, self.m))
        self.nnz = int(indptr[-1])
        self.device = _device_index(device)
        self.flags = int(flags)
        lb = None if lb is None else _np_f64(lb, self.n, "lb")
        ub = None if ub is None else _np_f64(ub, self.n, "ub")
        ylo = None if ylo is None else _np_f64(ylo, self.m, "ylo")
        yhi = None if yhi is None else _np_f64(yhi, self.m, "yhi")
        if (lb is None) != (ub is None) or (ylo is None) != (yhi is None):
            raise ValueError("lb/ub and ylo/yhi must be given in pairs")
        h = ctypes.c_void_p()
        rc = L.mllp_lp_create(self.m, self.n, self.nnz, _ptr(indptr), _ptr(indices), _ptr(values), _ptr(lb),
                              _ptr(ub), _ptr(ylo), _ptr(yhi), self.device, self.flags, ctypes.byref(h))
        _cabi.check(rc, "mllp_lp_create")
        self._h = h
        self._finalizer = weakref.finalize(self, L.mllp_lp_destroy, h)
        self._sigma_max = None
        self._sigma_robust = None
        # sqrt(||A||_1 ||A||_inf) >= ||A||_2: the guaranteed side of the step-size estimate
        absv = np.abs(values)
        row_sum = 0.0
        if self.nnz:
            row_sum = np.bincount(np.repeat(np.arange(self.m), np.diff(indptr)), weights=absv, minlength=self.m).max()
            col_sum = np.bincount(indices, weights=absv, minlength=self.n).max()
        else:
            col_sum = 0.0
        self.norm_upper = float(np.sqrt(row_sum * col_sum))
        self.preconditioned = bool(self.flags & _cabi.F_PRECONDITION)
        self.create_s = time.perf_counter() - t_create   # format build + geometry search + tuning rounds (synchronous)

    def scaling(self):
        """(dr, dc) of a preconditioned handle as numpy arrays in the caller's order (ones otherwise)."""
        import torch
        dev = torch.device("cuda", self.device)
        dr = torch.empty(self.m, dtype=torch.float64, device=dev)
        dc = torch.empty(self.n, dtype=torch.float64, device=dev)
        _cabi.check(_cabi.lib().mllp_lp_scaling(self.handle, dr.data_ptr(), dc.data_ptr(), _torch_stream(dev)), "mllp_lp_scaling")
        return dr.cpu().numpy(), dc.cpu().numpy()

    def close(self):
        self._finalizer()
        self._h = None

    @property
    def handle(self):
        if self._h is None:
            raise RuntimeError("DeviceLP is closed")
        return self._h

    def info(self):
        out = (ctypes.c_int64 * 16)()
        _cabi.check(_cabi.lib().mllp_lp_info(self.handle, out), "mllp_lp_info")
        keys = ("m", "n", "nnz", "tiles_A", "tiles_AT", "padded_A", "padded_AT", "split_rows_A", "split_rows_AT",
                "grid_ctas", "threads", "dyn_smem_bytes", "bytes_per_iter", "res_steps_A", "res_steps_AT",
                "ctas_per_sm")
        return dict(zip(keys, (int(v) for v in out)))

    def tune_info(self):
        """ns / iteration measured by mllp_lp_create before and after its tuning rounds."""
        out = (ctypes.c_double * 4)()
        _cabi.check(_cabi.lib().mllp_lp_tune_info(self.handle, out), "mllp_lp_tune_info")
        return {"ns_per_iter_first": out[0], "ns_per_iter_kept": out[1], "rounds": int(out[2])}

    def geometry(self):
        """Launch geometry of the persistent kernels (mllp_lp_geometry): cooperative grid, one cluster or one CTA."""
        out = (ctypes.c_double * 12)()
        _cabi.check(_cabi.lib().mllp_lp_geometry(self.handle, out), "mllp_lp_geometry")
        names = ("grid", "cluster16", "cluster8", "cluster4", "cta", "bcast16", "bcast8", "bcast4", "bcast1")
        return {"mode": ("grid", "cluster", "cta", "bcast")[int(out[0])], "ctas": int(out[1]),
                "ns_per_iter": {k: out[2 + i] for i, k in enumerate(names) if out[2 + i] > 0}}

    def blocks_info(self):
        """Block-angular structure found by mllp_lp_create and whether the parity kernel runs on it (mllp_lp_blocks_info)."""
        out = (ctypes.c_double * 8)()
        _cabi.check(_cabi.lib().mllp_lp_blocks_info(self.handle, out), "mllp_lp_blocks_info")
        return {"used": bool(out[0]), "found": bool(out[7]), "blocks": int(out[1]), "linking_rows": int(out[2]),
                "linking_nnz": int(out[3]), "smem_bytes": int(out[4]), "ns_per_iter_grid": out[5], "ns_per_iter_blocks": out[6]}

    def sigma_max(self, iters=50, stream=None):
        """||A||_2 estimate by power iteration on the device (cached)."""
        if self._sigma_max is None:
            s = ctypes.c_double(0.0)
            _cabi.check(_cabi.lib().mllp_estimate_norm(self.handle, int(iters), ctypes.byref(s), stream),
                        "mllp_estimate_norm")
            self._sigma_max = s.value
        return self._sigma_max

    def sigma_max_robust(self, rel_change=1e-4, max_iters=3200, stream=None):
        """||A||_2 for solve mode, where an underestimate makes the iteration diverge (50 steps
        are 2 % low on sc50b / ken-11): power iteration with doubling step counts until the estimate
        moves by less than ``rel_change``, inflated by 2 % and capped by sqrt(||A||_1 ||A||_inf)."""
        if self._sigma_robust is None:
            prev, iters = 0.0, 100
            while True:
                s = ctypes.c_double(0.0)
                _cabi.check(_cabi.lib().mllp_estimate_norm(self.handle, int(iters), ctypes.byref(s), stream),
                            "mllp_estimate_norm")
                cur = s.value
                if abs(cur - prev) <= rel_change * cur or iters >= max_iters:
                    break
                prev, iters = cur, iters * 2
            est = 1.02 * cur
            # the cap is a bound on the ORIGINAL matrix's norm: it does not apply to Dr A Dc
            self._sigma_robust = min(est, self.norm_upper) if (self.norm_upper > 0 and not self.preconditioned) else est
        return self._sigma_robust

    def _host_staging(self):
        """page-locked numpy views (x[n], y[m], scalars) for the host-buffer entry, allocated once per handle"""
        st = getattr(self, "_staging", None)
        if st is None:
            import torch
            pin = torch.cuda.is_available()
            mk = lambda k: (torch.empty(k, dtype=torch.float64).pin_memory() if pin else torch.empty(k, dtype=torch.float64))
            keep = (mk(max(self.n, 1)), mk(max(self.m, 1)), mk(_cabi.NUM_SCALARS))
            st = self._staging = (keep, (keep[0].numpy()[:self.n], keep[1].numpy()[:self.m], keep[2].numpy()))
        return st[1]

    # -- torch-tensor level helpers (device pointers, caller's stream) --------------------------
    def spmv(self, v, trans=False):
        import torch
        self._check_tensor(v, self.m if trans else self.n, "v")
        out = torch.empty(self.n if trans else self.m, dtype=torch.float64, device=v.device)
        _cabi.check(_cabi.lib().mllp_spmv(self.handle, int(bool(trans)), v.data_ptr(), out.data_ptr(),
                                          _torch_stream(v.device)), "mllp_spmv")
        return out

    def _check_tensor(self, t, size, name):
        import torch
        if not (isinstance(t, torch.Tensor) and t.is_cuda and t.dtype == torch.float64 and t.is_contiguous()
                and t.numel() == size and t.device.index == self.device):
            raise ValueError("%s must be a contiguous float64 CUDA tensor of %d entries on cuda:%d"
                             % (name, size, self.device))


def _torch_stream(device):
    import torch
    return ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)


# handles built by the loader (or lazily here), keyed by the identity of the weights array
_HANDLES = {}


def device_lp(constrs, constr_weights, rhs, coefs, lb=None, ub=None, ylo=None, yhi=None, device=0,
              flags=_cabi.F_DEFAULT, cache=True, precondition=False):
    """Return the DeviceLP of this instance, building (and caching) it on first use.  The cache
    is keyed on the identity of ``constr_weights`` -- the loader's tuple keeps that array alive.
    Bounds and row senses are baked into the handle when it is built, so an LP that passes any of them is
    never served from (or put into) the cache: two calls with the same matrix and different boxes get two handles."""
    if precondition:
        flags = int(flags) | _cabi.F_PRECONDITION
    key = (id(constr_weights), _device_index(device), int(flags))
    if lb is not None or ub is not None or ylo is not None or yhi is not None:
        cache = False
    if cache:
        hit = _HANDLES.get(key)
        if hit is not None and hit[0]() is constr_weights:
            return hit[1]
    lp = DeviceLP(constrs, constr_weights, len(rhs), len(coefs), lb, ub, ylo, yhi, device, flags)
    if cache:
        try:
            ref = weakref.ref(constr_weights, lambda _r, k=key: _HANDLES.pop(k, None))
            _HANDLES[key] = (ref, lp)
        except TypeError:
            pass  # not weak-referenceable (e.g. a list): no caching
    return lp


def estimate_step_size(lp, safety=0.9, iters=50):
    """eta = safety / sigma_max(A) (0.9/sigma_max by 50 power-iteration steps)."""
    s = lp.sigma_max(iters)
    return safety / s if s > 0 else 1.0


def _info_dict(scal):
    d = {k: float(v) for k, v in zip(SCALAR_NAMES, scal)}
    for k in ("iters", "restarts"):
        d[k] = int(d[k])
    d["converged"] = bool(d["converged"])
    del d["reserved"]
    return d


def _is_tensor(a):
    return type(a).__module__.startswith("torch") and hasattr(a, "data_ptr")


def pdhg_linear_program(constrs, constr_weights, rhs, coefs, *, num_iters, lb=None, ub=None, ylo=None, yhi=None,
                        x0=None, y0=None, tau=None, sigma=None, device=0, handle=None, flags=_cabi.F_DEFAULT,
                        verbose=False):
    """Parity mode: ``num_iters`` fixed-step PDHG iterations on  min c'x, Ax=b (or row senses via
    ylo/yhi), l<=x<=u  from (x0, y0) (default 0):

        g = c - A'y;  x+ = clip(x - tau g, l, u);  xbar = 2x+ - x;  y+ = clip(y + sigma (b - A xbar))

    Returns ``(objective, x, y, info)``.  numpy inputs give numpy outputs (host buffers go
    through ``mllp_pdhg_run_host``, copies included); float64 CUDA tensors for rhs/coefs give
    tensors on the same device with no host synchronisation (``mllp_pdhg_run`` on the current
    stream; ``objective`` is then a 0-d tensor).  ``tau``/``sigma`` default to
    0.9/sigma_max(A).  ``handle`` = a DeviceLP built earlier (e.g. by the loader)."""
    if num_iters < 0:
        raise ValueError("num_iters must be >= 0")
    lp = handle if handle is not None else device_lp(constrs, constr_weights, rhs, coefs, lb, ub, ylo, yhi,
                                                     device, flags)
    if tau is None or sigma is None:
        eta = estimate_step_size(lp)
        tau = eta if tau is None else tau
        sigma = eta if sigma is None else sigma
    L = _cabi.lib()
    if _is_tensor(rhs) or _is_tensor(coefs):
        import torch
        dev = torch.device("cuda", lp.device)
        b = rhs if _is_tensor(rhs) else torch.as_tensor(np.asarray(rhs, dtype=np.float64), device=dev)
        c = coefs if _is_tensor(coefs) else torch.as_tensor(np.asarray(coefs, dtype=np.float64), device=dev)
        lp._check_tensor(b, lp.m, "rhs")
        lp._check_tensor(c, lp.n, "coefs")
        start = lambda v, k, nm: (torch.zeros(k, dtype=torch.float64, device=dev) if v is None else
                                  v.clone() if _is_tensor(v) else torch.as_tensor(_np_f64(v, k, nm), device=dev))
        x, y = start(x0, lp.n, "x0"), start(y0, lp.m, "y0")
        lp._check_tensor(x, lp.n, "x0")
        lp._check_tensor(y, lp.m, "y0")
        scal = torch.empty(_cabi.NUM_SCALARS, dtype=torch.float64, device=dev)
        _cabi.check(L.mllp_pdhg_run(lp.handle, x.data_ptr(), y.data_ptr(), b.data_ptr(), c.data_ptr(), float(tau),
                                    float(sigma), int(num_iters), scal.data_ptr(), _torch_stream(dev)),
                    "mllp_pdhg_run")
        info = {"scalars": scal, "tau": float(tau), "sigma": float(sigma), "handle": lp}
        return scal[0], x, y, info
    b = _np_f64(rhs, lp.m, "rhs")
    c = _np_f64(coefs, lp.n, "coefs")
    # x / y are in-out for the C entry: the caller's x0 / y0 are never written, the results land in page-locked staging
    # arrays of the handle (allocated once), so both directions of the copy run at the pinned rate
    x, y, scal = lp._host_staging()
    x[:] = 0.0 if x0 is None else _np_f64(x0, lp.n, "x0")
    y[:] = 0.0 if y0 is None else _np_f64(y0, lp.m, "y0")
    _cabi.check(L.mllp_pdhg_run_host(lp.handle, _ptr(x), _ptr(y), _ptr(b), _ptr(c), float(tau), float(sigma),
                                     int(num_iters), _ptr(scal), None), "mllp_pdhg_run_host")
    info = _info_dict(scal)
    info.update(tau=float(tau), sigma=float(sigma), handle=lp)
    if verbose:
        print("pdhg: %d iters  pobj %.9g  dobj %.9g  rel_kkt %.3e" % (num_iters, scal[0], scal[1], scal[8]))
    return float(scal[0]), x.copy(), y.copy(), info


def solve_linear_program(constrs, constr_weights, rhs, coefs, *, tol=1e-6, max_iters=200000, check_every=64,
                         lb=None, ub=None, ylo=None, yhi=None, x0=None, y0=None, eta=None, primal_weight=1.0,
                         device=0, handle=None, flags=_cabi.F_DEFAULT, verbose=False, precondition=False):
    """Solve mode: reflected restarted Halpern PDHG on the device until the relative KKT error is
    <= ``tol`` (spec: oracle_pdhg_solve).  Returns ``(objective, x, y, info)``; numpy in/out.
    Non-convergence within ``max_iters`` is reported in ``info['converged']``, not raised.
    ``primal_weight=None`` selects PDLP's initial weight ||c|| / ||b|| (of the scaled LP on a preconditioned handle),
    computed on the device -- measured: it brings 4 of the 7 first-order-hard Netlib files (bnl1, pilot4, pilot.we, greenbea)
    to 1e-6 within 4e6 iterations where 1.0 brings 2, but it slows d2q06c and 25fv47 down several times, so 1.0 stays the
    default.  ``precondition=True`` builds (and caches) a handle with the device-side Ruiz + Pock-Chambolle scaling: typically
    2-4x fewer iterations; x, y, the objective and the KKT error -- hence termination -- refer to the ORIGINAL LP."""
    import torch
    lp = handle if handle is not None else device_lp(constrs, constr_weights, rhs, coefs, lb, ub, ylo, yhi,
                                                     device, flags, precondition=precondition)
    if eta is None:
        sm = lp.sigma_max_robust()
        eta = 0.99 / sm if sm > 0 else 1.0
    dev = torch.device("cuda", lp.device)
    as_t = lambda a, n, name: torch.as_tensor(_np_f64(a, n, name), device=dev)
    tensors_in = _is_tensor(rhs)
    b = rhs if tensors_in else as_t(rhs, lp.m, "rhs")
    c = coefs if _is_tensor(coefs) else as_t(coefs, lp.n, "coefs")
    x = torch.zeros(lp.n, dtype=torch.float64, device=dev) if x0 is None else (x0.clone() if _is_tensor(x0) else as_t(x0, lp.n, "x0"))
    y = torch.zeros(lp.m, dtype=torch.float64, device=dev) if y0 is None else (y0.clone() if _is_tensor(y0) else as_t(y0, lp.m, "y0"))
    for t, n, nm in ((b, lp.m, "rhs"), (c, lp.n, "coefs"), (x, lp.n, "x0"), (y, lp.m, "y0")):
        lp._check_tensor(t, n, nm)
    scal = torch.empty(_cabi.NUM_SCALARS, dtype=torch.float64, device=dev)
    _cabi.check(_cabi.lib().mllp_pdhg_solve(lp.handle, x.data_ptr(), y.data_ptr(), b.data_ptr(), c.data_ptr(),
                                            float(eta), 0.0 if primal_weight is None else float(primal_weight), int(max_iters), int(check_every),
                                            float(tol), scal.data_ptr(), _torch_stream(dev)), "mllp_pdhg_solve")
    if tensors_in:
        return scal[0], x, y, {"scalars": scal, "eta": float(eta), "handle": lp}
    s = scal.cpu().numpy()
    info = _info_dict(s)
    info.update(eta=float(eta), handle=lp)
    if verbose:
        print("solve: %d iters, %d restarts, converged=%s, pobj %.9g, rel_kkt %.3e"
              % (info["iters"], info["restarts"], info["converged"], s[0], s[8]))
    return float(s[0]), x.cpu().numpy(), y.cpu().numpy(), info


# ---------------------------------------------------------------------------------------------
# batched multi-instance mode: many LPs in one launch


class BatchLP:
    """Device formats of a batch of LP instances (``mllp_batch_t``).

    ``instances``: list of ``(constrs, constr_weights, rhs, coefs)`` in the loader's
    representation (the per-instance tuple of linear_program_experiment.py:123 without name and
    labels), or -- ``shared=True`` -- ONE such matrix used by ``count`` instances that differ
    only in b and c (BASELINE.json configs[4])."""

    def __init__(self, instances, shared=False, count=None, device=0, precondition=False):
        L = _cabi.lib()
        self.device = _device_index(device)
        self.preconditioned = bool(precondition)
        self.shared = bool(shared)
        mats = [instances[0]] if self.shared else list(instances)
        self.count = int(count if self.shared else len(mats))
        if self.count < 1:
            raise ValueError("empty batch")
        ms, ns, ips, iis, vvs = [], [], [], [], []
        for constrs, weights, rhs, coefs in mats:
            m, n = len(rhs) if np.ndim(rhs) == 1 else np.shape(rhs)[-1], len(coefs) if np.ndim(coefs) == 1 else np.shape(coefs)[-1]
            ip, ii, vv = csr_from_constrs(constrs, weights, n)
            if ip.shape[0] != m + 1:
                raise ValueError("constrs has %d rows, rhs has %d" % (ip.shape[0] - 1, m))
            ms.append(m); ns.append(n); ips.append(ip); iis.append(ii); vvs.append(vv)
        self.m = np.array(ms * (self.count if self.shared else 1), dtype=np.int64)
        self.n = np.array(ns * (self.count if self.shared else 1), dtype=np.int64)
        h_m, h_n = np.array(ms, dtype=np.int32), np.array(ns, dtype=np.int32)
        ip_off = np.concatenate([[0], np.cumsum([a.shape[0] for a in ips])[:-1]]).astype(np.int64)
        nz_off = np.concatenate([[0], np.cumsum([a.shape[0] for a in iis])[:-1]]).astype(np.int64)
        ip_all = np.ascontiguousarray(np.concatenate(ips), dtype=np.int32)
        ii_all = np.ascontiguousarray(np.concatenate(iis), dtype=np.int32) if sum(a.shape[0] for a in iis) else np.zeros(1, np.int32)
        vv_all = np.ascontiguousarray(np.concatenate(vvs), dtype=np.float64) if sum(a.shape[0] for a in vvs) else np.zeros(1)
        h = ctypes.c_void_p()
        rc = L.mllp_batch_create(self.count, int(self.shared), _ptr(h_m), _ptr(h_n), _ptr(ip_off), _ptr(nz_off),
                                 _ptr(ip_all), _ptr(ii_all), _ptr(vv_all), self.device,
                                 _cabi.F_PRECONDITION if precondition else 0, ctypes.byref(h))
        _cabi.check(rc, "mllp_batch_create")
        self._h = h
        self._finalizer = weakref.finalize(self, L.mllp_batch_destroy, h)
        self.x_off = np.concatenate([[0], np.cumsum(self.n)])
        self.y_off = np.concatenate([[0], np.cumsum(self.m)])
        self._sigma = None
        self._sigma_robust = None
        # sqrt(||A||_1 ||A||_inf) >= ||A||_2 per distinct matrix: the guaranteed side of the step-size estimate
        ups = []
        for ip, ii, vv, m_, n_ in zip(ips, iis, vvs, ms, ns):
            if vv.shape[0] == 0:
                ups.append(0.0)
                continue
            av = np.abs(vv)
            rs = np.bincount(np.repeat(np.arange(m_), np.diff(ip)), weights=av, minlength=m_).max()
            cs = np.bincount(ii, weights=av, minlength=n_).max()
            ups.append(float(np.sqrt(rs * cs)))
        self.norm_upper = np.array(ups * (self.count if self.shared else 1), dtype=np.float64)

    @property
    def handle(self):
        return self._h

    def close(self):
        self._finalizer()
        self._h = None

    def info(self):
        out = (ctypes.c_int64 * 16)()
        _cabi.check(_cabi.lib().mllp_batch_info(self.handle, out), "mllp_batch_info")
        keys = ("count", "sum_m", "sum_n", "sum_nnz", "grid_ctas", "threads", "dyn_smem_bytes", "bytes_per_iter",
                "instances_per_cta", "instances_per_cta_solve", "res_steps_A", "res_steps_AT", "res_steps_A_solve",
                "res_steps_AT_solve", "dyn_smem_bytes_solve", "grid_ctas_solve")
        return dict(zip(keys, (int(v) for v in out)))

    def sigma_max(self, iters=50):
        """per-instance ||A_k||_2 estimates (device tensor, cached per step count)."""
        import torch
        if self._sigma is None:
            self._sigma = {}
        if iters not in self._sigma:
            dev = torch.device("cuda", self.device)
            s = torch.zeros(self.count, dtype=torch.float64, device=dev)
            _cabi.check(_cabi.lib().mllp_batch_estimate_norm(self.handle, int(iters), s.data_ptr(), _torch_stream(dev)),
                        "mllp_batch_estimate_norm")
            self._sigma[iters] = s
        return self._sigma[iters]

    def sigma_max_robust(self, rel_change=1e-4, max_iters=3200):
        """solve mode, where an underestimate of ||A_k||_2 makes an instance diverge: power iteration with doubling step
        counts until no instance's estimate moves by more than ``rel_change`` (as DeviceLP.sigma_max_robust), inflated by
        2 % and capped by sqrt(||A_k||_1 ||A_k||_inf) where that bound is known (``norm_upper``, per instance)."""
        import torch
        if getattr(self, "_sigma_robust", None) is None:
            prev, iters = None, 100
            while True:
                cur = self.sigma_max(iters)
                if prev is not None and bool(((cur - prev).abs() <= rel_change * cur).all()):
                    break
                if iters >= max_iters:
                    break
                prev, iters = cur, iters * 2
            est = 1.02 * cur
            if self.norm_upper is not None and not self.preconditioned:   # the bound is on the original matrices
                ub = torch.as_tensor(self.norm_upper, device=est.device)
                est = torch.where(ub > 0, torch.minimum(est, ub), est)
            self._sigma_robust = est
        return self._sigma_robust

    # device-tensor level entries (concatenated vectors, caller's stream, no host sync)
    def run(self, x, y, b, c, tau, sigma, num_iters, scalars=None):
        _cabi.check(_cabi.lib().mllp_batch_run(self.handle, x.data_ptr(), y.data_ptr(), b.data_ptr(), c.data_ptr(),
                                               tau.data_ptr(), sigma.data_ptr(), int(num_iters),
                                               None if scalars is None else scalars.data_ptr(),
                                               _torch_stream(x.device)), "mllp_batch_run")

    def solve(self, x, y, b, c, eta, scalars, w0=1.0, max_iters=200000, check_every=64, tol=1e-6):
        """w0 = 0: every instance starts from PDLP's initial primal weight ||c|| / ||b|| (computed in the kernel)"""
        _cabi.check(_cabi.lib().mllp_batch_solve(self.handle, x.data_ptr(), y.data_ptr(), b.data_ptr(), c.data_ptr(),
                                                 eta.data_ptr(), float(w0), int(max_iters), int(check_every), float(tol),
                                                 scalars.data_ptr(), _torch_stream(x.device)), "mllp_batch_solve")


def _batch_vectors(bt, instances, rhs_batch, coefs_batch, x0, y0):
    import torch
    dev = torch.device("cuda", bt.device)
    if bt.shared:
        b = np.ascontiguousarray(rhs_batch, dtype=np.float64).reshape(bt.count, -1)
        c = np.ascontiguousarray(coefs_batch, dtype=np.float64).reshape(bt.count, -1)
        if b.shape[1] != bt.m[0] or c.shape[1] != bt.n[0]:
            raise ValueError("rhs/coefs batches must be (count, m) and (count, n)")
        b, c = b.reshape(-1), c.reshape(-1)
    else:
        b = np.concatenate([_np_f64(i[2], int(m), "rhs") for i, m in zip(instances, bt.m)])
        c = np.concatenate([_np_f64(i[3], int(n), "coefs") for i, n in zip(instances, bt.n)])
    nx, ny = int(bt.x_off[-1]), int(bt.y_off[-1])
    x = np.zeros(nx) if x0 is None else np.concatenate([np.asarray(v, dtype=np.float64).reshape(-1) for v in x0])
    y = np.zeros(ny) if y0 is None else np.concatenate([np.asarray(v, dtype=np.float64).reshape(-1) for v in y0])
    if x.shape[0] != nx or y.shape[0] != ny:
        raise ValueError("x0 / y0 sizes do not match the batch")
    t = lambda a: torch.as_tensor(a, device=dev)
    return t(b), t(c), t(x), t(y), dev


def _batch_results(bt, x, y, scal, extra):
    xs, ys, sc = x.cpu().numpy(), y.cpu().numpy(), scal.cpu().numpy().reshape(bt.count, _cabi.NUM_SCALARS)
    out = []
    for k in range(bt.count):
        info = _info_dict(sc[k])
        info.update({key: float(val[k]) for key, val in extra.items()})
        out.append((float(sc[k, 0]), xs[bt.x_off[k]:bt.x_off[k + 1]], ys[bt.y_off[k]:bt.y_off[k + 1]], info))
    return out


def pdhg_linear_program_batch(instances, *, num_iters, x0=None, y0=None, tau=None, sigma=None, device=0, handle=None,
                              shared=False, rhs_batch=None, coefs_batch=None):
    """Parity mode on a whole batch in ONE launch (one CTA per LP).  ``instances`` = list of
    ``(constrs, constr_weights, rhs, coefs)``; with ``shared=True`` one matrix
    ``instances[0]`` and ``rhs_batch`` (B, m) / ``coefs_batch`` (B, n).  ``tau`` / ``sigma``:
    scalars, per-instance arrays, or None (0.9 / sigma_max(A_k) computed on the device).
    Returns a list of ``(objective, x, y, info)`` per instance."""
    import torch
    bt = handle if handle is not None else BatchLP(instances, shared=shared,
                                                   count=None if not shared else len(rhs_batch), device=device)
    b, c, x, y, dev = _batch_vectors(bt, instances, rhs_batch, coefs_batch, x0, y0)
    if tau is None or sigma is None:
        eta = 0.9 / bt.sigma_max()
    as_arr = lambda v: eta.clone() if v is None else torch.as_tensor(np.broadcast_to(np.asarray(v, dtype=np.float64), (bt.count,)).copy(), device=dev)
    tau_t, sigma_t = as_arr(tau), as_arr(sigma)
    scal = torch.zeros(bt.count * _cabi.NUM_SCALARS, dtype=torch.float64, device=dev)
    bt.run(x, y, b, c, tau_t, sigma_t, num_iters, scal)
    return _batch_results(bt, x, y, scal, {"tau": tau_t.cpu().numpy(), "sigma": sigma_t.cpu().numpy()})


def solve_linear_program_batch(instances, *, tol=1e-6, max_iters=200000, check_every=64, x0=None, y0=None, eta=None,
                               primal_weight=1.0, device=0, handle=None, shared=False, rhs_batch=None, coefs_batch=None,
                               scale=False):
    """Solve mode on a whole batch in one launch; every instance restarts and terminates on
    its own.  Returns a list of ``(objective, x, y, info)``.  ``scale=True``: every distinct matrix is preconditioned
    (Ruiz + Pock-Chambolle, computed on the device by mllp_batch_create with MLLP_F_PRECONDITION) -- typically 2-4x fewer
    iterations on the Netlib instances; x, y, the objective and the KKT error, hence every instance's termination,
    refer to the ORIGINAL LP (``info['rel_kkt_original']`` repeats ``info['rel_kkt']``)."""
    import torch
    bt = handle if handle is not None else BatchLP(instances, shared=shared, count=None if not shared else len(rhs_batch),
                                                   device=device, precondition=scale)
    b, c, x, y, dev = _batch_vectors(bt, instances, rhs_batch, coefs_batch, x0, y0)
    eta_t = 0.99 / bt.sigma_max_robust() if eta is None else torch.as_tensor(
        np.broadcast_to(np.asarray(eta, dtype=np.float64), (bt.count,)).copy(), device=dev)
    scal = torch.zeros(bt.count * _cabi.NUM_SCALARS, dtype=torch.float64, device=dev)
    bt.solve(x, y, b, c, eta_t, scal, 0.0 if primal_weight is None else float(primal_weight), max_iters, check_every, tol)
    res = _batch_results(bt, x, y, scal, {"eta": eta_t.cpu().numpy()})
    for r in res:
        r[3]["rel_kkt_original"] = r[3]["rel_kkt"]
    return res
