"""Multi-GPU host logic: one process per GPU, torch.distributed for the plumbing.

Two ways the path shards:

* independent LP instances (batches): contiguous blocks of instances per rank, NO data-path
  collective; one gather of the per-instance results at the end
  (``shard_range`` / ``solve_batch_data_parallel``);
* ONE large LP (ken-18, osa-60, pds-20): row partition of A inside the C library, the A' phase
  replicated, ONE exchange of the y slices per iteration -- tagged words through peer mailboxes
  over NVLink inside the persistent kernel, or one NCCL all-gather per iteration
  (``RowPartLP`` / ``pdhg_linear_program_rowpart``).  All ranks' results are identical to the
  single-GPU path up to summation order.
"""
import ctypes

import numpy as np

from . import _cabi
from .linear_program_methods import (DeviceLP, _device_index, _info_dict, _np_f64, _ptr, _torch_stream,
                                     csr_from_constrs, pdhg_linear_program_batch, solve_linear_program_batch)


def shard_range(count, rank, world):
    """Contiguous block [lo, hi) of `count` items owned by `rank` (sizes differ by at most 1)."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("bad rank/world")
    base, extra = divmod(count, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def _dist():
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        raise RuntimeError("torch.distributed is not initialised")
    return dist


def gather_results(local_results, count):
    """All ranks get the full list of per-instance results in instance order."""
    dist = _dist()
    world, rank = dist.get_world_size(), dist.get_rank()
    parts = [None] * world
    dist.all_gather_object(parts, (rank, local_results))
    out = []
    for r, res in sorted(parts, key=lambda t: t[0]):
        lo, hi = shard_range(count, r, world)
        if len(res) != hi - lo:
            raise RuntimeError("rank %d returned %d results for a shard of %d" % (r, len(res), hi - lo))
        out.extend(res)
    return out


def solve_batch_data_parallel(instances, *, mode="solve", device=None, compute=None, shared=False, rhs_batch=None,
                              coefs_batch=None, single_process=False, count=None, **kwargs):
    """Shard independent LP instances over the ranks, run the batched kernel on each rank's
    block, gather.  ``mode`` = "solve" (to tolerance) or "run" (fixed ``num_iters``).
    ``shared=True``: ONE matrix ``instances[0]`` and (B, m) / (B, n) batches of right-hand sides and costs
    (BASELINE.json configs[4]); the rows of the batches are what is sharded.
    With ``count`` given, ``rhs_batch`` / ``coefs_batch`` hold only THIS rank's rows [lo, hi) of the ``count`` instances
    (each rank generated or loaded its own shard); otherwise every rank passes the whole batches.
    ``compute`` overrides the per-shard solver (used by the CPU tests of this host logic).
    ``single_process=True``: no process group (one GPU): the whole batch is this process's shard."""
    if single_process:
        world, rank = 1, 0
    else:
        dist = _dist()
        world, rank = dist.get_world_size(), dist.get_rank()
    local_rows = shared and count is not None
    count = int(count) if local_rows else (len(rhs_batch) if shared else len(instances))
    lo, hi = shard_range(count, rank, world)
    if local_rows and len(rhs_batch) != hi - lo:
        raise ValueError("rank %d holds %d rows of the batches, its shard has %d" % (rank, len(rhs_batch), hi - lo))
    if compute is None:
        fn = solve_linear_program_batch if mode == "solve" else pdhg_linear_program_batch
        dev = rank if device is None else device
        if shared:
            pick = (lambda a, sl: np.asarray(a)) if local_rows else (lambda a, sl: np.asarray(a)[sl])
            compute = lambda sl: fn(instances[:1], device=dev, shared=True, rhs_batch=pick(rhs_batch, sl),
                                    coefs_batch=pick(coefs_batch, sl), **kwargs) if sl.stop > sl.start else []
        else:
            compute = lambda shard: fn(shard, device=dev, **kwargs) if shard else []
    local = compute(slice(lo, hi) if shared else instances[lo:hi])
    # results carry numpy arrays only (picklable); drop handles
    local = [(o, x, y, {k: v for k, v in info.items() if k != "handle"}) for (o, x, y, info) in local]
    if single_process:
        return local
    return gather_results(local, count)


def broadcast_unique_id(src=0):
    """NCCL unique id from rank `src` to everyone, as 128 bytes."""
    import torch
    dist = _dist()
    buf = np.zeros(128, dtype=np.uint8)
    if dist.get_rank() == src:
        _cabi.check(_cabi.lib().mllp_nccl_unique_id(buf.ctypes.data), "mllp_nccl_unique_id")
    backend = dist.get_backend()
    t = torch.from_numpy(buf)
    if backend == "nccl":
        t = t.cuda()
    dist.broadcast(t, src=src)
    return t.cpu().numpy().copy()


class RowPartLP(DeviceLP):
    """Row-partitioned handle of one large LP; collective constructor (all ranks call it)."""

    def __init__(self, constrs, constr_weights, num_rows, num_cols, lb=None, ub=None, ylo=None, yhi=None, device=None,
                 flags=_cabi.F_DEFAULT, p2p=True):
        import weakref
        dist = _dist()
        L = _cabi.lib()
        self.rank, self.world = dist.get_rank(), dist.get_world_size()
        self.m, self.n = int(num_rows), int(num_cols)
        indptr, indices, values = csr_from_constrs(constrs, constr_weights, self.n)
        if indptr.shape[0] != self.m + 1:
            raise ValueError("constrs has %d rows, rhs has %d" % (indptr.shape[0] - 1, self.m))
        self.nnz = int(indptr[-1])
        self.device = _device_index(self.rank if device is None else device)
        self.flags = int(flags)
        f = lambda a, k, nm: None if a is None else _np_f64(a, k, nm)
        lb, ub, ylo, yhi = f(lb, self.n, "lb"), f(ub, self.n, "ub"), f(ylo, self.m, "ylo"), f(yhi, self.m, "yhi")
        uid = broadcast_unique_id(0) if self.world > 1 else np.zeros(128, dtype=np.uint8)
        h = ctypes.c_void_p()
        rc = L.mllp_lp_create_rowpart(self.m, self.n, self.nnz, _ptr(indptr), _ptr(indices), _ptr(values), _ptr(lb),
                                      _ptr(ub), _ptr(ylo), _ptr(yhi), self.device, self.flags, self.rank, self.world,
                                      uid.ctypes.data, ctypes.byref(h))
        _cabi.check(rc, "mllp_lp_create_rowpart")
        self._h = h
        self._finalizer = weakref.finalize(self, L.mllp_lp_destroy, h)
        self._sigma_max = None
        self._sigma_robust = None
        self.norm_upper = 0.0
        self.p2p = False
        if p2p and 1 < self.world <= 8:
            # in-kernel exchange over NVLink peer memory: swap the CUDA IPC handles of the ranks' mailboxes
            mine = np.zeros(64, dtype=np.uint8)
            _cabi.check(L.mllp_rowpart_ipc_export(h, mine.ctypes.data), "mllp_rowpart_ipc_export")
            parts = [None] * self.world
            dist.all_gather_object(parts, (self.rank, mine.tobytes()))
            blob = np.frombuffer(b"".join(b for _, b in sorted(parts)), dtype=np.uint8).copy()
            _cabi.check(L.mllp_rowpart_ipc_import(h, blob.ctypes.data), "mllp_rowpart_ipc_import")
            dist.barrier()
            self.p2p = True

    def sigma_max(self, iters=50, stream=None):
        raise RuntimeError("the power iteration is not available on a row-partitioned handle (mllp_estimate_norm); "
                           "estimate the step size on a single-GPU DeviceLP of the same matrix and pass tau / sigma")

    sigma_max_robust = sigma_max

    def exchange_error(self):
        f = ctypes.c_int32(0)
        _cabi.check(_cabi.lib().mllp_rowpart_error(self.handle, ctypes.byref(f)), "mllp_rowpart_error")
        return int(f.value)


def pdhg_linear_program_rowpart(lp, rhs, coefs, *, num_iters, tau, sigma, x0=None, y0=None):
    """Collective parity-mode run on a RowPartLP: every rank passes the same full-length rhs /
    coefs (numpy) and receives the full (objective, x, y, info)."""
    import torch
    dev = torch.device("cuda", lp.device)
    t = lambda a, k, nm: torch.as_tensor(_np_f64(a, k, nm), device=dev)
    b, c = t(rhs, lp.m, "rhs"), t(coefs, lp.n, "coefs")
    x = torch.zeros(lp.n, dtype=torch.float64, device=dev) if x0 is None else t(x0, lp.n, "x0")
    y = torch.zeros(lp.m, dtype=torch.float64, device=dev) if y0 is None else t(y0, lp.m, "y0")
    scal = torch.zeros(_cabi.NUM_SCALARS, dtype=torch.float64, device=dev)
    _cabi.check(_cabi.lib().mllp_pdhg_run(lp.handle, x.data_ptr(), y.data_ptr(), b.data_ptr(), c.data_ptr(), float(tau),
                                          float(sigma), int(num_iters), scal.data_ptr(), _torch_stream(dev)),
                "mllp_pdhg_run (row-partitioned)")
    s = scal.cpu().numpy()
    if lp.exchange_error():
        raise RuntimeError("mllp_b200: a cross-GPU exchange wait timed out on rank %d" % lp.rank)
    info = _info_dict(s)
    info.update(tau=float(tau), sigma=float(sigma), p2p=lp.p2p)
    return float(s[0]), x.cpu().numpy(), y.cpu().numpy(), info
