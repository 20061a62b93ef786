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
import os

import numpy as np

from . import _cabi
from .linear_program_methods import (DeviceLP, _device_index, _info_dict, _np_f64, _ptr, _torch_stream,
                                     csr_from_constrs, pdhg_linear_program_batch, solve_linear_program_batch)


def _dup_fd_from(pid, fd):
    """Duplicate file descriptor `fd` of process `pid` into this process (pidfd_open + pidfd_getfd, Linux >= 5.6; the
    ranks of one node run under one user).  Returns -1 if the kernel refuses."""
    libc = ctypes.CDLL(None, use_errno=True)
    SYS_pidfd_open, SYS_pidfd_getfd = 434, 438        # x86-64 and aarch64 share these numbers
    pfd = libc.syscall(SYS_pidfd_open, int(pid), 0)
    if pfd < 0:
        return -1
    try:
        return int(libc.syscall(SYS_pidfd_getfd, pfd, int(fd), 0))
    finally:
        os.close(pfd)


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
        self.multicast = False
        # NVSwitch multicast pays from 4 ranks on (8 GPUs: ken-18 15.7 -> 14.0 us, osa-60 15.2 -> 14.6); on 2 it only adds a
        # second copy of every word (ken-18 8.1 -> 8.5 us).  MLLP_ROWPART_MC = 0 / 1 forces it off / on.
        if p2p and 1 < self.world <= 8 and os.environ.get("MLLP_ROWPART_MC", "1" if self.world >= 4 else "0") != "0":
            self.multicast = self._setup_multicast(dist, L, h)
            self.p2p = self.multicast
        if p2p and 1 < self.world <= 8 and not self.multicast:
            # in-kernel exchange over NVLink peer memory: swap the CUDA IPC handles of the ranks' mailboxes
            mine = np.zeros(64, dtype=np.uint8)
            _cabi.check(L.mllp_rowpart_ipc_export(h, mine.ctypes.data), "mllp_rowpart_ipc_export")
            parts = [None] * self.world
            dist.all_gather_object(parts, (self.rank, mine.tobytes()))
            blob = np.frombuffer(b"".join(b for _, b in sorted(parts)), dtype=np.uint8).copy()
            _cabi.check(L.mllp_rowpart_ipc_import(h, blob.ctypes.data), "mllp_rowpart_ipc_import")
            dist.barrier()
            self.p2p = True

    def _setup_multicast(self, dist, L, h):
        """NVSwitch multicast mailbox (mllp_rowpart_mc_*): one multimem.st per dual value instead of world - 1 peer stores.
        Collective; returns False (on every rank) when any rank cannot do it -- the caller then swaps IPC handles instead."""
        import torch
        dev = torch.device("cuda", self.device)

        def all_ok(flag):
            t = torch.tensor([1 if flag else 0], dtype=torch.int32, device=dev if dist.get_backend() == "nccl" else "cpu")
            dist.all_reduce(t, op=dist.ReduceOp.MIN)
            return bool(int(t[0]))

        sup = ctypes.c_int32(0)
        ok = L.mllp_rowpart_mc_supported(h, ctypes.byref(sup)) == 0 and sup.value == 1
        if not all_ok(ok):
            return False
        fd = ctypes.c_int32(-1)
        ok = True
        if self.rank == 0:
            ok = L.mllp_rowpart_mc_create(h, ctypes.byref(fd)) == 0
        box = [(os.getpid(), int(fd.value)) if ok else None]
        dist.broadcast_object_list(box, src=0)
        if box[0] is None:
            return False
        my_fd = int(fd.value)
        if self.rank != 0:
            my_fd = _dup_fd_from(box[0][0], box[0][1])
        ok = my_fd >= 0 and L.mllp_rowpart_mc_attach(h, my_fd) == 0
        if not all_ok(ok):          # also the barrier "every device has been added"
            return False
        ok = L.mllp_rowpart_mc_bind(h) == 0
        if not all_ok(ok):          # ... and "every mailbox is bound"
            raise RuntimeError("mllp_b200: binding the multicast mailbox failed on some rank: %s" % _cabi.last_error())
        if my_fd >= 0:
            os.close(my_fd)         # the driver holds its own reference to the object
        return True

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
