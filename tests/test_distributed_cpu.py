"""CPU tests (gloo, world_size 2) of the multi-GPU host logic, plus the host-side emulation of
the row partition (all ranks replayed on the CPU by mllp_rowpart_selfcheck)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import mllp_b200.linear_program_data as D
from mllp_b200 import _cabi
from mllp_b200.distributed import shard_range


def test_shard_range_covers_everything():
    for count in (0, 1, 5, 8, 4096, 4097):
        for world in (1, 2, 3, 8):
            spans = [shard_range(count, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == count
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_range(4, 2, 2)


@pytest.mark.parametrize("name,nranks", [("afiro", 2), ("pilot87", 2), ("pilot87", 8), ("ken-18", 4), ("osa-60", 2),
                                         ("pds-20", 8)])
def test_row_partition_emulated_on_cpu(name, nranks):
    A, _, _ = D.load_csr(name)
    m, n = A.shape
    ip, ii, vv = A.indptr.astype(np.int32), A.indices.astype(np.int32), np.ascontiguousarray(A.data)
    out = np.zeros(4)
    rc = _cabi.lib().mllp_rowpart_selfcheck(m, n, A.nnz, ip.ctypes.data, ii.ctypes.data, vv.ctypes.data, 148, nranks,
                                            out.ctypes.data)
    assert rc == 0
    assert out[0] < 1e-12                        # every row of A v (partitioned) and A' w (replicated) exactly once
    assert out[1] >= m and out[1] % (2 * nranks) == 0 and out[2] == n
    assert out[3] < 1.25                         # nonzeros balanced over ranks


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from mllp_b200.distributed import solve_batch_data_parallel, gather_results
        instances = [("inst%d" % k, k) for k in range(7)]

        def fake_compute(shard):   # stands in for the batched GPU kernel: (objective, x, y, info)
            return [(float(k), np.full(3, k, dtype=np.float64), np.full(2, -k, dtype=np.float64),
                     {"iters": 10 * k, "rank": dist.get_rank(), "handle": object()}) for _, k in shard]

        res = solve_batch_data_parallel(instances, compute=fake_compute)
        ok = len(res) == 7 and all(r[0] == float(k) and r[1][0] == k and r[3]["iters"] == 10 * k for k, r in enumerate(res))
        ok = ok and [r[3]["rank"] for r in res] == [0, 0, 0, 0, 1, 1, 1] and all("handle" not in r[3] for r in res)
        # shared matrix (BASELINE.json configs[4]): the rows of the (B, m) / (B, n) batches are what is sharded
        bb, cb = np.arange(10, dtype=np.float64).reshape(5, 2), np.arange(15, dtype=np.float64).reshape(5, 3)

        def fake_shared(sl):
            return [(float(cb[i].sum()), cb[i].copy(), bb[i].copy(), {"iters": i}) for i in range(sl.start, sl.stop)]

        res2 = solve_batch_data_parallel([("A", None, None, None)], compute=fake_shared, shared=True, rhs_batch=bb, coefs_batch=cb)
        ok = ok and len(res2) == 5 and all(r[3]["iters"] == i and np.array_equal(r[1], cb[i]) for i, r in enumerate(res2))
        # ... also when every rank passes only its own rows
        lo, hi = (0, 3) if dist.get_rank() == 0 else (3, 5)
        res3 = solve_batch_data_parallel([("A", None, None, None)], compute=fake_shared, shared=True, rhs_batch=bb[lo:hi],
                                         coefs_batch=cb[lo:hi], count=5)
        ok = ok and [r[3]["iters"] for r in res3] == [0, 1, 2, 3, 4]
        # a wrong shard size is detected
        try:
            gather_results([1, 2, 3] if rank == 0 else [1], 7)
            ok = False
        except RuntimeError:
            pass
        q.put((rank, ok))
    finally:
        dist.destroy_process_group()


def test_data_parallel_sharding_and_gather_gloo_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert sorted(results) == [(0, True), (1, True)]
