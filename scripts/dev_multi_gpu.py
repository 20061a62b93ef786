"""Multi-GPU dev check (torchrun): row-partitioned parity + timing, data-parallel batch."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
import mllp_b200 as M
from mllp_b200 import _cabi
from mllp_b200.distributed import RowPartLP, pdhg_linear_program_rowpart, solve_batch_data_parallel
from oracle import pdhg_oracle as O

rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); local = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
def p0(*a):
    if rank == 0: print(*a, flush=True)

for name, K in (("afiro", 200), ("pilot87", 200), ("osa-60", 100), ("ken-18", 100), ("pds-20", 100)):
    A, b, c = M.load_csr(name); m, n = A.shape
    eta = 0.9 / O.power_iteration(A, 50)
    lp = RowPartLP(A, A.data, m, n, device=local)
    obj, x, y, info = pdhg_linear_program_rowpart(lp, b, c, num_iters=K, tau=eta, sigma=eta)
    xo, yo = O.pdhg_run(A, b, c, np.zeros(n), np.zeros(m), eta, eta, K)
    kk = O.kkt(A, b, c, xo, yo)
    ex, ey = np.linalg.norm(x - xo) / np.linalg.norm(xo), np.linalg.norm(y - yo) / np.linalg.norm(yo)
    print("rank %d %s rowpart parity x %.2e y %.2e obj %.9g (oracle %.9g) kkt %.3e (oracle %.3e)" % (rank, name, ex, ey, obj, kk[0], info["rel_kkt"], kk[8]), flush=True)
    assert ex < 1e-9 and ey < 1e-9
    # timing
    dev = torch.device("cuda", local)
    bt, ct = torch.tensor(b, device=dev), torch.tensor(c, device=dev)
    xt = torch.zeros(n, dtype=torch.float64, device=dev); yt = torch.zeros(m, dtype=torch.float64, device=dev)
    L = _cabi.lib(); s = torch.cuda.current_stream().cuda_stream
    KK = 300
    def run():
        _cabi.check(L.mllp_pdhg_run(lp.handle, xt.data_ptr(), yt.data_ptr(), bt.data_ptr(), ct.data_ptr(), eta, eta, KK, None, s), "run")
    run(); torch.cuda.synchronize(); dist.barrier()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(); run(); e1.record(); torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1)], device=dev); dist.all_reduce(t, op=dist.ReduceOp.MAX)
    p0("  %s row-partitioned over %d GPUs: %.2f us/iter" % (name, world, float(t[0]) * 1e3 / KK))
    lp.close()

# data-parallel batch: 64 perturbed sc105 instances, solve mode
A, b, c = M.load_csr("sc105")
insts = []
for i in range(64):
    g = np.random.default_rng(1234 + i)
    insts.append((A, A.data, b * (1 + 0.1 * g.uniform(0, 1, b.shape[0])) if False else b, c * (1 + 0.05 * g.uniform(-1, 1, c.shape[0]))))
t0 = time.time()
res = solve_batch_data_parallel(insts, mode="solve", device=local, tol=1e-6, max_iters=200000)
p0("data-parallel batch: %d instances, converged %d, wall %.2f s, first objs %s" % (len(res), sum(r[3]["converged"] for r in res), time.time() - t0, [round(r[0], 5) for r in res[:3]]))
single = M.solve_linear_program_batch(insts[:2], device=local, tol=1e-6, max_iters=200000)
assert abs(single[1][0] - res[1][0]) < 1e-12
dist.destroy_process_group()
