import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import mllp_b200 as M
from oracle import pdhg_oracle as O
names = ['sc50a', 'sc105', 'adlittle', 'blend', 'share2b', 'kb2']
insts, mats = [], []
for nm in names:
    A, b, c = M.load_csr(nm); insts.append((A, A.data, b, c)); mats.append((A, b, c))
bt = M.BatchLP(insts); print(bt.info())
sig = bt.sigma_max().cpu().numpy(); print('sigma', sig, [O.power_iteration(A, 50) for A, _, _ in mats])
K = 1000
res = M.pdhg_linear_program_batch(insts, num_iters=K, handle=bt)
for (A, b, c), (obj, x, y, info), s in zip(mats, res, sig):
    eta = 0.9 / s
    xo, yo = O.pdhg_run(A, b, c, np.zeros(A.shape[1]), np.zeros(A.shape[0]), eta, eta, K)
    print(' parity x %.2e y %.2e obj %.6f kkt %.2e' % (np.linalg.norm(x - xo) / np.linalg.norm(xo), np.linalg.norm(y - yo) / np.linalg.norm(yo), obj, info['rel_kkt']))
t = time.time(); res = M.solve_linear_program_batch(insts[:5], tol=1e-6, max_iters=400000); dt = time.time() - t
for nm, (obj, x, y, info) in zip(names, res): print(' solve', nm, obj, info['iters'], info['restarts'], info['converged'], info['rel_kkt'])
print(' batch solve wall %.3f s' % dt)
# shared-A batch: 25fv47 with perturbed b, c
A, b, c = M.load_csr('25fv47'); m, n = A.shape
B = 1024
rng = np.random.default_rng(0)
cb = c[None, :] * (1 + 0.1 * rng.uniform(-1, 1, (B, n))); bb = np.tile(b, (B, 1))
bts = M.BatchLP([(A, A.data, b, c)], shared=True, count=B); print(bts.info())
res = M.pdhg_linear_program_batch([(A, A.data, b, c)], num_iters=200, handle=bts, shared=True, rhs_batch=bb, coefs_batch=cb)
eta = 0.9 / O.power_iteration(A, 50)
for k in (0, 17, B - 1):
    xo, yo = O.pdhg_run(A, bb[k], cb[k], np.zeros(n), np.zeros(m), eta, eta, 200)
    print(' shared parity', k, np.linalg.norm(res[k][1] - xo) / np.linalg.norm(xo), np.linalg.norm(res[k][2] - yo) / np.linalg.norm(yo))
# timing
dev = torch.device('cuda')
for (bt_, nx, ny, label) in ((bt, int(bt.x_off[-1]), int(bt.y_off[-1]), 'small6'), (bts, B * n, B * m, 'shared 1024 x 25fv47')):
    x = torch.zeros(nx, dtype=torch.float64, device=dev); y = torch.zeros(ny, dtype=torch.float64, device=dev)
    bv = torch.randn(ny, dtype=torch.float64, device=dev); cv = torch.randn(nx, dtype=torch.float64, device=dev)
    tau = torch.full((bt_.count,), 0.1, dtype=torch.float64, device=dev)
    KK = 2000
    bt_.run(x, y, bv, cv, tau, tau, KK); torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(); bt_.run(x, y, bv, cv, tau, tau, KK); e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    print(' %s: %d iters in %.3f ms -> %.3f us per batch-iteration, %.3g LP-iterations/s, %.1f GB/s algorithmic' % (label, KK, ms, ms * 1e3 / KK, bt_.count * KK / ms * 1e3, bt_.info()['bytes_per_iter'] * KK / ms / 1e6))
