"""Dev tool: LP-iterations/s of the shared-matrix batch (4096 perturbed 25fv47) for R = 1..3 and residency on/off."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import mllp_b200 as M

def run(name, B, iters, **env):
    for k, v in env.items(): os.environ[k] = str(v)
    A, b, c = M.load_csr(name); m, n = A.shape
    rng = np.random.default_rng(0)
    dev = torch.device('cuda', 0)
    bt = M.BatchLP([(A, A.data, b, c)], shared=True, count=B)
    bb = torch.tensor(np.tile(b, B) * (1 + 0.1 * rng.uniform(0, 1, B * m)), device=dev)
    cb = torch.tensor(np.tile(c, B) * (1 + 0.1 * rng.uniform(-1, 1, B * n)), device=dev)
    eta = (0.9 / bt.sigma_max()).contiguous()
    x = torch.zeros(B * n, dtype=torch.float64, device=dev); y = torch.zeros(B * m, dtype=torch.float64, device=dev)
    bt.run(x, y, bb, cb, eta, eta, 10); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); bt.run(x, y, bb, cb, eta, eta, iters); e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    info = bt.info()
    print('%s B=%d %s: %.3e LP-it/s  (%.1f us / batch iteration) R=%d res A %d AT %d smem %d grid %d threads %d' % (
        name, B, env, B * iters / (ms * 1e-3), ms * 1e3 / iters, info['instances_per_cta'], info['res_steps_A'], info['res_steps_AT'],
        info['dyn_smem_bytes'], info['grid_ctas'], info['threads']), flush=True)
    # solve mode
    etas = (0.99 / bt.sigma_max_robust()).contiguous()
    scal = torch.zeros(B * 16, dtype=torch.float64, device=dev)
    x.zero_(); y.zero_()
    e0.record(); bt.solve(x, y, bb, cb, etas, scal, 1.0, 200000, 64, 1e-6); e1.record(); torch.cuda.synchronize()
    sc = scal.cpu().numpy().reshape(B, 16)
    print('   solve: %.1f LPs/s  converged %.3f  mean iters %.0f (min %.0f max %.0f)  R=%d' % (
        B / (e0.elapsed_time(e1) * 1e-3), sc[:, 12].mean(), sc[:, 10].mean(), sc[:, 10].min(), sc[:, 10].max(), info['instances_per_cta_solve']), flush=True)
    # solve loop at a fixed iteration count: without checks, and with the KKT check every 64 iterations
    for ce in (1000000, 64):
        x.zero_(); y.zero_()
        e0.record(); bt.solve(x, y, bb, cb, etas, scal, 1.0, 3000, ce, 0.0); e1.record(); torch.cuda.synchronize()
        print('   solve loop, 3000 iterations, check_every %d: %.3e LP-it/s' % (ce, B * 3000 / (e0.elapsed_time(e1) * 1e-3)), flush=True)
    bt.close()
    for k in env: os.environ.pop(k, None)

if __name__ == '__main__':
    name = sys.argv[1] if len(sys.argv) > 1 else '25fv47'
    B = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
    for spec in (sys.argv[3:] or ['MLLP_BATCH_R=1,MLLP_BATCH_RES=0', 'MLLP_BATCH_R=1', 'MLLP_BATCH_R=2', 'MLLP_BATCH_R=3']):
        run(name, B, 200, **dict(kv.split('=') for kv in spec.split(',') if kv))
