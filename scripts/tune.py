"""Dev tool: sweep launch/format knobs (env vars read at mllp_lp_create) and time parity-mode iterations."""
import os, sys, itertools
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import mllp_b200 as M
from mllp_b200 import _cabi

def time_lp(name, K=500, flags=0, **env):
    for k, v in env.items():
        os.environ[k] = str(v)
    A, b, c = M.load_csr(name); m, n = A.shape
    lp = M.DeviceLP(A, A.data, m, n, flags=flags)
    eta = 0.9 / lp.sigma_max()
    bt, ct = torch.tensor(b, device='cuda'), torch.tensor(c, device='cuda')
    x = torch.zeros(n, dtype=torch.float64, device='cuda'); y = torch.zeros(m, dtype=torch.float64, device='cuda')
    L = _cabi.lib(); s = torch.cuda.current_stream().cuda_stream
    def run():
        _cabi.check(L.mllp_pdhg_run(lp.handle, x.data_ptr(), y.data_ptr(), bt.data_ptr(), ct.data_ptr(), eta, eta, K, None, s), 'run')
    run(); torch.cuda.synchronize()
    best = 1e9
    for _ in range(3):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); run(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    info = lp.info(); lp.close()
    for k in env: os.environ.pop(k, None)
    return best * 1e3 / K, info

if __name__ == '__main__':
    names = sys.argv[1].split(',') if len(sys.argv) > 1 else ['afiro', 'ken-18', 'osa-60']
    grid = [dict(MLLP_THREADS=t, MLLP_CTAS_PER_SM=c, MLLP_PREF_STEPS=p, MLLP_MAX_STEPS=mx, MLLP_RES_STEPS=rs)
            for t, c in ((512, 2), (1024, 1))
            for p, mx in ((4, 4), (4, 8), (4, 16), (8, 16), (4, 32))
            for rs in (0, 100000)]
    for name in names:
        for env in grid:
            try:
                us, info = time_lp(name, **env)
                print('%-8s %s -> %.2f us/iter (G=%d tilesA=%d tilesAT=%d)' % (name, ' '.join('%s=%s' % (k[5:], v) for k, v in env.items()), us, info['grid_ctas'], info['tiles_A'], info['tiles_AT']), flush=True)
            except Exception as e:
                print(name, env, 'ERR', e, flush=True)
