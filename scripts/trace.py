"""Dev tool: per-CTA barrier timeline of the persistent parity kernel."""
import os, sys, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import mllp_b200 as M
from mllp_b200 import _cabi

def trace(name, iters=40, **env):
    for k, v in env.items(): os.environ[k] = str(v)
    A, b, c = M.load_csr(name); m, n = A.shape
    lp = M.DeviceLP(A, A.data, m, n)
    eta = 0.9 / lp.sigma_max()
    M.pdhg_linear_program(A, A.data, b, c, num_iters=50, tau=eta, sigma=eta, handle=lp)  # loads b, c; warm
    G = lp.info()['grid_ctas']
    out = np.zeros(iters * G * 4, dtype=np.uint64)
    L = _cabi.lib(); L.mllp_debug_trace.argtypes = [ctypes.c_void_p, ctypes.c_double, ctypes.c_double, ctypes.c_int32, ctypes.c_void_p]
    _cabi.check(L.mllp_debug_trace(lp.handle, eta, eta, iters, out.ctypes.data), 'trace')
    t = out.reshape(iters, G, 4).astype(np.int64)
    t = t[10:]  # skip the first iterations (smem fill, cold L2)
    # phase durations: release of previous barrier -> CTA arrival
    rel_prev = np.concatenate([t[:-1, :, 3][None].transpose(1, 2, 0)[..., 0][None] if False else t[:-1, :, 3]], axis=0)
    at_work = t[1:, :, 0] - rel_prev            # A' phase: from previous release to this CTA's arrival
    at_wait = t[1:, :, 1] - t[1:, :, 0]         # wait in barrier 1
    a_work = t[1:, :, 2] - t[1:, :, 1]
    a_wait = t[1:, :, 3] - t[1:, :, 2]
    it_time = (t[1:, :, 3] - t[:-1, :, 3]).mean()
    f = lambda x: 'mean %.0f min %.0f p50 %.0f max %.0f' % (x.mean(), x.min(axis=1).mean(), np.median(x, axis=1).mean(), x.max(axis=1).mean())
    print('%s %s: iteration %.0f ns  %s' % (name, env, it_time, {k: lp.info()[k] for k in ('dyn_smem_bytes', 'res_steps_A', 'res_steps_AT', 'tiles_A', 'tiles_AT')}))
    print('   tune: %s' % (lp.tune_info(),))
    print('   A\' work  (ns): ' + f(at_work)); print('   barrier1 wait: ' + f(at_wait))
    print('   A  work  (ns): ' + f(a_work)); print('   barrier2 wait: ' + f(a_wait))
    # barrier latency proper: last arrival -> mean release
    last1 = t[1:, :, 0].max(axis=1); relm1 = t[1:, :, 1]
    last2 = t[1:, :, 2].max(axis=1); relm2 = t[1:, :, 3]
    print('   barrier1 release after last arrival: mean %.0f max %.0f ; barrier2: mean %.0f max %.0f' % (
        (relm1 - last1[:, None]).mean(), (relm1 - last1[:, None]).max(axis=1).mean(), (relm2 - last2[:, None]).mean(), (relm2 - last2[:, None]).max(axis=1).mean()))
    aw = a_work.mean(axis=0); order = np.argsort(-aw)[:10]
    print('   slowest CTAs in A phase: ' + ' '.join('%d:%.0f' % (g, aw[g]) for g in order) + '  | fastest: ' + ' '.join('%d:%.0f' % (g, aw[g]) for g in np.argsort(aw)[:5]))
    aw2 = at_work.mean(axis=0); order = np.argsort(-aw2)[:6]
    print('   slowest CTAs in A\' phase: ' + ' '.join('%d:%.0f' % (g, aw2[g]) for g in order))
    lp.close()
    for k in env: os.environ.pop(k, None)

if __name__ == '__main__':
    for name in (sys.argv[1].split(',') if len(sys.argv) > 1 else ['afiro', 'ken-18', 'osa-60']):
        if len(sys.argv) > 2:
            for spec in sys.argv[2:]:
                trace(name, **dict(kv.split('=') for kv in spec.split(',') if kv))
        else:
            trace(name)
            trace(name, MLLP_THREADS=512, MLLP_CTAS_PER_SM=2)
