"""Dev tool: solve mode on the larger data instances; objective vs the HiGHS known answers."""
import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import mllp_b200 as M
HIGHS = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests', 'golden', 'highs_objectives.json')))
names = sys.argv[1].split(',') if len(sys.argv) > 1 else ['25fv47', 'd2q06c', 'pilot87', 'dfl001', 'osa-60', 'pds-20']
maxit = int(sys.argv[2]) if len(sys.argv) > 2 else 1000000
for nm in names:
    A, b, c = M.load_csr(nm)
    t = time.time()
    obj, x, y, info = M.solve_linear_program(A, A.data, b, c, tol=1e-6, max_iters=maxit, check_every=128)
    dt = time.time() - t
    ref = HIGHS.get(nm)
    print('%-8s obj %.9g ref %s relerr %s iters %d restarts %d conv %s kkt %.2e  %.2fs (%.2f us/iter)' % (
        nm, obj, ref, ('%.2e' % (abs(obj - ref) / (1 + abs(ref)))) if ref is not None else '-', info['iters'], info['restarts'],
        info['converged'], info['rel_kkt'], dt, dt / max(info['iters'], 1) * 1e6), flush=True)
