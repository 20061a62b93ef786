"""Dev tool (CPU only): gather-line statistics of the built format, with/without clustering."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import mllp_b200.linear_program_data as D
from mllp_b200 import _cabi
L = _cabi.lib()
for name in sys.argv[1:]:
    A, _, _ = D.load_csr(name); m, n = A.shape
    ip = A.indptr.astype(np.int32); ii = A.indices.astype(np.int32); v = A.data
    for G, ps, ms in ((148, 4, 8),):
        for cl in (0, 1):
            out = np.zeros(6)
            rc = L.mllp_format_gather_lines(m, n, A.nnz, ip.ctypes.data, ii.ctypes.data, v.ctypes.data, G, ps, ms, cl, out.ctypes.data)
            print('%-8s G=%d cluster=%d rc=%d lines A %.0f (max/CTA %.0f, mean %.0f, per instr %.1f)  A\' %.0f (max/CTA %.0f, mean %.0f, per instr %.1f)' % (
                name, G, cl, rc, out[0], out[2], out[0] / G, out[0] / out[4], out[1], out[3], out[1] / G, out[1] / out[5]))
