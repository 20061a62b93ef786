"""Dev tool: per-CTA timeline of the block kernel (MLLP_BLOCKS_DEBUG=2): local work, part -> finisher hop, dual -> CTA hop."""
import os, sys, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["MLLP_BLOCKS"] = "1"; os.environ["MLLP_BLOCKS_DEBUG"] = "2"
import numpy as np, torch
import mllp_b200 as M
from mllp_b200 import _cabi
name = sys.argv[1] if len(sys.argv) > 1 else "ken-18"
A, b, c = M.load_csr(name); m, n = A.shape
lp = M.DeviceLP(A, A.data, m, n)
eta = 0.9 / lp.sigma_max()
K = 64
M.pdhg_linear_program(A, A.data, b, c, num_iters=K, tau=eta, sigma=eta, handle=lp)
L = _cabi.lib()
it, G = ctypes.c_int32(0), ctypes.c_int32(0)
L.mllp_debug_blocks_trace.argtypes = [ctypes.c_void_p, ctypes.c_int64, ctypes.POINTER(ctypes.c_int32), ctypes.POINTER(ctypes.c_int32)]
buf = np.zeros(K * 160 * 4, dtype=np.uint64)
L.mllp_debug_blocks_trace(buf.ctypes.data, buf.size, ctypes.byref(it), ctypes.byref(G))
t = buf[: it.value * G.value * 4].reshape(it.value, G.value, 4).astype(np.int64)[8:]
f = lambda x: "mean %.0f min %.0f max %.0f" % (x.mean(), x.min(axis=1).mean(), x.max(axis=1).mean())
print(name, "iteration %.0f ns" % ((t[1:, :, 0] - t[:-1, :, 0]).mean()))
print("  start -> own parts published:", f(t[:, :, 1] - t[:, :, 0]))
print("  last CTA's parts published -> own row finished and sent (finishers): %.0f" % (t[:, :, 2] - t[:, :, 1].max(axis=1)[:, None]).mean())
print("  last finisher sent -> duals received by the last warp (mean over CTAs): %.0f" % (t[:, :, 3] - t[:, :, 2].max(axis=1)[:, None]).mean())
print("  duals received -> next iteration starts (CTA barrier): %.0f" % (t[1:, :, 0] - t[:-1, :, 3]).mean())
print("  spread of iteration starts over CTAs: %.0f" % (t[:, :, 0].max(axis=1) - t[:, :, 0].min(axis=1)).mean())
