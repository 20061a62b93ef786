"""numpy/scipy restatement of the diagonal preconditioner the library computes on the device (mllp_b200/csrc/scaling.cu,
precondition_device: MLLP_F_PRECONDITION) -- TEST INFRASTRUCTURE: the checker of tests/test_gpu_precondition.py.

Ruiz equilibration (`ruiz_iters` rounds: rows and columns divided by the square root of their largest scaled magnitude,
both taken from the same scaled matrix) followed by one Pock-Chambolle pass with alpha = 1 (square root of the scaled
absolute row / column sums): the PDLP recipe.  Not from the reference (it has no solver)."""
import numpy as np
import scipy.sparse as sp


def ruiz_pock_chambolle(A, ruiz_iters=10):
    A = sp.csr_matrix(A, dtype=np.float64)
    m, n = A.shape
    dr, dc = np.ones(m), np.ones(n)
    absA = abs(A)
    for _ in range(ruiz_iters):
        B = sp.diags(dr) @ absA @ sp.diags(dc)
        rn = np.sqrt(np.asarray(B.max(axis=1).todense()).ravel())
        cn = np.sqrt(np.asarray(B.max(axis=0).todense()).ravel())
        rn[rn == 0] = 1.0
        cn[cn == 0] = 1.0
        dr /= rn
        dc /= cn
    B = sp.diags(dr) @ absA @ sp.diags(dc)
    rn = np.sqrt(np.asarray(B.sum(axis=1)).ravel())
    cn = np.sqrt(np.asarray(B.sum(axis=0)).ravel())
    rn[rn == 0] = 1.0
    cn[cn == 0] = 1.0
    return dr / rn, dc / cn
