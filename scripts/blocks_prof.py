"""Dev tool: two 1000-iteration parity launches of ken-18 on the block-angular kernel (the target of the ncu capture in
profiles/r01_blocks_ncu_full.md)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import mllp_b200 as M
A, b, c = M.load_csr("ken-18"); m, n = A.shape
lp = M.DeviceLP(A, A.data, m, n)
assert lp.blocks_info()["used"], lp.blocks_info()
eta = 0.9 / lp.sigma_max()
bt, ct = torch.tensor(b, device="cuda"), torch.tensor(c, device="cuda")
for _ in range(2):
    M.pdhg_linear_program(A, A.data, bt, ct, num_iters=1000, tau=eta, sigma=eta, handle=lp)
torch.cuda.synchronize()
