"""ncu target: a few training steps of the device GNNModel on one instance (scripts/gnn_train_bench.py without the timing)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import mllp_b200.linear_program_data as D
import mllp_b200.gnn as GN
from mllp_b200.gnn_train import TrainableGNNModel

name = sys.argv[1] if len(sys.argv) > 1 else "ken-18"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
A, b, c = D.load_csr(name)
g = GN.BipartiteGraph(np.split(A.indices, A.indptr)[1:-1], A.data, b, c)
model = TrainableGNNModel(seed=1)
dout = torch.randn(A.shape[1], device="cuda") / A.shape[1]
for _ in range(steps):
    model(g).backward(dout)
    model.flat.grad = None
torch.cuda.synchronize()
print("done", name)
