"""Dev tool: time of one training step of the device GNNModel (forward + BCEWithLogitsLoss + mllp_gnn_backward + Adam step)
per instance, CUDA events, and the split forward / backward.  Writes a markdown table (profiles/r02_gnn_backward.md)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import mllp_b200.linear_program_data as D
import mllp_b200.gnn as GN
from mllp_b200 import _cabi
from mllp_b200.gnn_train import TrainableGNNModel


def timed(fn, reps=10):
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    return float(np.median(ts))


def main(names, out=None):
    dev = torch.device("cuda", 0)
    rows = []
    for name in names:
        A, b, c = D.load_csr(name)
        m, n = A.shape
        g = GN.BipartiteGraph(np.split(A.indices, A.indptr)[1:-1], A.data, b, c)
        model = TrainableGNNModel(seed=1)
        opt = torch.optim.Adam(model.parameters(), lr=1e-3)
        crit = torch.nn.BCEWithLogitsLoss()
        target = torch.as_tensor((np.random.default_rng(0).random(n) < 0.4).astype(np.float32), device=dev)
        dout = torch.randn(n, device=dev) / n

        def step():
            loss = crit(model(g), target)
            loss.backward()
            opt.step()
            opt.zero_grad()

        def fwd():
            with torch.no_grad():
                model(g)

        state = {}

        def fwd_keep():
            state["out"] = model(g)

        def bwd():
            state["out"].backward(dout)

        for _ in range(3):
            step()
        torch.cuda.synchronize()
        t_step = timed(step)
        t_fwd = timed(fwd)
        # backward alone: a forward (untimed) before each timed backward
        ts = []
        L0 = _cabi.lib().mllp_launch_count()
        fwd_keep(); L1 = _cabi.lib().mllp_launch_count()
        bwd(); L2 = _cabi.lib().mllp_launch_count()
        model.flat.grad = None
        for _ in range(10):
            fwd_keep()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); bwd(); e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1) * 1e3)
            model.flat.grad = None
        t_bwd = float(np.median(ts))
        rows.append((name, m, n, A.nnz, g.to_var.group, g.to_con.group, t_fwd, t_bwd, t_step, L1 - L0, L2 - L1))
        print("%-8s m %6d n %6d nnz %7d groups %d/%d: forward %.1f us, backward %.1f us, training step %.1f us (launches: forward %d, backward %d)"
              % rows[-1], flush=True)
    if out:
        with open(out, "w") as f:
            f.write("# r02 -- device GNNModel training step (forward + BCEWithLogitsLoss + mllp_gnn_backward + Adam), one B200\n\n"
                    "`python scripts/gnn_train_bench.py`: CUDA events, median of 10, warm L2.  Forward and backward are each ONE CUDA-graph\n"
                    "launch (`mllp_gnn_train_plan_create` = pack + 5-11 conv kernels, `mllp_gnn_backward_plan_create` = pack + 32-44 kernels:\n"
                    "row walk / dense maps / second sweep per conv, source passes over the transposed structure, parameter-gradient partial\n"
                    "sums, unpack) plus a copy in and a clone out.  The loss and Adam are torch ops on the (n,) logits / the 4721 parameters --\n"
                    "on the small graphs they are most of the step.  History of the backward on ken-18 / osa-60 / pds-20 (us): first version\n"
                    "1096 / 3647 / 738 (one CTA per cut row, one 255-register kernel per destination pass); cut rows through the forward's\n"
                    "items 1153 / 1329 / 766; destination pass split into row walk (80 registers) + dense maps (16 lanes per node) + sweep,\n"
                    "float4-prefetched parameter-gradient tiles, parallel partial sums 836 / 1010 / 593; as CUDA graphs 739 / 921 / 504; four\n"
                    "consecutive entries per thread in the parameter-gradient kernel (one 128-bit shared-memory load per node instead of eight\n"
                    "scalar ones: it was bound by shared-memory loads) and 128-bit stores of the per-node vectors 666 / 828 / 458; parameter-gradient\n"
                    "sums on a parallel branch of the captured graph (joined before the unpack: afiro 139 -> 127, 25fv47 181 -> 158): the table.\n\n"
                    "| instance | m | n | nnz | lanes per row (A' / A) | forward us | backward us | whole training step us | launches fwd / bwd |\n"
                    "|---|---:|---:|---:|---|---:|---:|---:|---|\n")
            for r in rows:
                f.write("| %s | %d | %d | %d | %d / %d | %.1f | %.1f | %.1f | %d / %d |\n" % r)


if __name__ == "__main__":
    args = [a for a in sys.argv[1:] if not a.startswith("--out=")]
    out = [a[6:] for a in sys.argv[1:] if a.startswith("--out=")]
    main(args or ["afiro", "25fv47", "pilot87", "ken-18", "osa-60", "pds-20"], out[0] if out else None)
