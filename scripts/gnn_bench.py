"""Dev tool: time of one GNNModel forward (5 TransformerConv passes, fc folded into the last) per instance, CUDA events.

Bytes counted per forward (algorithmic, every array touched once per conv): per conv 12 B/nnz (fp64 value + int32
index) + 4 B/row indptr + 4*din B per source and destination node (features in) + 64 B per destination node (features
out; 4 B for the last conv, whose output is the logit); the gathered source rows (4*din B per edge) are served by
L2 / L1 and reported separately."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import mllp_b200.linear_program_data as D
import mllp_b200.gnn as GN


def bytes_per_forward(m, n, nnz):
    tot = 0
    for k, (nd, ns) in enumerate([(n, m), (m, n), (n, m), (m, n), (n, m)]):
        din = 1 if k < 2 else 16
        tot += 12 * nnz + 4 * (nd + 1) + 4 * din * (ns + nd) + (64 if k < 4 else 4) * nd
    return tot


def gather_bytes(nnz):
    return nnz * (2 * 4 + 3 * 64)


def main(names):
    dev = torch.device("cuda", 0)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    for name in names:
        A, b, c = D.load_csr(name)
        m, n = A.shape
        g = GN.BipartiteGraph(np.split(A.indices, A.indptr)[1:-1], A.data, b, c)
        model = GN.GNNModel(seed=1)
        for _ in range(3):
            model(g)
        torch.cuda.synchronize()
        ts = []
        for _ in range(10):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); model(g); e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        ms = float(np.median(ts))
        B = bytes_per_forward(m, n, A.nnz)
        ts = []
        for _ in range(10):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); model(g); e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        warm = float(np.median(ts))
        ts = []
        for _ in range(10):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); model.forward(g, use_plan=False); e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        plain = float(np.median(ts))
        print("%-8s m %6d n %6d nnz %7d groups %d/%d: %.1f us / forward (L2 flushed; %.1f us warm; %.1f us as separate launches), %.0f GB/s algorithmic, gather traffic %.0f GB/s"
              % (name, m, n, A.nnz, g.to_var.group, g.to_con.group, ms * 1e3, warm * 1e3, plain * 1e3, B / ms / 1e6, gather_bytes(A.nnz) / ms / 1e6), flush=True)


if __name__ == "__main__":
    main(sys.argv[1:] or ["afiro", "25fv47", "pilot87", "ken-18", "osa-60", "pds-20"])
