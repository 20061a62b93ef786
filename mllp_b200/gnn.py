"""Host side of the bipartite message-passing forward: the reference's
``GNNModel`` (linear_program_methods.py:199-251) over device-resident CSR arrays of A and A'.

``BipartiteGraph`` takes the loader's arrays in the argument order of
``build_graph_from_weights_sets(constrs, constr_weights, rhs, coefs, device)`` (:89) and keeps what the kernels
need on the device (no Python loop over nonzeros, no [2, nnz] edge list); ``GNNModel`` keeps the reference module's
parameter names, so a ``state_dict()`` of the reference model loads unchanged, and ``forward(g)`` returns the logit
per variable (:250).  fp32 like the reference.  All arithmetic is in mllp_b200/csrc/gnn_kernels.cu behind the C ABI;
there is no PyTorch fallback.
"""
import numpy as np

from . import _cabi
from .linear_program_methods import _device_index, _torch_stream, csr_from_constrs

C = 16
CONVS = ("gconv1_w2s", "gconv1_s2w", "gconv2_w2s", "gconv2_s2w", "gconv3_w2s")
CHUNK = 2048   # edges per warp before a row is cut into items


def default_state(seed=0):
    """Random float32 parameters in the shapes and default init scale (U(-1/sqrt(fan_in), 1/sqrt(fan_in))) of the
    reference module, under its state_dict names."""
    rng = np.random.default_rng(seed)
    st = {}
    for cv in CONVS + ("gconv3_s2w",):
        din = 1 if cv.startswith("gconv1") else C
        for part in ("lin_key", "lin_query", "lin_value", "lin_skip", "lin_edge"):
            fin = 1 if part == "lin_edge" else din
            bound = 1.0 / np.sqrt(fin)
            st["%s.%s.weight" % (cv, part)] = rng.uniform(-bound, bound, (C, fin)).astype(np.float32)
            if part != "lin_edge":
                st["%s.%s.bias" % (cv, part)] = rng.uniform(-bound, bound, C).astype(np.float32)
    st["fc.weight"] = rng.uniform(-0.25, 0.25, (1, C)).astype(np.float32)
    st["fc.bias"] = rng.uniform(-0.25, 0.25, 1).astype(np.float32)
    return st


class _Side:
    """CSR of one direction (rows = destination nodes) + the long-row tables, on the device."""

    def __init__(self, M, dev, torch):
        M = M.tocsr()
        M.sort_indices()
        self.nd, self.ns = M.shape
        ip = np.ascontiguousarray(M.indptr, dtype=np.int32)
        self.indptr = torch.as_tensor(ip, device=dev)
        self.indices = torch.as_tensor(np.ascontiguousarray(M.indices, dtype=np.int32), device=dev)
        self.values = torch.as_tensor(np.ascontiguousarray(M.data, dtype=np.float64), device=dev)
        lens = np.diff(ip)
        long_rows = np.nonzero(lens > CHUNK)[0].astype(np.int32)
        items, first = [], [0]
        for r in long_rows:
            k = -(-int(lens[r]) // CHUNK)
            step = -(-int(lens[r]) // k)
            step = (step + 31) & ~31
            for e0 in range(int(ip[r]), int(ip[r + 1]), step):
                items.append((int(r), e0, min(e0 + step, int(ip[r + 1]))))
            first.append(len(items))
        self.nlong, self.nitems = len(long_rows), len(items)
        z = lambda a: torch.as_tensor(np.ascontiguousarray(a, dtype=np.int32).reshape(-1) if len(a) else np.zeros(1, np.int32), device=dev)
        self.long_rows, self.first, self.items = z(long_rows), z(first), z(items)
        self.scratch = torch.empty(max(1, 96 * self.nitems), dtype=torch.float32, device=dev)
        self.nnz = int(ip[-1])


class BipartiteGraph:
    """The LP as the reference's bipartite graph: x1 = coefs (variables), x2 = rhs (constraints), one edge per
    nonzero with attribute a_ij (linear_program_methods.py:89-103), held as CSR of A and of A'."""

    def __init__(self, constrs, constr_weights, rhs, coefs, device=0):
        import scipy.sparse as sp
        import torch
        self.device = _device_index(device)
        dev = torch.device("cuda", self.device)
        n, m = len(coefs), len(rhs)
        ip, ii, vv = csr_from_constrs(constrs, constr_weights, n)
        if ip.shape[0] != m + 1:
            raise ValueError("constrs has %d rows, rhs has %d" % (ip.shape[0] - 1, m))
        A = sp.csr_matrix((vv, ii, ip), shape=(m, n))
        self.m, self.n, self.nnz = m, n, int(A.nnz)
        self.to_con = _Side(A, dev, torch)            # variable -> constraint ("s2w"): rows of A
        self.to_var = _Side(A.T.tocsr(), dev, torch)  # constraint -> variable ("w2s"): rows of A'
        self.x1 = torch.as_tensor(np.asarray(coefs, dtype=np.float32).reshape(n, 1), device=dev)
        self.x2 = torch.as_tensor(np.asarray(rhs, dtype=np.float32).reshape(m, 1), device=dev)


class GNNModel:
    """Parameters under the reference module's names; ``forward(g)`` = linear_program_methods.py:238-251."""

    def __init__(self, state_dict=None, device=0, seed=0):
        import torch
        self.device = _device_index(device)
        self.load_state_dict(default_state(seed) if state_dict is None else state_dict)

    def load_state_dict(self, state_dict):
        import torch
        dev = torch.device("cuda", self.device)
        f = lambda a: np.asarray(a.detach().cpu().numpy() if hasattr(a, "detach") else a, dtype=np.float32)
        self.state = {k: f(v) for k, v in state_dict.items()}
        t = lambda a: torch.as_tensor(np.ascontiguousarray(a, dtype=np.float32).reshape(-1), device=dev)
        self.conv_params, self.proj_params, self.din = {}, {}, {}
        for cv in CONVS:
            g = lambda part, kind: self.state["%s.%s.%s" % (cv, part, kind)]
            din = g("lin_query", "weight").shape[1]
            self.din[cv] = din
            for part in ("lin_key", "lin_query", "lin_value", "lin_skip"):
                if g(part, "weight").shape != (C, din) or g(part, "bias").shape != (C,):
                    raise ValueError("%s.%s: expected weight (16, %d) and bias (16,)" % (cv, part, din))
            if g("lin_edge", "weight").shape != (C, 1):
                raise ValueError("%s.lin_edge.weight: expected (16, 1) (edge_dim = 1, no bias)" % cv)
            self.conv_params[cv] = t(np.concatenate([g("lin_query", "weight").T.reshape(-1), g("lin_query", "bias"),
                                                     g("lin_skip", "weight").T.reshape(-1), g("lin_skip", "bias"),
                                                     g("lin_edge", "weight").reshape(-1)]))
            self.proj_params[cv] = t(np.concatenate([g("lin_key", "weight").T.reshape(-1), g("lin_key", "bias"),
                                                     g("lin_value", "weight").T.reshape(-1), g("lin_value", "bias")]))
        if self.state["fc.weight"].shape != (1, C):
            raise ValueError("fc.weight: expected (1, 16)")
        self.fc = t(np.concatenate([self.state["fc.weight"].reshape(-1), self.state["fc.bias"].reshape(-1)]))

    def _conv(self, cv, side, h_src, h_dst, relu=True):
        import torch
        L = _cabi.lib()
        dev = h_dst.device
        st = _torch_stream(dev)
        din = self.din[cv]
        if h_src.shape != (side.ns, din) or h_dst.shape != (side.nd, din):
            raise ValueError("%s: feature shapes do not match the graph" % cv)
        kv = torch.empty(side.ns * 32, dtype=torch.float32, device=dev)
        _cabi.check(L.mllp_gnn_project(side.ns, h_src.data_ptr(), din, self.proj_params[cv].data_ptr(), kv.data_ptr(), st),
                    "mllp_gnn_project")
        out = torch.empty(side.nd, C, dtype=torch.float32, device=dev)
        _cabi.check(L.mllp_gnn_conv(side.nd, side.indptr.data_ptr(), side.indices.data_ptr(), side.values.data_ptr(),
                                    h_dst.data_ptr(), din, kv.data_ptr(), self.conv_params[cv].data_ptr(), out.data_ptr(),
                                    int(relu), CHUNK, side.nlong, side.long_rows.data_ptr(), side.first.data_ptr(),
                                    side.nitems, side.items.data_ptr(), side.scratch.data_ptr(), st), "mllp_gnn_conv")
        return out

    def forward(self, g):
        """logit per variable, float32 tensor (n,) on the graph's device; no host synchronisation."""
        import torch
        if not isinstance(g, BipartiteGraph):
            raise TypeError("GNNModel.forward expects a BipartiteGraph")   # the reference asserts its type too (:239)
        x1, x2 = g.x1, g.x2
        n1 = self._conv("gconv1_w2s", g.to_var, x2, x1)
        n2 = self._conv("gconv1_s2w", g.to_con, x1, x2)
        x1, x2 = n1, n2
        n1 = self._conv("gconv2_w2s", g.to_var, x2, x1)
        n2 = self._conv("gconv2_s2w", g.to_con, x1, x2)
        x1, x2 = n1, n2
        n1 = self._conv("gconv3_w2s", g.to_var, x2, x1)
        out = torch.empty(g.n, dtype=torch.float32, device=n1.device)
        _cabi.check(_cabi.lib().mllp_gnn_fc(g.n, n1.data_ptr(), self.fc.data_ptr(), out.data_ptr(), _torch_stream(n1.device)),
                    "mllp_gnn_fc")
        return out

    __call__ = forward
