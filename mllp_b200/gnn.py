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
CHUNK = 1024   # rows up to this many edges decide the lanes per row
ITEM = 512     # edges per item (one warp) of the rows that are cut


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
    """CSR of one direction (rows = destination nodes) + the long-row tables, on the device, and the
    ``mllp_gnn_side`` struct that describes them to the library."""

    def __init__(self, M, dev, torch, group=None):
        M = M.tocsr()
        M.sort_indices()
        self.nd, self.ns = M.shape
        ip = np.ascontiguousarray(M.indptr, dtype=np.int32)
        self.indptr = torch.as_tensor(ip, device=dev)
        self.indices = torch.as_tensor(np.ascontiguousarray(M.indices, dtype=np.int32), device=dev)
        self.values = torch.as_tensor(np.ascontiguousarray(M.data, dtype=np.float64), device=dev)
        lens = np.diff(ip)
        # lanes per destination row, like the lanes-per-row choice of the LP format: a lane walks its edges two at a
        # time, so one lane serves rows of a handful of edges (A' of ken-18 / pds-20 / osa-60: 2 - 6 per row) with no
        # idle lanes and no merge; S = 2^k >= p90(row length) / 8.  Rows above row_max edges (ken-18's A has 151 rows of
        # ~300 edges among 105 127 of ~3: on 4 lanes they were a 100 us tail) are cut into items of at most ITEM edges,
        # one warp each, whose partial softmax states are merged in a fixed order; the two extra launches only pay above
        # ~128 edges when the group is already wide.
        reg = lens[(lens > 0) & (lens <= CHUNK)]
        p90 = float(np.percentile(reg, 90)) if reg.size else 1.0
        self.group = 1
        while self.group < 32 and 8 * self.group < p90:
            self.group *= 2
        # a small graph does not fill the GPU's lanes: wider groups then only shorten the per-row edge chain
        while self.group < 32 and self.group < p90 and 2 * self.group * self.nd <= 148 * 768:
            self.group *= 2
        if group is not None:   # tests: force the lanes per row
            if group not in (1, 2, 4, 8, 16, 32):
                raise ValueError("group must be 1, 2, 4, 8, 16 or 32")
            self.group = int(group)
        self.row_max = 32 * self.group if self.group <= 2 else max(16 * self.group, 128)
        long_rows = np.nonzero(lens > self.row_max)[0].astype(np.int32)
        items, first = [], [0]
        for r in long_rows:
            k = -(-int(lens[r]) // ITEM)
            step = -(-int(lens[r]) // k)
            step = (step + 63) & ~63
            for e0 in range(int(ip[r]), int(ip[r + 1]), step):
                items.append((int(r), e0, min(e0 + step, int(ip[r + 1]))))
            first.append(len(items))
        self.nlong, self.nitems = len(long_rows), len(items)
        z = lambda a: torch.as_tensor(np.ascontiguousarray(a, dtype=np.int32).reshape(-1) if len(a) else np.zeros(1, np.int32), device=dev)
        self.long_rows, self.first, self.items = z(long_rows), z(first), z(items)
        self.scratch = torch.empty(max(1, 20 * self.nitems), dtype=torch.float32, device=dev)
        self.nnz = int(ip[-1])
        self.c = _cabi.GnnSide(self.nd, self.ns, self.group, self.row_max, self.indptr.data_ptr(), self.indices.data_ptr(),
                               self.values.data_ptr(), self.nlong, self.nitems, self.long_rows.data_ptr(),
                               self.first.data_ptr(), self.items.data_ptr(), self.scratch.data_ptr())


class BipartiteGraph:
    """The LP as the reference's bipartite graph: x1 = coefs (variables), x2 = rhs (constraints), one edge per
    nonzero with attribute a_ij (linear_program_methods.py:89-103), held as CSR of A and of A'."""

    def __init__(self, constrs, constr_weights, rhs, coefs, device=0, groups=None):
        import ctypes
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
        gv, gc = groups if groups is not None else (None, None)
        self.to_con = _Side(A, dev, torch, gc)            # variable -> constraint ("s2w"): rows of A
        self.to_var = _Side(A.T.tocsr(), dev, torch, gv)  # constraint -> variable ("w2s"): rows of A'
        self.x1 = torch.as_tensor(np.asarray(coefs, dtype=np.float32).reshape(n, 1), device=dev)
        self.x2 = torch.as_tensor(np.asarray(rhs, dtype=np.float32).reshape(m, 1), device=dev)
        self.work = torch.empty(int(_cabi.lib().mllp_gnn_workspace_floats(n, m)), dtype=torch.float32, device=dev)
        self._plans = {}   # parameter buffer -> (mllp_gnn_plan_t, output buffer, parameters)

    def close(self):
        """release the captured forward plans (the device arrays go with the object)"""
        for entry in self._plans.values():   # entry[0] is the mllp_gnn_plan_t, the rest keeps its buffers alive
            _cabi.lib().mllp_gnn_plan_destroy(entry[0])
        self._plans = {}

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def pack_conv(state, cv):
    """(din, parameter block of one conv as float32) in the layout of include/mllp_b200.h: the layer is evaluated
    without materialising query / key / value rows, so the products Wq'Wk, Wk'bq, Wq'We and We.bq (all with the
    1/sqrt(16) of the attention score folded in) are formed here, in float64, and rounded once.  lin_key.bias adds the
    same amount to every score of a destination node and cancels in the softmax."""
    g = lambda part, kind: np.asarray(state["%s.%s.%s" % (cv, part, kind)], dtype=np.float32)
    din = g("lin_query", "weight").shape[1]
    for part in ("lin_key", "lin_query", "lin_value", "lin_skip"):
        if g(part, "weight").shape != (C, din) or g(part, "bias").shape != (C,):
            raise ValueError("%s.%s: expected weight (16, %d) and bias (16,)" % (cv, part, din))
    if g("lin_edge", "weight").shape != (C, 1):
        raise ValueError("%s.lin_edge.weight: expected (16, 1) (edge_dim = 1, no bias)" % cv)
    d = lambda part, kind: g(part, kind).astype(np.float64)
    Wq, bq, Wk = d("lin_query", "weight"), d("lin_query", "bias"), d("lin_key", "weight")
    We = d("lin_edge", "weight").reshape(-1)
    head = np.concatenate([(Wq.T @ Wk).reshape(-1) / 4.0, (Wk.T @ bq) / 4.0, (Wq.T @ We) / 4.0, [(We @ bq) / 4.0]])
    head = np.concatenate([head, np.zeros(-len(head) % 4)])
    tail = np.concatenate([d("lin_value", "weight").T.reshape(-1), d("lin_value", "bias"),
                           d("lin_skip", "weight").T.reshape(-1), d("lin_skip", "bias"), We])
    blk = np.concatenate([head, tail]).astype(np.float32)
    if len(blk) != int(_cabi.lib().mllp_gnn_conv_param_floats(din)):
        raise ValueError("%s: %d input channels are not supported (1 or 16)" % (cv, din))
    return din, blk


class GNNModel:
    """Parameters under the reference module's names; ``forward(g)`` = linear_program_methods.py:238-251."""

    def __init__(self, state_dict=None, device=0, seed=0):
        self.device = _device_index(device)
        self.load_state_dict(default_state(seed) if state_dict is None else state_dict)

    def load_state_dict(self, state_dict):
        import torch
        dev = torch.device("cuda", self.device)
        f = lambda a: np.asarray(a.detach().cpu().numpy() if hasattr(a, "detach") else a, dtype=np.float32)
        self.state = {k: f(v) for k, v in state_dict.items()}
        parts, self.din = [], {}
        for k, cv in enumerate(CONVS):
            din, blk = pack_conv(self.state, cv)
            if din != (1 if k < 2 else C):
                raise ValueError("%s: expected %d input channels" % (cv, 1 if k < 2 else C))
            self.din[cv] = din
            parts.append(blk)
        if self.state["fc.weight"].shape != (1, C):
            raise ValueError("fc.weight: expected (1, 16)")
        parts += [self.state["fc.weight"].reshape(-1), self.state["fc.bias"].reshape(-1)]
        self.params = torch.as_tensor(np.ascontiguousarray(np.concatenate(parts), dtype=np.float32), device=dev)

    def forward(self, g, use_plan=True):
        """logit per variable, float32 tensor (n,) on the graph's device; no host synchronisation.  With ``use_plan`` the
        launches are captured once per (graph, parameters) into a CUDA graph (mllp_gnn_plan_*: the two convs of a layer on
        parallel branches) and replayed with one launch; otherwise one mllp_gnn_forward call (5 - 11 launches)."""
        import ctypes
        import torch
        if not isinstance(g, BipartiteGraph):
            from .graph import BipartiteData
            if not isinstance(g, BipartiteData):   # the reference asserts the type too (:239)
                raise TypeError("GNNModel.forward expects the BipartiteData of build_graph_from_weights_sets or a BipartiteGraph")
            g = g.bipartite_graph()
        if g.device != self.device:
            raise ValueError("graph and model live on different devices")
        dev = g.x1.device
        L = _cabi.lib()
        if not use_plan:
            out = torch.empty(g.n, dtype=torch.float32, device=dev)
            with torch.cuda.device(dev):   # the library launches on the calling thread's current device
                _cabi.check(L.mllp_gnn_forward(ctypes.byref(g.to_var.c), ctypes.byref(g.to_con.c), g.x1.data_ptr(),
                                               g.x2.data_ptr(), self.params.data_ptr(), g.work.data_ptr(), out.data_ptr(),
                                               _torch_stream(dev)), "mllp_gnn_forward")
            return out
        key = self.params.data_ptr()
        entry = g._plans.get(key)
        if entry is None:
            buf = torch.empty(g.n, dtype=torch.float32, device=dev)
            plan = ctypes.c_void_p()
            with torch.cuda.device(dev):
                _cabi.check(L.mllp_gnn_plan_create(ctypes.byref(g.to_var.c), ctypes.byref(g.to_con.c), g.x1.data_ptr(),
                                                   g.x2.data_ptr(), self.params.data_ptr(), g.work.data_ptr(), buf.data_ptr(),
                                                   ctypes.byref(plan)), "mllp_gnn_plan_create")
            entry = g._plans[key] = (plan, buf, self.params)   # the plan holds these pointers: keep the tensors alive
        with torch.cuda.device(dev):
            _cabi.check(L.mllp_gnn_plan_run(entry[0], _torch_stream(dev)), "mllp_gnn_plan_run")
        return entry[1].clone()

    __call__ = forward
