"""Device-side LP -> bipartite graph builder.

Replaces the Python double loop of the reference's ``build_graph_from_weights_sets``
(linear_program_methods.py:89-103, executed per instance per epoch at
linear_program_experiment.py:124) by one CUDA kernel over the CSR arrays
(``mllp_graph_edges``): ``edge_index[0, k]`` = variable (column) of nonzero k,
``edge_index[1, k]`` = constraint (row) of nonzero k -- the same (var, constr) orientation and the
same nonzero order as the reference -- ``edge_attr[k, 0]`` = a_ij as float32, ``x1`` = coefs
(n, 1) and ``x2`` = rhs (m, 1) as float32.

The return value has the shape of the reference's ``BipartiteData`` (linear_program_methods.py:60-72, returned at
:103): attributes ``edge_index``, ``x1``, ``x2``, ``edge_attr``.  Where torch_geometric is installed it IS a
``torch_geometric.data.Data`` (with the reference's ``__inc__`` rule for batching); otherwise a plain object with the
same attributes -- torch_geometric is a third-party dependency of the reference, not of this package.
"""
import ctypes

import numpy as np

from . import _cabi
from .linear_program_methods import _device_index, csr_from_constrs

try:   # the reference derives BipartiteData from torch_geometric.data.Data (:60); keep that when it is there
    import torch_geometric as _pyg
    _Base = _pyg.data.Data
except Exception:   # not installed (this image): same attributes on a plain object
    _pyg = None
    _Base = object


class BipartiteData(_Base):
    """``BipartiteData(edge_index, x_src, x_dst, edge_attr)`` -> ``.edge_index``, ``.x1``, ``.x2``, ``.edge_attr``
    (reference linear_program_methods.py:60-66)."""

    def __init__(self, edge_index=None, x_src=None, x_dst=None, edge_attr=None):
        super().__init__()
        self.edge_index = edge_index
        self.x1 = x_src
        self.x2 = x_dst
        self.edge_attr = edge_attr

    def __inc__(self, key, value, *args, **kwargs):   # reference :68-72
        import torch
        if key == "edge_index":
            return torch.tensor([[self.x1.size(0)], [self.x2.size(0)]])
        return super().__inc__(key, value, *args, **kwargs)

    # round-1 callers indexed the returned dict
    _ALIASES = {"x_src": "x1", "x_tgt": "x2"}

    def __getitem__(self, key):
        if isinstance(key, str) and key in ("edge_index", "edge_attr", "x1", "x2", "x_src", "x_tgt"):
            return getattr(self, self._ALIASES.get(key, key))
        if _pyg is not None:
            return super().__getitem__(key)
        raise KeyError(key)

    def to(self, device):
        if _pyg is not None:
            return super().to(device)
        for k in ("edge_index", "x1", "x2", "edge_attr"):
            setattr(self, k, getattr(self, k).to(device))
        return self

    def bipartite_graph(self):
        """The device CSR form of the same graph that the message-passing kernels walk (mllp_b200.gnn.BipartiteGraph),
        built once from the arrays this object was made from."""
        g = self.__dict__.get("_mllp_graph")
        if g is None:
            src = self.__dict__.get("_mllp_source")
            if src is None:
                raise ValueError("this BipartiteData was not made by mllp_b200's build_graph_from_weights_sets")
            g = _cached_graph(src)
            self.__dict__["_mllp_graph"] = g
        return g


# The reference's training loop rebuilds the graph of every instance in every epoch (linear_program_experiment.py:124) from
# the SAME dataset arrays: the device CSR form (and the CUDA-graph plans captured on it) is kept per (arrays, device), so an
# unchanged loop pays for it once.  Bounded; the arrays are held so that their identities stay valid.
_GRAPH_CACHE = {}
_GRAPH_CACHE_MAX = 256


def _cached_graph(src):
    from .gnn import BipartiteGraph
    constrs, constr_weights, rhs, coefs, device = src
    key = (id(constr_weights), id(rhs), id(coefs), device)
    # identity alone would miss an in-place edit of the arrays: a content fingerprint (sums, O(nnz)) rides along
    fp = tuple(float(np.asarray(a, dtype=np.float64).sum()) for a in (constr_weights, rhs, coefs)) + (len(rhs), len(coefs))
    hit = _GRAPH_CACHE.get(key)
    if hit is not None and hit[1] is constr_weights and hit[2] is rhs and hit[3] is coefs and hit[4] == fp:
        return hit[0]
    g = BipartiteGraph(*src)
    if hit is not None:
        _GRAPH_CACHE.pop(key)[0].close()
    elif len(_GRAPH_CACHE) >= _GRAPH_CACHE_MAX:
        old_key = next(iter(_GRAPH_CACHE))
        _GRAPH_CACHE.pop(old_key)[0].close()
    _GRAPH_CACHE[key] = (g, constr_weights, rhs, coefs, fp)
    return g


def clear_graph_cache():
    """release the cached device graphs (and their captured plans)"""
    while _GRAPH_CACHE:
        _GRAPH_CACHE.popitem()[1][0].close()


def build_graph_from_weights_sets(constrs, constr_weights, rhs, coefs, device=0):
    """Same signature and return shape as the reference (:89-103); the edge list is written by one kernel."""
    import torch
    dev = torch.device("cuda", _device_index(device))
    n, m = len(coefs), len(rhs)
    indptr, indices, values = csr_from_constrs(constrs, constr_weights, n)
    nnz = int(indptr[-1])
    d_indptr = torch.as_tensor(indptr, device=dev)
    d_indices = torch.as_tensor(indices, device=dev)
    d_values = torch.as_tensor(values, device=dev)
    edge_index = torch.empty((2, nnz), dtype=torch.int64, device=dev)
    edge_attr = torch.empty((nnz, 1), dtype=torch.float32, device=dev)
    stream = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
    _cabi.check(_cabi.lib().mllp_graph_edges(m, nnz, d_indptr.data_ptr(), d_indices.data_ptr(), d_values.data_ptr(),
                                             edge_index.data_ptr(), edge_attr.data_ptr(), stream), "mllp_graph_edges")
    x_src = torch.as_tensor(np.asarray(coefs, dtype=np.float32), device=dev).unsqueeze(-1)
    x_tgt = torch.as_tensor(np.asarray(rhs, dtype=np.float32), device=dev).unsqueeze(-1)
    g = BipartiteData(edge_index, x_src, x_tgt, edge_attr)
    g.__dict__["_mllp_source"] = (constrs, constr_weights, rhs, coefs, _device_index(device))
    return g
