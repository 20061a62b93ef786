"""Device-side LP -> bipartite graph builder.

Replaces the Python double loop of the reference's ``build_graph_from_weights_sets``
(linear_program_methods.py:89-103, executed per instance per epoch at
linear_program_experiment.py:124) by one CUDA kernel over the CSR arrays
(``mllp_graph_edges``): ``edge_index[0, k]`` = variable (column) of nonzero k,
``edge_index[1, k]`` = constraint (row) of nonzero k -- the same (var, constr) orientation and the
same nonzero order as the reference -- ``edge_attr[k, 0]`` = a_ij as float32, ``x_src`` = coefs
(n, 1) and ``x_tgt`` = rhs (m, 1) as float32.

Returns plain tensors (a dict); wrapping them in the reference's ``BipartiteData`` needs
torch_geometric, which is a third-party dependency of the reference and not of this package.
"""
import ctypes

import numpy as np

from . import _cabi
from .linear_program_methods import _device_index, csr_from_constrs


def build_graph_from_weights_sets(constrs, constr_weights, rhs, coefs, device=0):
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
    return {"edge_index": edge_index, "edge_attr": edge_attr, "x_src": x_src, "x_tgt": x_tgt}
