"""CPU oracle (test infrastructure, not a product path) of the reference's GNNModel forward.

Restates, in numpy float64, what linear_program_methods.py:238-251 computes with torch_geometric's
TransformerConv (:199-204: in_channels (a, b), out_channels 16, heads 1, edge_dim 1, defaults concat=True,
beta=False, dropout=0, bias=True, root_weight=True).  torch_geometric is neither vendored in the reference nor
installed here and its version is not pinned, so PARITY IS UNPINNED: the layer follows
torch_geometric's published definition,

    query = lin_query(x_dst)      key = lin_key(x_src)      value = lin_value(x_src)       (all with bias)
    e     = lin_edge(edge_attr)                                                            (no bias)
    alpha = softmax over the incoming edges of each destination node of  query_i . (key_j + e_ij) / sqrt(C)
    out_i = sum_j alpha_ij (value_j + e_ij)  +  lin_skip(x_dst_i)

and the graph follows build_graph_from_weights_sets (:89-103): edge_index = [variable; constraint] in CSR
nonzero order, edge_attr = float32(a_ij), x1 = coefs[:, None], x2 = rhs[:, None], all float32.
Parameter names are those of the reference module's state_dict ("gconv1_w2s.lin_key.weight", ..., "fc.bias").
"""
import numpy as np

CONVS = ("gconv1_w2s", "gconv1_s2w", "gconv2_w2s", "gconv2_s2w", "gconv3_w2s")   # gconv3_s2w is unused (:247)
C = 16


def init_state(seed=0, dtype=np.float32):
    """Random parameters in the shapes (and default init scale) of the reference module."""
    rng = np.random.default_rng(seed)
    st = {}

    def lin(name, fin, fout, bias=True):
        bound = 1.0 / np.sqrt(fin)
        st[name + ".weight"] = rng.uniform(-bound, bound, (fout, fin)).astype(dtype)
        if bias:
            st[name + ".bias"] = rng.uniform(-bound, bound, fout).astype(dtype)

    for cv in CONVS + ("gconv3_s2w",):
        din = 1 if cv.startswith("gconv1") else C
        for part in ("lin_key", "lin_query", "lin_value", "lin_skip"):
            lin(cv + "." + part, din, C)
        lin(cv + ".lin_edge", 1, C, bias=False)
    lin("fc", C, 1)
    return st


def transformer_conv(st, name, x_src, x_dst, src, dst, attr):
    """One TransformerConv((din, din), 16, edge_dim=1): edges src[e] -> dst[e] with attribute attr[e]."""
    f = lambda a: np.asarray(a, dtype=np.float64)
    W = lambda part: f(st[name + "." + part + ".weight"])
    B = lambda part: f(st[name + "." + part + ".bias"])
    q = f(x_dst) @ W("lin_query").T + B("lin_query")
    k = f(x_src) @ W("lin_key").T + B("lin_key")
    v = f(x_src) @ W("lin_value").T + B("lin_value")
    e = f(attr).reshape(-1, 1) @ W("lin_edge").T
    s = np.einsum("ec,ec->e", q[dst], k[src] + e) / np.sqrt(C)
    nd = x_dst.shape[0]
    mx = np.full(nd, -np.inf)
    np.maximum.at(mx, dst, s)
    p = np.exp(s - mx[dst])
    den = np.zeros(nd)
    np.add.at(den, dst, p)
    alpha = p / (den[dst] + 1e-16)
    out = np.zeros((nd, C))
    np.add.at(out, dst, alpha[:, None] * (v[src] + e))
    return out + f(x_dst) @ W("lin_skip").T + B("lin_skip")


def gnn_forward(st, A, rhs, coefs):
    """logit per variable (n,), float64 arithmetic on float32-rounded inputs."""
    A = A.tocsr()
    m, n = A.shape
    var = A.indices.astype(np.int64)                            # edge_index[0]: variable of each nonzero (CSR order)
    con = np.repeat(np.arange(m, dtype=np.int64), np.diff(A.indptr))
    attr = A.data.astype(np.float32)
    x1 = np.asarray(coefs, dtype=np.float32).reshape(n, 1)
    x2 = np.asarray(rhs, dtype=np.float32).reshape(m, 1)
    relu = lambda a: np.maximum(a, 0.0)
    n1 = relu(transformer_conv(st, "gconv1_w2s", x2, x1, con, var, attr))
    n2 = relu(transformer_conv(st, "gconv1_s2w", x1, x2, var, con, attr))
    x1, x2 = n1, n2
    n1 = relu(transformer_conv(st, "gconv2_w2s", x2, x1, con, var, attr))
    n2 = relu(transformer_conv(st, "gconv2_s2w", x1, x2, var, con, attr))
    x1, x2 = n1, n2
    n1 = relu(transformer_conv(st, "gconv3_w2s", x2, x1, con, var, attr))
    return (n1 @ np.asarray(st["fc.weight"], dtype=np.float64).T + np.asarray(st["fc.bias"], dtype=np.float64)).reshape(-1)
