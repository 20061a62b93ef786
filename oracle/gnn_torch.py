"""CPU oracle (test infrastructure, not a product path): the reference's GNNModel as plain PyTorch ops with autograd,
and a numpy restatement of the BACKWARD pass in exactly the factorisation the CUDA kernels use.

`torch_model_loss_and_grads` is the checker of the device backward (mllp_gnn_backward): the model of
linear_program_methods.py:199-251 written edge by edge with index_add_ / scatter_reduce (as torch_geometric's
TransformerConv runs it; that package is neither vendored nor installed, PARITY UNPINNED as for the forward), in
float64, differentiated by torch.autograd; the loss is the reference's criterion BCEWithLogitsLoss
(linear_program_experiment.py:41, :140).

`backward_numpy` restates the device algorithm (mllp_b200/csrc/gnn_backward.cu) in numpy float64: per conv a pass over
the destination rows (softmax statistics recomputed, d score, d query-side vectors), a pass over the SOURCE rows along
the transposed structure (no scatter, no atomics), the parameter gradients as sums of per-node outer products, and the
chain rule through the fused parameter block (MQ = Wq'Wk / 4, ...).  tests/test_gnn_backward.py holds it to autograd.
"""
import numpy as np

from .gnn_numpy import C, CONVS

ALL_CONVS = CONVS + ("gconv3_s2w",)
PARTS = ("lin_key", "lin_query", "lin_value", "lin_edge", "lin_skip")   # torch_geometric's registration order


def _edges(A):
    A = A.tocsr()
    m, n = A.shape
    var = A.indices.astype(np.int64)
    con = np.repeat(np.arange(m, dtype=np.int64), np.diff(A.indptr))
    return m, n, var, con, A.data.astype(np.float32)


def torch_forward(params, A, rhs, coefs, dtype=None):
    """logits (n,) as a torch tensor; `params`: dict name -> torch tensor (may require grad)"""
    import torch
    dtype = dtype or torch.float64
    m, n, var, con, attr = _edges(A)
    var, con = torch.as_tensor(var), torch.as_tensor(con)
    attr = torch.as_tensor(attr).to(dtype).unsqueeze(-1)
    P = lambda k: params[k].to(dtype)

    def conv(name, xs, xd, src, dst):
        lin = lambda part, x: x @ P(name + "." + part + ".weight").T + P(name + "." + part + ".bias")
        q, k, v = lin("lin_query", xd), lin("lin_key", xs), lin("lin_value", xs)
        e = attr @ P(name + ".lin_edge.weight").T
        s = (q[dst] * (k[src] + e)).sum(-1) / 4.0
        mx = torch.full((xd.shape[0],), -float("inf"), dtype=dtype).scatter_reduce(0, dst, s.detach(), reduce="amax")
        p = torch.exp(s - mx[dst])
        den = torch.zeros(xd.shape[0], dtype=dtype).index_add(0, dst, p)
        alpha = p / (den[dst] + 1e-16)
        out = torch.zeros(xd.shape[0], C, dtype=dtype).index_add(0, dst, alpha.unsqueeze(-1) * (v[src] + e))
        return out + lin("lin_skip", xd)

    x1 = torch.as_tensor(np.asarray(coefs, dtype=np.float32)).to(dtype).unsqueeze(-1)
    x2 = torch.as_tensor(np.asarray(rhs, dtype=np.float32)).to(dtype).unsqueeze(-1)
    n1, n2 = torch.relu(conv("gconv1_w2s", x2, x1, con, var)), torch.relu(conv("gconv1_s2w", x1, x2, var, con))
    x1, x2 = n1, n2
    n1, n2 = torch.relu(conv("gconv2_w2s", x2, x1, con, var)), torch.relu(conv("gconv2_s2w", x1, x2, var, con))
    x1, x2 = n1, n2
    n1 = torch.relu(conv("gconv3_w2s", x2, x1, con, var))
    return (n1 @ P("fc.weight").T + P("fc.bias")).reshape(-1)


def torch_model_loss_and_grads(state, A, rhs, coefs, target=None, dout=None):
    """(logits, loss, {name: grad}) in float64.  With `target` the loss is BCEWithLogitsLoss(logits, target); with
    `dout` it is sum(logits * dout) (so d loss / d logits = dout)."""
    import torch
    params = {k: torch.tensor(np.asarray(v, dtype=np.float64), requires_grad=True) for k, v in state.items()}
    out = torch_forward(params, A, rhs, coefs)
    if target is not None:
        loss = torch.nn.BCEWithLogitsLoss()(out, torch.as_tensor(np.asarray(target, dtype=np.float64)))
    else:
        loss = (out * torch.as_tensor(np.asarray(dout, dtype=np.float64))).sum()
    loss.backward()
    grads = {k: (p.grad.numpy() if p.grad is not None else np.zeros(p.shape)) for k, p in params.items()}
    return out.detach().numpy(), float(loss.detach()), grads


# ----------------------------------------------------------------------------------------------------------------
# the device algorithm, restated

def pack(state, cv):
    """the fused parameter block of one conv as a dict (float64)"""
    g = lambda part, kind: np.asarray(state["%s.%s.%s" % (cv, part, kind)], dtype=np.float64)
    Wq, bq, Wk = g("lin_query", "weight"), g("lin_query", "bias"), g("lin_key", "weight")
    We = g("lin_edge", "weight").reshape(-1)
    return dict(MQ=Wq.T @ Wk / 4.0, vq=Wk.T @ bq / 4.0, wq=Wq.T @ We / 4.0, sq=We @ bq / 4.0,
                WvT=g("lin_value", "weight").T.copy(), bv=g("lin_value", "bias"), WsT=g("lin_skip", "weight").T.copy(),
                bs=g("lin_skip", "bias"), We=We)


def unpack_grads(state, cv, d):
    """chain rule through pack(): gradients of the module's tensors from those of the fused block"""
    g = lambda part, kind: np.asarray(state["%s.%s.%s" % (cv, part, kind)], dtype=np.float64)
    Wq, bq, Wk = g("lin_query", "weight"), g("lin_query", "bias"), g("lin_key", "weight")
    We = g("lin_edge", "weight").reshape(-1)
    dWq = (Wk @ d["MQ"].T + np.outer(We, d["wq"])) / 4.0
    dWk = (Wq @ d["MQ"] + np.outer(bq, d["vq"])) / 4.0
    dbq = (Wk @ d["vq"] + d["sq"] * We) / 4.0
    dWe = d["We"] + (Wq @ d["wq"] + d["sq"] * bq) / 4.0
    return {cv + ".lin_query.weight": dWq, cv + ".lin_query.bias": dbq, cv + ".lin_key.weight": dWk,
            cv + ".lin_key.bias": np.zeros(C), cv + ".lin_value.weight": d["WvT"].T, cv + ".lin_value.bias": d["bv"],
            cv + ".lin_skip.weight": d["WsT"].T, cv + ".lin_skip.bias": d["bs"], cv + ".lin_edge.weight": dWe.reshape(C, 1)}


def conv_forward(p, M, hdst, hsrc):
    """pre-activation output of one conv along the rows of M (rows = destination nodes) + what the backward recomputes"""
    M = M.tocsr()
    nd = M.shape[0]
    dst = np.repeat(np.arange(nd), np.diff(M.indptr))
    src, a = M.indices, M.data.astype(np.float32).astype(np.float64)
    qt = hdst @ p["MQ"] + p["vq"]
    qe = hdst @ p["wq"] + p["sq"]
    s = np.einsum("ed,ed->e", qt[dst], hsrc[src]) + a * qe[dst]
    mx = np.full(nd, -np.inf)
    np.maximum.at(mx, dst, s)
    pe = np.exp(s - mx[dst])
    l = np.zeros(nd)
    np.add.at(l, dst, pe)
    any_ = l > 0
    lse = np.where(any_, mx + np.log(np.where(any_, l, 1.0)), 0.0)
    alpha = np.exp(s - lse[dst])
    xbar = np.zeros_like(hdst)
    np.add.at(xbar, dst, alpha[:, None] * hsrc[src])
    abar = np.zeros(nd)
    np.add.at(abar, dst, alpha * a)
    out = xbar @ p["WvT"] + any_[:, None] * p["bv"] + np.outer(abar, p["We"]) + hdst @ p["WsT"] + p["bs"]
    return out, dict(dst=dst, src=src, a=a, qt=qt, qe=qe, lse=lse, alpha=alpha, xbar=xbar, abar=abar, any=any_)


def conv_backward(p, M, hdst, hsrc, gout):
    """gout = d loss / d (pre-activation output).  Returns (d hdst, d hsrc, gradients of the fused block)."""
    _, f = conv_forward(p, M, hdst, hsrc)
    dst, src, a, alpha = f["dst"], f["src"], f["a"], f["alpha"]
    # destination pass
    dxbar = gout @ p["WvT"].T
    dabar = gout @ p["We"]
    D = np.einsum("nd,nd->n", dxbar, f["xbar"]) + dabar * f["abar"]
    dalpha = np.einsum("ed,ed->e", dxbar[dst], hsrc[src]) + dabar[dst] * a
    ds = alpha * (dalpha - D[dst])
    dqt = np.zeros_like(hdst)
    np.add.at(dqt, dst, ds[:, None] * hsrc[src])
    dqe = np.zeros(hdst.shape[0])
    np.add.at(dqe, dst, ds * a)
    dhdst = gout @ p["WsT"].T + dqt @ p["MQ"].T + np.outer(dqe, p["wq"])
    # source pass (the kernels walk the transposed structure: every source row owns its gradient row)
    dhsrc = np.zeros_like(hsrc)
    np.add.at(dhsrc, src, alpha[:, None] * dxbar[dst] + ds[:, None] * f["qt"][dst])
    # parameter gradients: sums over the nodes of outer products of per-node vectors
    d = dict(MQ=hdst.T @ dqt, vq=dqt.sum(0), wq=hdst.T @ dqe, sq=dqe.sum(), WvT=f["xbar"].T @ gout,
             bv=(gout * f["any"][:, None]).sum(0), WsT=hdst.T @ gout, bs=gout.sum(0), We=f["abar"] @ gout)
    return dhdst, dhsrc, d


def backward_numpy(state, A, rhs, coefs, dout):
    """(logits, {name: grad}) by the device algorithm, float64"""
    A = A.tocsr()
    AT = A.T.tocsr()
    n, m = A.shape[1], A.shape[0]
    x1 = np.asarray(coefs, dtype=np.float32).astype(np.float64).reshape(n, 1)
    x2 = np.asarray(rhs, dtype=np.float32).astype(np.float64).reshape(m, 1)
    P = {cv: pack(state, cv) for cv in CONVS}
    relu = lambda z: np.maximum(z, 0.0)
    z1a, _ = conv_forward(P["gconv1_w2s"], AT, x1, x2)
    z2a, _ = conv_forward(P["gconv1_s2w"], A, x2, x1)
    h1a, h2a = relu(z1a), relu(z2a)
    z1b, _ = conv_forward(P["gconv2_w2s"], AT, h1a, h2a)
    z2b, _ = conv_forward(P["gconv2_s2w"], A, h2a, h1a)
    h1b, h2b = relu(z1b), relu(z2b)
    z1c, _ = conv_forward(P["gconv3_w2s"], AT, h1b, h2b)
    h1c = relu(z1c)
    fw = np.asarray(state["fc.weight"], dtype=np.float64).reshape(-1)
    out = h1c @ fw + float(np.asarray(state["fc.bias"]).reshape(-1)[0])
    dout = np.asarray(dout, dtype=np.float64)
    grads = {"fc.weight": (dout @ h1c).reshape(1, C), "fc.bias": np.array([dout.sum()])}
    g1c = np.outer(dout, fw) * (z1c > 0)
    dh1b, dh2b, d = conv_backward(P["gconv3_w2s"], AT, h1b, h2b, g1c)
    grads.update(unpack_grads(state, "gconv3_w2s", d))
    dh1a, dh2a_s, d = conv_backward(P["gconv2_w2s"], AT, h1a, h2a, dh1b * (z1b > 0))
    grads.update(unpack_grads(state, "gconv2_w2s", d))
    dh2a, dh1a_s, d = conv_backward(P["gconv2_s2w"], A, h2a, h1a, dh2b * (z2b > 0))
    grads.update(unpack_grads(state, "gconv2_s2w", d))
    dh1a, dh2a = dh1a + dh1a_s, dh2a + dh2a_s
    _, _, d = conv_backward(P["gconv1_w2s"], AT, x1, x2, dh1a * (z1a > 0))
    grads.update(unpack_grads(state, "gconv1_w2s", d))
    _, _, d = conv_backward(P["gconv1_s2w"], A, x2, x1, dh2a * (z2a > 0))
    grads.update(unpack_grads(state, "gconv1_s2w", d))
    for k, v in state.items():
        if k.startswith("gconv3_s2w"):
            grads[k] = np.zeros(np.asarray(v).shape)
    return out, grads
