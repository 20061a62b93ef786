"""Bipartite message-passing forward (reference GNNModel, linear_program_methods.py:199-251).

torch_geometric is not available (nor pinned by the reference), so parity is UNPINNED for this row: the numpy oracle
(oracle/gnn_numpy.py) restates TransformerConv from its published definition, a plain-PyTorch fp32 reference written
here with index_add_ / scatter-reduce agrees with it (CPU), and the CUDA kernels are compared with the oracle (GPU).
Tolerance (fp32 arithmetic, fp64 oracle): 2e-4 relative to the largest output magnitude."""
import numpy as np
import pytest
import scipy.sparse as sp

import mllp_b200.linear_program_data as D
from oracle import gnn_numpy as G

TOL = 2e-4


def torch_reference(st, A, rhs, coefs):
    """the same forward in plain float32 PyTorch ops on the CPU (edge-list formulation, as PyG would run it)"""
    import torch
    A = A.tocsr()
    m, n = A.shape
    var = torch.as_tensor(A.indices.astype(np.int64))
    con = torch.as_tensor(np.repeat(np.arange(m), np.diff(A.indptr)).astype(np.int64))
    attr = torch.as_tensor(A.data.astype(np.float32)).unsqueeze(-1)
    T = lambda k: torch.as_tensor(st[k])

    def conv(name, xs, xd, src, dst):
        lin = lambda part, x: x @ T(name + "." + part + ".weight").T + T(name + "." + part + ".bias")
        q, k, v = lin("lin_query", xd), lin("lin_key", xs), lin("lin_value", xs)
        e = attr @ T(name + ".lin_edge.weight").T
        s = (q[dst] * (k[src] + e)).sum(-1) / 4.0
        mx = torch.full((xd.shape[0],), -float("inf")).scatter_reduce(0, dst, s, reduce="amax")
        p = torch.exp(s - mx[dst])
        den = torch.zeros(xd.shape[0]).index_add_(0, dst, p)
        alpha = p / (den[dst] + 1e-16)
        out = torch.zeros(xd.shape[0], 16).index_add_(0, dst, alpha.unsqueeze(-1) * (v[src] + e))
        return out + lin("lin_skip", xd)

    x1 = torch.as_tensor(np.asarray(coefs, dtype=np.float32)).unsqueeze(-1)
    x2 = torch.as_tensor(np.asarray(rhs, dtype=np.float32)).unsqueeze(-1)
    n1, n2 = torch.relu(conv("gconv1_w2s", x2, x1, con, var)), torch.relu(conv("gconv1_s2w", x1, x2, var, con))
    x1, x2 = n1, n2
    n1, n2 = torch.relu(conv("gconv2_w2s", x2, x1, con, var)), torch.relu(conv("gconv2_s2w", x1, x2, var, con))
    x1, x2 = n1, n2
    n1 = torch.relu(conv("gconv3_w2s", x2, x1, con, var))
    return (n1 @ T("fc.weight").T + T("fc.bias")).squeeze().numpy()


def close(a, b):
    return np.max(np.abs(np.asarray(a, dtype=np.float64) - b)) <= TOL * max(1.0, np.max(np.abs(b)))


@pytest.mark.parametrize("name", ["afiro", "sc105", "25fv47"])
def test_oracle_agrees_with_plain_torch_fp32(name):
    A, b, c = D.load_csr(name)
    st = G.init_state(3)
    assert close(torch_reference(st, A, b, c), G.gnn_forward(st, A, b, c))


def test_oracle_edge_cases():
    # a constraint without entries and a variable without entries keep only the root (skip) term
    A = sp.csr_matrix(np.array([[1.0, 0.0, 2.0], [0.0, 0.0, 0.0], [3.0, 0.0, -1.0]]))
    st = G.init_state(1)
    out = G.gnn_forward(st, A, np.array([1.0, 2.0, 3.0]), np.array([0.5, -0.5, 0.25]))
    assert out.shape == (3,) and np.all(np.isfinite(out))
    assert close(torch_reference(st, A, np.array([1.0, 2.0, 3.0]), np.array([0.5, -0.5, 0.25])), out)
    # attention weights are invariant to the order of a node's edges: permuting the columns permutes the output
    A2, b2, c2 = D.load_csr("afiro")
    perm = np.random.default_rng(0).permutation(A2.shape[1])
    o1 = G.gnn_forward(st, A2, b2, c2)
    o2 = G.gnn_forward(st, A2[:, perm].tocsr(), b2, c2[perm])
    assert np.allclose(o2, o1[perm], rtol=0, atol=1e-12)


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["afiro", "sc105", "25fv47", "pilot87", "ken-18", "osa-60"])
def test_gpu_forward_matches_oracle(name):
    """ken-18 and osa-60 exercise the cut rows (rows of up to 173 366 edges, items merged in a fixed order)"""
    import mllp_b200.gnn as GN
    A, b, c = D.load_csr(name)
    st = G.init_state(5)
    g = GN.BipartiteGraph(np.split(A.indices, A.indptr)[1:-1], A.data, b, c)
    assert (g.to_con.nlong > 0) == (np.diff(A.indptr).max() > g.to_con.row_max)
    assert name not in ("ken-18", "osa-60") or g.to_con.nlong > 0   # rows of 325 and 173 366 edges
    model = GN.GNNModel(st)
    out = model(g)
    assert out.shape == (A.shape[1],) and out.is_cuda and out.dtype.is_floating_point
    ref = G.gnn_forward(st, A, b, c)
    assert close(out.cpu().numpy(), ref)
    # deterministic: same launch twice gives the same bits
    assert np.array_equal(model(g).cpu().numpy(), out.cpu().numpy())


@pytest.mark.gpu
def test_gpu_forward_edge_cases_and_errors():
    import mllp_b200.gnn as GN
    A = sp.csr_matrix(np.array([[1.0, 0.0, 2.0], [0.0, 0.0, 0.0], [3.0, 0.0, -1.0]]))
    b, c = np.array([1.0, 2.0, 3.0]), np.array([0.5, -0.5, 0.25])
    st = G.init_state(1)
    g = GN.BipartiteGraph(A, A.data, b, c)
    assert close(GN.GNNModel(st)(g).cpu().numpy(), G.gnn_forward(st, A, b, c))
    with pytest.raises(TypeError):
        GN.GNNModel(st)(object())
    bad = dict(st)
    bad["gconv2_w2s.lin_key.weight"] = np.zeros((16, 3), np.float32)
    with pytest.raises(ValueError):
        GN.GNNModel(bad)
    # default parameters load and run
    assert np.all(np.isfinite(GN.GNNModel(seed=2)(g).cpu().numpy()))


@pytest.mark.gpu
@pytest.mark.parametrize("din,name", [(1, "25fv47"), (16, "25fv47"), (16, "pilot87")])
def test_gpu_single_conv_matches_oracle(din, name):
    """mllp_gnn_conv (one TransformerConv along the rows of A, no ReLU) on random features against the oracle's layer"""
    import torch
    import mllp_b200.gnn as GN
    from mllp_b200 import _cabi
    import ctypes
    A, b, c = D.load_csr(name)
    m, n = A.shape
    rng = np.random.default_rng(din)
    st = G.init_state(9)
    cv = "gconv1_s2w" if din == 1 else "gconv2_s2w"
    xs, xd = rng.standard_normal((n, din)).astype(np.float32), rng.standard_normal((m, din)).astype(np.float32)
    g = GN.BipartiteGraph(np.split(A.indices, A.indptr)[1:-1], A.data, b, c)
    d, blk = GN.pack_conv(st, cv)
    assert d == din
    t = lambda a: torch.as_tensor(a, device="cuda")
    dxs, dxd, dp = t(xs), t(xd), t(blk)
    out = torch.full((m, 16), float("nan"), dtype=torch.float32, device="cuda")
    _cabi.check(_cabi.lib().mllp_gnn_conv(ctypes.byref(g.to_con.c), din, dxd.data_ptr(), dxs.data_ptr(), dp.data_ptr(),
                                          out.data_ptr(), 0, None), "mllp_gnn_conv")
    torch.cuda.synchronize()
    var = A.indices.astype(np.int64)
    con = np.repeat(np.arange(m, dtype=np.int64), np.diff(A.indptr))
    ref = G.transformer_conv(st, cv, xs, xd, var, con, A.data.astype(np.float32))
    assert close(out.cpu().numpy(), ref)


@pytest.mark.gpu
@pytest.mark.parametrize("group", [1, 2, 4, 8, 16, 32])
def test_gpu_forward_every_lane_group(group):
    """every lanes-per-row instantiation of the conv kernel (the graph normally picks it from the row lengths), with rows
    above the group's limit going through the cut-row kernels"""
    import mllp_b200.gnn as GN
    A, b, c = D.load_csr("25fv47")
    st = G.init_state(group)
    g = GN.BipartiteGraph(np.split(A.indices, A.indptr)[1:-1], A.data, b, c, groups=(group, group))
    assert g.to_var.group == group and g.to_con.group == group
    out = GN.GNNModel(st)(g).cpu().numpy()
    assert close(out, G.gnn_forward(st, A, b, c))


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["afiro", "pilot87", "ken-18"])
def test_gpu_forward_plan_replays_bitwise(name):
    """the captured forward (CUDA graph, the two convs of a layer on parallel branches) gives the bits of the plain
    launches, replays after the inputs changed in place, and one graph serves several parameter sets"""
    import torch
    import mllp_b200.gnn as GN
    A, b, c = D.load_csr(name)
    g = GN.BipartiteGraph(np.split(A.indices, A.indptr)[1:-1], A.data, b, c)
    m1, m2 = GN.GNNModel(G.init_state(5)), GN.GNNModel(G.init_state(6))
    plain1, plain2 = m1.forward(g, use_plan=False).cpu().numpy(), m2.forward(g, use_plan=False).cpu().numpy()
    for _ in range(3):
        assert np.array_equal(m1(g).cpu().numpy(), plain1) and np.array_equal(m2(g).cpu().numpy(), plain2)
    assert close(plain1, G.gnn_forward(m1.state, A, b, c))
    g.x2.mul_(0.5)   # new right-hand sides in place: the plan reads the same buffers
    torch.cuda.synchronize()
    assert close(m1(g).cpu().numpy(), G.gnn_forward(m1.state, A, 0.5 * b, c))
    g.close()


GOLD = __import__("os").path.join(__import__("os").path.dirname(__file__), "golden", "gnn_forward_seed5.npz")


@pytest.mark.parametrize("name", ["afiro", "sc105"])
def test_oracle_reproduces_committed_fixture(name):
    """tests/golden/gnn_forward_seed5.npz (make_golden.py gnn): the oracle's logits and a plain-torch fp32 run, committed"""
    g = np.load(GOLD)
    A, b, c = D.load_csr(name)
    out = G.gnn_forward(G.init_state(5), A, b, c)
    assert np.allclose(out, g[name + "_oracle"], rtol=0, atol=1e-12)
    assert close(g[name + "_torch_fp32"], out)


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["afiro", "sc105"])
def test_gpu_forward_matches_committed_fixture(name):
    import mllp_b200.gnn as GN
    g = np.load(GOLD)
    A, b, c = D.load_csr(name)
    gr = GN.BipartiteGraph(np.split(A.indices, A.indptr)[1:-1], A.data, b, c)
    out = GN.GNNModel(G.init_state(5))(gr).cpu().numpy()
    assert close(out, g[name + "_oracle"]) and close(out, g[name + "_torch_fp32"].astype(np.float64))
    gr.close()
