"""Backward pass of the message-passing model (the reference trains GNNModel: linear_program_experiment.py:115-157).

Checker: torch.autograd through a plain-PyTorch float64 restatement of the model (oracle/gnn_torch.py; torch_geometric is
not available, PARITY UNPINNED as for the forward).  CPU: the numpy restatement of the DEVICE algorithm (destination pass,
source pass along the transposed structure, outer-product parameter gradients, chain rule through the fused block) equals
autograd to rounding.  GPU: mllp_gnn_backward through the C ABI against autograd; tolerance 2e-3 of the largest entry of
each gradient tensor (fp32 kernels, fp64 checker)."""
import numpy as np
import pytest
import scipy.sparse as sp

import mllp_b200.linear_program_data as D
from oracle import gnn_numpy as G
from oracle import gnn_torch as T

GTOL = 2e-3


def grad_errors(got, ref):
    """per tensor: max |got - ref| / (max |ref| of the tensor, floored at 1e-4 of the largest gradient entry overall)"""
    top = max(np.abs(v).max() for v in ref.values())
    return {k: float(np.abs(np.asarray(got[k], dtype=np.float64) - ref[k]).max() / max(np.abs(ref[k]).max(), 1e-4 * top)) for k in ref}


@pytest.mark.parametrize("name", ["afiro", "sc105", "25fv47"])
def test_device_algorithm_restated_in_numpy_equals_autograd(name):
    A, b, c = D.load_csr(name)
    st = G.init_state(3)
    dout = np.random.default_rng(1).standard_normal(A.shape[1])
    out, _, ref = T.torch_model_loss_and_grads(st, A, b, c, dout=dout)
    assert np.allclose(out, G.gnn_forward(st, A, b, c), rtol=0, atol=1e-12)
    out2, got = T.backward_numpy(st, A, b, c, dout)
    assert np.allclose(out2, out, rtol=0, atol=1e-12)
    assert set(got) == set(ref)
    assert max(grad_errors(got, ref).values()) < 1e-9


def test_backward_numpy_edge_cases():
    # a constraint and a variable without entries (only the root term reaches them)
    A = sp.csr_matrix(np.array([[1.0, 0.0, 2.0], [0.0, 0.0, 0.0], [3.0, 0.0, -1.0]]))
    b, c = np.array([1.0, 2.0, 3.0]), np.array([0.5, -0.5, 0.25])
    st = G.init_state(1)
    dout = np.array([0.3, -1.0, 2.0])
    _, _, ref = T.torch_model_loss_and_grads(st, A, b, c, dout=dout)
    _, got = T.backward_numpy(st, A, b, c, dout)
    assert max(grad_errors(got, ref).values()) < 1e-9
    # the unused conv and the key biases get no gradient
    assert not np.any(got["gconv3_s2w.lin_value.weight"]) and not np.any(got["gconv2_w2s.lin_key.bias"])


def test_flat_layout_matches_the_library():
    from mllp_b200 import _cabi
    from mllp_b200.gnn_train import flat_layout
    layout, total = flat_layout()
    assert total == int(_cabi.lib().mllp_gnn_flat_param_floats()) == 4721
    names = [n for n, _, _ in layout]
    assert set(names) == set(G.init_state(0).keys())
    offs = [o for _, o, _ in layout]
    assert offs == sorted(offs) and offs[0] == 0


def _device_grads(st, A, b, c, dout=None, target=None, groups=None, use_plans=True):
    import torch
    import mllp_b200.gnn as GN
    from mllp_b200.gnn_train import TrainableGNNModel
    g = GN.BipartiteGraph(np.split(A.indices, A.indptr)[1:-1], A.data, b, c, groups=groups)
    model = TrainableGNNModel(st, use_plans=use_plans)
    out = model(g)
    if dout is not None:
        loss = (out * torch.as_tensor(np.asarray(dout, dtype=np.float32), device=out.device)).sum()
    else:
        loss = torch.nn.BCEWithLogitsLoss()(out, torch.as_tensor(np.asarray(target, dtype=np.float32), device=out.device))
    loss.backward()
    grads = {k: v.detach().cpu().numpy() for k, v in model.named_gradients().items()}
    return out.detach().cpu().numpy(), float(loss.detach()), grads, model, g


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["afiro", "sc105", "25fv47", "pilot87", "ken-18", "osa-60"])
def test_gpu_backward_matches_autograd(name):
    """ken-18 / osa-60: rows above the chunk size in BOTH passes (one CTA per row; osa-60 has a row of 173 366 edges)"""
    A, b, c = D.load_csr(name)
    st = G.init_state(5)
    dout = np.random.default_rng(2).standard_normal(A.shape[1]) / np.sqrt(A.shape[1])
    out_ref, _, ref = T.torch_model_loss_and_grads(st, A, b, c, dout=dout)
    out, _, got, _, g = _device_grads(st, A, b, c, dout=dout)
    assert name not in ("ken-18", "osa-60") or g.to_con.nlong > 0
    assert np.max(np.abs(out - out_ref)) <= 2e-4 * max(1.0, np.max(np.abs(out_ref)))
    err = grad_errors(got, ref)
    assert max(err.values()) < GTOL, sorted(err.items(), key=lambda kv: -kv[1])[:4]


@pytest.mark.gpu
@pytest.mark.parametrize("groups", [(1, 1), (2, 2), (4, 8), (8, 4), (16, 32), (32, 16)])
def test_gpu_backward_every_lane_group(groups):
    """every lanes-per-row variant of the destination and source kernels (25fv47 forced onto each; rows above the forced
    group's chunk go to the CTA-per-row kernels)"""
    A, b, c = D.load_csr("25fv47")
    st = G.init_state(7)
    target = (np.random.default_rng(3).random(A.shape[1]) < 0.4).astype(np.float64)
    _, loss_ref, ref = T.torch_model_loss_and_grads(st, A, b, c, target=target)
    _, loss, got, _, _ = _device_grads(st, A, b, c, target=target, groups=groups)
    assert abs(loss - loss_ref) <= 1e-4 * max(1.0, abs(loss_ref))
    err = grad_errors(got, ref)
    assert max(err.values()) < GTOL, sorted(err.items(), key=lambda kv: -kv[1])[:4]


@pytest.mark.gpu
def test_gpu_backward_edge_cases_determinism_and_errors():
    import torch
    A = sp.csr_matrix(np.array([[1.0, 0.0, 2.0], [0.0, 0.0, 0.0], [3.0, 0.0, -1.0]]))
    b, c = np.array([1.0, 2.0, 3.0]), np.array([0.5, -0.5, 0.25])
    st = G.init_state(1)
    dout = np.array([0.3, -1.0, 2.0])
    _, _, ref = T.torch_model_loss_and_grads(st, A, b, c, dout=dout)
    _, _, got, model, g = _device_grads(st, A, b, c, dout=dout)
    assert max(grad_errors(got, ref).values()) < GTOL
    assert not np.any(got["gconv3_s2w.lin_value.weight"]) and not np.any(got["gconv1_w2s.lin_key.bias"])
    # no atomics: the same backward twice gives the same bits, replayed as a CUDA graph or launched kernel by kernel
    A2, b2, c2 = D.load_csr("25fv47")
    d2 = np.random.default_rng(0).standard_normal(A2.shape[1])
    _, _, g1, _, _ = _device_grads(st, A2, b2, c2, dout=d2)
    _, _, g2, _, _ = _device_grads(st, A2, b2, c2, dout=d2)
    _, _, g3, _, _ = _device_grads(st, A2, b2, c2, dout=d2, use_plans=False)
    assert all(np.array_equal(g1[k], g2[k]) and np.array_equal(g1[k], g3[k]) for k in g1)
    # a plan is replayed with new parameters and a new upstream gradient (the pointers stay, the contents change)
    import mllp_b200.gnn as GN
    from mllp_b200.gnn_train import TrainableGNNModel
    gg = GN.BipartiteGraph(np.split(A2.indices, A2.indptr)[1:-1], A2.data, b2, c2)
    mm = TrainableGNNModel(G.init_state(8))
    mm(gg).backward(torch.as_tensor(d2.astype(np.float32), device="cuda"))     # captures both plans with other contents
    mm.flat.grad = None
    mm.load_state_dict(st)
    mm(gg).backward(torch.as_tensor(d2.astype(np.float32), device="cuda"))
    assert all(np.array_equal(mm.named_gradients()[k].cpu().numpy(), g1[k]) for k in g1)
    gg.close()
    # a second forward on the same graph overwrites the activations: backward of the first one must refuse
    o1 = model(g)
    model(g)
    with pytest.raises(RuntimeError):
        o1.sum().backward()
    with pytest.raises(TypeError):
        model(object())
    # state_dict speaks the reference's names and round-trips through the forward-only model
    import mllp_b200.gnn as GN
    sd = model.state_dict()
    assert set(sd) == set(st) and all(np.array_equal(sd[k].cpu().numpy(), st[k]) for k in st)
    fwd_only = GN.GNNModel(sd)(g).cpu().numpy()
    with torch.no_grad():
        assert np.allclose(model(g).cpu().numpy(), fwd_only, rtol=0, atol=1e-5)


@pytest.mark.gpu
def test_gpu_training_loop_as_the_reference_runs_it():
    """linear_program_experiment.py:115-157 on afiro: build_graph -> model(graph) -> BCEWithLogitsLoss -> backward -> Adam;
    the loss falls and the trajectory follows the same loop run with autograd on the CPU checker (float64)"""
    import torch
    from mllp_b200.graph import build_graph_from_weights_sets
    from mllp_b200.gnn_train import TrainableGNNModel
    A, b, c = D.load_csr("afiro")
    constrs = np.split(A.indices, A.indptr)[1:-1]
    basis = (np.random.default_rng(4).random(A.shape[1]) < 0.5).astype(np.float32)
    st = G.init_state(11)
    model = TrainableGNNModel(st)
    opt = torch.optim.Adam(model.parameters(), lr=1e-2)
    crit = torch.nn.BCEWithLogitsLoss()
    graph = build_graph_from_weights_sets(constrs, A.data, b, c, 0)
    # the reference rebuilds the graph every epoch (:124): the device form is cached per (arrays, device)
    assert build_graph_from_weights_sets(constrs, A.data, b, c, 0).bipartite_graph() is graph.bipartite_graph()
    losses = []
    for _ in range(30):
        graph = build_graph_from_weights_sets(constrs, A.data, b, c, 0)
        obj = crit(model(graph), torch.as_tensor(basis, device="cuda"))
        obj.backward()
        opt.step()
        opt.zero_grad()
        losses.append(float(obj))
    # the checker's loop
    params = {k: torch.tensor(np.asarray(v, dtype=np.float64), requires_grad=True) for k, v in st.items()}
    opt2 = torch.optim.Adam(list(params.values()), lr=1e-2)
    ref = []
    for _ in range(30):
        obj = crit(T.torch_forward(params, A, b, c), torch.as_tensor(basis.astype(np.float64)))
        obj.backward()
        opt2.step()
        opt2.zero_grad()
        ref.append(float(obj))
    assert losses[-1] < losses[0]
    assert np.max(np.abs(np.array(losses) - np.array(ref))) < 2e-3


def test_device_graph_cache_follows_identity_and_content(monkeypatch):
    """build_graph_from_weights_sets is called per instance per epoch by the reference's loop (:124): the device form is kept
    per (arrays, device) and rebuilt when an array was edited in place"""
    import mllp_b200.gnn as GN
    import mllp_b200.graph as GR

    class Stub:
        made = 0

        def __init__(self, *a):
            Stub.made += 1
            self.closed = False

        def close(self):
            self.closed = True

    monkeypatch.setattr(GN, "BipartiteGraph", Stub)
    monkeypatch.setattr(GR, "_GRAPH_CACHE", {})
    w, b, c = np.ones(5), np.ones(2), np.ones(3)
    g1, g2 = GR._cached_graph(([], w, b, c, 0)), GR._cached_graph(([], w, b, c, 0))
    assert g1 is g2 and Stub.made == 1
    assert GR._cached_graph(([], w, b, c.copy(), 0)) is not g1          # another array object
    c[0] = 2.0                                                          # edited in place
    g3 = GR._cached_graph(([], w, b, c, 0))
    assert g3 is not g1 and g1.closed
    GR.clear_graph_cache()
    assert g3.closed and not GR._GRAPH_CACHE


@pytest.mark.gpu
@pytest.mark.parametrize("seed", range(8))
def test_gpu_backward_on_random_graphs(seed):
    """random sparse matrices: empty rows and columns, a dense row and a dense column (cut into items when the forced lane
    group makes them 'long'), sizes that are not multiples of the rows per warp"""
    rng = np.random.default_rng(100 + seed)
    m, n = int(rng.integers(1, 90)), int(rng.integers(1, 140))
    dens = float(rng.choice([0.02, 0.1, 0.4]))
    Ad = (rng.random((m, n)) < dens) * rng.standard_normal((m, n))
    if seed % 2 == 0 and m > 2 and n > 2:
        Ad[rng.integers(0, m)] = rng.standard_normal(n)        # a dense row
        Ad[:, rng.integers(0, n)] = rng.standard_normal(m)     # a dense column
    if m > 1:
        Ad[rng.integers(0, m)] = 0.0                           # an empty row
    A = sp.csr_matrix(Ad)
    b, c = rng.standard_normal(m), rng.standard_normal(n)
    st = G.init_state(seed)
    dout = rng.standard_normal(n)
    groups = (int(rng.choice([1, 2, 4, 8, 16, 32])), int(rng.choice([1, 2, 4, 8, 16, 32])))
    out_ref, _, ref = T.torch_model_loss_and_grads(st, A, b, c, dout=dout)
    out, _, got, _, g = _device_grads(st, A, b, c, dout=dout, groups=groups, use_plans=bool(seed % 2))
    assert np.max(np.abs(out - out_ref)) <= 2e-4 * max(1.0, np.max(np.abs(out_ref)))
    err = grad_errors(got, ref)
    assert max(err.values()) < GTOL, (m, n, dens, groups, sorted(err.items(), key=lambda kv: -kv[1])[:4])
    g.close()


def test_backward_restatement_property_random_graphs():
    """hypothesis: on random bipartite graphs (empty rows / columns, duplicate-free random patterns, random upstream gradients)
    the numpy restatement of the device algorithm equals autograd through the plain-PyTorch model"""
    from hypothesis import given, settings, strategies as hs

    @settings(max_examples=25, deadline=None)
    @given(hs.integers(1, 12), hs.integers(1, 16), hs.floats(0.05, 0.9), hs.integers(0, 10 ** 6))
    def check(m, n, dens, seed):
        rng = np.random.default_rng(seed)
        Ad = (rng.random((m, n)) < dens) * rng.standard_normal((m, n))
        A = sp.csr_matrix(Ad)
        b, c = rng.standard_normal(m), rng.standard_normal(n)
        st = G.init_state(seed % 97)
        dout = rng.standard_normal(n)
        out, _, ref = T.torch_model_loss_and_grads(st, A, b, c, dout=dout)
        out2, got = T.backward_numpy(st, A, b, c, dout)
        assert np.allclose(out, out2, rtol=0, atol=1e-11)
        assert max(grad_errors(got, ref).values()) < 1e-8

    check()
