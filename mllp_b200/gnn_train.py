"""Trainable form of the reference's ``GNNModel`` (linear_program_methods.py:199-251) on the device kernels.

The reference trains the model (linear_program_experiment.py:115-157: ``model = GNNModel().to(device)``,
``Adam(model.parameters())``, ``criterion(model(graph), basis_opt).backward()``, ``step()``).  This module keeps that
calling convention: a ``torch.nn.Module`` whose ``forward(graph)`` returns the logit per variable and whose backward runs
the hand-written kernels of mllp_b200/csrc/gnn_backward.cu through the C ABI (``mllp_gnn_backward``) -- torch supplies the
autograd plumbing, the optimiser and the loss on the (n,) logits, not the message passing.  There is no PyTorch fallback.

The parameters live in ONE flat ``torch.nn.Parameter`` (``model.flat``, layout of include/mllp_b200.h); ``state_dict()`` /
``load_state_dict()`` speak the reference module's tensor names ("gconv1_w2s.lin_key.weight", ..., "fc.bias"), so a
checkpoint of the reference model loads unchanged and vice versa.  Adam is elementwise, so stepping the flat vector is
the same update as stepping the named tensors.
"""
import collections
import ctypes

import numpy as np
import torch

from . import _cabi
from .gnn import C, BipartiteGraph, default_state
from .linear_program_methods import _device_index, _torch_stream

ALL_CONVS = ("gconv1_w2s", "gconv1_s2w", "gconv2_w2s", "gconv2_s2w", "gconv3_w2s", "gconv3_s2w")


def flat_layout():
    """[(name, offset, shape)] of the flat parameter vector (mirrors ``Flat<DIN>`` of gnn_backward.cu)"""
    out, off = [], 0
    for cv in ALL_CONVS:
        din = 1 if cv.startswith("gconv1") else C
        for part, shape in (("lin_key.weight", (C, din)), ("lin_key.bias", (C,)), ("lin_query.weight", (C, din)),
                            ("lin_query.bias", (C,)), ("lin_value.weight", (C, din)), ("lin_value.bias", (C,)),
                            ("lin_edge.weight", (C, 1)), ("lin_skip.weight", (C, din)), ("lin_skip.bias", (C,))):
            out.append(("%s.%s" % (cv, part), off, shape))
            off += int(np.prod(shape))
    out.append(("fc.weight", off, (1, C)))
    off += C
    out.append(("fc.bias", off, (1,)))
    return out, off + 1


def _plans(g, model, flat_c):
    """the (graph, parameter storage) pair's captured forward / backward (mllp_gnn_train_plan_create,
    mllp_gnn_backward_plan_create) and the fixed buffers they read and write; kept on the graph, released by g.close()"""
    key = ("train", flat_c.data_ptr(), model._packed.data_ptr())
    entry = g._plans.get(key)
    if entry is None:
        L = _cabi.lib()
        dev = flat_c.device
        out = torch.empty(g.n, dtype=torch.float32, device=dev)
        dout = torch.zeros(g.n, dtype=torch.float32, device=dev)
        dflat = torch.empty_like(flat_c)
        bwork = torch.empty(int(L.mllp_gnn_backward_workspace_floats(g.n, g.m)), dtype=torch.float32, device=dev)
        fwd, bwd = ctypes.c_void_p(), ctypes.c_void_p()
        with torch.cuda.device(dev):
            _cabi.check(L.mllp_gnn_train_plan_create(ctypes.byref(g.to_var.c), ctypes.byref(g.to_con.c), g.x1.data_ptr(),
                                                     g.x2.data_ptr(), flat_c.data_ptr(), model._packed.data_ptr(),
                                                     g.work.data_ptr(), out.data_ptr(), ctypes.byref(fwd)),
                        "mllp_gnn_train_plan_create")
            _cabi.check(L.mllp_gnn_backward_plan_create(ctypes.byref(g.to_var.c), ctypes.byref(g.to_con.c), g.x1.data_ptr(),
                                                        g.x2.data_ptr(), flat_c.data_ptr(), model._packed.data_ptr(),
                                                        g.work.data_ptr(), bwork.data_ptr(), dout.data_ptr(), dflat.data_ptr(),
                                                        ctypes.byref(bwd)), "mllp_gnn_backward_plan_create")
        # the plans hold these pointers: the tensors stay alive with the entry
        entry = g._plans[key] = (fwd, bwd, out, dout, dflat, bwork, flat_c, model._packed)
        g._plans[key + ("bwd",)] = (bwd,)   # so that g.close() destroys it too
    return entry


class _Forward(torch.autograd.Function):
    @staticmethod
    def forward(ctx, flat, model, g):
        L = _cabi.lib()
        dev = flat.device
        stream = _torch_stream(dev)
        flat_c = flat.detach()
        if not flat_c.is_contiguous():
            raise ValueError("the flat parameter vector must be contiguous")
        with torch.cuda.device(dev):   # the library launches on the calling thread's current device
            if model.use_plans:
                entry = _plans(g, model, flat_c)
                _cabi.check(L.mllp_gnn_plan_run(entry[0], stream), "mllp_gnn_plan_run")
                out = entry[2].clone()
            else:
                _cabi.check(L.mllp_gnn_pack_params(flat_c.data_ptr(), model._packed.data_ptr(), stream), "mllp_gnn_pack_params")
                out = torch.empty(g.n, dtype=torch.float32, device=dev)
                _cabi.check(L.mllp_gnn_forward(ctypes.byref(g.to_var.c), ctypes.byref(g.to_con.c), g.x1.data_ptr(), g.x2.data_ptr(),
                                               model._packed.data_ptr(), g.work.data_ptr(), out.data_ptr(), stream), "mllp_gnn_forward")
        g._forward_serial = getattr(g, "_forward_serial", 0) + 1
        ctx.g, ctx.model, ctx.serial = g, model, g._forward_serial
        ctx.save_for_backward(flat_c)
        return out

    @staticmethod
    def backward(ctx, dout):
        g, model = ctx.g, ctx.model
        (flat_c,) = ctx.saved_tensors
        if getattr(g, "_forward_serial", 0) != ctx.serial:
            raise RuntimeError("mllp_b200: another forward ran on this graph before backward(); the activations of this "
                               "forward live in the graph's workspace and are gone (call backward() first)")
        L = _cabi.lib()
        dev = flat_c.device
        stream = _torch_stream(dev)
        dout = dout.to(torch.float32).contiguous()
        with torch.cuda.device(dev):   # the library launches on the calling thread's current device
            if model.use_plans:
                entry = _plans(g, model, flat_c)
                entry[3].copy_(dout)
                _cabi.check(L.mllp_gnn_plan_run(entry[1], stream), "mllp_gnn_plan_run")
                return entry[4].clone(), None, None
            # the fused blocks of THIS forward's parameters (the model may have packed others since)
            _cabi.check(L.mllp_gnn_pack_params(flat_c.data_ptr(), model._packed.data_ptr(), stream), "mllp_gnn_pack_params")
            need = int(L.mllp_gnn_backward_workspace_floats(g.n, g.m))
            bw = getattr(g, "_bwork", None)
            if bw is None or bw.numel() < need:
                bw = g._bwork = torch.empty(need, dtype=torch.float32, device=dev)
            dflat = torch.empty_like(flat_c)
            _cabi.check(L.mllp_gnn_backward(ctypes.byref(g.to_var.c), ctypes.byref(g.to_con.c), g.x1.data_ptr(), g.x2.data_ptr(),
                                            flat_c.data_ptr(), model._packed.data_ptr(), g.work.data_ptr(), bw.data_ptr(),
                                            dout.data_ptr(), dflat.data_ptr(), stream), "mllp_gnn_backward")
            return dflat, None, None


class TrainableGNNModel(torch.nn.Module):
    """``GNNModel`` with a backward pass.  ``TrainableGNNModel().to(device)``, ``model(graph)`` with the ``BipartiteData``
    of ``build_graph_from_weights_sets`` (or a ``BipartiteGraph``), ``model.parameters()`` for the optimiser.  With
    ``use_plans`` (default) the two halves of a step are captured once per (graph, model) into CUDA graphs: the small Netlib
    graphs are launch-bound (the backward is 32 - 44 kernels)."""

    def __init__(self, state_dict=None, device=0, seed=0, use_plans=True):
        super().__init__()
        self.use_plans = bool(use_plans)   # forward / backward replayed as CUDA graphs (one launch each) or launched kernel by kernel
        layout, total = flat_layout()
        if total != int(_cabi.lib().mllp_gnn_flat_param_floats()):
            raise RuntimeError("flat parameter layout of gnn_train.py and the library disagree")
        self._layout = layout
        dev = torch.device("cuda", _device_index(device))
        self.flat = torch.nn.Parameter(torch.zeros(total, dtype=torch.float32, device=dev))
        self.register_buffer("_packed", torch.zeros(int(_cabi.lib().mllp_gnn_packed_param_floats()), dtype=torch.float32,
                                                    device=dev), persistent=False)
        self.load_state_dict(default_state(seed) if state_dict is None else state_dict)

    # the reference module's names <-> the flat vector
    def state_dict(self, *args, **kwargs):
        out = collections.OrderedDict()
        flat = self.flat.detach()
        for name, off, shape in self._layout:
            out[name] = flat[off:off + int(np.prod(shape))].reshape(shape).clone()
        return out

    def load_state_dict(self, state_dict, strict=True):
        have = set(state_dict.keys())
        want = set(n for n, _, _ in self._layout)
        if strict and (want - have):
            raise KeyError("missing keys in state_dict: %s" % sorted(want - have))
        with torch.no_grad():
            for name, off, shape in self._layout:
                if name not in state_dict:
                    continue
                v = state_dict[name]
                v = v.detach().to(torch.float32) if hasattr(v, "detach") else torch.as_tensor(np.asarray(v, dtype=np.float32))
                if tuple(v.shape) != tuple(shape):
                    raise ValueError("%s: expected shape %s, got %s" % (name, shape, tuple(v.shape)))
                self.flat[off:off + v.numel()] = v.reshape(-1).to(self.flat.device)

    def named_gradients(self):
        """{reference tensor name: gradient} views of ``flat.grad`` (None before the first backward)"""
        if self.flat.grad is None:
            return None
        return {name: self.flat.grad[off:off + int(np.prod(shape))].reshape(shape) for name, off, shape in self._layout}

    def forward(self, g):
        if not isinstance(g, BipartiteGraph):
            from .graph import BipartiteData
            if not isinstance(g, BipartiteData):   # the reference asserts the type too (:239)
                raise TypeError("GNNModel.forward expects the BipartiteData of build_graph_from_weights_sets or a BipartiteGraph")
            g = g.bipartite_graph()
        if not self.flat.is_cuda or g.device != self.flat.device.index:
            raise ValueError("graph and model live on different devices (the model needs a CUDA device: there is no CPU path)")
        return _Forward.apply(self.flat, self, g)
