"""Drop-in module surface: a module NAMED ``linear_program_methods`` that the reference's driver can star-import
(``from linear_program_methods import *``, reference linear_program_experiment.py:1) with the B200 path behind it.

What the driver takes from that import (linear_program_experiment.py:19, 50, 83, 117, 124, 168 and the test section):
``torch``, ``np``, ``set_seed``, ``InvariantModel``, ``AngleModel``, ``GNNModel``, ``get_netlib_dataloader``,
``build_graph_from_weights_sets``, ``compute_obj_differentiable`` and the ``*_max_covering`` routines.

* ON the hot path, provided here (mllp_b200, hand-written sm_100a kernels behind the C ABI, no CPU fallback):
  ``build_graph_from_weights_sets`` (same signature and ``BipartiteData``-shaped return, edge list written on the device),
  ``BipartiteData``, and the solve functions the reference lacks: ``pdhg_linear_program``, ``solve_linear_program``,
  their ``*_batch`` twins, ``DeviceLP`` / ``device_lp``, ``DeviceGNNModel`` (the message-passing forward) and
  ``TrainableGNNModel`` (the same model as a ``torch.nn.Module`` with a device backward, for the driver's training loop).
* ``set_seed`` -- same effect as the reference's (:15-24).
* Everything else (the basis-prediction models and the max-covering learners / solvers, SURVEY.md section 2 rows 5-12, OUT
  of scope) is NOT rebuilt: those names resolve lazily to the reference's own definitions when a reference checkout is
  reachable (``MLLP_REFERENCE_DIR``, default ``/root/reference``) and its third-party imports (torch_geometric,
  gumbel_sinkhorn_topk, perturbations, blackbox_diff, lap_solvers -- reference :7, :9-12) are installed; otherwise to a
  placeholder that raises ``ImportError`` naming what is missing WHEN USED, so the star-import itself always succeeds.
  ``GNNModel`` is the reference's module when that is available and the device model (``TrainableGNNModel``) otherwise.
"""
import importlib.util
import os
import random
import sys

import numpy as np
import torch

from mllp_b200.graph import BipartiteData, build_graph_from_weights_sets
from mllp_b200.gnn import GNNModel as DeviceGNNModel
from mllp_b200.gnn_train import TrainableGNNModel
from mllp_b200.linear_program_methods import (BatchLP, DeviceLP, device_lp, estimate_step_size, pdhg_linear_program,
                                              pdhg_linear_program_batch, solve_linear_program,
                                              solve_linear_program_batch)

_OURS = ["torch", "np", "set_seed", "BipartiteData", "build_graph_from_weights_sets", "DeviceGNNModel", "TrainableGNNModel", "DeviceLP", "BatchLP",
         "device_lp", "estimate_step_size", "pdhg_linear_program", "pdhg_linear_program_batch", "solve_linear_program",
         "solve_linear_program_batch"]
# names of the reference module that are outside the hot path (resolved lazily, see the module docstring)
_REFERENCE_ONLY = ["compute_objective", "compute_obj_differentiable", "has_nan", "cosine_similarity", "get_netlib_dataloader",
                   "build_graph_from_Q_sets", "InvariantModel", "AngleModel", "GNNModel", "egn_max_covering",
                   "sinkhorn_max_covering", "lml_max_covering", "gumbel_max_covering", "blackbox_max_covering",
                   "greedy_max_covering", "ortools_max_covering", "gurobi_max_covering"]
__all__ = _OURS + _REFERENCE_ONLY


def set_seed(seed: int = 42) -> None:
    """Seed every generator the driver relies on and make cuDNN deterministic (reference :15-24)."""
    for seeder in (random.seed, np.random.seed, torch.manual_seed, torch.cuda.manual_seed):
        seeder(seed)
    torch.backends.cudnn.benchmark = False
    torch.backends.cudnn.deterministic = True
    os.environ["PYTHONHASHSEED"] = str(seed)


_reference_state = {"tried": False, "module": None, "why": None}


def _reference_module():
    st = _reference_state
    if not st["tried"]:
        st["tried"] = True
        path = os.path.join(os.environ.get("MLLP_REFERENCE_DIR", "/root/reference"), "linear_program_methods.py")
        if not os.path.exists(path) or os.path.abspath(path) == os.path.abspath(__file__):
            st["why"] = "no reference checkout at %s (set MLLP_REFERENCE_DIR)" % os.path.dirname(path)
        else:
            try:
                spec = importlib.util.spec_from_file_location("_mllp_reference_linear_program_methods", path)
                mod = importlib.util.module_from_spec(spec)
                spec.loader.exec_module(mod)
                st["module"] = mod
            except Exception as e:   # a missing third-party package of the reference (torch_geometric, ...)
                st["why"] = "the reference module at %s does not import here: %s: %s" % (path, type(e).__name__, e)
    return st["module"], st["why"]


class _Unavailable:
    """Stands in for a reference-only name; usable as a function or class, raises when used."""

    def __init__(self, name, why):
        self.__name__ = name
        self._why = why

    def _fail(self, *a, **k):
        raise ImportError("%s is outside the B200 hot path and is taken from the reference's own module, which is not usable "
                          "here (%s)" % (self.__name__, self._why))

    __call__ = _fail

    def __getattr__(self, item):
        if item.startswith("__"):
            raise AttributeError(item)
        self._fail()

    def __repr__(self):
        return "<unavailable reference object %s: %s>" % (self.__name__, self._why)


def __getattr__(name):   # PEP 562: called for the lazily resolved names (also by `from ... import *`)
    if name not in _REFERENCE_ONLY:
        raise AttributeError("module %r has no attribute %r" % (__name__, name))
    mod, why = _reference_module()
    if mod is not None and hasattr(mod, name):
        obj = getattr(mod, name)
    elif name == "GNNModel":
        obj = TrainableGNNModel   # the device model: forward and backward in hand-written kernels (mllp_b200/gnn_train.py)
    else:
        obj = _Unavailable(name, why or "the reference module has no such name")
    globals()[name] = obj
    return obj
