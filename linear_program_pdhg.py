#!/usr/bin/env python
"""Runner of the solve path in the reference driver's conventions (reference linear_program_experiment.py:17-48, :123):

    python linear_program_pdhg.py --cfg linear_program_netlib.yaml [--instances afiro sc50a ...]

reads the same yaml (``--cfg``, like config.py:8-17; unknown keys accepted), loads ``get_netlib_dataset(normalize=True)``
from the working directory's ``netlib_mps/`` + ``dataset/`` exactly as the driver does, and walks
``for method_name in cfg.methods`` -- handling the ONE method name the stock driver silently skips (it has no ``else``,
:45-48, :158): ``'pdhg'``.  For every instance tuple ``(name, constrs, constr_weights, coefs, rhs, basis_opt)`` it calls
``solve_linear_program`` (B200 kernels) and logs objective / iterations / KKT error to ``pdhg_log.json``, the way the
driver logs to ``train_log.json`` (:77-78).  Objectives are also given in the MPS file's units (x ||c_raw||_2, SURVEY
App. A.3) when ``dataset/netlib_mps/<name>_coefs.npy`` is there.

``'gs-topk'`` / ``'soft-topk'`` (the supervised basis-prediction training of ``GNNModel``, :115-157) are handled too, with
the reference's loop -- ``build_graph_from_weights_sets`` -> ``model(graph)`` -> ``BCEWithLogitsLoss`` against ``basis_opt`` ->
``backward`` -> ``Adam.step`` per instance, ``train_iter`` epochs, ``train_log.json``, ``state_dict`` saved to
``linear_program_<train_data_type>_<method>.pt`` (:46, :185) -- on the device model (forward AND backward in hand-written
kernels, mllp_b200/gnn_train.py); the stock driver cannot run it without torch_geometric.

Extra yaml keys (all optional): ``pdhg_tol`` (1e-6), ``pdhg_max_iters`` (400000), ``pdhg_check_every`` (64),
``pdhg_instances`` (list of names; default: all listed instances), ``pdhg_scale`` (true: Ruiz + Pock-Chambolle preconditioning).
"""
import argparse
import json
import os
import sys

import numpy as np
import yaml

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


class Config(dict):
    """yaml mapping with attribute access (what the reference gets from EasyDict, config.py:19-27)."""
    __getattr__ = dict.get

    @classmethod
    def load(cls, path):
        with open(path, "r") as f:
            raw = yaml.full_load(f) or {}
        conv = lambda v: cls({k: conv(x) for k, x in v.items()}) if isinstance(v, dict) else v
        return conv(raw)


def train_gnn(cfg, method_name, train_dataset, train_dict, device):
    """the reference's training branch for 'gs-topk' / 'soft-topk' (linear_program_experiment.py:115-157, :185-186)"""
    import torch
    from sklearn.metrics import f1_score
    from linear_program_methods import TrainableGNNModel, build_graph_from_weights_sets
    print(f"Training the model weights for {method_name}...")
    dev = torch.device(device)
    model = TrainableGNNModel(device=dev.index or 0)
    criterion = torch.nn.BCEWithLogitsLoss()
    train_optimizer = torch.optim.Adam(model.parameters(), lr=float(cfg.train_lr))
    graphs = {}
    for epoch in range(int(cfg.train_iter)):
        obj_sum = 0
        for name, constrs, constr_weights, coefs, rhs, basis_opt in train_dataset:
            if name not in graphs:   # the reference rebuilds the graph every epoch (:124); the device graph is kept
                graphs[name] = build_graph_from_weights_sets(constrs, constr_weights, rhs, coefs, dev)
            latent_vars = model(graphs[name])
            target = torch.tensor(basis_opt, dtype=torch.float, device=dev)
            obj = criterion(latent_vars, target)
            obj.backward()
            obj_sum += obj.mean()
            train_optimizer.step()
            train_optimizer.zero_grad()
            pred_indices = torch.topk(latent_vars, k=rhs.shape[0])[-1].cpu().detach().numpy()
            pred = np.zeros([coefs.shape[0]])
            pred[pred_indices] = 1
            f1 = f1_score(np.asarray(basis_opt), pred)
            correct_num = pred @ np.asarray(basis_opt)
            print("%8d, %8d, %8d, %5f" % (correct_num, rhs.shape[0], coefs.shape[0], f1))
            train_dict[name].append(float(correct_num))
        train_dict["obj"].append(float(obj_sum) / len(train_dataset))
        with open("train_log.json", "w") as json_file:
            json.dump(train_dict, json_file)
        print(f"epoch {epoch}, obj={obj_sum / len(train_dataset)}")
    model_path = f"linear_program_{cfg.train_data_type}_{method_name}.pt"
    torch.save(model.state_dict(), model_path)
    print(f"Model saved to {model_path}.")
    return model


def main(argv=None):
    ap = argparse.ArgumentParser(description="PDHG solve runner (B200)")
    ap.add_argument("--cfg", "--config", dest="cfg_file", default=None)
    ap.add_argument("--instances", nargs="*", default=None)
    ap.add_argument("--device", default="cuda:0")
    args = ap.parse_args(argv)
    if args.cfg_file is None:
        raise ValueError("Please specify path to the configuration file!")      # as config.py:11-12
    cfg = Config.load(args.cfg_file)
    if cfg.train_data_type != "netlib":
        raise ValueError(f"Unknown training dataset {cfg.train_data_type}!")    # as linear_program_experiment.py:38-39

    from linear_program_data import get_netlib_dataset
    from linear_program_methods import solve_linear_program, set_seed
    set_seed()
    names = args.instances if args.instances is not None else cfg.pdhg_instances
    train_dataset, train_dict = get_netlib_dataset(normalize=True, names=names)
    log = {"obj": []}
    for method_name in cfg.methods or []:
        if method_name in ("gs-topk", "soft-topk"):
            train_gnn(cfg, method_name, train_dataset, train_dict, args.device)
            continue
        if method_name != "pdhg":
            continue    # the other stock methods belong to the stock driver
        print("Solving the LPs with pdhg...")
        for name, constrs, constr_weights, coefs, rhs, basis_opt in train_dataset:
            scale = bool(cfg.pdhg_scale) if cfg.pdhg_scale is not None else True
            kw = dict(tol=float(cfg.pdhg_tol or 1e-6), max_iters=int(cfg.pdhg_max_iters or 400000),
                      check_every=int(cfg.pdhg_check_every or 64), device=args.device)
            if scale:
                import scipy.sparse as sp
                from mllp_b200.linear_program_methods import csr_from_constrs
                from mllp_b200.scaling import solve_scaled
                ip, ii, vv = csr_from_constrs(constrs, constr_weights, len(coefs))
                A = sp.csr_matrix((vv, ii, ip), shape=(len(rhs), len(coefs)))
                obj, x, y, info = solve_scaled(A, rhs, coefs, **kw)
            else:
                obj, x, y, info = solve_linear_program(constrs, constr_weights, rhs, coefs, **kw)
            raw = os.path.join("dataset", "netlib_mps", name + "_coefs.npy")
            unit = float(np.linalg.norm(np.load(raw))) if os.path.exists(raw) else None
            rec = {"objective": float(obj), "objective_netlib_units": None if unit is None else float(obj) * unit,
                   "iters": int(info["iters"]), "converged": bool(info["converged"]),
                   "rel_kkt": float(info.get("rel_kkt_original", info["rel_kkt"]))}
            if basis_opt is not None and len(basis_opt) == len(x):
                # how much of the solution's support the reference's optimal-basis labels cover (its learning target, :137-143)
                k = int(np.asarray(basis_opt).sum())
                pred = np.zeros(len(x)); pred[np.argsort(-np.abs(x))[:k]] = 1
                rec["support_in_basis"] = float(pred @ basis_opt) / max(k, 1)
            log[name] = rec
            log["obj"].append(rec["objective"])
            print("%-16s obj %.9g%s  iters %d  rel_kkt %.2e  %s" % (
                name, rec["objective"], "" if unit is None else " (%.9g in file units)" % rec["objective_netlib_units"],
                rec["iters"], rec["rel_kkt"], "converged" if rec["converged"] else "NOT converged"))
            with open("pdhg_log.json", "w") as json_file:
                json.dump(log, json_file)
    return log


if __name__ == "__main__":
    main()
