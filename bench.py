#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200 primal-dual LP path.

Metric (BASELINE.json): PDHG iterations/s (+ fraction of the HBM roofline) on a large
Netlib instance.  N=1 workload: osa-60 (`_norm` arrays, 10280 x 243246, nnz 1408073), parity
mode, fp64.  A "step" is ONE solve call of `--iters-per-step` fused iterations (one persistent
kernel launch).  N>1: one process per GPU, each rank iterates its own perturbed copy of the
instance (data-parallel over independent LPs, no data-path collective) -> weak scaling,
value = total iterations/s over all ranks.

  value : inputs resident in HBM, timed with CUDA events around each step on the launch stream
  e2e   : the public Python call pdhg_linear_program() with HOST (pinned) buffers: per step
          H2D of b, c, x0, y0 and D2H of x, y and the KKT scalars, wall clock around the calls
  --impl reference : the CPU oracle (the only "reference implementation" of this path that
          exists) on all host cores, same metric and config.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="osa-60")
    ap.add_argument("--iters-per-step", type=int, default=1000)
    ap.add_argument("--cpu-baseline-seconds", type=float, default=12.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the secondary measurements (ken-18, batches)")
    return ap.parse_args()


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def measured_traffic(workload, iters_per_step, kernel):
    """DRAM bytes per launch of the dominant kernel from the committed ncu --set full capture of THE SAME launch
    (profiles/r02_*_traffic.json, written by scripts/summarize_kernel_profile.py, which refuses a capture whose duration
    does not match the CUDA-event time of the benchmarked launch); None unless workload, launch size and kernel match."""
    for f in ("r02_persistent_traffic.json", "r02_blocks_traffic.json"):
        try:
            d = json.load(open(os.path.join(ROOT, "profiles", f)))
            if d["workload"] == workload and int(d["iters_per_step"]) == int(iters_per_step) and str(d["kernel"]) in kernel:
                return int(d["dram_bytes_per_launch"])
        except Exception:
            pass
    return None


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.lines, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for nm, v in zip(names, f[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def perturbed(b, c, rank):
    """rank 0 keeps the Netlib data; other ranks get the ML-data style perturbation of
    BASELINE.json configs[4] (c scaled by 1 +- 10 %, b by 1 + U(0, 10 %))."""
    if rank == 0:
        return b, c
    rng = np.random.default_rng(1234 + rank)
    return b * (1.0 + 0.1 * rng.random(b.shape[0])), c * (1.0 + 0.1 * rng.uniform(-1, 1, c.shape[0]))


def host_threads():
    """all host threads this process may use (torchrun exports OMP_NUM_THREADS=1, so ask the OS)"""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


def scipy_csr_rate(A, b, c, eta, seconds=3.0):
    """iterations/s of the same iteration written with SciPy CSR products on ONE host thread ('also report a
    1-thread SciPy CSR figure'); bounded sample."""
    import scipy.sparse as sp
    A = sp.csr_matrix(A)
    At = A.T.tocsr()
    m, n = A.shape
    x, y = np.zeros(n), np.zeros(m)

    def it():
        nonlocal x, y
        xn = np.maximum(x - eta * (c - At @ y), 0.0)
        xbar = 2.0 * xn - x
        x = xn
        y = y + eta * (b - A @ xbar)
    it()
    t0 = time.perf_counter()
    k = 0
    while time.perf_counter() - t0 < seconds:
        it()
        k += 1
    return k / (time.perf_counter() - t0), k


def slack_rows(A, c):
    """rows of a `_norm` LP that own a slack column (slack columns are appended, one nonzero, cost 0)"""
    Ac = A.tocsc()
    rows = []
    for j in range(A.shape[1] - 1, -1, -1):
        if Ac.indptr[j + 1] - Ac.indptr[j] != 1 or c[j] != 0.0:
            break
        rows.append(int(Ac.indices[Ac.indptr[j]]))
    return np.array(sorted(rows), dtype=np.int64)


def config5_batches(A, b, c, lo, hi, variant="bounded"):
    """BASELINE.json configs[4] / SURVEY 8d config 5: instance i of the 4096 perturbs the fixed matrix's data with
    torch.Generator().manual_seed(1234 + i) on the CPU in fp64; b_i = b (1 + 0.1 U(0, 1)) on the rows that own a slack
    column, equality rows unchanged.  The cost perturbation:
      variant "survey"  : c_i = c (1 + 0.1 U(-1, 1)), as SURVEY 8d writes it.  On 25fv47 (`_norm` form: bounds dropped, x >= 0)
                          this makes the LP UNBOUNDED for about nine instances in ten (HiGHS: status 3; the unperturbed LP has
                          recession directions with c'd = 0, any sign change of c'd opens them) -- there is nothing to solve;
      variant "bounded" : c_ij = c_j (1 + 0.1 u) for c_j > 0 and c_j (1 - 0.1 u) for c_j < 0, u ~ U(0, 1) from the same stream:
                          c'd can only grow on every direction d >= 0, so every instance stays bounded (and, with the same b
                          rule, feasible: checked with HiGHS in tests/test_oracle.py).  This is what the bench solves."""
    import torch
    m, n = A.shape
    sr = slack_rows(A, c)
    cb, bb = np.empty((hi - lo, n)), np.tile(b, (hi - lo, 1))
    for k, i in enumerate(range(lo, hi)):
        g = torch.Generator().manual_seed(1234 + i)
        u = torch.rand(n, generator=g, dtype=torch.float64).numpy()
        if variant == "survey":
            cb[k] = c * (1.0 + 0.1 * (2.0 * u - 1.0))
        else:
            cb[k] = np.where(c > 0, c * (1.0 + 0.1 * u), c * (1.0 - 0.1 * u))
        bb[k, sr] = b[sr] * (1.0 + 0.1 * torch.rand(sr.shape[0], generator=g, dtype=torch.float64).numpy())
    return bb, cb


def cpu_oracle_rate(A, b, c, eta, seconds, min_iters=5):
    """iterations/s of the CPU oracle (all host threads) on a bounded sample."""
    from oracle import pdhg_oracle as O
    csr = O.CSR(A)
    m, n = A.shape
    nt = host_threads()
    O.pdhg_run(csr, b, c, np.zeros(n), np.zeros(m), eta, eta, min_iters, nthreads=nt)  # warm-up (thread pool, page faults)
    t0 = time.perf_counter()
    O.pdhg_run(csr, b, c, np.zeros(n), np.zeros(m), eta, eta, 4 * min_iters, nthreads=nt)
    per_it = max((time.perf_counter() - t0) / (4 * min_iters), 1e-7)
    iters = int(max(min_iters, min(20000, seconds / per_it)))
    t0 = time.perf_counter()
    O.pdhg_run(csr, b, c, np.zeros(n), np.zeros(m), eta, eta, iters, nthreads=nt)
    dt = time.perf_counter() - t0
    return iters / dt, iters, dt, nt


def measure_extras(M, torch, dev, local, rank, world, dist):
    """Secondary numbers of BASELINE.json's metric, outside the headline's timed region:
    ken-18 iterations/s, batched LP-iterations/s (4096 perturbed instances of one Netlib matrix per
    rank) and LPs solved/s (solve mode to 1e-6 on perturbed small Netlib instances).  Whole-job sums."""
    from mllp_b200 import _cabi
    L = _cabi.lib()
    out = {}

    def timed(fn, reps=3):
        fn(); torch.cuda.synchronize(dev)
        best = 1e30
        for _ in range(reps):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); fn(); e1.record(); torch.cuda.synchronize(dev)
            best = min(best, e0.elapsed_time(e1))
        t = torch.tensor([best], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t[0]) * 1e-3

    from oracle import pdhg_oracle as O   # the checker of the in-run parity asserts below (never the thing measured)
    rel = lambda a, r: float(np.linalg.norm(a - r) / max(np.linalg.norm(r), 1e-300))
    hbm, _ = peaks()
    sp = torch.cuda.current_stream(dev).cuda_stream

    # (1) the other large config instances, parity mode, 1000 fused iterations per launch; every record carries what
    # mllp_lp_create chose by measurement (geometry, block-angular kernel), its cost, and an in-run parity check against
    # the oracle (K = 100, 1e-9) on the very handle that is timed
    for name in ("ken-18", "pds-20"):
        A, b, c = M.load_csr(name)
        m, n = A.shape
        lp = M.DeviceLP(A, A.data, m, n, device=local)
        eta = 0.9 / lp.sigma_max()
        _, xk, yk, _ = M.pdhg_linear_program(A, A.data, b, c, num_iters=100, tau=eta, sigma=eta, handle=lp)
        xo, yo = O.pdhg_run(O.CSR(A), b, c, np.zeros(n), np.zeros(m), eta, eta, 100, nthreads=host_threads())
        ex, ey = rel(xk, xo), rel(yk, yo)
        parity_check(ex < 1e-9 and ey < 1e-9, "%s: iterates differ from the oracle (%.2e, %.2e)" % (name, ex, ey))
        bt, ct = torch.tensor(b, device=dev), torch.tensor(c, device=dev)
        xt, yt = torch.zeros(n, dtype=torch.float64, device=dev), torch.zeros(m, dtype=torch.float64, device=dev)
        sec = timed(lambda: _cabi.check(L.mllp_pdhg_run(lp.handle, xt.data_ptr(), yt.data_ptr(), bt.data_ptr(), ct.data_ptr(),
                                                        eta, eta, 1000, None, sp), "run"))
        bpi = lp.info()["bytes_per_iter"]
        blk, geo = lp.blocks_info(), lp.geometry()
        out[name] = {"iterations_per_sec": world * 1000 / sec, "us_per_iteration": sec * 1e3,
                     "roofline_frac_of_measured_hbm": bpi * 1000 / sec / 1e9 / hbm, "bytes_per_iter": bpi,
                     "kernel": "k_pdhg_blocks" if blk["used"] else "k_pdhg_persistent", "geometry": geo["mode"], "ctas": geo["ctas"],
                     "block_angular": {"blocks": blk["blocks"], "linking_rows": blk["linking_rows"]} if blk["used"] else None,
                     "create_s": lp.create_s, "parity_vs_oracle_K100": {"x": ex, "y": ey, "tol": 1e-9}}
        lp.close()

    # (2) shared-matrix batch: 4096 perturbed (b, c) instances of 25fv47 per rank, parity mode
    A, b, c = M.load_csr("25fv47")
    m, n = A.shape
    B = 4096
    bs = M.BatchLP([(A, A.data, b, c)], shared=True, count=B, device=local)
    g = torch.Generator(device="cpu").manual_seed(1234 + rank)
    cb = torch.tensor(c).repeat(B, 1) * (1 + 0.1 * (2 * torch.rand(B, n, generator=g, dtype=torch.float64) - 1))
    bb = torch.tensor(b).repeat(B, 1) * (1 + 0.1 * torch.rand(B, m, generator=g, dtype=torch.float64))
    cb, bb = cb.reshape(-1).to(dev), bb.reshape(-1).to(dev)
    xb, yb = torch.zeros(B * n, dtype=torch.float64, device=dev), torch.zeros(B * m, dtype=torch.float64, device=dev)
    etab = (0.9 / bs.sigma_max()).contiguous()
    K = 200
    # in-run parity: instances 0, 1 and B - 1 of this very batch against the oracle (K = 50)
    xb.zero_(); yb.zero_()
    bs.run(xb, yb, bb, cb, etab, etab, 50)
    eh = etab.cpu().numpy()
    perr = 0.0
    for i in (0, 1, B - 1):
        xo, yo = O.pdhg_run(O.CSR(A), bb[i * m:(i + 1) * m].cpu().numpy(), cb[i * n:(i + 1) * n].cpu().numpy(), np.zeros(n), np.zeros(m),
                            float(eh[i]), float(eh[i]), 50)
        perr = max(perr, rel(xb[i * n:(i + 1) * n].cpu().numpy(), xo), rel(yb[i * m:(i + 1) * m].cpu().numpy(), yo))
    parity_check(perr < 1e-9, "batch iterates differ from the oracle (%.2e)" % perr)
    xb.zero_(); yb.zero_()
    sec = timed(lambda: bs.run(xb, yb, bb, cb, etab, etab, K), reps=2)
    out["batch_4096x25fv47"] = {"lp_iterations_per_sec": world * B * K / sec, "instances_per_rank": B,
                                "us_per_batch_iteration": sec / K * 1e6, "algorithmic_GBps_per_gpu": bs.info()["bytes_per_iter"] * K / sec / 1e9,
                                "roofline_frac_of_measured_hbm": bs.info()["bytes_per_iter"] * K / sec / 1e9 / hbm,
                                "instances_per_cta": bs.info()["instances_per_cta"], "kernel": "k_batch_run_r<%d>" % bs.info()["instances_per_cta"],
                                "parity_vs_oracle_K50": {"max_rel": perr, "instances": [0, 1, B - 1], "tol": 1e-9}}
    bs.close()
    del cb, bb, xb, yb

    # (2b) shared-matrix batch solved to 1e-6: 4096 cost-perturbed instances of sc105 per rank (two instances per CTA
    # share every matrix step; a slot that converges pulls the next instance at once)
    A, b, c = M.load_csr("sc105")
    m, n = A.shape
    bs = M.BatchLP([(A, A.data, b, c)], shared=True, count=B, device=local)
    g = torch.Generator(device="cpu").manual_seed(4321 + rank)
    cb = (torch.tensor(c).repeat(B, 1) * (1 + 0.05 * (2 * torch.rand(B, n, generator=g, dtype=torch.float64) - 1))).reshape(-1).to(dev)
    bb = torch.tensor(b).repeat(B).to(dev)
    etas = (0.99 / bs.sigma_max_robust()).contiguous()
    scal = torch.zeros(B * _cabi.NUM_SCALARS, dtype=torch.float64, device=dev)

    def solve_shared():
        xs, ys = torch.zeros(B * n, dtype=torch.float64, device=dev), torch.zeros(B * m, dtype=torch.float64, device=dev)
        bs.solve(xs, ys, bb, cb, etas, scal, 1.0, 200000, 64, 1e-6)

    sec = timed(solve_shared, reps=2)
    sc = scal.cpu().numpy().reshape(B, _cabi.NUM_SCALARS)
    out["solve_4096x_sc105_shared"] = {"lps_solved_per_sec": world * B / sec, "instances_per_rank": B,
                                       "converged_fraction": float(sc[:, 12].mean()), "mean_iterations": float(sc[:, 10].mean()),
                                       "instances_per_cta": bs.info()["instances_per_cta_solve"],
                                       "kernel": "k_batch_solve_warp (one warp per instance)" if bs.info()["instances_per_cta_solve"] == 4 and bs.info()["res_steps_A_solve"] == 0 else "k_batch_solve_r<%d>" % bs.info()["instances_per_cta_solve"],
                                       "tol": 1e-6, "seconds": sec}
    bs.close()
    del cb, bb

    # (3) LPs solved per second: 1024 perturbed copies each of sc50a / sc105 / blend per rank, solve mode to 1e-6
    insts = []
    for nm in ("sc50a", "sc105", "blend"):
        A, b, c = M.load_csr(nm)
        for i in range(1024):
            rg = np.random.default_rng(99991 * rank + i)
            insts.append((A, A.data, b, c * (1 + 0.05 * rg.uniform(-1, 1, c.shape[0]))))
    bsol = M.BatchLP(insts, device=local)
    bvec = torch.tensor(np.concatenate([i[2] for i in insts]), device=dev)
    cvec = torch.tensor(np.concatenate([i[3] for i in insts]), device=dev)
    nx, ny = int(bsol.x_off[-1]), int(bsol.y_off[-1])
    etas = (0.99 / bsol.sigma_max_robust()).contiguous()
    scal = torch.zeros(len(insts) * _cabi.NUM_SCALARS, dtype=torch.float64, device=dev)

    def solve_all():
        xs, ys = torch.zeros(nx, dtype=torch.float64, device=dev), torch.zeros(ny, dtype=torch.float64, device=dev)
        bsol.solve(xs, ys, bvec, cvec, etas, scal, 1.0, 100000, 64, 1e-6)

    sec = timed(solve_all, reps=2)
    sc = scal.cpu().numpy().reshape(len(insts), _cabi.NUM_SCALARS)
    out["solve_3072_small_netlib"] = {"lps_solved_per_sec": world * len(insts) / sec, "instances_per_rank": len(insts),
                                      "converged_fraction": float(sc[:, 12].mean()), "mean_iterations": float(sc[:, 10].mean()),
                                      "tol": 1e-6, "max_iters": 100000, "seconds": sec}
    bsol.close()

    # (3b) the same 3072 LPs with every distinct matrix preconditioned BY THE LIBRARY (mllp_batch_create with
    # MLLP_F_PRECONDITION: Ruiz + Pock-Chambolle on the device, outside the timed region like the format build); the kernel
    # evaluates the KKT error of the ORIGINAL LPs and terminates on it, so the scalars below are those of the original LPs
    bscl = M.BatchLP(insts, device=local, precondition=True)
    etas = (0.99 / bscl.sigma_max_robust()).contiguous()
    xs = torch.zeros(nx, dtype=torch.float64, device=dev)
    ys = torch.zeros(ny, dtype=torch.float64, device=dev)

    def solve_scaled():
        xs.zero_(); ys.zero_()
        bscl.solve(xs, ys, bvec, cvec, etas, scal, 1.0, 100000, 64, 1e-6)

    sec = timed(solve_scaled, reps=2)
    sc = scal.cpu().numpy().reshape(len(insts), _cabi.NUM_SCALARS)
    conv = sc[:, 12] > 0
    bscl.close()
    # the checker's KKT error of three returned points, on the original LPs
    xh, yh = xs.cpu().numpy(), ys.cpu().numpy()
    kerr = 0.0
    for i in (0, 1024, 2048):
        Ai, _, bi, ci = insts[i]
        kk = O.kkt(O.CSR(Ai), bi, ci, xh[bsol.x_off[i]:bsol.x_off[i + 1]], yh[bsol.y_off[i]:bsol.y_off[i + 1]])
        kerr = max(kerr, abs(kk[8] - sc[i, 8]))
    parity_check(kerr <= 1e-9, "KKT scalars of the preconditioned batch differ from the oracle's on the original LPs (%.2e)" % kerr)
    out["solve_3072_small_netlib_preconditioned"] = {
        "lps_solved_per_sec": world * len(insts) / sec, "instances_per_rank": len(insts), "converged_fraction": float(conv.mean()),
        "mean_iterations": float(sc[:, 10].mean()), "tol": 1e-6, "max_iters": 100000, "seconds": sec,
        "rel_kkt_original_median": float(np.median(sc[:, 8])),
        "rel_kkt_original_max_over_converged": float(sc[conv, 8].max()) if conv.any() else None,
        "rel_kkt_original_max": float(sc[:, 8].max()),
        "kkt_scalars_vs_oracle_on_original_lps": kerr}

    # (4) BASELINE.json configs[4] as specified: 4096 perturbed instances of 25fv47 IN TOTAL, sharded data-parallel over the
    # ranks (4096 / world each; 512 per GPU on 8), solve mode to 1e-6 on the original LPs (library preconditioner),
    # through the public call with HOST buffers (solve_batch_data_parallel: H2D of every rank's b / c batches, D2H of x / y /
    # scalars, gather of the results): LPs solved per second end to end -- strong scaling over the ranks
    from mllp_b200.distributed import shard_range, solve_batch_data_parallel
    A, b, c = M.load_csr("25fv47")
    m, n = A.shape
    lo, hi = shard_range(4096, rank, world)
    bb5, cb5 = config5_batches(A, b, c, lo, hi)
    torch.cuda.synchronize(dev)
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    res = solve_batch_data_parallel([(A, A.data, b, c)], mode="solve", device=local, shared=True, rhs_batch=bb5, coefs_batch=cb5,
                                    count=4096, tol=1e-6, max_iters=200000, scale=True, single_process=(world == 1))
    wall = time.perf_counter() - t0
    tw = torch.tensor([wall], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tw, op=dist.ReduceOp.MAX)
    wall = float(tw[0])
    conv = np.array([r[3]["converged"] for r in res])
    its = np.array([r[3]["iters"] for r in res])
    kk5 = np.array([r[3]["rel_kkt"] for r in res])
    out["config5_4096x25fv47_solve_dp"] = {
        "instances_total": 4096, "instances_per_rank": hi - lo, "lps_solved_per_sec_e2e": 4096 / wall, "seconds_e2e": wall,
        "converged_fraction": float(conv.mean()), "mean_iterations": float(its.mean()), "max_iterations": int(its.max()),
        "rel_kkt_original_max_over_converged": float(kk5[conv].max()) if conv.any() else None, "tol": 1e-6,
        "h2d_bytes": 8 * (hi - lo) * (m + n) + 8 * (hi - lo) * (m + n), "d2h_bytes": 8 * (hi - lo) * (m + n + _cabi.NUM_SCALARS),
        "scaling": "strong (4096 instances in total over %d GPU(s))" % world,
        "perturbation": "seeds 1234+i, b (1 + 0.1 U(0,1)) on slack rows; costs moved AWAY from zero-cost recession directions (c_j > 0: x(1 + 0.1u), c_j < 0: x(1 - 0.1u)): SURVEY 8d's c (1 + 0.1 U(-1,1)) makes ~90 % of the 25fv47 instances unbounded (HiGHS), see config5_batches",
        "call": "mllp_b200.distributed.solve_batch_data_parallel(shared matrix, rhs_batch, coefs_batch, scale=True) with numpy batches; includes format build + device preconditioning"}
    if world == 1:
        out["solve_mode_large"] = measure_solve_large(M, torch, dev, local)
        out["gnn_training_step"] = measure_gnn_training(torch, dev)
    return out


def measure_solve_large(M, torch, dev, local, names=(("osa-60", 3.224480553), ("pds-20", 21985.28701))):
    """Solve mode (k_solve_persistent: reflected restarted Halpern PDHG, KKT test of the ORIGINAL LP on the device) on the
    large config-4 instances that are bounded in the dataset form: time per iteration of the solve kernel and the objective
    against HiGHS on the same arrays."""
    out = {}
    for name, target in names:
        A, b, c = M.load_csr(name)
        m, n = A.shape
        lp = M.DeviceLP(A, A.data, m, n, device=local, precondition=True)
        eta = 0.99 / lp.sigma_max_robust()
        bt, ct = torch.tensor(b, device=dev), torch.tensor(c, device=dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(dev)
        e0.record()
        _, xt, yt, inf = M.solve_linear_program(A, A.data, bt, ct, tol=1e-6, max_iters=400000, check_every=64, handle=lp, eta=eta)
        e1.record()
        torch.cuda.synchronize(dev)
        sec = e0.elapsed_time(e1) * 1e-3
        from mllp_b200.linear_program_methods import _info_dict
        info = _info_dict(inf["scalars"].cpu().numpy())
        obj = float(inf["scalars"][0])
        geo = lp.geometry()
        out[name] = {"iterations": int(info["iters"]), "converged": bool(info["converged"]), "rel_kkt_original": float(info["rel_kkt"]),
                     "objective": float(obj), "highs_on_the_same_arrays": target, "rel_error": abs(float(obj) - target) / (1 + abs(target)),
                     "seconds_on_device": sec, "us_per_iteration": 1e6 * sec / max(int(info["iters"]), 1),
                     "restarts": int(info.get("restarts", -1)), "geometry": geo["mode"], "create_s": lp.create_s,
                     "kernel": "k_solve_persistent" if geo["mode"] == "grid" else "k_solve_cluster"}
        parity_check(bool(info["converged"]) and out[name]["rel_error"] <= 1e-5,
                     "%s: solve mode did not reach the HiGHS objective (%r vs %r)" % (name, float(obj), target))
        lp.close()
    return out


def measure_gnn_training(torch, dev, name="ken-18"):
    """One training step of the reference's GNNModel on the device kernels (mllp_gnn_backward; the reference trains it,
    linear_program_experiment.py:115-157): times on `name`, and an in-run check of the device gradients against
    torch.autograd through the plain-PyTorch float64 checker on afiro."""
    import mllp_b200.gnn as GN
    import mllp_b200.linear_program_data as D
    from mllp_b200.gnn_train import TrainableGNNModel
    from oracle import gnn_numpy as GO, gnn_torch as GT   # checker only
    A, b, c = D.load_csr("afiro")
    st = GO.init_state(5)
    dout = np.random.default_rng(2).standard_normal(A.shape[1])
    _, _, ref = GT.torch_model_loss_and_grads(st, A, b, c, dout=dout)
    g = GN.BipartiteGraph(np.split(A.indices, A.indptr)[1:-1], A.data, b, c)
    model = TrainableGNNModel(st)
    (model(g) * torch.as_tensor(dout.astype(np.float32), device=dev)).sum().backward()
    got = {k: v.detach().cpu().numpy() for k, v in model.named_gradients().items()}
    top = max(np.abs(v).max() for v in ref.values())
    gerr = max(float(np.abs(got[k] - ref[k]).max() / max(np.abs(ref[k]).max(), 1e-4 * top)) for k in ref)
    parity_check(gerr < 2e-3, "device GNN gradients differ from autograd (%.2e)" % gerr)
    g.close()

    A, b, c = D.load_csr(name)
    n = A.shape[1]
    g = GN.BipartiteGraph(np.split(A.indices, A.indptr)[1:-1], A.data, b, c)
    model = TrainableGNNModel(seed=1)
    opt = torch.optim.Adam(model.parameters(), lr=1e-3)
    crit = torch.nn.BCEWithLogitsLoss()
    target = torch.as_tensor((np.random.default_rng(0).random(n) < 0.4).astype(np.float32), device=dev)

    def step():
        loss = crit(model(g), target)
        loss.backward()
        opt.step()
        opt.zero_grad()

    def timed(fn, reps=10):
        ts = []
        for _ in range(reps):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); fn(); e1.record(); torch.cuda.synchronize(dev)
            ts.append(e0.elapsed_time(e1) * 1e3)
        return float(np.median(ts))

    for _ in range(3):
        step()
    torch.cuda.synchronize(dev)
    t_step = timed(step)
    with torch.no_grad():
        t_fwd = timed(lambda: model(g))
    g.close()
    return {"instance": name, "us_forward": t_fwd, "us_training_step": t_step,
            "what": "forward + BCEWithLogitsLoss + backward (mllp_gnn_backward, one CUDA-graph launch) + Adam step",
            "gradients_vs_autograd_afiro": {"max_rel": gerr, "tol": 2e-3}}


def measure_rowpart(M, torch, dist, dev, local, rank, world, names=("osa-60", "ken-18", "pds-20"), K=1000, parity_K=100,
                    nccl_K=200, trace=False):
    """BASELINE.json configs[3]: ONE large LP row-partitioned over all `world` GPUs (A by rows, A' phase replicated, one
    exchange of the y slices per iteration).  Per instance: in-run parity against the CPU oracle (K=100, 1e-9, asserted
    on every rank), us/iteration of the in-kernel NVLink exchange and of the NCCL all-gather variant (device time, max
    over ranks), and the same run's one-GPU time of the same instance (best rank) for the speed-up."""
    from mllp_b200 import _cabi
    from mllp_b200.distributed import RowPartLP, pdhg_linear_program_rowpart
    from oracle import pdhg_oracle as O   # the checker of the in-run parity assert
    L = _cabi.lib()
    sp = torch.cuda.current_stream(dev).cuda_stream
    out = {}

    def timed(fn, reps=3):
        fn(); torch.cuda.synchronize(dev); dist.barrier()
        best = 1e30
        for _ in range(reps):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            dist.barrier()
            e0.record(); fn(); e1.record(); torch.cuda.synchronize(dev)
            t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            best = min(best, float(t[0]))
        return best * 1e-3

    for name in names:
        A, b, c = M.load_csr(name)
        m, n = A.shape
        bt, ct = torch.tensor(b, device=dev), torch.tensor(c, device=dev)
        xt, yt = torch.zeros(n, dtype=torch.float64, device=dev), torch.zeros(m, dtype=torch.float64, device=dev)

        def run_on(handle, iters):
            xt.zero_(); yt.zero_()
            _cabi.check(L.mllp_pdhg_run(handle, xt.data_ptr(), yt.data_ptr(), bt.data_ptr(), ct.data_ptr(), eta, eta, iters,
                                        None, sp), "mllp_pdhg_run")

        one = M.DeviceLP(A, A.data, m, n, device=local)
        eta = 0.9 / one.sigma_max()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        run_on(one.handle, K); torch.cuda.synchronize(dev)
        e0.record(); run_on(one.handle, K); e1.record(); torch.cuda.synchronize(dev)
        t1 = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
        dist.all_reduce(t1, op=dist.ReduceOp.MIN)
        us_one = float(t1[0]) * 1e3 / K
        one_kernel = "k_pdhg_blocks" if one.blocks_info()["used"] else "k_pdhg_persistent (%s)" % one.geometry()["mode"]
        one.close()

        lp = RowPartLP(A, A.data, m, n, device=local)
        # in-run parity: K=100 against the oracle (computed on rank 0, broadcast), asserted on every rank
        obj, x, y, info = pdhg_linear_program_rowpart(lp, b, c, num_iters=parity_K, tau=eta, sigma=eta)
        ref = torch.zeros(n + m, dtype=torch.float64, device=dev)
        if rank == 0:
            xo, yo = O.pdhg_run(O.CSR(A), b, c, np.zeros(n), np.zeros(m), eta, eta, parity_K, nthreads=host_threads())
            ref.copy_(torch.tensor(np.concatenate([xo, yo])))
        dist.broadcast(ref, src=0)
        refh = ref.cpu().numpy()
        ex = float(np.linalg.norm(x - refh[:n]) / max(np.linalg.norm(refh[:n]), 1e-300))
        ey = float(np.linalg.norm(y - refh[n:]) / max(np.linalg.norm(refh[n:]), 1e-300))
        parity_check(ex < 1e-9 and ey < 1e-9, "row-partitioned %s on rank %d: iterates differ from the oracle (%.2e, %.2e)" % (name, rank, ex, ey))
        perr = torch.tensor([ex, ey], dtype=torch.float64, device=dev)
        dist.all_reduce(perr, op=dist.ReduceOp.MAX)

        sec = timed(lambda: run_on(lp.handle, K))
        os.environ["MLLP_ROWPART_NCCL"] = "1"
        try:
            sec_nccl = timed(lambda: run_on(lp.handle, nccl_K), reps=2)
        finally:
            os.environ.pop("MLLP_ROWPART_NCCL", None)
        parity_check(lp.exchange_error() == 0, "row-partitioned %s on rank %d: an exchange wait timed out" % (name, rank))
        rec = {"n_gpus": world, "us_per_iteration": sec * 1e6 / K, "iterations_per_sec": K / sec,
               "us_per_iteration_nccl_allgather": sec_nccl * 1e6 / nccl_K,
               "us_per_iteration_one_gpu": us_one, "one_gpu_kernel": one_kernel, "speedup_vs_one_gpu": us_one / (sec * 1e6 / K),
               "iters_timed": K, "parity_vs_oracle_K%d" % parity_K: {"x": float(perr[0]), "y": float(perr[1]), "tol": 1e-9},
               "multicast": bool(lp.multicast),
               "exchange": ("tagged 16-byte words, ONE multimem.st per dual replicated by NVSwitch multicast into every rank's mailbox" if lp.multicast else
                            "tagged 16-byte words stored into every peer's mailbox over NVLink (world - 1 stores per dual)") +
                           ", inside one cooperative launch; A' phase replicated, one exchange per iteration"}
        if trace:
            rec["timeline_us"] = rowpart_timeline(lp, eta, dist, torch, dev)
        out[name] = rec
        lp.close()
    return out


def rowpart_timeline(lp, eta, dist, torch, dev, iters=64, skip=8):
    """Per-phase split of one row-partitioned iteration from in-kernel timestamps (mllp_debug_trace_rowpart, dev tool):
    mean over iterations of the per-rank maximum over CTAs, then max over ranks."""
    import ctypes
    from mllp_b200 import _cabi
    L = ctypes.CDLL(_cabi.SO_PATH)
    G = lp.info()["grid_ctas"]
    buf = np.zeros(iters * G * 6, dtype=np.uint64)
    dist.barrier()
    rc = L.mllp_debug_trace_rowpart(ctypes.c_void_p(lp.handle.value), ctypes.c_double(eta), ctypes.c_double(eta),
                                    ctypes.c_int32(iters), ctypes.c_void_p(buf.ctypes.data))
    assert rc == 0, _cabi.last_error()
    t = buf.reshape(iters, G, 6).astype(np.int64)
    prev_rel = t[skip - 1:-1, :, 4].max(axis=1)            # previous iteration's release (last CTA)
    cur = t[skip:]
    seg = {
        "At_phase": (cur[:, :, 0].max(axis=1) - prev_rel),                              # slowest CTA done with A'
        "barrier_1": (cur[:, :, 1].max(axis=1) - cur[:, :, 0].max(axis=1)),
        "A_phase": (cur[:, :, 2].max(axis=1) - cur[:, :, 1].max(axis=1)),               # slowest CTA done with its rows of A
        "unpack_wait": (cur[:, :, 3].max(axis=1) - cur[:, :, 2].max(axis=1)),           # peers' words arrived and unpacked
        "barrier_2": (cur[:, :, 4].max(axis=1) - cur[:, :, 3].max(axis=1)),
        "iteration": (cur[:, :, 4].max(axis=1) - prev_rel),
    }
    vals = torch.tensor([float(v.mean()) * 1e-3 for v in seg.values()], dtype=torch.float64, device=dev)
    dist.all_reduce(vals, op=dist.ReduceOp.MAX)
    return {k: round(float(v), 3) for k, v in zip(seg.keys(), vals)}


def run_reference(args, rank, world):
    """--impl reference: the CPU oracle timed on the host cores (rank 0 only)."""
    if rank != 0:
        return
    import mllp_b200.linear_program_data as D
    from oracle import pdhg_oracle as O
    A, b, c = D.load_csr(args.workload)
    m, n = A.shape
    csr = O.CSR(A)
    nt = host_threads()
    eta = 0.9 / O.power_iteration(csr, 50, nthreads=nt)
    # bounded sample per step, sized so steps+warmup end within a few minutes
    O.pdhg_run(csr, b, c, np.zeros(n), np.zeros(m), eta, eta, 5, nthreads=nt)
    t0 = time.perf_counter()
    O.pdhg_run(csr, b, c, np.zeros(n), np.zeros(m), eta, eta, 10, nthreads=nt)
    per_it = (time.perf_counter() - t0) / 10
    budget = 120.0 / max(1, args.steps + args.warmup)
    iters = int(max(5, min(args.iters_per_step, budget / per_it)))
    x, y = np.zeros(n), np.zeros(m)
    for _ in range(args.warmup):
        x, y = O.pdhg_run(csr, b, c, x, y, eta, eta, iters, nthreads=nt)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        x, y = O.pdhg_run(csr, b, c, x, y, eta, eta, iters, nthreads=nt)
    dt = time.perf_counter() - t0
    val = args.steps * iters / dt
    info_bytes = 24 * A.nnz + 36 * m + 44 * n + 8
    line = {
        "impl": "reference", "metric": "pdhg_iterations_per_sec", "value": val, "unit": "iterations/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "netlib " + args.workload + " (_norm arrays)",
        "config": {"workload": args.workload, "m": m, "n": n, "nnz": int(A.nnz), "mode": "parity (fixed step PDHG)",
                   "iters_per_step": iters, "bytes_per_iter": info_bytes},
        "cpu_baseline": {"value": val, "unit": "iterations/s", "cores": nt, "kind": "port",
                         "sample": "%d steps x %d iterations of %s on the CPU oracle (OpenMP, all threads)" % (args.steps, iters, args.workload)},
        "e2e": {"value": val, "unit": "iterations/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "the reference has no implementation of this path (SURVEY.md section 0); this arm times the repo's CPU oracle",
    }
    emit(line)


_JSON_OUT = None
# in-run parity checks that failed: the JSON line is still printed (with "parity_failures"), then the process exits non-zero
PARITY_FAILURES = []


def parity_check(ok, msg):
    if not ok:
        PARITY_FAILURES.append(msg)
        sys.stderr.write("PARITY FAILURE: " + msg + "\n")
    return bool(ok)


def quiet_stdout():
    """The contract is ONE JSON line on stdout: libraries that write to file descriptor 1 on their own (NCCL prints
    its version there when a process group is created) are sent to stderr; the JSON line goes to the saved descriptor."""
    global _JSON_OUT
    sys.stdout.flush()
    _JSON_OUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)


def emit(line):
    out = _JSON_OUT if _JSON_OUT is not None else sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def main():
    args = parse()
    quiet_stdout()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    import mllp_b200 as M
    from mllp_b200 import _cabi

    assert torch.cuda.is_available(), "bench.py needs a GPU (there is no CPU fallback)"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    A, b0, c0 = M.load_csr(args.workload)
    m, n = A.shape
    b, c = perturbed(b0, c0, rank)
    constrs = np.split(A.indices, A.indptr)[1:-1]   # the loader's representation
    weights = A.data
    lp = M.device_lp(constrs, weights, b, c, device=local)   # built once, as in the loader
    info = lp.info()
    geom, blocks = lp.geometry(), lp.blocks_info()
    eta = 0.9 / lp.sigma_max()
    KI = args.iters_per_step
    L = _cabi.lib()
    stream = torch.cuda.current_stream(dev)
    sp = stream.cuda_stream

    bt, ct = torch.tensor(b, device=dev), torch.tensor(c, device=dev)
    xt, yt = torch.zeros(n, dtype=torch.float64, device=dev), torch.zeros(m, dtype=torch.float64, device=dev)
    scal = torch.zeros(_cabi.NUM_SCALARS, dtype=torch.float64, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)   # > 126 MB L2

    def step_device():
        xt.zero_(); yt.zero_()
        _cabi.check(L.mllp_pdhg_run(lp.handle, xt.data_ptr(), yt.data_ptr(), bt.data_ptr(), ct.data_ptr(), eta, eta,
                                    KI, scal.data_ptr(), sp), "mllp_pdhg_run")

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    for _ in range(max(args.warmup, 3)):
        step_device()
    barrier()
    sampler = ClockSampler(local)
    sampler.start()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    barrier()
    launches0 = int(L.mllp_launch_count())
    wall0 = time.perf_counter()
    for e0, e1 in evs:
        flush.fill_(1)            # evict L2 between steps (outside the event pair)
        xt.zero_(); yt.zero_()
        e0.record(stream)
        _cabi.check(L.mllp_pdhg_run(lp.handle, xt.data_ptr(), yt.data_ptr(), bt.data_ptr(), ct.data_ptr(), eta, eta,
                                    KI, scal.data_ptr(), sp), "mllp_pdhg_run")
        e1.record(stream)
    barrier()
    wall = time.perf_counter() - wall0
    launches = int(L.mllp_launch_count()) - launches0     # counted by the library at every launch site
    clocks = sampler.stop()
    dev_ms = sum(e0.elapsed_time(e1) for e0, e1 in evs)
    t = torch.tensor([dev_ms, wall * 1e3], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_ms, wall_ms = float(t[0]), float(t[1])
    total_iters = world * args.steps * KI
    value = total_iters / (dev_ms * 1e-3)
    final_scal = scal.cpu().numpy()

    # ---- end to end through the public API with host (pinned) buffers ---------------------------
    e2e = None
    if not args.no_e2e:
        def pinned(a):
            t_ = torch.empty(a.shape[0], dtype=torch.float64).pin_memory()
            t_.numpy()[:] = a
            return t_.numpy()
        hb, hc, hx0, hy0 = pinned(b), pinned(c), pinned(np.zeros(n)), pinned(np.zeros(m))
        for _ in range(2):
            M.pdhg_linear_program(constrs, weights, hb, hc, num_iters=KI, tau=eta, sigma=eta, x0=hx0, y0=hy0, device=local)
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            obj, xh, yh, inf = M.pdhg_linear_program(constrs, weights, hb, hc, num_iters=KI, tau=eta, sigma=eta,
                                                     x0=hx0, y0=hy0, device=local)
        barrier()
        e2e_s = time.perf_counter() - t0
        te = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(te, op=dist.ReduceOp.MAX)
        e2e = {"value": total_iters / float(te[0]), "unit": "iterations/s",
               "h2d_bytes_per_step": 8 * (2 * n + 2 * m), "d2h_bytes_per_step": 8 * (n + m + _cabi.NUM_SCALARS),
               "ms_per_step": 1e3 * float(te[0]) / args.steps,
               "call": "mllp_b200.pdhg_linear_program(constrs, constr_weights, rhs, coefs, num_iters=%d) with numpy (pinned) arrays; device formats cached by the loader" % KI}
        # the end-to-end call starts from the same point with the same data as the device-timed step: same scalars
        parity_check(abs(inf["pobj"] - final_scal[0]) <= 1e-9 * (1 + abs(final_scal[0])),
                     "end-to-end call and device-timed step disagree on the objective (%r vs %r)" % (inf["pobj"], final_scal[0]))

    extras = None
    if not args.no_extras:
        extras = measure_extras(M, torch, dev, local, rank, world, dist)
        if world > 1:
            extras["rowpart"] = measure_rowpart(M, torch, dist, dev, local, rank, world)

    if rank == 0:
        peak, peak_src = peaks()
        bytes_iter = info["bytes_per_iter"]
        per_launch_ms = dev_ms / args.steps
        achieved = bytes_iter * KI / (per_launch_ms * 1e-3) / 1e9
        line = {
            "metric": "pdhg_iterations_per_sec", "value": value, "unit": "iterations/s", "n_gpus": world,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": per_launch_ms, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "netlib %s (_norm arrays from the reference's dataset; ranks>0 perturb b,c)" % args.workload,
            "config": {"workload": args.workload, "m": m, "n": n, "nnz": int(A.nnz), "mode": "parity (fixed step PDHG)",
                       "iters_per_step": KI, "bytes_per_iter": bytes_iter, "parallelism": "dp%d (independent LPs, no collective)" % world,
                       "l2": "flushed between steps (256 MiB write); inside a step the iterations reuse the L2-resident matrix by design",
                       "grid_ctas": info["grid_ctas"], "threads": info["threads"], "geometry": geom["mode"],
                       "create_s": lp.create_s, "create_note": "mllp_lp_create (format build, geometry search, tuning rounds, block-kernel timing): once per instance, in the loader, outside value and e2e",
                       "block_angular": ({"blocks": blocks["blocks"], "linking_rows": blocks["linking_rows"],
                                          "ns_per_iter_grid_kernel": blocks["ns_per_iter_grid"],
                                          "ns_per_iter_block_kernel": blocks["ns_per_iter_blocks"]} if blocks["used"] else None),
                       "final_pobj": float(final_scal[0]),
                       "final_rel_kkt": float(final_scal[8])},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": measured_traffic(args.workload, KI, "k_pdhg_blocks" if blocks["used"] else "k_pdhg_persistent"),
                         "algorithmic_bytes_per_launch": bytes_iter * KI,
                         "yardstick": "ALGORITHMIC bytes (SURVEY 8d) / time against the measured HBM copy peak, as north_star defines the roofline; it is NOT measured DRAM throughput: the working set is L2 / shared-memory resident, `traffic` is the DRAM bytes ncu measured for the same launch",
                         "kernel": "%s (one launch = %d iterations)" % (
                             "k_pdhg_blocks" if blocks["used"] else "k_pdhg_persistent" if geom["mode"] == "grid" else "k_pdhg_cluster", KI),
                         "peak_source": peak_src, "frac_of_8TBs_spec": achieved / 8000.0,
                         "note": "achieved = (24 nnz + 36 m + 44 n + 8) B x iterations / CUDA-event time of the step; the working set is L2-resident so DRAM traffic is far below the algorithmic bytes"},
            "clocks": clocks, "gpu_launches": launches, "wall_ms_per_step": wall_ms / args.steps,
            "gpu_launches_note": "counted by the library (mllp_launch_count) around the timed region on rank 0: per step 4 gathers, the persistent kernel, 2 KKT kernels, 2 scatters",
        }
        if e2e is not None:
            line["e2e"] = e2e
        if extras is not None:
            line["extras"] = extras
        if world == 1 and not args.no_cpu_baseline:
            from oracle import pdhg_oracle as O  # the checker, used here only as the timed CPU baseline
            rate, its, dt, cores = cpu_oracle_rate(A, b, c, eta, args.cpu_baseline_seconds)
            r1, k1 = scipy_csr_rate(A, b, c, eta)
            line["cpu_baseline"] = {"value": rate, "unit": "iterations/s", "cores": cores, "kind": "port",
                                    "sample": "%d iterations of %s on the CPU oracle (OpenMP, %d threads, %.1f s)" % (its, args.workload, cores, dt),
                                    "scipy_csr_1_thread": {"value": r1, "unit": "iterations/s", "cores": 1,
                                                           "sample": "%d iterations of the same update with scipy.sparse CSR products on one thread" % k1}}
        line["parity_checks"] = "in-run asserts against the CPU oracle: ken-18 / pds-20 K=100, batch K=50, preconditioned batch KKT scalars, row partition K=100 on every rank (N > 1), e2e vs device objective, solve mode on osa-60 / pds-20 vs HiGHS, GNN gradients vs autograd"
        if PARITY_FAILURES:
            line["parity_failures"] = PARITY_FAILURES
        emit(line)
    nfail = torch.tensor([len(PARITY_FAILURES)], dtype=torch.int64, device=dev)
    if world > 1:
        dist.all_reduce(nfail)
        dist.destroy_process_group()
    if int(nfail[0]) > 0:
        sys.exit(3)


if __name__ == "__main__":
    main()
