"""Target of the ncu captures under profiles/: sets up ONE benchmarked launch, times it with CUDA events (no profiler),
writes gpurun_out/<what>_expected.json, then repeats the same launch inside cudaProfilerStart/Stop so that
`ncu --profile-from-start off` captures exactly that launch and nothing of the handle's creation (geometry search and
tuning rounds launch the same kernels with 48 iterations; round 1's capture landed on one of those).

    python scripts/prof_target.py persistent|blocks|batch_run|batch_solve --write-expected
    ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:<kernel> -o gpurun_out/<name> \
        python scripts/prof_target.py <what>
"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import mllp_b200 as M
from mllp_b200 import _cabi

what = sys.argv[1]
dev = torch.device("cuda", 0)
L = _cabi.lib()
sp = torch.cuda.current_stream(dev).cuda_stream
meta = {"what": what}

if what in ("persistent", "blocks"):
    name = "osa-60" if what == "persistent" else "ken-18"
    A, b, c = M.load_csr(name)
    m, n = A.shape
    lp = M.DeviceLP(A, A.data, m, n, device=0)
    if what == "blocks":
        assert lp.blocks_info()["used"], lp.blocks_info()
    eta = 0.9 / lp.sigma_max()
    bt, ct = torch.tensor(b, device=dev), torch.tensor(c, device=dev)
    xt, yt = torch.zeros(n, dtype=torch.float64, device=dev), torch.zeros(m, dtype=torch.float64, device=dev)
    K = 1000
    meta.update(workload=name, iters=K, bytes_per_iter=lp.info()["bytes_per_iter"], geometry=lp.geometry()["mode"],
                kernel="k_pdhg_blocks" if what == "blocks" else "k_pdhg_persistent")

    def step():
        xt.zero_(); yt.zero_()
        _cabi.check(L.mllp_pdhg_run(lp.handle, xt.data_ptr(), yt.data_ptr(), bt.data_ptr(), ct.data_ptr(), eta, eta, K, None, sp), "run")
elif what in ("batch_run", "batch_solve"):
    name = "25fv47" if what == "batch_run" else "sc105"
    A, b, c = M.load_csr(name)
    m, n = A.shape
    B = 4096
    bs = M.BatchLP([(A, A.data, b, c)], shared=True, count=B, device=0)
    g = torch.Generator(device="cpu").manual_seed(1234)
    cb = (torch.tensor(c).repeat(B, 1) * (1 + 0.1 * (2 * torch.rand(B, n, generator=g, dtype=torch.float64) - 1))).reshape(-1).to(dev)
    bb = (torch.tensor(b).repeat(B, 1) * (1 + 0.1 * torch.rand(B, m, generator=g, dtype=torch.float64))).reshape(-1).to(dev) \
        if what == "batch_run" else torch.tensor(b).repeat(B).to(dev)
    xb, yb = torch.zeros(B * n, dtype=torch.float64, device=dev), torch.zeros(B * m, dtype=torch.float64, device=dev)
    scal = torch.zeros(B * _cabi.NUM_SCALARS, dtype=torch.float64, device=dev)
    info = bs.info()
    if what == "batch_run":
        eta = (0.9 / bs.sigma_max()).contiguous()
        K = 200
        meta.update(workload="4096 x 25fv47 (shared matrix)", iters=K, bytes_per_iter=info["bytes_per_iter"],
                    instances_per_cta=info["instances_per_cta"], kernel="k_batch_run")

        def step():
            xb.zero_(); yb.zero_()
            bs.run(xb, yb, bb, cb, eta, eta, K)
    else:
        eta = (0.99 / bs.sigma_max_robust()).contiguous()
        meta.update(workload="4096 x sc105 (shared matrix, cost-perturbed), solve to 1e-6", instances_per_cta=info["instances_per_cta_solve"],
                    kernel="k_batch_solve_warp" if info["instances_per_cta_solve"] == 4 and info["res_steps_A_solve"] == 0 else "k_batch_solve")

        def step():
            xb.zero_(); yb.zero_()
            bs.solve(xb, yb, bb, cb, eta, scal, 1.0, 200000, 64, 1e-6)
else:
    raise SystemExit("unknown target " + what)

for _ in range(2):
    step()
torch.cuda.synchronize()
ts = []
for _ in range(3):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); step(); e1.record(); torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1))
meta["expected_ms_per_launch"] = float(np.median(ts))
os.makedirs("gpurun_out", exist_ok=True)
if "--write-expected" in sys.argv:   # the run WITHOUT the profiler
    json.dump(meta, open(os.path.join("gpurun_out", "r02_%s_expected.json" % what), "w"))
print("EXPECTED", json.dumps(meta), flush=True)
torch.cuda.profiler.start()
step()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
