"""Dev tool: LPs solved / s of small-LP batches under launch-geometry knobs (MLLP_BATCH_THREADS = threads per CTA,
MLLP_BATCH_R): (a) 4096 cost-perturbed sc105 sharing one matrix, (b) 3072 perturbed sc50a / sc105 / blend with their own matrices."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import mllp_b200 as M


def timed(fn):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); fn(); e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e-3


def run(**env):
    for k, v in env.items(): os.environ[k] = str(v)
    dev = torch.device("cuda", 0)
    A, b, c = M.load_csr("sc105"); m, n = A.shape
    B = 4096
    bs = M.BatchLP([(A, A.data, b, c)], shared=True, count=B)
    g = torch.Generator(device="cpu").manual_seed(4321)
    cb = (torch.tensor(c).repeat(B, 1) * (1 + 0.05 * (2 * torch.rand(B, n, generator=g, dtype=torch.float64) - 1))).reshape(-1).to(dev)
    bb = torch.tensor(b).repeat(B).to(dev)
    etas = (0.99 / bs.sigma_max_robust()).contiguous()
    scal = torch.zeros(B * 16, dtype=torch.float64, device=dev)
    xs, ys = torch.zeros(B * n, dtype=torch.float64, device=dev), torch.zeros(B * m, dtype=torch.float64, device=dev)
    def solve_shared():
        xs.zero_(); ys.zero_()
        bs.solve(xs, ys, bb, cb, etas, scal, 1.0, 200000, 64, 1e-6)
    sec = timed(solve_shared)
    sc = scal.cpu().numpy().reshape(B, 16)
    inf = bs.info()
    print("%s\n   shared sc105 x4096: %.0f LPs/s conv %.3f mean iters %.0f -> %.3e LP-it/s  (threads %d grid %d R %d smem %d) objsum %.12g"
          % (env, B / sec, sc[:, 12].mean(), sc[:, 10].mean(), sc[:, 10].sum() / sec, inf["threads"], inf["grid_ctas"],
             inf["instances_per_cta_solve"], inf["dyn_smem_bytes"], sc[:, 0].sum()), flush=True)
    bs.close()
    insts = []
    for nm in ("sc50a", "sc105", "blend"):
        A, b, c = M.load_csr(nm)
        for i in range(1024):
            rg = np.random.default_rng(i)
            insts.append((A, A.data, b, c * (1 + 0.05 * rg.uniform(-1, 1, c.shape[0]))))
    for pre in (False, True):
        bsol = M.BatchLP(insts, precondition=pre)
        bvec = torch.tensor(np.concatenate([i[2] for i in insts]), device=dev)
        cvec = torch.tensor(np.concatenate([i[3] for i in insts]), device=dev)
        nx, ny = int(bsol.x_off[-1]), int(bsol.y_off[-1])
        etas = (0.99 / bsol.sigma_max_robust()).contiguous()
        scal = torch.zeros(len(insts) * 16, dtype=torch.float64, device=dev)
        xs, ys = torch.zeros(nx, dtype=torch.float64, device=dev), torch.zeros(ny, dtype=torch.float64, device=dev)
        def solve_all():
            xs.zero_(); ys.zero_()
            bsol.solve(xs, ys, bvec, cvec, etas, scal, 1.0, 100000, 64, 1e-6)
        sec = timed(solve_all)
        sc = scal.cpu().numpy().reshape(len(insts), 16)
        inf = bsol.info()
        print("   3072 small LPs%s: %.0f LPs/s conv %.4f mean iters %.0f -> %.3e LP-it/s (threads %d grid %d) objsum %.12g"
              % (" preconditioned" if pre else "", len(insts) / sec, sc[:, 12].mean(), sc[:, 10].mean(), sc[:, 10].sum() / sec,
                 inf["threads"], inf["grid_ctas"], sc[sc[:, 12] > 0, 0].sum()), flush=True)
        bsol.close()
    for k in env: os.environ.pop(k, None)


if __name__ == "__main__":
    for spec in (sys.argv[1:] or ["", "MLLP_BATCH_THREADS=128", "MLLP_BATCH_THREADS=64", "MLLP_BATCH_THREADS=32", "MLLP_BATCH_THREADS=32,MLLP_BATCH_R=1"]):
        run(**dict(kv.split("=") for kv in spec.split(",") if kv))
