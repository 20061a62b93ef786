"""Row-partitioned large LPs (BASELINE.json configs[3]) under torchrun: in-run parity against the oracle, us/iteration of
the in-kernel NVLink exchange and of the NCCL all-gather variant, the same run's one-GPU time, and the per-phase timeline.
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P scripts/rowpart_bench.py [names...]
Rank 0 prints one JSON object; every rank asserts parity (1e-9)."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

import bench
import mllp_b200 as M

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
names = [a for a in sys.argv[1:] if not a.startswith("-")] or ["afiro", "pilot87", "osa-60", "ken-18", "pds-20"]
K = 1000 if "--quick" not in sys.argv else 300
res = bench.measure_rowpart(M, torch, dist, dev, local, rank, world, names=names, K=K, trace=True)
assert not bench.PARITY_FAILURES, bench.PARITY_FAILURES      # every rank holds the oracle's iterates to 1e-9
for nm in names:
    print("rank %d %s rowpart parity x %.2e y %.2e" % (rank, nm, res[nm]["parity_vs_oracle_K100"]["x"], res[nm]["parity_vs_oracle_K100"]["y"]), flush=True)
if rank == 0:
    print("ROWPART_JSON " + json.dumps(res), flush=True)
dist.destroy_process_group()
