"""Multi-GPU tests: the row-partitioned run (BASELINE.json configs[3]) at every world size the box offers (2, 4, 8) and
bench.py under torchrun.  On a box with at least two GPUs nothing here is skipped; on a one-GPU box the tests are skipped
LOUDLY (the reason names the missing hardware) -- the CPU side of the partition is covered by tests/test_distributed_cpu.py."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ngpu():
    import torch
    return torch.cuda.device_count()


def _torchrun(args, port, nproc, timeout=900):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(nproc), "--master-addr",
           "127.0.0.1", "--master-port", str(port)] + args
    return subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=timeout)


@pytest.mark.parametrize("world", [2, 4, 8])
def test_row_partition_matches_oracle(world):
    n = _ngpu()
    if n < 2:
        pytest.skip("ONE GPU on this box: the row-partitioned path needs at least two (not run, not verified here)")
    if n < world:
        pytest.skip("box has %d GPUs" % n)
    names = ["afiro", "pilot87", "osa-60", "ken-18", "pds-20"]
    r = _torchrun(["scripts/rowpart_bench.py", "--quick"] + names, 29540 + world, world)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert r.stdout.count("rowpart parity") == len(names) * world   # every rank asserted < 1e-9 inside
    line = [l for l in r.stdout.splitlines() if l.startswith("ROWPART_JSON ")]
    assert len(line) == 1
    d = json.loads(line[0][len("ROWPART_JSON "):])
    for nm in names:
        assert d[nm]["n_gpus"] == world and d[nm]["us_per_iteration"] > 0
        assert d[nm]["parity_vs_oracle_K100"]["x"] < 1e-9 and d[nm]["parity_vs_oracle_K100"]["y"] < 1e-9


def test_bench_two_gpus_prints_one_json_line():
    if _ngpu() < 2:
        pytest.skip("ONE GPU on this box: bench.py --gpus 2 not run here")
    r = _torchrun(["bench.py", "--gpus", "2", "--steps", "3", "--warmup", "3", "--no-cpu-baseline"], 29549, 2)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["n_gpus"] == 2 and d["scaling"] == "weak" and d["value"] > 0 and d["e2e"]["value"] > 0
    rp = d["extras"]["rowpart"]
    assert set(rp) == {"osa-60", "ken-18", "pds-20"} and all(v["parity_vs_oracle_K100"]["x"] < 1e-9 for v in rp.values())
