"""2-GPU tests (skipped on a 1-GPU box): row-partitioned run and bench.py under torchrun."""
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


def _torchrun(args, port, timeout=600):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", str(port)] + args
    return subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=timeout)


def test_row_partition_two_gpus_matches_oracle():
    if _ngpu() < 2:
        pytest.skip("needs 2 GPUs")
    r = _torchrun(["scripts/dev_multi_gpu.py"], 29541)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert r.stdout.count("rowpart parity") == 10   # 5 instances x 2 ranks, each asserted < 1e-9 inside


def test_bench_two_gpus_prints_one_json_line():
    if _ngpu() < 2:
        pytest.skip("needs 2 GPUs")
    r = _torchrun(["bench.py", "--gpus", "2", "--steps", "3", "--warmup", "3", "--no-cpu-baseline"], 29542)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["n_gpus"] == 2 and d["scaling"] == "weak" and d["value"] > 0 and d["e2e"]["value"] > 0
