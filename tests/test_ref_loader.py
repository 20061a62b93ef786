"""The loader against the REFERENCE's own loader (reference linear_program_data.py:58-80): bit for bit.

oracle/_ref for this path is the reference's get_netlib_dataset itself -- pure Python, nothing to compile: it is imported
from /root/reference where that exists (tests/golden/make_ref_pins.py holds the recipe: scratch working directory with
netlib_mps/ and dataset/ linked) and compared live on all 97 instances it lists; the digests of its tuples are committed
(tests/golden/ref_loader.json) so that the comparison also runs, on the instances carried in data/, where it is absent."""
import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import make_ref_pins as R  # noqa: E402

import mllp_b200.linear_program_data as D  # noqa: E402

PINS = json.load(open(os.path.join(ROOT, "tests", "golden", "ref_loader.json")))
HAVE_REF = os.path.isdir("/root/reference/dataset/netlib_mps_norm")


def test_tuple_digests_of_carried_instances_match_the_reference():
    carried = [n for n in D.list_instances(True, os.path.join(ROOT, "data")) if n in PINS["instances"]]
    assert len(carried) >= 11
    ds, td = D.get_netlib_dataset(normalize=True, names=carried, root=os.path.join(ROOT, "data"))
    assert sorted(td.keys()) == sorted(["obj"] + carried)
    for inst in ds:
        assert R.tuple_digest(inst) == PINS["instances"][inst[0]], inst[0]


@pytest.mark.skipif(not HAVE_REF, reason="/root/reference is not present on this box (the committed digests cover the carried instances)")
def test_live_against_the_reference_loader_all_97():
    ref_ds, ref_td = R.reference_dataset()
    assert len(ref_ds) == 97
    names = [inst[0] for inst in ref_ds]
    ds, td = D.get_netlib_dataset(normalize=True, names=names, root="/root/reference/dataset")
    assert list(td.keys()) == list(ref_td.keys()) and all(v == [] for v in td.values())
    for got, ref in zip(ds, ref_ds):
        assert got[0] == ref[0] and len(got) == len(ref) == 6
        assert len(got[1]) == len(ref[1]) and all(np.array_equal(a, b) and a.dtype == b.dtype for a, b in zip(got[1], ref[1]))
        for k in (2, 3, 4, 5):
            assert got[k].dtype == ref[k].dtype and np.array_equal(got[k], ref[k]), (got[0], k)
        assert R.tuple_digest(ref) == PINS["instances"][ref[0]]
