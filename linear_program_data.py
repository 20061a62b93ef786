"""Drop-in module surface for the loader: a module NAMED ``linear_program_data`` exporting ``get_netlib_dataset`` with the
reference's signature and return value (reference linear_program_data.py:58-80; bit-identical tuples, tests/test_ref_loader.py)
backed by mllp_b200.linear_program_data (which can also build the device-resident formats while loading).

Like the reference, the default instance list is the content of ``netlib_mps/`` in the working directory (:59-60) when
that directory exists and ``dataset/netlib_mps_norm/`` holds the arrays; otherwise every instance of the dataset directory.
The other loaders of the reference (random / facebook / twitch / OR-Lib set cover, SURVEY.md section 2: out of scope) resolve
lazily to the reference's own functions when a checkout is reachable (``MLLP_REFERENCE_DIR``)."""
import importlib.util
import os

from mllp_b200.linear_program_data import get_netlib_dataset as _get_netlib_dataset
from mllp_b200.linear_program_data import load_csr, load_instance  # noqa: F401

_REFERENCE_ONLY = ["get_random_dataset", "get_netlib_dataset_dense", "get_facebook_dataset", "get_twitch_dataset",
                   "get_orlib_dataset"]
__all__ = ["get_netlib_dataset", "load_csr", "load_instance"] + _REFERENCE_ONLY


def get_netlib_dataset(normalize=True, names=None, root=None, device=None):
    if names is None and os.path.isdir("netlib_mps") and os.path.isdir(os.path.join("dataset", "netlib_mps_norm")):
        names = [f for f in os.listdir("netlib_mps")]          # the reference's enumeration (:59-61), same order
        root = "dataset" if root is None else root
    return _get_netlib_dataset(normalize=normalize, names=names, root=root, device=device)


def __getattr__(name):
    if name not in _REFERENCE_ONLY:
        raise AttributeError("module %r has no attribute %r" % (__name__, name))
    path = os.path.join(os.environ.get("MLLP_REFERENCE_DIR", "/root/reference"), "linear_program_data.py")
    if os.path.exists(path) and os.path.abspath(path) != os.path.abspath(__file__):
        spec = importlib.util.spec_from_file_location("_mllp_reference_linear_program_data", path)
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        if hasattr(mod, name):
            return getattr(mod, name)

    def missing(*a, **k):
        raise ImportError("%s is outside the B200 hot path; it is taken from the reference's linear_program_data.py, "
                          "which was not found (set MLLP_REFERENCE_DIR)" % name)
    missing.__name__ = name
    return missing
