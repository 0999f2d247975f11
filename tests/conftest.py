import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
# fixed-hyper-parameter cases; notebook_* fixtures (the reference's own executed outputs, full fit) have their own test module
GOLDEN_CASES = sorted(f[:-4] for f in os.listdir(GOLDEN_DIR) if f.endswith(".npz") and not f.startswith("notebook_")) if os.path.isdir(GOLDEN_DIR) else []


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu on the GPU box")
    config.addinivalue_line("markers", "slow: larger sizes")


def load_golden(name):
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"), allow_pickle=False)
    d = {k: z[k] for k in z.files}
    d["meta"] = json.loads(str(d["meta"]))
    return d


@pytest.fixture(scope="session")
def lib_built():
    """Make sure the in-tree shared library exists (builds it with nvcc if needed)."""
    import __graft_entry__ as ge

    ge.build()
    from gumbi_b200 import _lib

    return _lib.load()


@pytest.fixture(scope="session")
def engine(lib_built):
    from gumbi_b200 import GPEngine

    eng = GPEngine(0)
    yield eng
    eng.close()


def relmax(a, b):
    a, b = np.asarray(a), np.asarray(b)
    return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300))
