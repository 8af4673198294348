import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def golden(name):
    path = os.path.join(GOLDEN, name)
    return np.load(path) if name.endswith(".npy") else np.load(path + ".npz")


def problem_from(z):
    """Rebuild the problem dict a golden fixture was generated with."""
    D = int(z["D"])
    p = {"num_vars": D, "names": [f"x{i + 1}" for i in range(D)]}
    p["bounds"] = z["bounds"].tolist() if "bounds" in z.files else [[0.0, 1.0]] * D
    if "groups" in z.files and len(z["groups"]):
        p["groups"] = [str(g) for g in z["groups"]]
    return p


@pytest.fixture(scope="session", autouse=True)
def _built_library():
    """The product has no CPU path; make sure the shared library exists before anything imports it."""
    import __graft_entry__ as g

    if not os.path.exists(g.LIB):
        g.build()
