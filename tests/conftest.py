import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def unhex(rows):
    return np.array([[int(x, 16) for x in r] for r in rows], dtype=np.uint64).reshape(len(rows), -1)


@pytest.fixture(scope="session")
def oracle():
    from oracle import binding
    binding.build()
    return binding


@pytest.fixture(scope="session")
def ntt_params():
    return np.load(os.path.join(GOLDEN, "ntt_params.npz"))


@pytest.fixture(scope="session")
def poseidon_kat():
    return json.load(open(os.path.join(GOLDEN, "poseidon_kat.json")))


@pytest.fixture(scope="session")
def model_anchors():
    return json.load(open(os.path.join(GOLDEN, "model_anchors.json")))


@pytest.fixture(scope="session")
def oracle_commits():
    return json.load(open(os.path.join(GOLDEN, "oracle_commits.json")))


@pytest.fixture(scope="session")
def V():
    import vfhe_b200
    return vfhe_b200


@pytest.fixture(scope="session")
def ctx(V):
    """A device context; fails loudly (no skip, no fallback) if the CUDA library is unusable."""
    V.build.build()
    c = V.Context(0)
    yield c
    c.close()
