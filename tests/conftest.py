import json
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def _ensure_built():
    """The built artefacts are git-ignored (they travel to the GPU box with the snapshot); build them when a fresh checkout
    has none.  nvcc cross-compiles sm_100a without a GPU."""
    import subprocess
    pkg = os.path.join(ROOT, "clique_b200")
    if not all(os.path.exists(os.path.join(pkg, f)) for f in ("libclq.so", "libclq_host.so", "clq_align")):
        subprocess.run(["make", "-C", os.path.join(pkg, "csrc")], check=True, capture_output=True)


def pytest_sessionstart(session):
    _ensure_built()


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def goldens():
    with open(os.path.join(ROOT, "tests", "golden", "reference_goldens.json")) as f:
        return json.load(f)
