import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def golden():
    """Reference fixture data/EMresults.rda, extracted by tests/golden/make_golden.py."""
    z = np.load(os.path.join(ROOT, "tests", "golden", "emresults_chr2.npz"))
    return {k: z[k] for k in z.files}


@pytest.fixture(scope="session", autouse=True)
def _built():
    """Make sure the CPU checkers exist (cheap; no-op when already built)."""
    import subprocess
    subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "all"], check=True,
                   stdout=subprocess.DEVNULL)
