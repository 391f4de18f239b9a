import os
import sys

import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if REPO not in sys.path:
    sys.path.insert(0, REPO)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def golden():
    import numpy as np
    path = os.path.join(REPO, "tests", "golden", "paule_golden.npz")
    return dict(np.load(path, allow_pickle=False))


@pytest.fixture(scope="session")
def golden_branches():
    """vectors of the speech-classifier / somatosensory branches from the real reference (make_branches_golden.py)"""
    import numpy as np
    path = os.path.join(REPO, "tests", "golden", "branches_golden.npz")
    return dict(np.load(path, allow_pickle=False))


@pytest.fixture(scope="session")
def golden_grad():
    """fp64 d(loss)/d(cp) of the real reference's plan_resynth, total and model-path part (make_grad_golden.py)"""
    import numpy as np
    path = os.path.join(REPO, "tests", "golden", "grad_golden.npz")
    return dict(np.load(path, allow_pickle=False))
