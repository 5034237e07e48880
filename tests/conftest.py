import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def golden():
    import numpy as np

    def load(name):
        return np.load(os.path.join(GOLDEN, name + ".npz"))
    return load


_MEASURED = {}


@pytest.fixture(scope="session")
def measured():
    """record(name, value): keeps the MAX of a measured error per name in gpurun_out/r2_test_measurements.json (copied to
    profiles/), so that tolerances are set from measurements instead of round numbers."""
    import json

    def record(name, value):
        _MEASURED[name] = max(float(value), _MEASURED.get(name, 0.0))
        path = os.path.join(ROOT, "gpurun_out", "r2_test_measurements.json")
        os.makedirs(os.path.dirname(path), exist_ok=True)
        with open(path, "w") as f:
            json.dump(_MEASURED, f, indent=1, sort_keys=True)
        return float(value)
    return record
