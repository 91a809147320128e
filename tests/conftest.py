import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    config.addinivalue_line("markers", "slow: longer CPU-only checks")
    config.addinivalue_line("markers", "gpu_unverified: GPU test of device code that has not run on hardware yet; ordered last")


# GPU tests of rows whose device code could not be run on a B200 in the round it was written (marker gpu_unverified)
# come LAST, so that a failure there cannot mask the results of the suites that were validated on hardware
# (pytest -x stops at the first failure).
def pytest_collection_modifyitems(config, items):
    tail = [it for it in items if it.get_closest_marker("gpu_unverified")]
    if tail:
        chosen = set(map(id, tail))
        items[:] = [it for it in items if id(it) not in chosen] + tail
