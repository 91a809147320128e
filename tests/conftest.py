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


# Order of the GPU run (pytest -x stops at the first failure): first the suites that have been green on a B200, oldest
# evidence first; then suites that feed new INPUTS to device code already validated on hardware; last (marker
# gpu_unverified) the tests of device code that could not be run on a B200 in the round it was written, so that a failure
# there cannot mask the results before it.
HARDWARE_PROVEN = ["test_gpu_parity", "test_gpu_full_size", "test_dimacs", "test_next_candidate_list", "test_next_scaling",
                   "test_io_and_adapter", "test_properties"]


def pytest_collection_modifyitems(config, items):
    def rank(it):
        if it.get_closest_marker("gpu_unverified"):
            return len(HARDWARE_PROVEN) + 1
        if not it.get_closest_marker("gpu"):
            return -1  # CPU tests keep their place in front
        name = it.module.__name__.split(".")[-1]
        return HARDWARE_PROVEN.index(name) if name in HARDWARE_PROVEN else len(HARDWARE_PROVEN)

    items.sort(key=rank)  # stable
