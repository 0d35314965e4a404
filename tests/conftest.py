import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    # The suites load the in-tree libvidmem.so / libvm_oracle.so; build them (nvcc / gcc, no GPU needed)
    # when a fresh checkout does not have them yet.
    lib = os.path.join(ROOT, "real-time-brain-inspired-video-memory_b200", "libvidmem.so")
    if not os.path.exists(lib):
        import __graft_entry__
        __graft_entry__.build()


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN
