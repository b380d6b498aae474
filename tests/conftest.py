import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for _p in (ROOT, os.path.join(ROOT, "tests")):
    if _p not in sys.path:
        sys.path.insert(0, _p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def ref():
    """The unmodified reference minimap2 C (oracle/_ref/libmm2ref.so); built by __graft_entry__.build()."""
    from oracle import refmm2
    if not os.path.exists(refmm2.REF_SO):
        pytest.skip("oracle/_ref/libmm2ref.so not built (needs /root/reference at build time)")
    return refmm2.load_ref()
